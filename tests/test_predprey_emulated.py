"""The predator-prey engine (agenttorch_b200/predprey.py) driven against csrc/at_predprey.cu COMPILED FOR THE HOST
(tests/host_emulation/pp_emulation.cpp: same kernels, same C-ABI wrappers, host pointers): the kernel source and the engine's
launch protocol are checked against the reference's golden trajectories and the oracle on a machine without a GPU.  The
-m gpu tests (tests/test_gpu_predprey.py) run the same comparisons on the real library."""
import dataclasses
import os
import subprocess

import numpy as np
import pytest
import torch

from agenttorch_b200 import _native_pp, predprey
from tests.helpers import PREDPREY_CASES, PREDPREY_KEYS, load_predprey_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emulated_library(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("pp_emu") / "libat_predprey_emulated.so")
    cmd = ["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "host_emulation", "pp_emulation.cpp"), "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return out


@pytest.fixture()
def engine_on_host(emulated_library, monkeypatch):
    """Point the ctypes binding at the emulated library and lift the engine's CUDA checks for the duration of a test."""
    monkeypatch.setattr(_native_pp, "LIB_PATH", emulated_library)
    monkeypatch.setattr(_native_pp, "_lib", None)
    monkeypatch.setattr(predprey, "_require_cuda", lambda what, is_cuda: None)
    monkeypatch.setattr(predprey, "_stream", lambda: None)
    yield
    monkeypatch.setattr(_native_pp, "_lib", None)          # the next user loads the real library again


def _spec(g):
    pw, yw, pn, gn, rt = (float(v) for v in g["consts"])
    ap, ay = predprey.PredPreySpec.work_flags(g["update_outputs"])
    mx, my = int(g["bounds"][0]), int(g["bounds"][1])
    return predprey.PredPreySpec(mx, my, g["n_cells"], mx, pw, yw, pn, gn, rt, ap, ay)


def _state(spec, src, prefix="init_"):
    t = lambda k: torch.from_numpy(np.ascontiguousarray(src[prefix + k]))      # noqa: E731
    return predprey.PredPreyState.from_reference(spec, t("pred_pos"), t("pred_energy"), t("prey_pos"), t("prey_energy"),
                                                 t("growth_stage"), t("growth_countdown"))


def _assert_equal(st, spec, want, where):
    got = st.to_reference(spec)
    for k in PREDPREY_KEYS:
        a, b = got[k].numpy(), np.asarray(want[k])
        assert a.dtype == b.dtype, (where, k, a.dtype, b.dtype)
        assert np.array_equal(a, b), (where, k, int((a != b).sum()))


@pytest.mark.parametrize("case", PREDPREY_CASES)
@pytest.mark.parametrize("mode", ["substeps", "substeps_csr", "fused", "fused_chunks", "fused_sorted_tiled"])
def test_kernel_source_matches_reference_trajectory(engine_on_host, case, mode):
    g = load_predprey_golden(case)
    spec = _spec(g)
    if mode == "fused_sorted_tiled":
        spec = dataclasses.replace(spec, tiled=True)
    adj = predprey.CsrAdjacency(torch.from_numpy(g["adjacency"].astype(np.int64)), "cpu") if mode == "substeps_csr" else None
    eng = predprey.PredPreyEngine(spec, "cpu", adjacency=adj)
    st = _state(spec, g)
    T, seed = g["steps"], g["seed"]
    want = lambda t: {k: g["traj_" + k][t] for k in PREDPREY_KEYS}             # noqa: E731
    if mode.startswith("substeps"):
        for t in range(T):
            eng.move(st, t, seed); eng.eat(st); eng.hunt(st); eng.grow(st)
            _assert_equal(st, spec, want(t), (case, mode, t))
    elif mode == "fused_chunks":
        done = 0
        for chunk in (1, 3, 2, 5, T):
            chunk = min(chunk, T - done)
            if chunk <= 0:
                break
            eng.run(st, chunk, seed, t0=done, cuda_graph=False)
            done += chunk
            _assert_equal(st, spec, want(done - 1), (case, mode, done))
    else:
        if mode == "fused_sorted_tiled":
            st.sort_by_cell(spec)
        eng.run(st, T, seed, cuda_graph=False, resort_every=5 if mode == "fused_sorted_tiled" else 0)
        _assert_equal(st, spec, want(T - 1), (case, mode))
        alive_pred, alive_prey, grown = eng.census(st)
        ref = want(T - 1)
        assert alive_pred == int((ref["pred_energy"] > 0).sum()) and alive_prey == int((ref["prey_energy"] > 0).sum())
        assert grown == int(np.asarray(ref["growth_stage"]).sum())


@pytest.mark.parametrize("shape,n_pred,n_prey,f64", [((30, 50), 400, 1600, True), ((17, 9), 301, 503, False)])
def test_kernel_source_matches_oracle_on_seeded_grids(engine_on_host, shape, n_pred, n_prey, f64):
    from oracle.predprey_port import PredPreyConstants, PredPreyPort

    mx, my = shape
    rng = np.random.default_rng(mx * 7 + my)
    C = mx * my
    init = {
        "pred_pos": np.stack([rng.integers(0, mx, n_pred), rng.integers(0, my, n_pred)], 1),
        "prey_pos": np.stack([rng.integers(0, mx, n_prey), rng.integers(0, my, n_prey)], 1),
        "pred_energy": rng.uniform(-2, 30, (n_pred, 1)).astype(np.float32),
        "prey_energy": rng.uniform(-2, 40, (n_prey, 1)).astype(np.float32),
    }
    cd = np.where(rng.random(C) < 0.4, 0.0, rng.uniform(0.01, 12.0, C))
    init["growth_countdown"] = cd.reshape(-1, 1).astype(np.float64 if f64 else np.float32)
    init["growth_stage"] = (init["growth_countdown"] <= 0).astype(np.int64)
    consts = dict(pred_work=1.5, prey_work=2.25, prey_nutrition=6.5, grass_nutrition=3.3, regrowth_time=7.7)
    port = PredPreyPort(PredPreyConstants(mx, my, **consts), **init)
    spec = predprey.PredPreySpec(mx, my, C, mx, tiled=f64, **consts)
    eng = predprey.PredPreyEngine(spec, "cpu")
    st = _state(spec, init, prefix="")
    T, seed = 12, 2026
    for t in range(T):
        port.step(t, seed)
    if f64:
        st.sort_by_cell(spec)
    eng.run(st, T, seed, cuda_graph=False, resort_every=4 if f64 else 0)
    _assert_equal(st, spec, port.snapshot(), (shape, "fused"))


def test_the_real_binding_is_restored_after_emulation():
    assert _native_pp.LIB_PATH.endswith(os.path.join("lib", "libat_predprey.so"))
    assert predprey._require_cuda.__module__ == "agenttorch_b200.predprey"
    with pytest.raises(RuntimeError, match="CUDA"):
        predprey.PredPreyEngine(None, "cpu")
