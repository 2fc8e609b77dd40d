#!/usr/bin/env python
"""Differential fuzzing of the SEIRM kernels WITHOUT a GPU: random small populations, graphs (hubs, self-loops, duplicates,
isolated agents, rows longer than a TMA chunk), layouts and noise modes are run through the real engine on the emulated library
(tests/host_emulation/) and compared with the CPU oracle — stages / timers bit for bit, series 1e-5, gradients 1e-4.

    python tools/fuzz_emulated.py [--cases 50] [--seed 0]            (test infrastructure: uses oracle/, never the product path)
"""
import argparse
import contextlib
import os
import subprocess
import sys
import tempfile
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def build_library():
    out = os.path.join(tempfile.mkdtemp(prefix="at_fuzz_"), "libat_seirm_emulated.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "host_emulation", "seirm_emulation.cpp"), "-o", out], check=True)
    return out


def random_case(rng):
    n = int(rng.choice([33, 64, 97, 256, 500, 1024, 1500, 2500]))
    deg = float(rng.choice([0.3, 2.0, 8.0, 25.0, 60.0]))
    e = max(0, int(n * deg))
    src, dst = rng.integers(0, n, e), rng.integers(0, n, e)
    if e and rng.random() < 0.5:                                 # hubs: a few destinations collect a large share of the edges
        hubs = rng.integers(0, n, 3)
        pick = rng.random(e) < rng.choice([0.05, 0.3])
        dst = np.where(pick, rng.choice(hubs, e), dst)
    if e and rng.random() < 0.3:                                 # a block of isolated agents
        dst = np.where(dst >= n - n // 4, rng.integers(0, max(1, n - n // 4), e), dst)
    if e and rng.random() < 0.3:
        src[: e // 10] = dst[: e // 10]                          # self-loops
    edges = np.stack([src, dst], axis=1).astype(np.int64)
    fe, fi = rng.choice([0.0, 0.02, 0.2]), rng.choice([0.0, 0.02, 0.3])
    r = rng.random(n)
    stage0 = np.zeros(n)
    stage0[r < fe] = 1.0
    stage0[(r >= fe) & (r < fe + fi)] = 2.0
    weights = rng.uniform(0.1, 3.0, e).astype(np.float32) if e and rng.random() < 0.3 else None
    layout = rng.choice(["fused", "blocked", "hashed32", "hashed64", "stream", "stream", "stream"] if weights is None else
                        ["fused", "blocked", "hashed32", "hashed64"])
    kw = {"fused": {"layout": "bcsr"}, "blocked": {"layout": "bcsr", "block_size": 1024}, "hashed32": {"hashed": True, "hash_words": 32},
          "hashed64": {"hashed": True, "hash_words": 64},
          "stream": {"layout": "stream", "block_size": int(rng.choice([1024, 2048, 4096])),
                     "tile_rows": int(rng.choice([32, 64, 256, 2048])), "bank_aware": bool(rng.random() < 0.7),
                     "grid": int(rng.choice([0, 1, 3, 5]))}}[layout]
    if layout not in ("blocked", "stream") and rng.random() < 0.5:
        kw["edges_per_chunk"] = int(rng.choice([512, 768, 1024]))
    if rng.random() < 0.3:
        kw["tasks_per_group"] = int(rng.choice([1, 2, 8]))
    return dict(n=n, deg=deg, edges=edges, age=rng.integers(0, 6, n), stage0=stage0, weights=weights, layout=layout, kw=kw,
                mode=str(rng.choice(["hard", "soft", "none"])), packed=bool(rng.random() < 0.5), tensor_noise=bool(rng.random() < 0.3),
                r2=(rng.uniform(0.05, 6.0, (12, 1)) * rng.choice([0.2, 1.0, 3.0])).astype(np.float32),
                m=rng.uniform(0.01, 0.9, 1).astype(np.float32), seed=int(rng.integers(1, 2**31)), warps=int(rng.choice([1, 2, 4])),
                T=int(rng.choice([12, 21, 21, 43, 60, 64])), e2i=int(rng.choice([1, 5, 9])), i2r=int(rng.choice([1, 5, 12])),
                sfs=tuple(np.round(rng.uniform(0.2, 2.0, 9), 2).tolist()) if rng.random() < 0.5 else None,
                sfi=(0.0, float(np.round(rng.uniform(0.0, 1.0), 2)), float(np.round(rng.uniform(0.1, 1.0), 2)), 0.0, 0.0) if rng.random() < 0.5 else None)


def run_case(c):
    from agenttorch_b200.engine import NoiseSpec, SeirmEngine, SeirmSpec, SeirmState
    from agenttorch_b200.graph import build_contact_graph, mean_interactions_from_age
    from oracle.seirm_port import SeirmPort, episode_with_grads
    from tests.helpers import make_noise, rel_err

    from oracle.seirm_port import SeirmConstants

    T, n = c["T"], c["n"]
    weeks = (T + 6) // 7
    r2 = c["r2"][:weeks]
    consts = SeirmConstants(num_steps=T, num_weeks=weeks, exposed_to_infected_time=c["e2i"], infected_to_recovered_time=c["i2r"])
    spec_kw = dict(num_steps=T, exposed_to_infected_time=c["e2i"], infected_to_recovered_time=c["i2r"])
    if c["sfs"] is not None:
        consts.sf_susceptibility = c["sfs"]
        spec_kw["sfs"] = c["sfs"]
    if c["sfi"] is not None:
        consts.sf_infector = c["sfi"]
        spec_kw["sfi"] = c["sfi"]
    U, W = make_noise(c["seed"], c["mode"], n, T)
    gI, gD = np.linspace(0.5, 1.5, T, dtype=np.float32), np.linspace(2.0, 1.0, T, dtype=np.float32)
    try:
        ref = episode_with_grads(SeirmPort(c["age"], c["stage0"], c["edges"], consts=consts, edge_weight=c["weights"]), W, U, gI, gD, r2=r2,
                                 m=c["m"], num_steps=T)
    except IndexError:
        return None, -1.0          # the reference itself indexes past lam_gamma_integrals (an infection older than the episode)
    ei = torch.from_numpy(c["edges"]).t().contiguous().reshape(2, -1)
    attr = None if c["weights"] is None else torch.stack((torch.ones(len(c["edges"])), torch.from_numpy(c["weights"])))
    graph = build_contact_graph(ei, n, edge_attr=attr, mean_interactions=mean_interactions_from_age(torch.from_numpy(c["age"])))
    eng = SeirmEngine(graph, SeirmSpec(**spec_kw), warps_per_cta=c["warps"], **c["kw"])
    st = SeirmState.from_initial(c["stage0"].astype(np.float32), c["age"], T, torch.device("cpu"), packed=c["packed"])
    if c["tensor_noise"]:
        noise = NoiseSpec(seed=0, isolation="tensor", uniform="tensor", isolation_tensor=torch.from_numpy(W), uniform_tensor=torch.from_numpy(U))
    else:
        noise = NoiseSpec(seed=c["seed"], isolation=c["mode"])
    r2_t = torch.tensor(r2, requires_grad=True)
    eye = torch.eye(weeks)
    r_steps = torch.stack([(r2_t * eye[int(t / 7)]).sum() for t in range(T)])
    m_t = torch.tensor(c["m"])
    sv = eng.run_episode(st, r_steps.detach(), m_t, T, noise, want_grad=True)
    gr, gm = eng.episode_backward(sv, T, torch.from_numpy(gI), torch.from_numpy(gD), m_t)
    eng.check_errors()
    r_steps.backward(gr)
    problems = []
    if not np.array_equal(st.stage.numpy().astype(np.float64), ref["stage"]):
        problems.append("stage (%d agents differ)" % int((st.stage.numpy().astype(np.float64) != ref["stage"]).sum()))
    if not np.array_equal(st.next_stage_time.numpy().astype(np.float32), ref["next_stage_time"]):
        problems.append("next_stage_time")
    if not np.array_equal(st.infected_time.numpy().astype(np.float32), ref["infected_time"]):
        problems.append("infected_time")
    for key, got, tol in (("daily_infected", st.daily_infected.numpy(), 1e-5), ("daily_deaths", st.daily_deaths.numpy(), 1e-5),
                          ("grad_R2", r2_t.grad.numpy(), 1e-4), ("grad_M", gm.numpy(), 1e-4)):
        if np.abs(ref[key]).max() > 0 or np.abs(got).max() > 0:
            err = rel_err(ref[key], got)
            if not err <= tol:
                problems.append(f"{key} rel err {err:.2e}")
    return problems, float(ref["daily_infected"].sum())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=50)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    from agenttorch_b200 import _native, engine

    _native.LIB_PATH, _native._lib = build_library(), None
    engine._require_cuda = lambda message, is_cuda: None
    torch.cuda.device = lambda device: contextlib.nullcontext()
    torch.cuda.current_stream = lambda device=None: types.SimpleNamespace(cuda_stream=0)
    torch.cuda.synchronize = lambda device=None: None
    rng = np.random.default_rng(args.seed)
    failures = 0
    for i in range(args.cases):
        c = random_case(rng)
        t0 = time.time()
        desc = {k: c[k] for k in ("n", "deg", "T", "e2i", "i2r", "layout", "kw", "mode", "packed", "tensor_noise", "seed", "warps")}
        desc["tables"] = (c["sfs"] is not None, c["sfi"] is not None)
        desc["weighted"], desc["edges"] = c["weights"] is not None, len(c["edges"])
        try:
            problems, infected = run_case(c)
        except Exception as exc:                                    # refusals (e.g. too many in-edges per block) are reported too
            problems, infected = [f"{type(exc).__name__}: {exc}"], -1.0
        if problems is None:
            print(f"[{i}] skipped: the reference raises IndexError on this input  {desc}", flush=True)
            continue
        failures += bool(problems)
        print(f"[{i}] {'FAIL ' + '; '.join(problems) if problems else 'ok'}  infected={infected:.1f}  {time.time() - t0:.1f}s  {desc}", flush=True)
    print(f"{args.cases - failures}/{args.cases} cases agree with the oracle")
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
