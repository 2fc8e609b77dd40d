#!/usr/bin/env python
"""Differential fuzzing of the predator-prey kernels WITHOUT a GPU: random grids (degenerate shapes included), populations,
energies, countdown dtypes, storage layouts (plain / tiled + sorted, re-sorting), chunked runs and CSR adjacency through the real
engine on the host-compiled library (tests/host_emulation/pp_emulation.cpp) against the CPU oracle, bit for bit.

    python tools/fuzz_emulated_predprey.py [--cases 100] [--seed 0]        (test infrastructure: uses oracle/, never the product path)
"""
import argparse
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    from agenttorch_b200 import _native_pp, predprey
    from oracle.predprey_port import PredPreyConstants, PredPreyPort
    from tests.helpers import PREDPREY_KEYS

    lib = os.path.join(tempfile.mkdtemp(prefix="at_fuzz_pp_"), "libat_predprey_emulated.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "host_emulation", "pp_emulation.cpp"), "-o", lib], check=True)
    _native_pp.LIB_PATH, _native_pp._lib = lib, None
    predprey._require_cuda = lambda what, is_cuda: None
    predprey._stream = lambda: None
    rng = np.random.default_rng(args.seed)
    failures = 0
    for i in range(args.cases):
        mx, my = int(rng.choice([1, 2, 3, 7, 16, 17, 40, 65])), int(rng.choice([1, 2, 5, 16, 31, 64]))
        if mx * my == 1:
            my = 2                    # a single cell has no neighbour: the reference's random.choice([]) raises (move.py:88-90)
        C = mx * my
        n_pred, n_prey = int(rng.choice([0, 1, 5, 50, 400])), int(rng.choice([0, 1, 7, 90, 1500]))
        f64 = bool(rng.random() < 0.5)
        tiled = bool(rng.random() < 0.5)
        T = int(rng.choice([1, 3, 12]))
        chunks = bool(rng.random() < 0.3)
        init = {
            "pred_pos": np.stack([rng.integers(0, mx, n_pred), rng.integers(0, my, n_pred)], 1).reshape(n_pred, 2),
            "prey_pos": np.stack([rng.integers(0, mx, n_prey), rng.integers(0, my, n_prey)], 1).reshape(n_prey, 2),
            "pred_energy": rng.uniform(-2, 30, (n_pred, 1)).astype(np.float32),
            "prey_energy": rng.uniform(-2, 40, (n_prey, 1)).astype(np.float32),
        }
        cd = np.where(rng.random(C) < 0.4, 0.0, rng.uniform(0.01, 12.0, C))
        init["growth_countdown"] = cd.reshape(-1, 1).astype(np.float64 if f64 else np.float32)
        init["growth_stage"] = (init["growth_countdown"] <= 0).astype(np.int64)
        consts = dict(pred_work=float(rng.choice([0.0, 1.5])), prey_work=float(rng.choice([0.0, 2.25])), prey_nutrition=6.5,
                      grass_nutrition=3.3, regrowth_time=float(rng.choice([1.0, 7.7])))
        desc = dict(grid=(mx, my), n_pred=n_pred, n_prey=n_prey, f64=f64, tiled=tiled, T=T, chunks=chunks, **consts)
        t0 = time.time()
        try:
            port = PredPreyPort(PredPreyConstants(mx, my, **consts), **init)
            seed = int(rng.integers(1, 2**31))
            for t in range(T):
                port.step(t, seed)
            spec = predprey.PredPreySpec(mx, my, C, mx, tiled=tiled, **consts)
            eng = predprey.PredPreyEngine(spec, "cpu")
            t_ = lambda k: torch.from_numpy(np.ascontiguousarray(init[k]))                  # noqa: E731
            st = predprey.PredPreyState.from_reference(spec, t_("pred_pos"), t_("pred_energy"), t_("prey_pos"), t_("prey_energy"),
                                                       t_("growth_stage"), t_("growth_countdown"))
            if tiled:
                st.sort_by_cell(spec)
            if chunks:
                done = 0
                for chunk in (1, 2, T):
                    chunk = min(chunk, T - done)
                    if chunk > 0:
                        eng.run(st, chunk, seed, t0=done, cuda_graph=False, resort_every=2 if tiled else 0)
                        done += chunk
            else:
                eng.run(st, T, seed, cuda_graph=False, resort_every=int(rng.choice([0, 1, 4])) if tiled else 0)
            got, want = st.to_reference(spec), port.snapshot()
            bad = [k for k in PREDPREY_KEYS if not (got[k].numpy().dtype == np.asarray(want[k]).dtype and np.array_equal(got[k].numpy(), want[k]))]
            census = eng.census(st)
            if census != (int((port.pred_energy > 0).sum()), int((port.prey_energy > 0).sum()), int(port.growth_stage.sum())):
                bad.append("census")
        except Exception as exc:
            bad = [f"{type(exc).__name__}: {exc}"]
        failures += bool(bad)
        print(f"[{i}] {'FAIL ' + ', '.join(bad) if bad else 'ok'}  {time.time() - t0:.2f}s  {desc}", flush=True)
    print(f"{args.cases - failures}/{args.cases} cases agree with the oracle")
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
