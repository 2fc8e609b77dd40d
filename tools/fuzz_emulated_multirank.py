#!/usr/bin/env python
"""Differential fuzzing of the SHARDED SEIRM engine WITHOUT a GPU: random population sizes, degrees, epidemic strengths (from
epidemics that die out at once to ones that reach everybody), isolation modes, exact / hashed filters, 2-4 emulated ranks and
the three halo-exchange paths — every run must reproduce the one-rank stages and timers bit for bit and the daily series /
gradients to 1e-6 (each rank rounds its partial sums to float32 before the all-reduce) (tests/test_seirm_emulated_multirank.py
is the harness; one OS process per rank, gloo, POSIX shared memory standing in for symmetric memory).

    python tools/fuzz_emulated_multirank.py [--cases 12] [--seed 0]
"""
import argparse
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=12)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    import numpy as np

    rng = np.random.default_rng(args.seed)
    failures = 0
    for i in range(args.cases):
        world = int(rng.choice([2, 3, 4]))
        exchange = str(rng.choice(["allgather", "sparse", "dense"]))
        per = int(rng.choice([1024, 2048]))                       # ownership ranges are aligned to 1024: the last rank owns 1..per agents
        n_agents = (world - 1) * per + int(rng.integers(1, per + 1))
        env = dict(os.environ, AT_EMU_N=str(n_agents), AT_EMU_DEG=str(float(rng.choice([1.0, 6.0, 20.0]))),
                   AT_EMU_SEED=str(int(rng.integers(1, 10**6))), AT_EMU_R=str(float(rng.choice([0.5, 9.0, 40.0]))),
                   AT_EMU_EXPOSED=str(float(rng.choice([0.0, 0.002, 0.03, 0.3]))), AT_EMU_ISOLATION=str(rng.choice(["hard", "soft", "none"])),
                   AT_EMU_HASHED=str(int(rng.random() < 0.4)))
        desc = {k[7:].lower(): v for k, v in env.items() if k.startswith("AT_EMU_")}
        desc.update(world=world, exchange=exchange)
        code = (f"import sys, os, torch, numpy as np; sys.path.insert(0, {ROOT!r}); import tests.test_seirm_emulated_multirank as M\n"
                f"import torch.multiprocessing as mp, tempfile\n"
                f"out = tempfile.mkdtemp(); lib = sys.argv[1]\n"
                f"mp.spawn(M._episode, args=(1, 0, out, lib, 'single', 'fz'), nprocs=1, join=True)\n"
                f"mp.spawn(M._episode, args=({world}, M._free_port(), out, lib, {exchange!r}, 'fz' + str(os.getpid())), nprocs={world}, join=True)\n"
                f"one = torch.load(os.path.join(out, 'single_w1_r0.pt'))\n"
                f"parts = [torch.load(os.path.join(out, '{exchange}_w{world}_r%d.pt' % r)) for r in range({world})]\n"
                f"ok = all(torch.equal(torch.cat([p[k] for p in parts]), one[k]) for k in ('stage', 'nst', 'itime')) and np.allclose(parts[0]['di'].numpy(), one['di'].numpy(), rtol=1e-6, atol=0)\n"
                f"ok = ok and np.allclose(parts[0]['gr'].numpy(), one['gr'].numpy(), rtol=1e-6) and np.allclose(parts[0]['dd'].numpy(), one['dd'].numpy(), rtol=1e-6)\n"
                f"print('RESULT', ok, float(one['di'].sum()))\n")
        res = subprocess.run([sys.executable, "-c", code, LIB], capture_output=True, text=True, env=env, timeout=1200)
        line = [ln for ln in res.stdout.splitlines() if ln.startswith("RESULT")]
        ok = bool(line) and line[0].split()[1] == "True"
        failures += not ok
        print(f"[{i}] {'ok' if ok else 'FAIL'}  {line[0] if line else res.stderr[-400:]}  {desc}", flush=True)
    print(f"{args.cases - failures}/{args.cases} sharded runs reproduce the one-rank result")
    return 1 if failures else 0


if __name__ == "__main__":
    LIB = os.path.join(tempfile.mkdtemp(prefix="at_fuzz_mr_"), "libat_seirm_emulated.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "host_emulation", "seirm_emulation.cpp"), "-o", LIB], check=True)
    sys.exit(main())
