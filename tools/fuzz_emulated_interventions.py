#!/usr/bin/env python
"""Differential fuzzing of the quarantine / testing / SEIRMS-progression kernels WITHOUT a GPU: random population sizes (tails
of 1-3 agents), statuses (negative and fractional counters included), dates around the current step, probabilities 0 / 1 /
in between, fused in-place and split policy/transition calls, float64 stages, injected draws — through the torch ops on the
host-compiled library (tests/host_emulation/iv_emulation.cpp) against the CPU oracle, bit for bit.

    python tools/fuzz_emulated_interventions.py [--cases 300] [--seed 0]      (test infrastructure: uses oracle/)
"""
import argparse
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=300)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    from agenttorch_b200 import _native_iv
    from agenttorch_b200 import interventions as ops
    from oracle import noise as onoise
    from oracle.interventions_port import SeirmsPort, quarantine_step, testing_step
    from oracle.seirm_port import SeirmState

    lib = os.path.join(tempfile.mkdtemp(prefix="at_fuzz_iv_"), "libat_interventions_emulated.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "host_emulation", "iv_emulation.cpp"), "-o", lib], check=True)
    _native_iv.LIB_PATH, _native_iv._lib = lib, None
    ops._require_cuda = lambda name, t: None
    ops._stream = lambda t: None
    rng = np.random.default_rng(args.seed)
    f = np.float32
    failures = 0
    for i in range(args.cases):
        n = int(rng.choice([1, 2, 3, 5, 31, 64, 257, 1000, 4099]))
        seed, id0 = int(rng.integers(1, 2**31)), int(rng.choice([0, 7, 1024, 2**33]))
        p_start, p_break, p_comply = (float(rng.choice([0.0, 1.0, rng.random()])) for _ in range(3))
        days, inelig, delay, inf_t = int(rng.choice([0, 1, 4, 12])), int(rng.choice([0, 2])), int(rng.choice([0, 1, 3])), 130
        stage = rng.choice([0.0, 1.0, 2.0, 3.0, 4.0, 0.37], size=n).astype(f)
        q = rng.choice([-2.0, -1.0, 0.0, 0.0, 1.0, 0.5], size=n).astype(f)
        d = rng.choice([0.0, 1.0, 2.0, 3.0, 130.0], size=n).astype(f)
        aw = rng.choice([0.0, 0.0, 1.0, -1.0], size=n).astype(f)
        rd = rng.choice([2.0, 3.0, 4.0, 5.0, 130.0], size=n).astype(f)
        re = rng.choice([0.0, 3.0, 4.0, 5.0, 6.0, 130.0], size=n).astype(f)
        split, f64, inject = bool(rng.random() < 0.5), bool(rng.random() < 0.3), bool(rng.random() < 0.3)
        ids = np.arange(n) + id0
        tq, td, taw, trd, tre = (torch.from_numpy(x.copy()) for x in (q, d, aw, rd, re))
        bad = []
        for t in (2, 3, 4):
            us, ub = onoise.stream_uniforms(seed, t, ids, onoise.STREAM_QUARANTINE)
            uc = onoise.stream_uniforms(seed, t, ids, onoise.STREAM_TESTING)[0]
            q, d = quarantine_step(t, stage, q, d, us, ub, p_start, p_break, days, inf_t)
            aw, rd, re = testing_step(t, q, aw, rd, re, uc, p_comply, inelig, delay, inf_t)
            ts = torch.from_numpy(stage.astype(np.float64) if f64 else stage)
            kseed = 99 if inject else seed                       # injected draws must make the seed irrelevant
            if split:
                ius, iub, iuc = (torch.from_numpy(x) for x in (us, ub, uc)) if inject else (None, None, None)
                start, _ = ops.quarantine_decide(ts, tq, p_start, 0.0, t, kseed, id0, ius, None)
                _, brk = ops.quarantine_decide(ts, tq, 0.0, p_break, t, kseed, id0, None, iub)
                tq, td = ops.quarantine_update(tq, td, start, brk, t, float(days), float(inf_t))
                acc = ops.testing_accept(tq, p_comply, t, kseed, id0, iuc)
                taw, trd, tre = ops.testing_update(taw, trd, tre, acc, t, float(inelig), float(delay), float(inf_t))
            else:
                ops.quarantine_step(ts, tq, td, p_start, p_break, t, seed, id0, float(days), float(inf_t))
                ops.testing_step(tq, taw, trd, tre, p_comply, t, seed, id0, float(inelig), float(delay), float(inf_t))
            for name, got, want in (("is_q", tq, q), ("start_date", td, d), ("awaiting", taw, aw), ("result_date", trd, rd), ("re_eligible", tre, re)):
                if not np.array_equal(got.numpy(), want):
                    bad.append(f"{name}@t={t}")
        # SEIRMS progression of the same population
        t = 5
        nst = rng.choice([3.0, 5.0, 6.0, 9.0, 41.0], size=n).astype(f)
        m = float(rng.choice([0.0, 1.0, rng.random()]))
        if n >= 2:
            port = SeirmsPort(np.zeros(n), stage, np.zeros((0, 2), dtype=np.int64))
            u = onoise.stream_uniforms(seed, t, ids, onoise.STREAM_MORTALITY)[0]
            ref = port.progression_seirms(SeirmState(torch.from_numpy(stage.astype(np.float64)), torch.from_numpy(nst), torch.zeros(n),
                                                     torch.zeros(21), torch.zeros(21)), t, torch.tensor([m]), u)
            s2, n2, dd2, code, counts = ops.seirms_progression(torch.from_numpy(stage), torch.from_numpy(nst), torch.zeros(21), torch.tensor([m]), t,
                                                               seed, id0, torch.from_numpy(u) if inject else None, 5.0, 41.0, 100.0)
            if not np.array_equal(s2.numpy().astype(np.float64), ref.stage.numpy()):
                bad.append("seirms stage")
            if not np.array_equal(n2.numpy(), ref.next_stage_time.numpy()):
                bad.append("seirms next_stage_time")
            if not torch.equal(dd2, ref.daily_deaths):
                bad.append("seirms daily_deaths")
        failures += bool(bad)
        if bad or i % 50 == 0:
            print(f"[{i}] {'FAIL ' + ', '.join(bad) if bad else 'ok'}  n={n} split={split} f64={f64} inject={inject} p=({p_start:.2f},{p_break:.2f},{p_comply:.2f}) "
                  f"days={days} inelig={inelig} delay={delay} m={m:.2f} id0={id0}", flush=True)
    print(f"{args.cases - failures}/{args.cases} cases agree with the oracle")
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
