"""Bench line for BASELINE configs[3]: predator-prey grid model, 10 M agents on a 4096 x 4096 grid, one B200.

Same contract as bench.py (one JSON line; `value` = device-resident, `e2e` = through the Runner API with the initial
population copied from pinned host memory and the alive counts read back every step), kept apart from bench.py because
the headline metric of this repo is the SEIRM episode.  A "step" here is one 20-step episode.

    python tools/bench_predprey.py [--steps K] [--warmup W] [--side 4096] [--predators 2000000] [--prey 8000000]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import ClockSampler  # noqa: E402

EPISODE = 20


def cpu_port_sample(side, n_pred, n_prey, steps=2):
    """oracle/predprey_port.py (numpy, one thread for the hot loops) on the same workload: seconds per step."""
    from oracle.predprey_port import PredPreyConstants, PredPreyPort

    rng = np.random.default_rng(0)
    C = side * side
    cd = np.where(rng.random(C) < 0.4, 0.0, rng.integers(1, 101, C)).astype(np.float32)
    port = PredPreyPort(PredPreyConstants(side, side), np.stack([rng.integers(0, side, n_pred), rng.integers(0, side, n_pred)], 1),
                        rng.uniform(30, 101, (n_pred, 1)), np.stack([rng.integers(0, side, n_prey), rng.integers(0, side, n_prey)], 1),
                        rng.uniform(40, 101, (n_prey, 1)), (cd <= 0).astype(np.int64), cd)
    t0 = time.perf_counter()
    for s in range(steps):
        port.step(s, 7)
    return (time.perf_counter() - t0) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--side", type=int, default=4096)
    ap.add_argument("--predators", type=int, default=2_000_000)
    ap.add_argument("--prey", type=int, default=8_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("bench_predprey.py needs a CUDA device (no CPU fallback)")
    from agenttorch_b200.models import predator_prey as pp
    from agenttorch_b200.models.predator_prey.synthetic import synthetic_config

    dev = torch.device("cuda:0")
    n = args.predators + args.prey
    C = args.side * args.side
    cfg = synthetic_config(args.side, args.side, args.predators, args.prey, seed=20261019, num_steps=EPISODE, device="cuda",
                           grass_coordinates=False)
    runner = pp.get_runner(cfg, pp.registry)
    runner.init()
    s = runner.state
    live = {"agents": {"predator": {k: s["agents"]["predator"][k] for k in ("coordinates", "energy")},
                       "prey": {k: s["agents"]["prey"][k] for k in ("coordinates", "energy")}},
            "objects": {"grass": {k: s["objects"]["grass"][k] for k in ("growth_stage", "growth_countdown")}}}
    # the initial population as a host snapshot in compact dtypes (int16 coordinates, uint8 stage): what a user restores per episode
    narrow = {"coordinates": torch.int16 if args.side <= 32767 else torch.int32, "growth_stage": torch.uint8}
    host = {a: {b: {k: v.to(narrow.get(k, v.dtype)).cpu().pin_memory() for k, v in d.items()} for b, d in grp.items()} for a, grp in live.items()}
    dev0 = {a: {b: {k: v.clone() for k, v in d.items()} for b, d in grp.items()} for a, grp in live.items()}
    h2d = sum(v.numel() * v.element_size() for grp in host.values() for d in grp.values() for v in d.values())
    counts = torch.zeros(3, dtype=torch.int64).pin_memory()

    def restore(src):
        runner.restore_state(src)
        for m in runner._pp_plan.movers if getattr(runner, "_pp_plan", None) else []:
            m.calls = 0

    def episode(src, read_back):
        restore(src)
        runner.step(EPISODE)
        if read_back:
            plan = runner._pp_plan
            c = plan.engine.census(plan.st)
            counts.copy_(torch.tensor(c))
        return None

    def timed(fn, k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k

    for _ in range(max(args.warmup, 3)):
        episode(dev0, False)
    plan = runner._pp_plan
    eng = plan.engine
    # kernels alone: the captured episode graph on the engine state (slots sorted by cell), restored on the device
    restore(dev0)
    plan._adopt(runner.state)
    st0 = plan.st.clone()
    sampler = ClockSampler(0)

    def kernels_only():
        plan.st.copy_(st0)
        eng.run(plan.st, EPISODE, plan.seed, t0=0, cuda_graph=True)

    for _ in range(3):
        kernels_only()
    before = eng.launch_count
    ms_value = timed(kernels_only, args.steps)
    launches = (eng.launch_count - before) // args.steps
    ms_copy = timed(lambda: plan.st.copy_(st0), args.steps)           # the device-side restore is not part of the step
    ms_value -= ms_copy
    ms_e2e = timed(lambda: episode(host, True), args.steps)
    clocks = sampler.stop()

    cd_bytes = st0.countdown.element_size()
    bytes_step = n * 20 + C * 2 * (cd_bytes + 1)          # cell + energy read and written, agent number read; info byte + countdown read and written
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    achieved = bytes_step * EPISODE / (ms_value * 1e-3) / 1e9
    line = {
        "metric": "agent-steps/s, predator-prey grid model (BASELINE configs[3])", "value": n * EPISODE / (ms_value * 1e-3), "unit": "agent-steps/s",
        "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32 cells / f32 energies / f32 countdowns / u8 cell info", "data": "synthetic",
        "config": {"workload": f"predator-prey, {args.predators} predators + {args.prey} prey on a {args.side}x{args.side} grid, "
                               f"{EPISODE}-step episode (move, eat, hunt, grow)", "l2_policy": "agent + countdown streams (>= 300 MB per step) exceed the 126 MB L2"},
        "e2e": {"value": n * EPISODE / (ms_e2e * 1e-3), "unit": "agent-steps/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 24},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "peak_source": peak_src, "algorithmic_bytes_per_step": bytes_step},
        "census_after_episode": [int(v) for v in counts.tolist()],
    }
    if not args.no_cpu_baseline:
        sample_pred, sample_prey, side = args.predators // 4, args.prey // 4, args.side // 2
        sec = cpu_port_sample(side, sample_pred, sample_prey)
        line["cpu_baseline"] = {"value": (sample_pred + sample_prey) / sec, "unit": "agent-steps/s", "cores": 1, "kind": "port",
                                "sample": f"oracle/predprey_port.py (numpy), 2 steps of {sample_pred + sample_prey} agents on a {side}x{side} grid"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
