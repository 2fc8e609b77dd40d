#!/usr/bin/env python
"""Throughput of the quarantine / testing / SEIRMS-progression kernels on one B200.

    python tools/bench_interventions.py [--agents 33554432] [--steps 10] [--warmup 3]

A "step" is one 21-step episode of the intervention substeps over the whole population, captured in a CUDA graph
(quarantine_step + testing_step per simulation step, in place; then 21 SEIRMS progressions).  Prints one JSON line per
kernel family with the roofline of the streaming pass: algorithmic bytes = the float32 state every agent-step must READ
(quarantine 12 B: stage, status, start date; testing 16 B: status, queue flag, two dates; SEIRMS 8 B: stage, timer) —
writes happen only for 16-byte groups that change, and are not counted."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
T = 21


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--agents", type=int, default=32 * 1024 * 1024)   # 134 MB per state array: every array exceeds the 126 MB L2
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import torch

    from agenttorch_b200 import interventions as ops

    if not torch.cuda.is_available():
        raise RuntimeError("needs a CUDA device: agenttorch_b200 has no CPU path")
    dev = torch.device("cuda", 0)
    n = args.agents
    gen = torch.Generator(device=dev).manual_seed(1)
    r = torch.rand(n, device=dev, generator=gen)
    stage0 = torch.zeros(n, device=dev)
    stage0[r < 0.10] = 3.0
    stage0[r < 0.08] = 2.0
    stage0[r < 0.04] = 1.0
    nst0 = torch.where((stage0 == 1) | (stage0 == 2), torch.randint(0, T, (n,), device=dev, generator=gen).float(),
                       torch.full((n,), 500.0, device=dev))
    state0 = {"q": torch.zeros(n, device=dev), "d": torch.full((n,), 130.0, device=dev), "aw": torch.zeros(n, device=dev),
              "rd": torch.full((n,), 130.0, device=dev), "re": torch.zeros(n, device=dev)}
    st = {k: v.clone() for k, v in state0.items()}
    stage, nst = stage0.clone(), nst0.clone()
    m = torch.tensor([0.12], device=dev)
    dd = torch.zeros(T, device=dev)

    def interventions_episode():
        for t in range(T):
            ops.quarantine_step(stage0, st["q"], st["d"], 0.7, 0.1, t, 7, 0, 12.0, 130.0)
            ops.testing_step(st["q"], st["aw"], st["rd"], st["re"], 0.95, t, 7, 0, 2.0, 3.0, 130.0)

    from agenttorch_b200 import _native_iv as nat
    counts = torch.zeros(4, dtype=torch.int64, device=dev)

    def seirms_episode():                                   # in place through the C ABI (the op returns new tensors)
        for t in range(T):
            a = nat.SeirmsArgs(abi_version=nat.ABI_VERSION, t=t, n=n, id_offset=0, seed=7, seed_ptr=None, infected_to_recovered_time=5.0,
                               infinity_time=41.0, recovered_to_susceptible_time=100.0, num_steps=T, m_value=m.data_ptr(), u_death=None,
                               stage_in=stage.data_ptr(), next_stage_time_in=nst.data_ptr(), stage_out=stage.data_ptr(),
                               next_stage_time_out=nst.data_ptr(), move_code=None, counts=counts.data_ptr(), daily_deaths=dd.data_ptr())
            nat.check(nat.lib().at_iv_seirms_progression(a, torch.cuda.current_stream().cuda_stream))

    def reset():
        for k, v in state0.items():
            st[k].copy_(v)
        stage.copy_(stage0); nst.copy_(nst0); dd.zero_(); counts.zero_()

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0
    for name, fn, bytes_per_agent_step, launches in (("quarantine+testing", interventions_episode, 28.0, 2 * T),
                                                     ("seirms_progression", seirms_episode, 8.0, 2 * T)):
        reset()
        fn()                                                # warm the library / allocator before capture
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        reset()
        with torch.cuda.graph(graph):
            fn()
        for _ in range(args.warmup):
            reset(); graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        total = 0.0
        for _ in range(args.steps):
            reset()                                         # untimed: every episode starts from the same state
            torch.cuda.synchronize()
            e0.record(); graph.replay(); e1.record()
            torch.cuda.synchronize()
            total += e0.elapsed_time(e1)
        ms = total / args.steps
        achieved = n * T * bytes_per_agent_step / (ms / 1e3) / 1e9
        print(json.dumps({
            "metric": f"agent-steps/s, covid {name} substeps", "value": n * T / (ms / 1e3), "unit": "agent-steps/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{n} agents, {T}-step episode, 4 % exposed / 4 % infected / 2 % recovered",
                       "l2_policy": f"inputs larger than L2: every launch streams 2-4 float32 arrays of {n * 4 / 1e6:.0f} MB each (L2 is 126 MB)"},
            "gpu_launches": launches * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "algorithmic_bytes_per_agent_step": bytes_per_agent_step,
                         "note": "bytes = compulsory float32 reads only; conditional 16-byte writes not counted"},
            "check": {"quarantined": int((st["q"] != 0).sum()), "awaiting": int((st["aw"] != 0).sum()),
                      "deaths": float(dd.sum()), "stage_hist": torch.bincount(stage.long(), minlength=5).tolist()}}), flush=True)


if __name__ == "__main__":
    main()
