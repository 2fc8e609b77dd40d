#!/usr/bin/env python
"""bench.py — SEIRM agent-update hot path: agent-steps/s, forward+backward, 8.4M-agent synthetic population.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    (N>1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

One "step" = one 21-step SEIRM episode, forward + backward (calibration gradients dL/dR2, dL/dM), over the
whole population: N_agents x 21 agent-steps.  Workload = BASELINE.json configs[2] (8.4M agents, average
contact degree 20, T=21), which fits one B200.
  value  : agent-steps/s with the population, graph and initial state resident in HBM (device-side state
           restore + CUDA-graph replay of the 22 forward launches + the fused backward kernel).
  e2e    : the same metric through the public API (Runner.restore_state from PINNED HOST buffers (compact dtypes) ->
           runner.step(21) -> loss.backward() -> series and gradients read back to the host), every step.
  roofline: the step kernel (seirm_kernel<MODE_HASHED>, one launch per simulation step): algorithmic bytes per launch (in-kernel noise,
           dne save: N*(4d+41)) / its average launch duration measured with CUDA events around the forward graph.
  cpu_baseline / --impl reference: the CPU oracle port of the reference's PyTorch path (oracle/seirm_port.py,
           reference op sequence incl. the (E,7) float64 gathers) on a bounded sample of the same generator.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_STEPS = 21
SEED = 20261019


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--agents", type=int, default=8_400_000)
    ap.add_argument("--degree", type=float, default=20.0)
    ap.add_argument("--family", default="rand", choices=["rand", "comm"])
    ap.add_argument("--cpu-sample-agents", type=int, default=200_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ CPU arm
_CPU_INPUTS = {}


def cpu_port_episode(n_agents, degree, family, threads, fwd_bwd=True):
    """One 21-step episode (forward+backward) of the CPU oracle port on a synthetic sample; returns seconds.  The sample
    (population, graph, noise) is generated once per process and outside the timed region."""
    import numpy as np
    import torch

    from agenttorch_b200.graph import synthetic_contact_graph, synthetic_population
    from oracle import noise
    from oracle.seirm_port import SeirmPort, episode_with_grads

    torch.set_num_threads(threads)
    key = (n_agents, degree, family)
    if key not in _CPU_INPUTS:
        _CPU_INPUTS.clear()
        age, stage0 = synthetic_population(n_agents, SEED, device="cpu")
        g = synthetic_contact_graph(n_agents, degree, SEED + 1, family=family, device="cpu")
        ids = np.arange(n_agents)
        U = np.stack([noise.exposure_uniform(7, t, ids) for t in range(T_STEPS)])
        W = np.stack([noise.isolation_decision(7, t, ids, "soft") for t in range(T_STEPS)])
        _CPU_INPUTS[key] = (age.numpy(), stage0.numpy(), g.edge_index().numpy(), U, W)
    age, stage0, edges, U, W = _CPU_INPUTS[key]
    port = SeirmPort(age, stage0, edges, faithful_layout=True)
    gI = np.linspace(0.5, 1.5, T_STEPS, dtype=np.float32)
    gD = np.linspace(2.0, 1.0, T_STEPS, dtype=np.float32)
    t0 = time.perf_counter()
    episode_with_grads(port, W, U, gI, gD)
    return time.perf_counter() - t0


def cpu_baseline_sample(n_agents, degree, family, threads, min_seconds=10.0, max_episodes=12):
    """Bounded CPU sample for the GPU arm's `cpu_baseline`: whole episodes of the port until ~min_seconds of CPU work
    (the sample size is capped by host memory: the reference op sequence keeps (E,7) float64 gathers of all 21 steps alive
    for autograd, 9.4 GB at 200 k agents).  Returns (mean seconds per episode, episodes)."""
    times = []
    while sum(times) < min_seconds and len(times) < max_episodes:
        times.append(cpu_port_episode(n_agents, degree, family, threads))
    return sum(times) / len(times), len(times)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.cpu_sample_agents
    for _ in range(args.warmup):
        cpu_port_episode(n, args.degree, args.family, threads)
    times = [cpu_port_episode(n, args.degree, args.family, threads) for _ in range(max(1, args.steps))]
    sec = sum(times) / len(times)
    value = n * T_STEPS / sec
    sample = (f"oracle port (reference op sequence, torch CPU) on N={n} agents, d={args.degree}, {T_STEPS} steps fwd+bwd, "
              f"same generator as the GPU workload; throughput per agent-step is size-independent (per-edge work)")
    line = {
        "impl": "reference", "metric": "agent-steps/s fwd+bwd, 8.4M-agent SEIRM", "value": value, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, note="CPU arm runs a bounded sample of this workload"),
        "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, note=None):
    cfg = {"workload": f"synthetic SEIRM, {args.agents} agents, avg in-degree {args.degree:g} ({args.family} graph, Poisson degrees), "
                       f"{T_STEPS}-step episode forward+backward (BASELINE configs[2])",
           "agents": args.agents, "avg_degree": args.degree, "steps_per_episode": T_STEPS, "graph_family": args.family,
           "noise": "in-kernel counter hash (soft isolation)", "l2_policy": "inputs larger than L2 (col_idx alone is "
           f"{args.agents * args.degree * 4 / 1e6:.0f} MB vs 126 MB L2); no flush needed"}
    if note:
        cfg["note"] = note
    return cfg


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.proc, self.path = None, f"/tmp/clocks_{os.getpid()}.csv"
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nme, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


# ------------------------------------------------------------------------------------------ GPU arm
def ours(args):
    import torch
    import torch.distributed as dist

    from agenttorch_b200 import _native
    from agenttorch_b200.core import Runner
    from agenttorch_b200.engine import NoiseSpec, SeirmEngine, SeirmSpec, SeirmState
    from agenttorch_b200.graph import (mean_interactions_from_age, shard_bounds, shard_code_len, synthetic_contact_graph,
                                       synthetic_population)
    from agenttorch_b200.models import covid
    from agenttorch_b200.models.covid.synthetic import synthetic_config

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: agenttorch_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    _native.lib()
    N, d, T = args.agents, args.degree, T_STEPS
    gI = torch.linspace(0.5, 1.5, T, device=dev)
    gD = torch.linspace(2.0, 1.0, T, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, k):
        """K calls bracketed by barrier+sync; device time by CUDA events; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    fwd_ms_acc = []
    if world == 1:
        # ---- public API: Runner built from the (synthetic) covid config
        cfg = synthetic_config(N, d, seed=SEED, family=args.family, device="cuda", noise_seed=SEED)
        runner = Runner(cfg, covid.registry)
        runner.init()
        tr = runner.initializer.transition_function["0"]["new_transmission"]
        pg_mod = runner.initializer.transition_function["1"]["seirm_progression"]
        cit = runner.state["agents"]["citizens"]
        env = runner.state["environment"]
        names = ("disease_stage", "next_stage_time", "infected_time")
        init_dev = {"agents": {"citizens": {k: cit[k].clone() for k in names}},
                    "environment": {"daily_infected": env["daily_infected"].clone(), "daily_deaths": env["daily_deaths"].clone()}}
        # host snapshot of the initial condition in compact dtypes (stage uint8, timers int16): 5 B/agent over PCIe
        compact = {"disease_stage": torch.uint8, "next_stage_time": torch.int16, "infected_time": torch.int16}
        init_host = {"agents": {"citizens": {k: v.to(compact[k]).cpu().pin_memory() for k, v in init_dev["agents"]["citizens"].items()}},
                     "environment": {k: v.cpu().pin_memory() for k, v in init_dev["environment"].items()}}
        for k, v in init_host["agents"]["citizens"].items():
            assert torch.equal(v.to(torch.float32), init_dev["agents"]["citizens"][k].cpu())
        r2_host = torch.full((3, 1), 4.75).pin_memory()
        m_host = torch.full((1,), 0.12).pin_memory()
        out_host = torch.empty(2 * T + 4).pin_memory()
        h2d = sum(v.numel() * v.element_size() for grp in init_host.values() for v in
                  (grp["citizens"].values() if "citizens" in grp else grp.values())) + 16
        d2h = out_host.numel() * 4

        def episode(snapshot, to_host):
            runner.restore_state(snapshot)
            r2 = r2_host.to(dev, non_blocking=True).requires_grad_(True)
            m = m_host.to(dev, non_blocking=True).requires_grad_(True)
            runner._set_parameters({"initializer.transition_function.0.new_transmission.learnable_args.R2": r2,
                                    "initializer.transition_function.1.seirm_progression.learnable_args.M": m})
            runner.step(T)
            e = runner.state["environment"]
            loss = (gI * e["daily_infected"].reshape(-1)).sum() + (gD * e["daily_deaths"].reshape(-1)).sum()
            loss.backward()
            if to_host:
                out_host.copy_(torch.cat([e["daily_infected"].detach().reshape(-1), e["daily_deaths"].detach().reshape(-1),
                                          r2.grad.reshape(-1), m.grad.reshape(-1)]), non_blocking=True)
                torch.cuda.current_stream().synchronize()
            return loss

        engine = None
        for _ in range(args.warmup):
            episode(init_dev, False)
        from agenttorch_b200.models.covid.context import get_context

        engine = get_context(cfg, runner.state, tr.input_variables).engine
        sampler = ClockSampler(local)
        for _ in range(3):
            episode(init_dev, False)          # the sampler needs >= 1 period under load before the timed region ends
        ms = timed(lambda: episode(init_dev, False), args.steps)
        clocks = sampler.stop()
        launches = args.steps * engine.kernels_per_episode(T)
        # forward-only graph time -> average step-kernel launch duration (includes the 1 encode launch: conservative)
        sv = engine.saved_buffers(T, True)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fwd = []
        cit = runner.state["agents"]["citizens"]
        st = SeirmState(cit["disease_stage"].view(-1), cit["next_stage_time"].view(-1), cit["infected_time"].view(-1),
                        cit["age"].view(-1), torch.zeros(T, device=dev), torch.zeros(T, device=dev))
        for _ in range(max(4, args.steps // 2)):
            runner.restore_state(init_dev)
            st.daily_infected.zero_(); st.daily_deaths.zero_()
            torch.cuda.synchronize()
            f0.record()
            engine.run_episode(st, torch.full((T,), 14.25, device=dev), torch.tensor([0.12], device=dev), T,
                               NoiseSpec(seed=SEED, isolation="soft"), want_grad=True, cuda_graph=True)
            f1.record()
            torch.cuda.synchronize()
            fwd.append(f0.elapsed_time(f1))
        fwd_ms = statistics.median(fwd[1:]) if len(fwd) > 1 else fwd[0]
        for _ in range(args.warmup):
            episode(init_host, True)
        ms_e2e = timed(lambda: episode(init_host, True), args.steps)
        edges = engine.graph.num_edges
        final_di = runner.state["environment"]["daily_infected"].detach().sum().item()
    else:
        # ---- N>1: agents sharded by contiguous id range; per-step halo exchange of the infectiousness code/bitmap
        from agenttorch_b200.parallel import reduce_daily_series

        bounds = shard_bounds(N, world)
        r0, r1 = bounds[rank]
        age_all, stage_all = synthetic_population(N, SEED, device=dev)
        mi = mean_interactions_from_age(age_all)
        graph = synthetic_contact_graph(N, d, SEED + 1, family=args.family, device=dev, row_begin=r0, row_end=r1)
        full0 = synthetic_contact_graph(N, d, SEED + 1, family=args.family, device=dev, row_begin=0, row_end=min(1 << 20, N),
                                        mean_interactions=mi)
        graph.i_bar, graph.i_bar_agent = full0.i_bar, full0.i_bar_agent
        del full0
        engine = SeirmEngine(graph, SeirmSpec(num_steps=T), process_group=pg, code_len=shard_code_len(N, world))
        st0 = SeirmState.from_initial(stage_all[r0:r1], age_all[r0:r1], T, dev, packed=False)
        st = st0.clone()
        host = {k: getattr(st0, k).cpu().pin_memory() for k in ("stage", "next_stage_time", "infected_time")}
        r_steps = torch.full((T,), 14.25, device=dev)
        m_val = torch.tensor([0.12], device=dev)
        out_host = torch.empty(2 * T + T + 1).pin_memory()
        h2d = sum(v.numel() * v.element_size() for v in host.values()) + 16
        d2h = out_host.numel() * 4
        edges_t = torch.tensor([graph.num_edges], device=dev, dtype=torch.float64)
        dist.all_reduce(edges_t)
        edges = int(edges_t.item())

        def episode(from_host, to_host):
            for k in ("stage", "next_stage_time", "infected_time"):
                getattr(st, k).copy_(host[k] if from_host else getattr(st0, k), non_blocking=True)
            st.daily_infected.zero_(); st.daily_deaths.zero_()
            sv = engine.run_episode(st, r_steps, m_val, T, NoiseSpec(seed=SEED, isolation="soft"), want_grad=True,
                                    cuda_graph=engine.code_hdl is not None)
            di, dd = reduce_daily_series(st, pg)                 # one packed all-reduce
            gr, gm = engine.episode_backward(sv, T, gI, gD, m_val)
            if to_host:
                out_host.copy_(torch.cat([di, dd, gr, gm]), non_blocking=True)
                torch.cuda.current_stream().synchronize()
            return di

        for _ in range(args.warmup):
            episode(False, False)
        sampler = ClockSampler(local)
        for _ in range(3):
            episode(False, False)
        ms = timed(lambda: episode(False, False), args.steps)
        clocks = sampler.stop()
        launches = args.steps * engine.kernels_per_episode(T)
        # forward-only time of this rank's episode -> average step-kernel launch duration (max over ranks)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fwd = []
        for _ in range(max(4, args.steps // 2)):
            for k in ("stage", "next_stage_time", "infected_time"):
                getattr(st, k).copy_(getattr(st0, k))
            st.daily_infected.zero_(); st.daily_deaths.zero_()
            barrier()
            f0.record()
            engine.run_episode(st, r_steps, m_val, T, NoiseSpec(seed=SEED, isolation="soft"), want_grad=True,
                               cuda_graph=engine.code_hdl is not None)
            f1.record()
            torch.cuda.synchronize()
            fwd.append(f0.elapsed_time(f1))
        fwd_t = torch.tensor([statistics.median(fwd[1:])], device=dev)
        dist.all_reduce(fwd_t, op=dist.ReduceOp.MAX)
        fwd_ms = float(fwd_t.item())
        for _ in range(args.warmup):
            episode(True, True)
        ms_e2e = timed(lambda: episode(True, True), args.steps)
        final_di = episode(False, False).sum().item()
        engine.check_errors()                                  # e.g. a halo-exchange rendezvous that timed out

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    agent_steps = N * T
    value = agent_steps * args.steps / (ms / 1e3)
    e2e = agent_steps * args.steps / (ms_e2e / 1e3)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    d_act = edges / N
    bytes_per_launch = (N / world) * (4.0 * d_act + 41.0)
    roofline = None
    if fwd_ms is not None:
        launch_ms = fwd_ms / T
        achieved = bytes_per_launch / (launch_ms / 1e3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(f"{N}x{int(d)}")
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel": "seirm_kernel<MODE_HASHED>", "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": launch_ms,
                    "note": "bytes = (N/n_gpus)*(4*d+41) per GPU: SURVEY 8d forward 4d+45, in-kernel noise (-8), dne_dr save (+4); "
                            "launch time = forward episode time / 21 (includes the encode launch and, on several GPUs, "
                            "the per-step halo exchange)"}
    line = {
        "metric": "agent-steps/s fwd+bwd, 8.4M-agent SEIRM", "value": value, "unit": "agent-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args),
        "e2e": {"value": e2e, "unit": "agent-steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        "check": {"sum_daily_infected": final_di, "edges": int(edges)},
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_s = args.cpu_sample_agents
        sec, reps = cpu_baseline_sample(n_s, d, args.family, threads)
        line["cpu_baseline"] = {"value": n_s * T / sec, "unit": "agent-steps/s", "cores": threads, "kind": "port",
                                "sample": f"oracle port (reference op sequence, torch CPU, {threads} threads) on N={n_s} agents, "
                                          f"d={d:g}: {reps} {T}-step episodes fwd+bwd, {sec:.2f} s each ({reps * sec:.0f} s of CPU work)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        reference_arm(a)
    else:
        ours(a)
