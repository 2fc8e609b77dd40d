#!/usr/bin/env python
"""bench.py — SEIRM agent-update hot path: agent-steps/s on the 8.4M-agent synthetic population (BASELINE configs[2]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    (N>1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

One "step" = one 21-step SEIRM episode, forward + backward (calibration gradients dL/dR2, dL/dM), over the whole
population: N_agents x 21 agent-steps.  The same code path at every N: the public API — `Runner` on one GPU,
`DistributedRunner` (what `envs.create(distributed=True)` returns) on several — with a fixed noise seed, so that the
trajectories, and the `check` block of the JSON line, are the same for every N.
  value    : episodes with the population, graph and initial state resident in HBM (device-side state restore, CUDA-graph
             replay of the forward launches, fused backward kernel).
  e2e      : the same episodes with the initial state restored from ONE pinned host buffer (compact dtypes, 5 B/agent) and the
             daily series + gradients read back to the host, every episode.
  fwd_only : forward-only episodes (no saved-for-backward buffers, no backward), resident.
  roofline : per-launch timing of ONE eager episode with CUDA events around every step (gather + update launch of that
             step).  frac = algorithmic bytes per step / mean duration of the steps that have >= 1 infectious source (the
             others skip the graph and are reported separately as episode_frac / inert steps).
  hard     : a second workload block — hard isolation (p_isolate 0.3), where the epidemic spreads to most agents (dense
             phase): resident fwd+bwd value through the engine API and its own roofline numbers.
  cpu_baseline / --impl reference: the UNMODIFIED reference (baseline/_ref, driven by oracle/ref_harness) when it is
             installed, else the CPU oracle port, on the box's host cores, on a bounded sample of the same generator.
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
R2_VALUE, M_VALUE = 4.75, 0.12
METRIC = "agent-steps/s fwd+bwd, 8.4M-agent SEIRM"


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
    ap.add_argument("--no-hard", action="store_true", help="skip the hard-isolation workload block")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ CPU arm
_CPU_INPUTS = {}


def _cpu_inputs(n_agents, degree, family):
    import numpy as np

    from agenttorch_b200.graph import synthetic_contact_graph, synthetic_population
    from oracle import noise

    key = (n_agents, degree, family)
    if key not in _CPU_INPUTS:
        _CPU_INPUTS.clear()
        age, stage0 = synthetic_population(n_agents, SEED, device="cpu")
        g = synthetic_contact_graph(n_agents, degree, SEED + 1, family=family, device="cpu")
        ids = np.arange(n_agents)
        U = np.stack([noise.exposure_uniform(7, t, ids) for t in range(T_STEPS)])
        W = np.stack([noise.isolation_decision(7, t, ids, "soft") for t in range(T_STEPS)])
        _CPU_INPUTS[key] = (age.numpy(), stage0.numpy(), g.edge_index().numpy(), U, W)
    return _CPU_INPUTS[key]


def reference_root():
    """Where the unmodified reference lives: baseline/_ref (travels to the GPU box) or the build container's mount."""
    for cand in (os.path.join(ROOT, "baseline", "_ref"), os.environ.get("AGENT_TORCH_REFERENCE", ""), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "agent_torch", "models", "covid")):
            return cand
    return None


def cpu_episode(n_agents, degree, family, threads, kind):
    """One 21-step episode, forward+backward, on the host: kind='reference' runs the UNMODIFIED reference package through
    its own Runner (oracle/ref_harness/run_reference.py: dependency stand-ins + injected noise, nothing else), kind='port'
    the oracle restatement.  Returns seconds of the episode itself (inputs are generated once, outside)."""
    import numpy as np
    import torch

    torch.set_num_threads(threads)
    age, stage0, edges, U, W = _cpu_inputs(n_agents, degree, family)
    gI = np.linspace(0.5, 1.5, T_STEPS, dtype=np.float32)
    gD = np.linspace(2.0, 1.0, T_STEPS, dtype=np.float32)
    if kind == "reference":
        os.environ["AGENT_TORCH_REFERENCE"] = reference_root()
        from oracle.ref_harness import run_reference as rr

        rr.REFERENCE_ROOT = reference_root()
        timing = {}
        rr.run_reference(age, stage0.astype(np.float64), edges.T if edges.shape[0] == 2 else edges, U, W, num_steps=T_STEPS,
                         gI=gI, gD=gD, num_threads=threads, timing=timing)
        return timing["forward_s"] + timing.get("backward_s", 0.0)
    from oracle.seirm_port import SeirmPort, episode_with_grads

    port = SeirmPort(age, stage0, edges, faithful_layout=True)
    t0 = time.perf_counter()
    episode_with_grads(port, W, U, gI, gD)
    return time.perf_counter() - t0


def cpu_kind():
    return "reference" if reference_root() is not None else "port"


def cpu_sample_text(kind, n, degree, threads):
    what = ("UNMODIFIED reference (agent_torch from baseline/_ref, its own Runner / NewTransmission / SEIRMProgression, torch CPU; "
            "torch_geometric, omegaconf, dask stand-ins of oracle/ref_harness)" if kind == "reference"
            else "oracle port (reference op sequence, torch CPU)")
    return (f"{what} on N={n} agents, d={degree:g}, {T_STEPS} steps fwd+bwd, {threads} threads, same generator as the GPU "
            f"workload; per-edge work, so agent-steps/s is size-independent (host RAM bounds N: the reference keeps (E,7) float64 "
            f"gathers of all steps alive for autograd)")


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.cpu_sample_agents
    kind = cpu_kind()
    for _ in range(min(args.warmup, 1)):
        cpu_episode(n, args.degree, args.family, threads, kind)
    times, t_begin = [], time.perf_counter()
    while len(times) < max(1, args.steps) and (not times or time.perf_counter() - t_begin < 150.0):    # bounded: a few minutes
        times.append(cpu_episode(n, args.degree, args.family, threads, kind))
    sec = sum(times) / len(times)
    value = n * T_STEPS / sec
    cfg = workload_config(args, note="CPU arm runs a bounded sample of this workload")
    cfg["agents"] = n
    cfg["agents_gpu_arm"] = args.agents
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": threads, "kind": kind,
                         "sample": cpu_sample_text(kind, n, args.degree, threads)},
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, note=None):
    cfg = {"workload": f"synthetic SEIRM, {args.agents} agents, avg in-degree {args.degree:g} ({args.family} graph, Poisson degrees), "
                       f"{T_STEPS}-step episode forward+backward (BASELINE configs[2])",
           "agents": args.agents, "avg_degree": args.degree, "steps_per_episode": T_STEPS, "graph_family": args.family,
           "noise": "in-kernel counter hash, soft isolation (the reference's default policy), fixed seed",
           "l2_policy": "inputs larger than L2 (the edge stream alone is "
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
def state_checksum(stage, nst, itime, first_id, torch):
    """Order-independent 64-bit checksum of (stage, next_stage_time, infected_time) keyed by GLOBAL agent id: a sum over the
    agents of a mixed 64-bit word, so per-rank partial checksums add up (mod 2^63) to the one-GPU value."""
    ids = torch.arange(first_id, first_id + stage.numel(), device=stage.device, dtype=torch.int64)
    a = stage.reshape(-1).to(torch.float32).view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    b = nst.reshape(-1).to(torch.float32).to(torch.int64) & 0xFFFF
    c = itime.reshape(-1).to(torch.float32).to(torch.int64) & 0xFFFF
    w = (a << 32) | (b << 16) | c
    h = (w ^ (ids * -7046029254386353131)) * -4658895280553007687          # two odd 64-bit multipliers, wrap-around arithmetic
    h = h ^ (h >> 29)
    return h.sum()                                                           # int64 sum wraps: commutative and associative


def ours(args):
    import torch
    import torch.distributed as dist

    from agenttorch_b200 import _native
    from agenttorch_b200.core import Runner
    from agenttorch_b200.core.snapshot import PackedSnapshot
    from agenttorch_b200.distributed import DistributedRunner
    from agenttorch_b200.engine import NoiseSpec, SeirmState
    from agenttorch_b200.models import covid
    from agenttorch_b200.models.covid.synthetic import synthetic_config

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: agenttorch_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime

        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
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

    # ---- the public API: Runner (one GPU) / DistributedRunner (several) built from the synthetic covid config
    cfg = synthetic_config(N, d, seed=SEED, family=args.family, device="cuda", noise_seed=SEED)
    base = Runner(cfg, covid.registry)
    runner = DistributedRunner(base) if world > 1 else base
    runner.init()
    tr = runner.initializer.transition_function["0"]["new_transmission"]
    tr.noise_seed, tr.advance_seed = SEED, False                # the same noise stream in every episode and at every N
    if world > 1:
        engine, first_id = runner.engine, runner.agent_partition_map["start_idx"]
        snap_dev = runner.local_snapshot()
    else:
        from agenttorch_b200.models.covid.context import get_context

        engine, first_id = get_context(cfg, runner.state, tr.input_variables).engine, 0
        cit, env = runner.state["agents"]["citizens"], runner.state["environment"]
        snap_dev = {"agents": {"citizens": {k: cit[k].clone() for k in ("disease_stage", "next_stage_time", "infected_time")}},
                    "environment": {"daily_infected": env["daily_infected"].clone(), "daily_deaths": env["daily_deaths"].clone()}}
    # host snapshot of the initial condition: ONE pinned buffer, compact dtypes (stage uint8, timers int16): 5 B/agent over PCIe
    snap_host = PackedSnapshot(snap_dev, compact={"disease_stage": torch.uint8, "next_stage_time": torch.int16,
                                                  "infected_time": torch.int16})
    params_host = torch.tensor([R2_VALUE] * 3 + [M_VALUE]).pin_memory()       # the calibration parameters of one episode: R2 (3,1), M (1,)
    out_host = torch.empty(2 * T + 4).pin_memory()
    h2d = snap_host.payload_bytes() + params_host.numel() * 4
    d2h = out_host.numel() * 4
    prog = runner.initializer.transition_function["1"]["seirm_progression"]

    def episode(snapshot, to_host, grad=True):
        runner.restore_state(snapshot)
        params = params_host.to(dev, non_blocking=True)
        r2 = params[:3].view(3, 1).requires_grad_(grad)
        m = params[3:].requires_grad_(grad)
        tr.calibrate_R2 = r2                                     # what _set_parameters does (it insists on an unchanged requires_grad)
        prog.calibrate_M = m
        runner.step(T)
        e = runner.state["environment"]
        if grad:
            loss = torch.dot(gI, e["daily_infected"].reshape(-1)) + torch.dot(gD, e["daily_deaths"].reshape(-1))
            loss.backward()
        if to_host:
            out_host.copy_(torch.cat([e["daily_infected"].detach().reshape(-1), e["daily_deaths"].detach().reshape(-1),
                                      r2.grad.reshape(-1), m.grad.reshape(-1)]), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return (r2.grad, m.grad) if grad else (None, None)

    for _ in range(args.warmup):
        episode(snap_dev, False)
    sampler = ClockSampler(local)
    # nvidia-smi needs a few periods under load before the timed regions begin: ~0.6 s of episodes — the SAME number on every
    # rank (derived from an all-reduced time), a rank-local loop "until 0.6 s have passed" would desynchronise the collectives
    n_load = max(3, min(2000, int(600.0 / max(timed(lambda: episode(snap_dev, False), 3) / 3, 0.05))))
    for _ in range(n_load):
        episode(snap_dev, False)
    ms = timed(lambda: episode(snap_dev, False), args.steps)
    # kernels of this package inside the timed region: per episode 1 encode + T x (gather + update) [+ the halo exchanges] replayed
    # from the CUDA graph + the fused backward kernel (replays do not pass through the engine's launch counter)
    launches = args.steps * engine.kernels_per_episode(T)
    def episode_e2e():
        """The end-to-end episode: parameters from pinned host memory, step, loss, backward, results to pinned host memory, stream
        synchronised.  The initial condition of the NEXT episode is restored right after the forward pass (the backward pass reads
        only what the forward saved inside the engine), so its widening copies do not sit between the host's synchronisation and
        the next forward; its host->device copy was prefetched while the previous episode computed."""
        params = params_host.to(dev, non_blocking=True)
        r2 = params[:3].view(3, 1).requires_grad_(True)
        m = params[3:].requires_grad_(True)
        tr.calibrate_R2 = r2
        prog.calibrate_M = m
        runner.step(T)
        e = runner.state["environment"]
        di, dd = e["daily_infected"], e["daily_deaths"]
        loss = torch.dot(gI, di.reshape(-1)) + torch.dot(gD, dd.reshape(-1))
        runner.restore_state(snap_host)          # state of the next episode (one prefetched copy consumed ...
        snap_host.prefetch(dev)                  # ... and the one after it started: one H2D copy of 5 B per agent per episode)
        loss.backward()
        out_host.copy_(torch.cat([di.detach().reshape(-1), dd.detach().reshape(-1), r2.grad.reshape(-1), m.grad.reshape(-1)]),
                       non_blocking=True)
        torch.cuda.current_stream().synchronize()

    runner.restore_state(snap_host)
    snap_host.prefetch(dev)
    for _ in range(args.warmup):
        episode_e2e()
    ms_e2e = timed(episode_e2e, args.steps)
    for _ in range(args.warmup):
        episode(snap_dev, False, grad=False)
    ms_fwd = timed(lambda: episode(snap_dev, False, grad=False), args.steps)
    clocks = sampler.stop()                    # sampled across the three timed regions

    # ---- check block: one more episode, then checksums / sums that must not depend on N
    g_r2, g_m = episode(snap_dev, False)
    engine.check_errors()
    if world > 1:
        st = runner.local_state
        stage, nst, itime = st.stage, st.next_stage_time, st.infected_time
    else:
        cit = runner.state["agents"]["citizens"]
        stage, nst, itime = cit["disease_stage"], cit["next_stage_time"], cit["infected_time"]
    cks = state_checksum(stage, nst, itime, first_id, torch).reshape(1)
    hist = torch.bincount(stage.reshape(-1).round().long().clamp(0, 3), minlength=4)
    edges_t = torch.tensor([engine.graph.num_edges], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(cks); dist.all_reduce(hist); dist.all_reduce(edges_t)
    env = runner.state["environment"]
    check = {"state_checksum": int(cks.item()) & ((1 << 63) - 1), "stage_histogram": hist.tolist(),
             "sum_daily_infected": float(env["daily_infected"].detach().sum()), "sum_daily_deaths": float(env["daily_deaths"].detach().sum()),
             "grad_R2": [float(v) for v in g_r2.reshape(-1)], "grad_M": float(g_m.reshape(-1)[0]), "edges": int(edges_t.item()),
             "note": "state_checksum and stage_histogram are exact and must be identical at every N; the series and gradients are "
                     "float32 sums (per-rank rounding before the all-reduce: last-bit differences possible)"}
    edges = int(edges_t.item())

    # ---- roofline: per-step timing of one eager episode (events around the launches of every step), engine API
    def profile_episode(isolation):
        sv = engine.saved_buffers(T, True)
        if world > 1:
            st = runner.local_state
        else:
            cit = runner.state["agents"]["citizens"]
            st = SeirmState(cit["disease_stage"].view(-1), cit["next_stage_time"].view(-1), cit["infected_time"].view(-1),
                            cit["age"].view(-1), torch.zeros(T, device=dev), torch.zeros(T, device=dev))
        noise = NoiseSpec(seed=SEED, isolation=isolation, p_isolate=0.3)
        best = None
        for rep in range(3):
            runner.restore_state(snap_dev)
            st.daily_infected.zero_(); st.daily_deaths.zero_()
            sv["r_steps"].fill_(3 * R2_VALUE); sv["m_value"].fill_(M_VALUE); engine.seed_buf.fill_(SEED)
            sv["recovered_step"].fill_(255)
            ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(T)]
            ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(T)]
            flags = torch.zeros(T, dtype=torch.bool, device=dev)
            engine.over.zero_()
            engine._use_over = True                                  # as inside run_episode(): "epidemic over" short cuts enabled
            engine.encode(st, 0)
            barrier()
            for t in range(T):
                # "active" = some source agent (of any rank) is infectious in this step, i.e. the gather streams the graph
                # (evaluated on the device, outside the step's event pair; no host synchronisation inside the episode)
                flags[t] = (engine.hot[engine.cur][: engine.n_code // 32] != 0).any()
                ev0[t].record()
                engine.step(st, t, noise, sv["r_steps"], t, sv["m_value"], dne_dr=sv["dne_dr"][t], sum_dne_dr=sv["sum_dne_dr"],
                            death_count=sv["death_count"], recovered_step=sv["recovered_step"], seed_ptr=engine.seed_buf,
                            check_range=t + 1 < T)
                ev1[t].record()
            engine._use_over = False
            torch.cuda.synchronize(dev)
            active = flags.tolist()
            times = torch.tensor([ev0[t].elapsed_time(ev1[t]) for t in range(T)], device=dev)
            if world > 1:
                dist.all_reduce(times, op=dist.ReduceOp.MAX)
            times = times.tolist()
            if best is None or sum(times) < sum(best[0]):
                di_sum = st.daily_infected.sum().reshape(1).clone()
                if world > 1:
                    dist.all_reduce(di_sum)
                best = (times, active, float(di_sum))
        return best

    def resident_engine_episodes(isolation, k):
        """fwd+bwd episodes through the engine API (state restored on the device), for the workload the Runner's policy
        does not offer (hard isolation)."""
        if world > 1:
            st = runner.local_state
        else:
            cit = runner.state["agents"]["citizens"]
            st = SeirmState(cit["disease_stage"].view(-1), cit["next_stage_time"].view(-1), cit["infected_time"].view(-1),
                            cit["age"].view(-1), torch.zeros(T, device=dev), torch.zeros(T, device=dev))
        noise = NoiseSpec(seed=SEED, isolation=isolation, p_isolate=0.3)
        r_steps = torch.full((T,), 3 * R2_VALUE, device=dev)
        m_val = torch.tensor([M_VALUE], device=dev)

        def one():
            runner.restore_state(snap_dev)
            st.daily_infected.zero_(); st.daily_deaths.zero_()
            sv = engine.run_episode(st, r_steps, m_val, T, noise, want_grad=True, cuda_graph=engine.pg is None or engine.code_hdl is not None)
            if world > 1:
                from agenttorch_b200.parallel import reduce_daily_series

                reduce_daily_series(st, engine.pg)
            engine.episode_backward(sv, T, gI, gD, m_val)

        for _ in range(max(2, args.warmup)):
            one()
        ms_h = timed(one, k)
        hist_h = torch.bincount(st.stage.reshape(-1).round().long().clamp(0, 3), minlength=4)
        if world > 1:
            dist.all_reduce(hist_h)
        return ms_h, hist_h.tolist()

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    d_act = edges / N
    bytes_per_step = (N / world) * (4.0 * d_act + 41.0)

    def ncu_traffic():
        """DRAM bytes (read + written) of the two launches of one hot step, from the committed `ncu --set full` capture of this
        build's kernels on this workload (profiles/r02b_ncu_full_summary.csv); one GPU only, None if the file is absent."""
        path = os.path.join(ROOT, "profiles", "r02b_ncu_full_summary.csv")
        if world != 1 or N != 8_400_000 or not os.path.exists(path):
            return None
        import csv

        rows = {r[0]: r for r in csv.reader(open(path)) if r}
        try:
            cols = [i for i, h in enumerate(rows["metric"]) if "stream_gather_kernel" in h or "stream_update_kernel" in h]
            return sum(float(rows[m][i]) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum") for i in cols) * 1e6
        except (KeyError, ValueError, IndexError):
            return None

    def roofline_of(times, active):
        act = [x for x, a in zip(times, active) if a]
        inert = [x for x, a in zip(times, active) if not a]
        mean_act = sum(act) / max(len(act), 1)
        ach = bytes_per_step / (mean_act / 1e3) / 1e9 if act else None
        ach_all = bytes_per_step / (sum(times) / len(times) / 1e3) / 1e9
        return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                "traffic": traffic_bytes, "kernel": "stream_gather_kernel + stream_update_kernel (the two launches of one simulation step)",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_step,
                "active_launches": len(act), "active_launch_ms": {"mean": mean_act, "min": min(act) if act else None, "max": max(act) if act else None},
                "inert_launches": len(inert), "inert_launch_ms_mean": (sum(inert) / len(inert)) if inert else None,
                "episode_frac": ach_all / peak,
                "worst_active_frac": (bytes_per_step / (max(act) / 1e3) / 1e9 / peak) if act else None,
                "note": "bytes = (N/n_gpus)*(4*d+41) per GPU and step: SURVEY 8d forward 4d+45, in-kernel noise (-8), dne_dr save (+4). "
                        "frac = bytes / MEAN duration of the steps in which some source is infectious (the gather streams the whole "
                        "edge list in those; on several GPUs the step includes the halo exchange); steps without any infectious "
                        "source skip the graph and only appear in episode_frac. traffic: measured by ncu only, see profiles/ (the "
                        "gather reads the stream once: dram bytes = algorithmic col bytes + 3 % padding)"}

    times_s, active_s, _ = profile_episode("soft")
    traffic_bytes = ncu_traffic()
    roofline = roofline_of(times_s, active_s)
    traffic_bytes = None                      # (the capture is of the soft workload)
    hard = None
    if not args.no_hard:
        times_h, active_h, di_h = profile_episode("hard")
        ms_h, hist_h = resident_engine_episodes("hard", max(2, args.steps // 2))
        hard = {"workload": "the same population and graph, hard isolation (p_isolate 0.3): the epidemic reaches most agents",
                "value": N * T * max(2, args.steps // 2) / (ms_h / 1e3), "unit": "agent-steps/s",
                "ms_per_step": ms_h / max(2, args.steps // 2), "api": "SeirmEngine.run_episode + episode_backward (resident)",
                "stage_histogram": hist_h, "sum_daily_infected": di_h, "roofline": roofline_of(times_h, active_h),
                "step_ms": [round(x, 4) for x in times_h]}
    runner.restore_state(snap_dev)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    agent_steps = N * T
    line = {
        "metric": METRIC, "value": agent_steps * args.steps / (ms / 1e3), "unit": "agent-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args),
        "e2e": {"value": agent_steps * args.steps / (ms_e2e / 1e3), "unit": "agent-steps/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / args.steps,
                "api": "DistributedRunner" if world > 1 else "Runner",
                "input_pipeline": "PackedSnapshot.prefetch: every episode issues one pinned-host->device copy of an initial condition "
                                  "(copy stream, two staging buffers) and consumes the one issued by the episode before it (restored "
                                  "right after the forward pass, before loss.backward()); the result is read back to pinned host "
                                  "memory and the stream synchronised every episode"},
        "fwd_only": {"value": agent_steps * args.steps / (ms_fwd / 1e3), "unit": "agent-steps/s", "ms_per_step": ms_fwd / args.steps},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        "step_ms": [round(x, 4) for x in times_s], "hard": hard, "check": check,
        "layout": {"name": "stream" if engine.graph.layout == 1 else "bcsr", "source_blocks": engine.graph.n_blocks,
                   "block_size": engine.graph.block_size, "tile_rows": engine.graph.tile_rows},
    }
    if world > 1 and getattr(engine, "halo_mask", None) is not None:
        # share of the population's 32-agent groups this rank's edges refer to (its own included): what the halo exchange may pull;
        # 1.0 on the uniformly random graph of this workload, less on a graph with community structure (--family comm)
        line["halo"] = {"groups_referenced_fraction_rank0": engine.halo_fraction, "exchange": "sparse pull of hot 32-byte code sectors "
                        "inside the rank's halo mask" if engine.sparse_exchange and engine.code_hdl is not None else "whole slices"}
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_s = args.cpu_sample_agents
        kind = cpu_kind()
        times = []
        while sum(times) < 10.0 and len(times) < 6:
            times.append(cpu_episode(n_s, d, args.family, threads, kind))
        sec = sum(times) / len(times)
        line["cpu_baseline"] = {"value": n_s * T / sec, "unit": "agent-steps/s", "cores": threads, "kind": kind,
                                "sample": cpu_sample_text(kind, n_s, d, threads) + f": {len(times)} episodes, {sec:.2f} s each"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        reference_arm(a)
    else:
        ours(a)
