"""BASELINE configs[1]: synthetic 1M-agent SEIRM, average contact degree 20, FORWARD-ONLY on one B200 (CUDA-graph episode)."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from agenttorch_b200.engine import NoiseSpec, SeirmEngine, SeirmSpec, SeirmState
from agenttorch_b200.graph import synthetic_contact_graph, synthetic_population, mean_interactions_from_age

dev = torch.device("cuda:0"); T = 21; n = 1_000_000
age, stage0 = synthetic_population(n, 20261019, device=dev)
graph = synthetic_contact_graph(n, 20.0, 20261020, device=dev, mean_interactions=mean_interactions_from_age(age))
eng = SeirmEngine(graph, SeirmSpec(num_steps=T))
st0 = SeirmState.from_initial(stage0, age, T, dev); st = st0.clone()
r = torch.full((T,), 14.25, device=dev); m = torch.tensor([0.12], device=dev)
for mode in ("soft", "hard"):
    times = []
    for i in range(13):
        for k in ("stage", "next_stage_time", "infected_time", "daily_infected", "daily_deaths"):
            getattr(st, k).copy_(getattr(st0, k))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        eng.run_episode(st, r, m, T, NoiseSpec(seed=5, isolation=mode), want_grad=False, cuda_graph=True)
        e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = statistics.median(times[3:])
    print(f"C2 1M agents d=20 forward-only, isolation={mode}: {ms:.3f} ms/episode -> {n*T/ms/1e6:.1f} G agent-steps/s, "
          f"algorithmic {n*T*(4*graph.num_edges/n+37)/ms/1e6/1e3:.2f} TB/s, infected {float(st.daily_infected.sum()):.0f}")
