"""Developer probe: per-launch timing of the step kernel on a synthetic population (not the bench)."""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from agenttorch_b200.engine import NoiseSpec, SeirmEngine, SeirmSpec, SeirmState
from agenttorch_b200.graph import synthetic_contact_graph, synthetic_population, mean_interactions_from_age

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--d", type=float, default=20.0)
ap.add_argument("--steps", type=int, default=21)
ap.add_argument("--family", default="rand")
ap.add_argument("--block", type=int, default=0)
ap.add_argument("--tpg", type=int, default=0)
ap.add_argument("--hashed", type=int, default=-1)
ap.add_argument("--hmw", type=int, default=16)
ap.add_argument("--grad", type=int, default=1)
ap.add_argument("--r2", type=float, default=4.75)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--iso", default="soft")
ap.add_argument("--layout", default="auto")
ap.add_argument("--tile", type=int, default=0)
ap.add_argument("--epc", type=int, default=0)
ap.add_argument("--grid", type=int, default=0)
ap.add_argument("--bank", type=int, default=1)
ap.add_argument("--warps", type=int, default=0)
ap.add_argument("--lib", default="")          # a variant build of libat_seirm.so (kernel experiments)
ap.add_argument("--f32", type=int, default=0)  # reference layout: float32 timers and ages (what the Runner path carries)
ap.add_argument("--brief", type=int, default=0)
args = ap.parse_args()
if args.lib:
    from agenttorch_b200 import _native
    _native.LIB_PATH = os.path.abspath(args.lib)
dev = torch.device("cuda:0")
T = args.steps
age, stage0 = synthetic_population(args.n, 1234, device=dev)
mi = mean_interactions_from_age(age)
graph = synthetic_contact_graph(args.n, args.d, 99, family=args.family, device=dev, mean_interactions=mi)
print("N", args.n, "E", graph.num_edges, "ibar", graph.i_bar, flush=True)
eng = SeirmEngine(graph, SeirmSpec(num_steps=T), block_size=args.block or None, tasks_per_group=args.tpg, hashed=None if args.hashed < 0 else bool(args.hashed), hash_min_warps=args.hmw,
                  layout=args.layout, tile_rows=args.tile or None, edges_per_chunk=args.epc or None, warps_per_cta=args.warps, grid=args.grid, bank_aware=bool(args.bank))
print('units', int(eng.graph.block_units[-1]) if eng.graph.block_units is not None else 0, 'layout', eng.graph.layout, 'blocks', eng.graph.n_blocks, 'block_size', eng.graph.block_size, 'hash_words', eng.graph.hash_words, 'tile', eng.graph.tile_rows, 'epc', eng.graph.edges_per_chunk, flush=True)
r_steps = torch.full((T,), 3 * args.r2, device=dev)
m = torch.tensor([0.12], device=dev)
noise = NoiseSpec(seed=5, isolation=args.iso)
for rep in range(args.reps):
    st = SeirmState.from_initial(stage0, age, T, dev, packed=not args.f32)
    sv = eng.saved_buffers(T, bool(args.grad))
    sv["r_steps"].copy_(r_steps); sv["m_value"].copy_(m); eng.seed_buf.fill_(5)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(T + 2)]
    sv["recovered_step"].fill_(255)
    evs[0].record()
    eng.encode(st, 0)
    evs[1].record()
    for t in range(T):
        eng.step(st, t, noise, sv["r_steps"], t, sv["m_value"], dne_dr=sv["dne_dr"][t] if args.grad else None,
                 sum_dne_dr=sv["sum_dne_dr"], death_count=sv["death_count"], recovered_step=sv["recovered_step"], seed_ptr=eng.seed_buf)
        evs[t + 2].record()
    torch.cuda.synchronize()
    times = [evs[i].elapsed_time(evs[i + 1]) * 1e3 for i in range(T + 1)]
    if rep == args.reps - 1:
        sus = None
        print("encode us %.1f" % times[0])
        di = st.daily_infected.tolist()
        hot = [times[t + 1] for t in range(T) if times[t + 1] > 100.0]
        print("hot steps: mean %.1f us  min %.1f  max %.1f  (n=%d)   inert mean %.1f us" % (
            sum(hot) / max(len(hot), 1), min(hot) if hot else 0, max(hot) if hot else 0, len(hot),
            sum(x for x in times[1:] if x <= 100.0) / max(1, T - len(hot))))
        for t in range(T if not args.brief else 0):
            bytes_ = args.n * (4 * graph.num_edges / args.n + 41)
            print("t=%2d  %8.1f us  %6.2f TB/s(alg)  new_inf=%d" % (t, times[t + 1], bytes_ / times[t + 1] / 1e6, di[t]))
        tot = sum(times[1:])
        print("episode fwd %.3f ms  -> %.2f G agent-steps/s" % (tot / 1e3, args.n * T / tot / 1e3))
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        if args.grad:
            gi = torch.ones(T, device=dev); gd = torch.ones(T, device=dev)
            for prune in (False, True, True):
                # (run_episode() hands the backward pass a copy of infected_time, so that it skips dne_dr entries after an agent's exposure)
                sv["itime_copy"], sv["itime_valid"] = st.infected_time, prune
                e0.record(); gr, gm = eng.episode_backward(sv, T, gi, gd, m); e1.record(); torch.cuda.synchronize()
                print("bwd (pruned by infected_time: %s) %.1f us" % (prune, e0.elapsed_time(e1) * 1e3), gr[:4].tolist(), gm.tolist())
        print("checksum", int((st.stage.double() * torch.arange(1, args.n + 1, device=dev).double()).sum().item()) % 1000003, int(st.next_stage_time.long().sum()), int(st.infected_time.long().sum()))
        print("stage hist", torch.bincount(st.stage.round().long().clamp(0, 3), minlength=4).tolist())
