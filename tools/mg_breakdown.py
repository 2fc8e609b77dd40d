"""Where a multi-GPU SEIRM episode spends its time (run under torchrun on N >= 2 GPUs).
Eager (non-graph) episode with CUDA events around the two pieces of every step — step kernel, halo-exchange kernel
(rendezvous + sparse pull) — plus the per-episode extras (restore, all-reduces, backward) and the captured-graph total."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from agenttorch_b200.engine import NoiseSpec, SeirmEngine, SeirmSpec, SeirmState
from agenttorch_b200.graph import (mean_interactions_from_age, shard_bounds, shard_code_len, synthetic_contact_graph,
                                   synthetic_population)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
pg = dist.group.WORLD
N, d, T, SEED = 8_400_000, 20.0, 21, 20261019
r0, r1 = shard_bounds(N, world)[rank]
age_all, stage_all = synthetic_population(N, SEED, device=dev)
mi = mean_interactions_from_age(age_all)
graph = synthetic_contact_graph(N, d, SEED + 1, device=dev, row_begin=r0, row_end=r1)
full0 = synthetic_contact_graph(N, d, SEED + 1, device=dev, row_begin=0, row_end=min(1 << 20, N), mean_interactions=mi)
graph.i_bar, graph.i_bar_agent = full0.i_bar, full0.i_bar_agent
del full0
engine = SeirmEngine(graph, SeirmSpec(num_steps=T), process_group=pg, code_len=shard_code_len(N, world))
assert engine.code_hdl is not None, "symmetric memory not available"
st0 = SeirmState.from_initial(stage_all[r0:r1], age_all[r0:r1], T, dev, packed=False)
st = st0.clone()
r_steps = torch.full((T,), 14.25, device=dev); m_val = torch.tensor([0.12], device=dev)
gI = torch.linspace(0.5, 1.5, T, device=dev); gD = torch.linspace(1.5, 0.5, T, device=dev)
noise = NoiseSpec(seed=SEED, isolation="soft")

marks = []
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

orig_exchange = engine.exchange
def timed_exchange(buf_index):
    a = ev()
    b = a
    orig_exchange(buf_index)                      # one kernel: rendezvous + sparse pull + filter fold
    c = ev()
    marks.append((a, b, c))
orig_launch = engine._submit
lmarks = []
def timed_launch(*a, **k):
    x = ev(); orig_launch(*a, **k); y = ev(); lmarks.append((x, y))

def restore():
    for k in ("stage", "next_stage_time", "infected_time"):
        getattr(st, k).copy_(getattr(st0, k))
    st.daily_infected.zero_(); st.daily_deaths.zero_()

# ---- eager, instrumented
for rep in range(3):
    restore(); marks.clear(); lmarks.clear()
    engine.exchange, engine._submit = timed_exchange, timed_launch
    dist.barrier(); torch.cuda.synchronize()
    engine.run_episode(st, r_steps, m_val, T, noise, want_grad=True, cuda_graph=False)
    torch.cuda.synchronize()
    engine.exchange, engine._submit = orig_exchange, orig_launch
bar = [a.elapsed_time(b) * 1e3 for a, b, c in marks]; pull = [b.elapsed_time(c) * 1e3 for a, b, c in marks]
kern = [x.elapsed_time(y) * 1e3 for x, y in lmarks]
# ---- graph total and per-episode extras
def tm(fn, k=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize(); e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k * 1e3
sv = None
def fwd():
    global sv
    restore()
    sv = engine.run_episode(st, r_steps, m_val, T, noise, want_grad=True, cuda_graph=True)
for _ in range(3): fwd()
t_fwd = tm(fwd)
t_restore = tm(restore)
def reduces():
    from agenttorch_b200.parallel import reduce_daily_series
    reduce_daily_series(st, pg)
t_red = tm(reduces)
t_bwd = tm(lambda: engine.episode_backward(sv, T, gI, gD, m_val))
if rank == 0:
    f = lambda v: f"n={len(v)} med={statistics.median(v):.1f} min={min(v):.1f} max={max(v):.1f} sum={sum(v):.0f}"
    print(f"world {world}: step kernels  us: {f(kern)}")
    print(f"world {world}: exchange (rendezvous + sparse pull) us: {f(pull)}")
    print(f"world {world}: graph episode fwd incl restore {t_fwd:.0f} us, restore {t_restore:.0f} us, packed all-reduce of the series {t_red:.0f} us, backward (incl. its all-reduce) {t_bwd:.0f} us")
    print("kernels:", " ".join(f"{v:.0f}" for v in kern))
    print("pull:", " ".join(f"{v:.0f}" for v in pull))
dist.barrier()
dist.destroy_process_group()
