"""BASELINE configs[3]: predator-prey grid model, 10 M agents on a 4096 x 4096 grid, one B200.
Times the fused multi-step path (CUDA graph) and prints algorithmic bytes per step."""
import argparse, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from agenttorch_b200.predprey import PredPreyEngine, PredPreySpec, PredPreyState

ap = argparse.ArgumentParser()
ap.add_argument("--side", type=int, default=4096)
ap.add_argument("--pred", type=int, default=2_000_000)
ap.add_argument("--prey", type=int, default=8_000_000)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--reps", type=int, default=7)
ap.add_argument("--f64", type=int, default=0)
ap.add_argument("--grid", type=int, default=0)
ap.add_argument("--sort", type=int, default=1)
ap.add_argument("--resort", type=int, default=0)
ap.add_argument("--tiled", type=int, default=1)
args = ap.parse_args()
dev = "cuda:0"
mx = my = args.side
C = mx * my
g = torch.Generator(device=dev); g.manual_seed(1)
spec = PredPreySpec(mx, my, C, mx, tiled=bool(args.tiled and args.sort))
pos = lambda n: torch.stack([torch.randint(0, mx, (n,), generator=g, device=dev), torch.randint(0, my, (n,), generator=g, device=dev)], 1)
cd = torch.where(torch.rand(C, generator=g, device=dev) < 0.4, 0.0, torch.randint(1, 101, (C,), generator=g, device=dev).float())
if args.f64:
    cd = cd.double()
st0 = PredPreyState.from_reference(spec, pos(args.pred), 30 + 71 * torch.rand(args.pred, 1, generator=g, device=dev), pos(args.prey),
                                   40 + 61 * torch.rand(args.prey, 1, generator=g, device=dev), (cd <= 0).long(), cd)
if args.sort:
    st0.sort_by_cell(spec)
eng = PredPreyEngine(spec, dev, grid=args.grid)
st = st0.clone()
times = []
for i in range(args.reps + 2):
    st.copy_(st0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    eng.run(st, args.steps, 7, cuda_graph=True, resort_every=args.resort if args.sort else 0)
    e1.record(); torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
ms = statistics.median(times[2:])
n = args.pred + args.prey
cd_b = 8 if args.f64 else 4
bytes_step = n * (20 if args.sort else 16) + C * 2 * (cd_b + 1)      # agents: cell + energy read and written; cells: info byte + countdown read and written
print(f"sort={args.sort} tiled={args.tiled} resort={args.resort} predator-prey {n/1e6:.1f} M agents, {mx}x{my} cells, {args.steps} steps: {ms:.3f} ms -> {ms/args.steps*1e3:.1f} us/step, "
      f"{n*args.steps/ms/1e6:.1f} G agent-steps/s, algorithmic {bytes_step/1e6:.0f} MB/step -> {bytes_step*args.steps/ms/1e9:.2f} TB/s; census {eng.census(st)}")
