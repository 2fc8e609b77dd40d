"""How much of an episode is spent OUTSIDE the step kernels: host issue time vs device time of bench.py's `episode()` on one GPU,
and the launch list of one episode (torch.profiler / CUPTI).  Usage: python tools/probe_host.py --agents 1050000 [--no-grad]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from agenttorch_b200.core import Runner  # noqa: E402
from agenttorch_b200.models import covid  # noqa: E402
from agenttorch_b200.models.covid.synthetic import synthetic_config  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--agents", type=int, default=1_050_000)
p.add_argument("--degree", type=float, default=20.0)
p.add_argument("--k", type=int, default=50)
p.add_argument("--no-grad", action="store_true")
p.add_argument("--trace", action="store_true")
a = p.parse_args()
dev = torch.device("cuda", 0)
T = 21
cfg = synthetic_config(a.agents, a.degree, seed=20261019, family="rand", device="cuda", noise_seed=20261019)
runner = Runner(cfg, covid.registry)
runner.init()
tr = runner.initializer.transition_function["0"]["new_transmission"]
tr.noise_seed, tr.advance_seed = 20261019, False
cit, env = runner.state["agents"]["citizens"], runner.state["environment"]
snap = {"agents": {"citizens": {k: cit[k].clone() for k in ("disease_stage", "next_stage_time", "infected_time")}},
        "environment": {"daily_infected": env["daily_infected"].clone(), "daily_deaths": env["daily_deaths"].clone()}}
gI = torch.linspace(0.5, 1.5, T, device=dev)
gD = torch.linspace(2.0, 1.0, T, device=dev)
params_host = torch.tensor([4.75] * 3 + [0.12]).pin_memory()
prog = runner.initializer.transition_function["1"]["seirm_progression"]
grad = not a.no_grad


def episode():
    runner.restore_state(snap)
    params = params_host.to(dev, non_blocking=True)
    r2 = params[:3].view(3, 1).requires_grad_(grad)
    m = params[3:].requires_grad_(grad)
    tr.calibrate_R2 = r2
    prog.calibrate_M = m
    runner.step(T)
    e = runner.state["environment"]
    if grad:
        loss = torch.dot(gI, e["daily_infected"].reshape(-1)) + torch.dot(gD, e["daily_deaths"].reshape(-1))
        loss.backward()


for _ in range(5):
    episode()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(a.k):
    episode()
e1.record()
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"agents {a.agents} grad {grad}: device {e0.elapsed_time(e1) / a.k:.3f} ms/episode, host issue {t_issue / a.k * 1e3:.3f} ms/episode, "
      f"wall {t_all / a.k * 1e3:.3f} ms/episode", flush=True)

# host time of the sections of an episode (perf_counter between them, no synchronisation: the host runs ahead of the device
# only if it is faster, so at a size where the host bounds the episode these are issue times)
sec = {"restore": 0.0, "params": 0.0, "step": 0.0, "loss": 0.0, "backward": 0.0}
torch.cuda.synchronize()
for _ in range(200):
    t0 = time.perf_counter()
    runner.restore_state(snap)
    t1 = time.perf_counter()
    params = params_host.to(dev, non_blocking=True)
    r2 = params[:3].view(3, 1).requires_grad_(grad)
    m = params[3:].requires_grad_(grad)
    tr.calibrate_R2 = r2
    prog.calibrate_M = m
    t2 = time.perf_counter()
    runner.step(T)
    t3 = time.perf_counter()
    e = runner.state["environment"]
    t4 = t3
    if grad:
        loss = torch.dot(gI, e["daily_infected"].reshape(-1)) + torch.dot(gD, e["daily_deaths"].reshape(-1))
        t4 = time.perf_counter()
        loss.backward()
    t5 = time.perf_counter()
    for k, v in zip(sec, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
        sec[k] += v
torch.cuda.synchronize()
print("  host sections (us/episode):", {k: round(v / 200 * 1e6, 1) for k, v in sec.items()}, flush=True)

# is the batched evaluation of effective_r bit-identical to the per-week one on this device?
eye = torch.eye(3, device=dev)
bad = 0
for trial in range(2000):
    for shape in ((3, 1), (3,)):
        r = (torch.rand(shape, device=dev) * 9 + 0.1)
        per = torch.stack([(r * eye[w]).sum() for w in range(3)])
        bat = (r.unsqueeze(0) * eye.view(3, *([1] * (r.dim() - 1)), 3)).flatten(1).sum(1)
        bad += int(not torch.equal(per, bat))
print("batched effective_r differs from per-week in", bad, "of 4000 trials", flush=True)

# parts of the host time
def clock(fn, k=200):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k):
        fn()
    dt = (time.perf_counter() - t0) / k * 1e3
    torch.cuda.synchronize()
    return dt


print(f"  restore_state host {clock(lambda: runner.restore_state(snap)):.3f} ms", flush=True)
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        episode()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70), flush=True)
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=45, max_name_column_width=70), flush=True)
if a.trace:
    prof.export_chrome_trace("gpurun_out/probe_host_trace.json")
import cProfile  # noqa: E402
import pstats  # noqa: E402

pr = cProfile.Profile()
pr.enable()
for _ in range(100):
    episode()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(70)
