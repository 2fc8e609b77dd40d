"""Probe: symmetric memory + barrier + peer copies inside a CUDA graph (run under torchrun on >= 2 GPUs)."""
import os, time
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
per = 1 << 20
n = per * world
try:
    symm.enable_symm_mem_for_group(dist.group.WORLD.group_name)
except Exception as e:
    print(rank, "enable_symm_mem_for_group:", repr(e)[:200], flush=True)
buf = symm.empty(n, dtype=torch.uint8, device=dev)
hdl = symm.rendezvous(buf, dist.group.WORLD)
print(rank, "rendezvous ok", hdl.world_size, hdl.buffer_size, flush=True)
buf.zero_()
torch.cuda.synchronize(); dist.barrier()

def exchange(step):
    buf[rank * per:(rank + 1) * per].fill_(rank + 1 + step)
    hdl.barrier(channel=0)
    for p in range(world):
        if p != rank:
            buf[p * per:(p + 1) * per].copy_(hdl.get_buffer(p, (per,), torch.uint8, p * per))
    hdl.barrier(channel=1)

exchange(0)
torch.cuda.synchronize()
ok = all(int(buf[p * per]) == p + 1 for p in range(world))
print(rank, "eager exchange ok:", ok, flush=True)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s):
        for step in range(1, 4):
            exchange(step)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    g.replay()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20 / 3
ok = all(int(buf[p * per]) == p + 1 + 3 for p in range(world))
print(rank, "graph exchange ok:", ok, "us per exchange: %.1f" % (dt * 1e6), flush=True)
dist.barrier()
dist.destroy_process_group()
