"""world_size-2 gloo test of the multi-GPU host logic (no kernels): contiguous ownership, shard generation,
in-place halo exchange of the code/bitmap slices, per-episode reductions.  The lambda each rank would compute from
the exchanged arrays is emulated with torch ops and compared with the single-process result."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

N, DEG, SEED = 6144, 5.0, 17


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _lambda_like(graph, code):
    """sum over in-edges of the (float) code of the source: the gather pattern of the step kernel"""
    deg = (graph.row_ptr[1:] - graph.row_ptr[:-1]).long()
    dst = torch.repeat_interleave(torch.arange(graph.n_rows), deg)
    vals = code[graph.col_idx[: graph.num_edges].long()].double()
    return torch.zeros(graph.n_rows, dtype=torch.float64).index_add(0, dst, vals)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from agenttorch_b200.graph import shard_bounds, synthetic_contact_graph
        from agenttorch_b200.parallel import halo_exchange, reduce_episode_outputs

        bounds = shard_bounds(N, world, align=1024)
        r0, r1 = bounds[rank]
        n_code = (bounds[0][1] - bounds[0][0]) * world
        shard = synthetic_contact_graph(N, DEG, SEED, row_begin=r0, row_end=r1)
        assert shard.row_offset == r0 and shard.n_rows == r1 - r0 and shard.n_sources == N
        # every rank encodes ITS agents (here: a deterministic function of the global id), then the slices are exchanged
        code = torch.zeros(n_code, dtype=torch.uint8)
        ids = torch.arange(r0, r1)
        code[r0:r1] = ((ids * 7 + 3) % 11 == 0).to(torch.uint8) * ((ids % 5) + 1).to(torch.uint8)
        hot = torch.zeros(n_code // 32 + 4, dtype=torch.int32)
        bits = (code[r0:r1] != 0).view(-1, 32).to(torch.int64)
        words = (bits << torch.arange(32)).sum(1)
        hot[r0 // 32: r1 // 32] = torch.where(words >= 2**31, words - 2**32, words).to(torch.int32)
        halo_exchange(code, hot, n_code)
        lam = _lambda_like(shard, code)
        total = reduce_episode_outputs([lam.sum().reshape(1).clone()])[0]
        torch.save({"lam": lam, "code": code, "hot": hot, "total": total, "bounds": bounds}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_two_rank_halo_exchange_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from agenttorch_b200.graph import synthetic_contact_graph

    full = synthetic_contact_graph(N, DEG, SEED)
    ids = torch.arange(N)
    code = ((ids * 7 + 3) % 11 == 0).to(torch.uint8) * ((ids % 5) + 1).to(torch.uint8)
    ref = _lambda_like(full, code)
    outs = [torch.load(os.path.join(str(tmp_path), f"r{r}.pt")) for r in range(world)]
    for r, o in enumerate(outs):
        r0, r1 = o["bounds"][r]
        assert torch.equal(o["code"][:N], code)                                  # every rank sees every slice
        assert torch.equal(o["lam"], ref[r0:r1])                                 # partition-invariant result
        assert float(o["total"]) == float(ref.sum())                             # per-episode all-reduce
        bits = ((o["hot"][: N // 32].to(torch.int64).unsqueeze(1) >> torch.arange(32)) & 1).reshape(-1)
        assert torch.equal(bits.bool(), code != 0)
    assert torch.equal(outs[0]["hot"], outs[1]["hot"])
