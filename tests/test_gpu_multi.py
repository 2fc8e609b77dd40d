"""2-GPU parity (runs only where >= 2 CUDA devices are visible): the sharded engine with the NCCL halo exchange must
reproduce the single-GPU result bit for bit (noise is keyed by global agent id, sums are fixed-point)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N, DEG, T, SEED = 200_000, 12.0, 21, 5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _episode(rank, world, port, out_dir, cuda_graph):
    import torch.distributed as dist

    from agenttorch_b200.engine import NoiseSpec, SeirmEngine, SeirmSpec, SeirmState
    from agenttorch_b200.graph import (mean_interactions_from_age, shard_bounds, shard_code_len, synthetic_contact_graph,
                                       synthetic_population)

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    pg = None
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        pg = dist.group.WORLD
    bounds = shard_bounds(N, world)
    r0, r1 = bounds[rank]
    age, stage0 = synthetic_population(N, SEED, device=dev)
    full = synthetic_contact_graph(N, DEG, SEED + 1, device=dev, mean_interactions=mean_interactions_from_age(age))
    graph = synthetic_contact_graph(N, DEG, SEED + 1, device=dev, row_begin=r0, row_end=r1)
    graph.i_bar, graph.i_bar_agent = full.i_bar, full.i_bar_agent
    eng = SeirmEngine(graph, SeirmSpec(num_steps=T), process_group=pg, code_len=shard_code_len(N, world),
                      symmetric=None if cuda_graph else False)      # False: NCCL all-gather path
    st = SeirmState.from_initial(stage0[r0:r1], age[r0:r1], T, dev)
    r_steps = torch.full((T,), 14.25, device=dev)
    m = torch.tensor([0.12], device=dev)
    for _ in range(2 if cuda_graph else 1):      # second pass replays the captured graph
        st = SeirmState.from_initial(stage0[r0:r1], age[r0:r1], T, dev) if not cuda_graph else st
        if cuda_graph:
            fresh = SeirmState.from_initial(stage0[r0:r1], age[r0:r1], T, dev)
            for k in ("stage", "next_stage_time", "infected_time", "daily_infected", "daily_deaths"):
                getattr(st, k).copy_(getattr(fresh, k))
        sv = eng.run_episode(st, r_steps, m, T, NoiseSpec(seed=9, isolation="hard"), want_grad=True, cuda_graph=cuda_graph)
    di, dd = st.daily_infected.clone(), st.daily_deaths.clone()
    if world > 1:
        dist.all_reduce(di); dist.all_reduce(dd)
    gr, gm = eng.episode_backward(sv, T, torch.linspace(0.5, 1.5, T, device=dev), torch.linspace(2, 1, T, device=dev), m)
    torch.save({"stage": st.stage.cpu(), "nst": st.next_stage_time.cpu(), "di": di.cpu(), "dd": dd.cpu(), "gr": gr.cpu(), "gm": gm.cpu(),
                "bounds": bounds}, os.path.join(out_dir, f"w{world}_r{rank}.pt"))
    if world > 1:
        dist.destroy_process_group()


@pytest.mark.parametrize("cuda_graph", [False, True])
def test_two_gpus_reproduce_one_gpu(tmp_path, cuda_graph):
    """cuda_graph=True: symmetric-memory halo exchange (barrier + peer copies) replayed from one CUDA graph per rank"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    mp.spawn(_episode, args=(1, 0, str(tmp_path), False), nprocs=1, join=True)
    mp.spawn(_episode, args=(2, _free_port(), str(tmp_path), cuda_graph), nprocs=2, join=True)
    one = torch.load(os.path.join(str(tmp_path), "w1_r0.pt"))
    parts = [torch.load(os.path.join(str(tmp_path), f"w2_r{r}.pt")) for r in range(2)]
    assert torch.equal(torch.cat([p["stage"] for p in parts]), one["stage"])
    assert torch.equal(torch.cat([p["nst"] for p in parts]), one["nst"])
    assert torch.equal(parts[0]["di"], one["di"])
    # deaths: every rank adds f32(count_r) * M, so the all-reduced series differs from f32(count) * M by f32 rounding
    np.testing.assert_allclose(parts[0]["dd"].numpy(), one["dd"].numpy(), rtol=1e-6)
    assert float(one["di"].sum()) > 1000                        # a real epidemic happened
    np.testing.assert_allclose(parts[0]["gr"].numpy(), one["gr"].numpy(), rtol=1e-6)
    np.testing.assert_allclose(parts[0]["gm"].numpy(), one["gm"].numpy(), rtol=1e-6)
