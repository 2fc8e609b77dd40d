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
    # "comm": a graph with community structure (99.8 % of an agent's contacts inside its block of 1024 ids), exchanged like "sync":
    # the rank's edges refer to a part of the peer's agents only, the halo mask of the exchange leaves the rest alone
    comm = cuda_graph == "comm"
    gkw = {"family": "comm", "block": 1024, "p_local": 0.998} if comm else {}
    part = cuda_graph == "part"
    if part:
        # the community structure hidden behind RANDOM agent ids, recovered by the partitioner (agenttorch_b200/partition.py: balanced
        # label propagation into 2 parts, relabelled part by part with the boundary agents last) — the same relabelling on one GPU
        # and on two, so both simulate the same population; exchanged like "sync"
        from agenttorch_b200 import partition
        from agenttorch_b200.graph import build_contact_graph, shard_graph

        hidden = synthetic_contact_graph(N, DEG, SEED + 1, device=dev, family="comm", block=1024, p_local=0.998)
        rp = hidden.row_ptr.long()
        perm = torch.randperm(N, generator=torch.Generator().manual_seed(SEED + 2)).to(dev)
        edges = torch.stack([perm[hidden.col_idx[: hidden.num_edges].long()],
                             perm[torch.repeat_interleave(torch.arange(N, device=dev), rp[1:] - rp[:-1])]])
        labels = partition.label_propagation_partition(edges, N, 2, rounds=15)
        cut_before = partition.edge_cut_fraction(edges[0], edges[1], partition.label_propagation_partition(edges, N, 2, rounds=0))
        cut_after = partition.edge_cut_fraction(edges[0], edges[1], labels)
        edges, moved, _ = partition.reorder_population(edges, {"age": age, "stage0": stage0}, labels)
        age, stage0 = moved["age"], moved["stage0"]
        full = build_contact_graph(edges, N, mean_interactions=mean_interactions_from_age(age), device=dev)
        graph = shard_graph(full, r0, r1)
        comm = True                                                # the halo checks below apply
    else:
        full = synthetic_contact_graph(N, DEG, SEED + 1, device=dev, mean_interactions=mean_interactions_from_age(age), **gkw)
        graph = synthetic_contact_graph(N, DEG, SEED + 1, device=dev, row_begin=r0, row_end=r1, **gkw)
        graph.i_bar, graph.i_bar_agent = full.i_bar, full.i_bar_agent
    dense = cuda_graph == "dense"                                  # symmetric memory, barrier + whole-slice pull
    sync = cuda_graph in ("sync", "comm", "part")                  # sparse exchange on the critical path (no look-ahead encoding)
    cuda_graph = bool(cuda_graph)
    eng = SeirmEngine(graph, SeirmSpec(num_steps=T), process_group=pg, code_len=shard_code_len(N, world),
                      symmetric=None if cuda_graph else False,      # False: NCCL all-gather path
                      sparse_exchange=not dense, overlap_exchange=bool(cuda_graph) and not (sync or dense) and world > 1)
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
    eng.check_errors()
    halo = None
    if world > 1 and comm:
        groups = torch.zeros(eng.n_code // 32, dtype=torch.bool, device=dev)
        groups[eng.halo_groups] = True
        outside = ~groups.repeat_interleave(32)
        outside[r0:r1] = False                                     # the rank's own slice is not pulled
        halo = {"fraction": eng.halo_fraction, "cut": (cut_before, cut_after) if part else None,
                "stray": sum(int((buf[: eng.n_code][outside] != 0).sum()) for buf in eng.xbuf),
                "pulled": sum(int((buf[: eng.n_code][~outside] != 0).sum()) for buf in eng.xbuf)}
    torch.save({"stage": st.stage.cpu(), "nst": st.next_stage_time.cpu(), "di": di.cpu(), "dd": dd.cpu(), "gr": gr.cpu(), "gm": gm.cpu(),
                "bounds": bounds, "halo": halo}, os.path.join(out_dir, f"w{world}_r{rank}.pt"))
    if world > 1:
        dist.destroy_process_group()


@pytest.mark.parametrize("cuda_graph", [False, True, "sync", "dense", "comm", "part"])
def test_two_gpus_reproduce_one_gpu(tmp_path, cuda_graph):
    """False: NCCL all-gather, issued step by step.  True: symmetric memory, one kernel per step with the rendezvous
    inside and a bitmap-directed sparse pull, issued on a side stream for step t+2 while step t+1 computes (look-ahead
    encoding), the whole two-stream episode replayed from one CUDA graph per rank.  "sync": the same exchange on the critical path.
    "dense": symmetric memory with a separate barrier and a whole-slice pull (the round-1 first version)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    mp.spawn(_episode, args=(1, 0, str(tmp_path), cuda_graph if cuda_graph in ("comm", "part") else False), nprocs=1, join=True)
    mp.spawn(_episode, args=(2, _free_port(), str(tmp_path), cuda_graph), nprocs=2, join=True)
    one = torch.load(os.path.join(str(tmp_path), "w1_r0.pt"))
    parts = [torch.load(os.path.join(str(tmp_path), f"w2_r{r}.pt")) for r in range(2)]
    assert torch.equal(torch.cat([p["stage"] for p in parts]), one["stage"])
    assert torch.equal(torch.cat([p["nst"] for p in parts]), one["nst"])
    assert torch.equal(parts[0]["di"], one["di"])
    # deaths: every rank adds f32(count_r) * M, so the all-reduced series differs from f32(count) * M by f32 rounding
    np.testing.assert_allclose(parts[0]["dd"].numpy(), one["dd"].numpy(), rtol=1e-6)
    assert float(one["di"].sum()) > 1000                        # a real epidemic happened
    if cuda_graph in ("comm", "part"):
        for p in parts:                                         # nothing outside the halo mask was ever written, something inside was
            assert p["halo"]["stray"] == 0 and p["halo"]["pulled"] > 0 and p["halo"]["fraction"] < 0.95
            if cuda_graph == "part":                            # random ids cut half of the edges; the partitioner finds the communities
                assert p["halo"]["cut"][0] > 0.45 and p["halo"]["cut"][1] < 0.1
    np.testing.assert_allclose(parts[0]["gr"].numpy(), one["gr"].numpy(), rtol=1e-6)
    np.testing.assert_allclose(parts[0]["gm"].numpy(), one["gm"].numpy(), rtol=1e-6)


def _distributed_runner(rank, world, port, out_dir, pop_dir):
    import torch.distributed as dist

    from agenttorch_b200.core import envs
    from agenttorch_b200.models import covid
    from tests.helpers import load_golden

    g = load_golden("synth3k_soft")
    torch.cuda.set_device(rank)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    runner = envs.create(model=covid, population=pop_dir, distributed=True, world_size=world)
    dev = runner.initializer.device
    r2 = torch.tensor(g["r2"], dtype=torch.float32, device=dev, requires_grad=True)
    m = torch.tensor(g["m"], dtype=torch.float32, device=dev, requires_grad=True)
    runner._set_parameters({"initializer.transition_function.0.new_transmission.learnable_args.R2": r2,
                            "initializer.transition_function.1.seirm_progression.learnable_args.M": m})
    runner.tr.noise_seed, runner.tr.advance_seed = g["seed"], False
    runner.step(g["num_steps"])
    env = runner.state_trajectory[-1][-1]["environment"]
    loss = (torch.from_numpy(g["gI"]).to(dev) * env["daily_infected"].reshape(-1)).sum() + \
           (torch.from_numpy(g["gD"]).to(dev) * env["daily_deaths"].reshape(-1)).sum()
    loss.backward()
    final = runner.gather_final_state()
    assert runner.agent_partition_map["local_size"] == runner.state["agents"]["citizens"]["disease_stage"].shape[0]
    torch.save({"stage": final["disease_stage"].cpu(), "nst": final["next_stage_time"].cpu(), "di": env["daily_infected"].detach().cpu(),
                "dd": env["daily_deaths"].detach().cpu(), "gr2": r2.grad.cpu(), "gm": m.grad.cpu()}, os.path.join(out_dir, f"dr{rank}.pt"))
    dist.destroy_process_group()


def test_distributed_runner_api_matches_reference_golden(tmp_path):
    """envs.create(distributed=True) on 2 GPUs reproduces the REAL reference's golden outputs"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from agenttorch_b200.populations import write_population_folder
    from tests.helpers import load_golden, rel_err

    g = load_golden("synth3k_soft")
    pop = write_population_folder(str(tmp_path / "pop"), g["age"], g["stage0"], g["edges"])
    mp.spawn(_distributed_runner, args=(2, _free_port(), str(tmp_path), pop), nprocs=2, join=True)
    for r in range(2):
        o = torch.load(os.path.join(str(tmp_path), f"dr{r}.pt"))
        assert np.array_equal(o["stage"].numpy().reshape(-1).astype(np.float64), g["ref_stage"])
        assert np.array_equal(o["nst"].numpy().reshape(-1).astype(np.int64), g["ref_next_stage_time"].astype(np.int64))
        assert tuple(o["dd"].shape) == (1, g["num_steps"])
        assert rel_err(g["ref_daily_infected"], o["di"].numpy()) <= 1e-5 and rel_err(g["ref_daily_deaths"], o["dd"].numpy()) <= 1e-5
        assert rel_err(g["ref_grad_R2"], o["gr2"].numpy()) <= 1e-4 and rel_err(g["ref_grad_M"], o["gm"].numpy()) <= 1e-4
