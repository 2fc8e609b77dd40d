"""The SHARDED SEIRM engine on several emulated "GPUs": one OS process per rank (gloo process group), every rank loads
csrc/at_seirm.cu compiled for the host (tests/host_emulation/), and the per-step halo exchange runs

  * "allgather": through torch.distributed (the engine's NCCL fallback path) + `hot_from_code_kernel`;
  * "sparse" / "dense": through the engine's symmetric-memory path — the real `exchange_sparse_kernel` (rendezvous INSIDE the
    kernel via release/acquire flags in the peers' buffers, bitmap-directed pull) and `pull_codes_kernel` — with
    `torch.distributed._symmetric_memory` replaced by buffers in POSIX shared memory that every rank maps (peer "pointers" are
    ordinary addresses of the mapped peers' buffers).

Every variant must reproduce the single-rank result bit for bit (noise is keyed by global agent id, the daily sums are fixed
point).  The 2-GPU tests on real hardware are tests/test_gpu_multi.py."""
import contextlib
import os
import socket
import subprocess
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# (overridable through the environment by tools/fuzz_emulated_multirank.py; spawned workers re-import this module)
N, DEG, T, SEED = int(os.environ.get("AT_EMU_N", 4096)), float(os.environ.get("AT_EMU_DEG", 6.0)), 21, int(os.environ.get("AT_EMU_SEED", 5))
R_VALUE, EXPOSED0 = float(os.environ.get("AT_EMU_R", 9.0)), float(os.environ.get("AT_EMU_EXPOSED", 0.03))
ISOLATION, HASHED = os.environ.get("AT_EMU_ISOLATION", "hard"), os.environ.get("AT_EMU_HASHED", "0") == "1"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.fixture(scope="module")
def emulated_library(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("seirm_emu_mr") / "libat_seirm_emulated.so")
    cmd = ["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "host_emulation", "seirm_emulation.cpp"), "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return out


class _SharedSymmetricMemory(types.ModuleType):
    """Stand-in for torch.distributed._symmetric_memory on the host: `empty` allocates in a file under /dev/shm,
    `rendezvous` maps every peer's file; `buffer_ptrs` are the addresses of those mappings in THIS process."""

    def __init__(self, tag, rank, world):
        super().__init__("torch.distributed._symmetric_memory")
        self.tag, self.rank, self.world, self.count, self.keep = tag, rank, world, 0, []

    def _path(self, rank, index):
        return f"/dev/shm/at_emu_{self.tag}_{index}_r{rank}"

    def empty(self, nbytes, dtype=torch.uint8, device=None):
        assert dtype == torch.uint8
        path = self._path(self.rank, self.count)
        t = torch.from_file(path, shared=True, size=int(nbytes), dtype=torch.uint8)
        t._at_symm_index = self.count
        self.count += 1
        return t

    def rendezvous(self, tensor, group):
        import torch.distributed as dist

        dist.barrier(group=group)                                        # every rank has created its file
        index, nbytes = tensor._at_symm_index, tensor.numel()
        peers = [tensor if r == self.rank else torch.from_file(self._path(r, index), shared=True, size=nbytes, dtype=torch.uint8)
                 for r in range(self.world)]
        self.keep.append(peers)
        return types.SimpleNamespace(buffer_ptrs=[p.data_ptr() for p in peers],
                                     barrier=lambda channel=0: dist.barrier(group=group))

    def cleanup(self):
        for index in range(self.count):
            with contextlib.suppress(OSError):
                os.remove(self._path(self.rank, index))


def _episode(rank, world, port, out_dir, lib_path, exchange, tag):
    sys.path.insert(0, ROOT)
    # "<exchange>+comm": a contact graph with community structure (99.5 % of an agent's contacts inside its block of 256 ids), so
    # that a rank's edges refer to a part of the peers' agents only and the halo mask of the exchange has something to leave out
    label = exchange
    exchange, _, family = exchange.partition("+")
    gkw = {"family": "comm", "block": 256, "p_local": 0.995} if family == "comm" else {}
    import torch.distributed as dist

    from agenttorch_b200 import _native, engine
    from agenttorch_b200.engine import NoiseSpec, SeirmEngine, SeirmSpec, SeirmState
    from agenttorch_b200.graph import (mean_interactions_from_age, shard_bounds, shard_code_len, synthetic_contact_graph,
                                       synthetic_population)

    # the emulated library instead of libat_seirm.so; no CUDA checks, streams or device contexts in this process
    _native.LIB_PATH, _native._lib = lib_path, None
    engine._require_cuda = lambda message, is_cuda: None
    torch.cuda.device = lambda device: contextlib.nullcontext()
    torch.cuda.current_stream = lambda device=None: types.SimpleNamespace(cuda_stream=0)
    torch.cuda.synchronize = lambda device=None: None
    torch.set_num_threads(1)
    pg, symm = None, None
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        pg = dist.group.WORLD
        if exchange != "allgather":
            symm = _SharedSymmetricMemory(tag, rank, world)
            sys.modules["torch.distributed._symmetric_memory"] = symm
            torch.distributed._symmetric_memory = symm
    try:
        bounds = shard_bounds(N, world)
        r0, r1 = bounds[rank]
        age, stage0 = synthetic_population(N, SEED, initial_exposed=EXPOSED0, device="cpu")
        if family == "part":
            # a community graph whose agents carry RANDOM ids (the structure is there, the numbering hides it), partitioned by
            # agenttorch_b200.partition into 4 parts and relabelled part by part, boundary agents last — the same relabelling at
            # every world size, so one rank and several ranks simulate the same population
            from agenttorch_b200 import partition
            from agenttorch_b200.graph import build_contact_graph, shard_graph

            hidden = synthetic_contact_graph(N, DEG, SEED + 1, device="cpu", family="comm", block=256, p_local=0.995)
            rp = hidden.row_ptr.long()
            perm = torch.randperm(N, generator=torch.Generator().manual_seed(SEED + 2))
            edges = torch.stack([perm[hidden.col_idx[: hidden.num_edges].long()],
                                 perm[torch.repeat_interleave(torch.arange(N), rp[1:] - rp[:-1])]])
            labels = partition.label_propagation_partition(edges, N, 4, rounds=15)
            edges, moved, _ = partition.reorder_population(edges, {"age": age, "stage0": stage0}, labels)
            age, stage0 = moved["age"], moved["stage0"]
            full = build_contact_graph(edges, N, mean_interactions=mean_interactions_from_age(age), device="cpu")
            graph = shard_graph(full, r0, r1)
        else:
            full = synthetic_contact_graph(N, DEG, SEED + 1, device="cpu", mean_interactions=mean_interactions_from_age(age), **gkw)
            graph = synthetic_contact_graph(N, DEG, SEED + 1, device="cpu", row_begin=r0, row_end=r1, **gkw)
            graph.i_bar, graph.i_bar_agent = full.i_bar, full.i_bar_agent
        eng = SeirmEngine(graph, SeirmSpec(num_steps=T), process_group=pg, code_len=shard_code_len(N, world), warps_per_cta=2,
                          symmetric=(exchange != "allgather") if world > 1 else None, sparse_exchange=exchange.startswith("sparse"),
                          overlap_exchange=(exchange == "sparse") if world > 1 else None,
                          hashed=True if HASHED else None, hash_words=32 if HASHED else None)
        if world > 1:
            assert (eng.code_hdl is not None) == (exchange != "allgather")
            assert eng.overlap == (exchange == "sparse")         # lam_gamma_integrals[0] == 0: the exchange of t+2 runs beside step t+1
        st = SeirmState.from_initial(stage0[r0:r1], age[r0:r1], T, torch.device("cpu"))
        r_steps, m = torch.full((T,), R_VALUE), torch.tensor([0.12])
        sv = eng.run_episode(st, r_steps, m, T, NoiseSpec(seed=9, isolation=ISOLATION), want_grad=True, cuda_graph=False)
        di, dd = st.daily_infected.clone(), st.daily_deaths.clone()
        if world > 1:
            dist.all_reduce(di); dist.all_reduce(dd)
        gr, gm = eng.episode_backward(sv, T, torch.linspace(0.5, 1.5, T), torch.linspace(2, 1, T), m)
        eng.check_errors()
        halo = None
        if world > 1 and exchange.startswith("sparse"):
            # the rank's copies of the peers' code bytes: nothing outside the halo mask may ever have been written
            groups = torch.zeros(eng.n_code // 32, dtype=torch.bool)
            groups[eng.halo_groups] = True
            outside = ~groups.repeat_interleave(32)
            outside[r0:r1] = False                                       # the rank's own slice is not pulled
            stray = sum(int((buf[: eng.n_code][outside] != 0).sum()) for buf in eng.xbuf)
            pulled_any = sum(int((buf[: eng.n_code][~outside] != 0).sum()) for buf in eng.xbuf)
            halo = {"fraction": eng.halo_fraction, "stray": stray, "pulled": pulled_any}
        torch.save({"stage": st.stage, "nst": st.next_stage_time, "itime": st.infected_time, "di": di, "dd": dd, "gr": gr, "gm": gm,
                    "halo": halo}, os.path.join(out_dir, f"{label}_w{world}_r{rank}.pt"))
    finally:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        if symm is not None:
            symm.cleanup()


@pytest.fixture(scope="module")
def single_rank(emulated_library, tmp_path_factory):
    import torch.multiprocessing as mp

    out = str(tmp_path_factory.mktemp("single"))
    mp.spawn(_episode, args=(1, 0, out, emulated_library, "single", "single"), nprocs=1, join=True)
    return torch.load(os.path.join(out, "single_w1_r0.pt"))


def test_halo_exchange_pulls_only_what_the_ranks_edges_refer_to(emulated_library, tmp_path):
    """a community graph hidden behind random agent ids, recovered by the partitioner (agenttorch_b200/partition.py) before the ranks
    take their ranges; 2 emulated ranks: same trajectories as one rank, and the sparse exchange neither reads nor writes the
    32-agent groups no edge of the rank refers to (include/at_seirm.h, halo_mask).  ("+comm", a community graph in its natural
    order, runs in tests/test_gpu_multi.py.)"""
    import torch.multiprocessing as mp

    out = str(tmp_path)
    mp.spawn(_episode, args=(1, 0, out, emulated_library, "single+part", "single"), nprocs=1, join=True)
    one = torch.load(os.path.join(out, "single+part_w1_r0.pt"))
    assert float(one["di"].sum()) > 50                                                       # an epidemic happened
    exchange, world = "sparse_sync+part", 2
    mp.spawn(_episode, args=(world, _free_port(), out, emulated_library, exchange, f"{os.getpid()}_halo"), nprocs=world, join=True)
    parts = [torch.load(os.path.join(out, f"{exchange}_w{world}_r{r}.pt")) for r in range(world)]
    assert torch.equal(torch.cat([p["stage"] for p in parts]), one["stage"])
    assert torch.equal(torch.cat([p["nst"] for p in parts]), one["nst"])
    assert torch.equal(torch.cat([p["itime"] for p in parts]), one["itime"])
    assert torch.equal(parts[0]["di"], one["di"])
    for p in parts:
        assert p["halo"]["stray"] == 0 and p["halo"]["pulled"] > 0
        assert p["halo"]["fraction"] < 0.95                              # the mask really leaves groups out


@pytest.mark.parametrize("exchange,world", [("allgather", 2), ("sparse", 2), ("sparse_sync", 2), ("dense", 2), ("sparse", 4)])
def test_sharded_engine_reproduces_single_rank(emulated_library, single_rank, tmp_path, exchange, world):
    import torch.multiprocessing as mp

    out, one = str(tmp_path), single_rank
    tag = f"{os.getpid()}_{exchange}_{world}"
    mp.spawn(_episode, args=(world, _free_port(), out, emulated_library, exchange, tag), nprocs=world, join=True)
    parts = [torch.load(os.path.join(out, f"{exchange}_w{world}_r{r}.pt")) for r in range(world)]
    assert torch.equal(torch.cat([p["stage"] for p in parts]), one["stage"])
    assert torch.equal(torch.cat([p["nst"] for p in parts]), one["nst"])
    assert torch.equal(torch.cat([p["itime"] for p in parts]), one["itime"])
    assert torch.equal(parts[0]["di"], one["di"]) and float(one["di"].sum()) > 100            # an epidemic happened
    np.testing.assert_allclose(parts[0]["dd"].numpy(), one["dd"].numpy(), rtol=1e-6)          # f32(count_r) * M per rank
    np.testing.assert_allclose(parts[0]["gr"].numpy(), one["gr"].numpy(), rtol=1e-6)
    np.testing.assert_allclose(parts[0]["gm"].numpy(), one["gm"].numpy(), rtol=1e-6)


# ------------------------------------------------------------------------------------------------------------------
# envs.create(distributed=True) — the reference's multi-GPU entry point — on two emulated ranks, against the real reference's
# golden outputs (the 2-GPU version is tests/test_gpu_multi.py::test_distributed_runner_api_matches_reference_golden).
def _distributed_runner(rank, world, port, out_dir, lib_path, pop_dir, tag, partition=None):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from agenttorch_b200 import _native, distributed, engine
    from agenttorch_b200.core import envs
    from agenttorch_b200.core.runner import Runner
    from agenttorch_b200.models import covid
    from agenttorch_b200.models.covid import context
    from tests.helpers import load_golden

    _native.LIB_PATH, _native._lib = lib_path, None
    no_check = lambda *args: None                                                # noqa: E731
    engine._require_cuda = context._require_cuda = distributed._require_cuda = no_check
    torch.cuda.device = lambda device: contextlib.nullcontext()
    torch.cuda.current_stream = lambda device=None: types.SimpleNamespace(cuda_stream=0)
    torch.cuda.synchronize = lambda device=None: None
    torch.set_num_threads(1)
    real_engine = engine.SeirmEngine
    distributed.SeirmEngine = lambda graph, spec, **kw: real_engine(graph, spec, warps_per_cta=2, **kw)
    eager_episode = real_engine.run_episode                                      # no CUDA graphs without CUDA
    real_engine.run_episode = lambda self, *args, cuda_graph=False, **kw: eager_episode(self, *args, cuda_graph=False, **kw)
    real_plan = Runner._fused_plan

    def plan_as_if_on_cuda(self):                            # the planner insists on a CUDA device
        real_device = self.initializer.device
        self.initializer.device = types.SimpleNamespace(type="cuda")
        try:
            return real_plan(self)
        finally:
            self.initializer.device = real_device

    Runner._fused_plan = plan_as_if_on_cuda
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    symm = _SharedSymmetricMemory(tag, rank, world)
    sys.modules["torch.distributed._symmetric_memory"] = symm
    torch.distributed._symmetric_memory = symm
    try:
        g = load_golden("synth3k_soft")
        runner = envs.create(model=covid, population=pop_dir, distributed=True, world_size=world,
                             distributed_config=None if partition is None else {"partition": partition})
        assert (runner.engine.code_hdl is not None) == (world > 1)               # the symmetric-memory exchange is in use
        r2 = torch.tensor(g["r2"], dtype=torch.float32, requires_grad=True)
        m = torch.tensor(g["m"], dtype=torch.float32, requires_grad=True)
        runner._set_parameters({"initializer.transition_function.0.new_transmission.learnable_args.R2": r2,
                                "initializer.transition_function.1.seirm_progression.learnable_args.M": m})
        runner.tr.noise_seed, runner.tr.advance_seed = g["seed"], False
        runner.step(g["num_steps"])
        env = runner.state_trajectory[-1][-1]["environment"]
        loss = (torch.from_numpy(g["gI"]) * env["daily_infected"].reshape(-1)).sum() + \
               (torch.from_numpy(g["gD"]) * env["daily_deaths"].reshape(-1)).sum()
        loss.backward()
        final = runner.gather_final_state()
        assert runner.agent_partition_map["local_size"] == runner.state["agents"]["citizens"]["disease_stage"].shape[0]
        torch.save({"stage": final["disease_stage"], "nst": final["next_stage_time"], "di": env["daily_infected"].detach(),
                    "dd": env["daily_deaths"].detach(), "gr2": r2.grad, "gm": m.grad,
                    "order": getattr(runner, "_order", None), "report": getattr(runner, "partition_report", None)},
                   os.path.join(out_dir, f"dr{rank}.pt" if partition is None else f"dr_w{world}_{rank}.pt"))
    finally:
        dist.barrier()
        dist.destroy_process_group()
        symm.cleanup()


def test_distributed_runner_api_on_two_emulated_ranks(emulated_library, tmp_path):
    import torch.multiprocessing as mp

    from agenttorch_b200.populations import write_population_folder
    from tests.helpers import load_golden, rel_err

    g = load_golden("synth3k_soft")
    pop = write_population_folder(str(tmp_path / "pop"), g["age"], g["stage0"], g["edges"])
    mp.spawn(_distributed_runner, args=(2, _free_port(), str(tmp_path), emulated_library, pop, f"{os.getpid()}_dr"), nprocs=2, join=True)
    for r in range(2):
        o = torch.load(os.path.join(str(tmp_path), f"dr{r}.pt"))
        assert np.array_equal(o["stage"].numpy().reshape(-1).astype(np.float64), g["ref_stage"])
        assert np.array_equal(o["nst"].numpy().reshape(-1).astype(np.int64), g["ref_next_stage_time"].astype(np.int64))
        assert tuple(o["dd"].shape) == (1, g["num_steps"])
        assert rel_err(g["ref_daily_infected"], o["di"].numpy()) <= 1e-5 and rel_err(g["ref_daily_deaths"], o["dd"].numpy()) <= 1e-5
        assert rel_err(g["ref_grad_R2"], o["gr2"].numpy()) <= 1e-4 and rel_err(g["ref_grad_M"], o["gm"].numpy()) <= 1e-4


def test_distributed_runner_partition_option_gives_the_same_population_at_every_world_size(emulated_library, tmp_path):
    """envs.create(distributed=True, distributed_config={"partition": {"parts": 2}}): the population is re-ordered by the contact-graph
    partitioner before the ranks take their ranges; one rank and two ranks then simulate the same (relabelled) population, and
    gather_final_state() returns the agents in the population's original order"""
    import torch.multiprocessing as mp

    from agenttorch_b200.populations import write_population_folder
    from tests.helpers import load_golden

    g = load_golden("synth3k_soft")
    pop = write_population_folder(str(tmp_path / "pop"), g["age"], g["stage0"], g["edges"])
    part = {"parts": 2, "rounds": 4}
    for world in (1, 2):
        mp.spawn(_distributed_runner, args=(world, _free_port(), str(tmp_path), emulated_library, pop, f"{os.getpid()}_pt{world}", part),
                 nprocs=world, join=True)
    one = torch.load(os.path.join(str(tmp_path), "dr_w1_0.pt"))
    n = one["stage"].shape[0]
    assert one["order"] is not None and not torch.equal(one["order"], torch.arange(n))     # a real relabelling happened
    assert torch.equal(torch.sort(one["order"]).values, torch.arange(n)) and 0.0 <= one["report"]["edge_cut"] <= 1.0
    assert float(one["di"].sum()) > 10                                                       # an epidemic happened
    for r in range(2):
        o = torch.load(os.path.join(str(tmp_path), f"dr_w2_{r}.pt"))
        assert torch.equal(o["order"], one["order"])
        assert torch.equal(o["stage"], one["stage"]) and torch.equal(o["nst"], one["nst"])  # original agent order, bit for bit
        np.testing.assert_allclose(o["di"].numpy(), one["di"].numpy(), rtol=1e-6)
        np.testing.assert_allclose(o["dd"].numpy(), one["dd"].numpy(), rtol=1e-6)
        np.testing.assert_allclose(o["gr2"].numpy(), one["gr2"].numpy(), rtol=1e-5)
        np.testing.assert_allclose(o["gm"].numpy(), one["gm"].numpy(), rtol=1e-5)


# ------------------------------------------------------------------------------------------------------------------
# A config WITH the quarantine / testing substeps on two emulated ranks: their per-agent state is sharded like every agent
# tensor, the draws are addressed by the global agent id, one in-place launch per block after every step on every rank.
def _distributed_interventions(rank, world, port, out_dir, seirm_lib, iv_lib, pop_dir, tag):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from agenttorch_b200 import _native, _native_iv, distributed, engine
    from agenttorch_b200 import interventions as iv_ops
    from agenttorch_b200.core.dataloader import DataLoader, LoadPopulation
    from agenttorch_b200.core.runner import Runner
    from agenttorch_b200.models import covid
    from agenttorch_b200.models.covid import context
    from agenttorch_b200.models.covid.interventions import add_intervention_substeps
    from tests.helpers import load_interventions_golden

    _native.LIB_PATH, _native._lib = seirm_lib, None
    _native_iv.LIB_PATH, _native_iv._lib = iv_lib, None
    no_check = lambda *args: None                                                # noqa: E731
    engine._require_cuda = context._require_cuda = distributed._require_cuda = iv_ops._require_cuda = no_check
    iv_ops._stream = lambda t: None
    torch.cuda.device = lambda device: contextlib.nullcontext()
    torch.cuda.current_stream = lambda device=None: types.SimpleNamespace(cuda_stream=0)
    torch.cuda.synchronize = lambda device=None: None
    torch.set_num_threads(1)
    real_engine = engine.SeirmEngine
    distributed.SeirmEngine = lambda graph, spec, **kw: real_engine(graph, spec, warps_per_cta=2, **kw)
    real_plan = Runner._fused_plan

    def plan_as_if_on_cuda(self):
        real_device = self.initializer.device
        self.initializer.device = types.SimpleNamespace(type="cuda")
        try:
            return real_plan(self)
        finally:
            self.initializer.device = real_device

    Runner._fused_plan = plan_as_if_on_cuda
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    symm = _SharedSymmetricMemory(tag, rank, world)
    sys.modules["torch.distributed._symmetric_memory"] = symm
    torch.distributed._symmetric_memory = symm
    try:
        g = load_interventions_golden("interventions_1500_softopen")
        ov = g["overrides"]
        loader = DataLoader(covid, LoadPopulation(pop_dir))
        for k, v in ov["meta"].items():
            loader.set_config_attribute(k, v)
        loader.set_config_attribute("noise_seed", g["seed"])
        for k, v in ov["initial"].items():
            loader.config["state"]["agents"]["citizens"]["properties"][k]["value"] = v
        for k, v in ov["env"].items():
            loader.config["state"]["environment"][k]["value"] = v
        runner = distributed.DistributedRunner(Runner(add_intervention_substeps(loader.get_config()), covid.get_registry()), world_size=world)
        runner.init()
        pf = runner.initializer.policy_function
        for mod in (runner.tr, pf["2"]["citizens"]["start_compliance"], pf["2"]["citizens"]["break_compliance"], pf["3"]["citizens"]["accept_test"]):
            mod.advance_seed = False
        assert pf["3"]["citizens"]["accept_test"].id_offset == runner.agent_partition_map["start_idx"]
        r2 = torch.tensor(g["r2"], dtype=torch.float32, requires_grad=True)
        m = torch.tensor(g["m"], dtype=torch.float32, requires_grad=True)
        runner._set_parameters({"initializer.transition_function.0.new_transmission.learnable_args.R2": r2,
                                "initializer.transition_function.1.seirm_progression.learnable_args.M": m})
        runner.step(g["num_steps"])
        env = runner.state_trajectory[-1][-1]["environment"]
        loss = (torch.from_numpy(g["gI"]) * env["daily_infected"].reshape(-1)).sum() + \
               (torch.from_numpy(g["gD"]) * env["daily_deaths"].reshape(-1)).sum()
        loss.backward()
        final = runner.gather_final_state()
        out = {k: v.detach() for k, v in final.items()}
        out.update(di=env["daily_infected"].detach(), dd=env["daily_deaths"].detach(), gr2=r2.grad, gm=m.grad)
        torch.save(out, os.path.join(out_dir, f"iv{rank}.pt"))
    finally:
        dist.barrier()
        dist.destroy_process_group()
        symm.cleanup()


def test_intervention_substeps_on_two_emulated_ranks(emulated_library, tmp_path):
    import torch.multiprocessing as mp

    from agenttorch_b200.populations import write_population_folder
    from tests.helpers import load_interventions_golden, rel_err

    iv_lib = str(tmp_path / "libat_interventions_emulated.so")
    res = subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                          os.path.join(ROOT, "tests", "host_emulation", "iv_emulation.cpp"), "-o", iv_lib], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    g = load_interventions_golden("interventions_1500_softopen")
    pop = write_population_folder(str(tmp_path / "pop"), g["age"], g["stage0"], g["edges"])
    mp.spawn(_distributed_interventions, args=(2, _free_port(), str(tmp_path), emulated_library, iv_lib, pop, f"{os.getpid()}_iv"),
             nprocs=2, join=True)
    T = g["num_steps"]
    for r in range(2):
        o = torch.load(os.path.join(str(tmp_path), f"iv{r}.pt"))
        for k in ("disease_stage", "next_stage_time", "is_quarantined", "quarantine_start_date", "awaiting_test_result",
                  "test_result_date", "test_re_eligble_date"):
            assert np.array_equal(o[k].numpy().reshape(-1).astype(np.float64), g["ref_" + k + "_history"][T - 1].astype(np.float64)), k
        assert rel_err(g["ref_daily_infected"], o["di"].numpy()) <= 1e-5 and rel_err(g["ref_daily_deaths"], o["dd"].numpy()) <= 1e-5
        assert rel_err(g["ref_grad_R2"], o["gr2"].numpy()) <= 1e-4 and rel_err(g["ref_grad_M"], o["gm"].numpy()) <= 1e-4


def test_a_rank_without_agents_is_refused_on_every_rank():
    """shard_bounds aligns the ownership ranges to 1024 agents: 2048 agents on 4 ranks would leave two ranks empty, and the
    others waiting for them in every rendezvous.  The engine refuses up front, from numbers every rank sees."""
    import torch.distributed as dist

    from agenttorch_b200 import engine as eng_mod
    from agenttorch_b200.engine import SeirmEngine, SeirmSpec
    from agenttorch_b200.graph import shard_bounds, shard_code_len, synthetic_contact_graph

    port = _free_port()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        class FourRanks:                                           # stands in for a 4-rank group: only its size is consulted
            pass

        real = dist.get_world_size
        dist.get_world_size = lambda group=None: 4 if isinstance(group, FourRanks) else real(group)
        real_check = eng_mod._require_cuda
        eng_mod._require_cuda = lambda message, is_cuda: None
        try:
            assert shard_bounds(2048, 4)[2] == (2048, 2048)
            graph = synthetic_contact_graph(2048, 4.0, 1, device="cpu", row_begin=0, row_end=1024)
            with pytest.raises(ValueError, match="without agents"):
                SeirmEngine(graph, SeirmSpec(), process_group=FourRanks(), code_len=shard_code_len(2048, 4))
        finally:
            dist.get_world_size = real
            eng_mod._require_cuda = real_check
    finally:
        dist.destroy_process_group()
