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
N, DEG, T, SEED = 4096, 6.0, 21, 5


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
        age, stage0 = synthetic_population(N, SEED, initial_exposed=0.03, device="cpu")
        full = synthetic_contact_graph(N, DEG, SEED + 1, device="cpu", mean_interactions=mean_interactions_from_age(age))
        graph = synthetic_contact_graph(N, DEG, SEED + 1, device="cpu", row_begin=r0, row_end=r1)
        graph.i_bar, graph.i_bar_agent = full.i_bar, full.i_bar_agent
        eng = SeirmEngine(graph, SeirmSpec(num_steps=T), process_group=pg, code_len=shard_code_len(N, world), warps_per_cta=2,
                          symmetric=(exchange != "allgather") if world > 1 else None, sparse_exchange=exchange == "sparse",
                          hashed=True if exchange == "sparse_hashed" else None, hash_words=32 if exchange == "sparse_hashed" else None)
        if world > 1:
            assert (eng.code_hdl is not None) == (exchange != "allgather")
        st = SeirmState.from_initial(stage0[r0:r1], age[r0:r1], T, torch.device("cpu"))
        r_steps, m = torch.full((T,), 9.0), torch.tensor([0.12])
        sv = eng.run_episode(st, r_steps, m, T, NoiseSpec(seed=9, isolation="hard"), want_grad=True, cuda_graph=False)
        di, dd = st.daily_infected.clone(), st.daily_deaths.clone()
        if world > 1:
            dist.all_reduce(di); dist.all_reduce(dd)
        gr, gm = eng.episode_backward(sv, T, torch.linspace(0.5, 1.5, T), torch.linspace(2, 1, T), m)
        eng.check_errors()
        torch.save({"stage": st.stage, "nst": st.next_stage_time, "itime": st.infected_time, "di": di, "dd": dd, "gr": gr, "gm": gm},
                   os.path.join(out_dir, f"{exchange}_w{world}_r{rank}.pt"))
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


@pytest.mark.parametrize("exchange,world", [("allgather", 2), ("sparse", 2), ("dense", 2), ("sparse", 4)])
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
