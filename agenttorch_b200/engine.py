"""SeirmEngine — host driver of the sm_100a SEIRM step kernel (libat_seirm.so).

Owns the device-resident working set of one shard (one process per GPU):
  * the destination-sorted CSR contact graph (graph.ContactGraph),
  * the packed SoA agent state (SeirmState: stage f32, timers i16, age u8),
  * the double-buffered infectiousness arrays (1-byte code + 1-bit "is infectious" bitmap per
    source agent) that one step writes for the next — the only per-step data another GPU needs,
  * the reduction workspace and saved-for-backward buffers,
and issues the launches on torch's current stream (so they can be captured into a CUDA graph).

Replaces the reference's per-substep ATen op chains (new_transmission/transition.py:131-252,
seirm_progression/transition.py:90-121) and its T-step autograd graph.
"""
import ctypes
from dataclasses import dataclass
from typing import Optional

import torch

from . import _native as nat


# ----------------------------------------------------------------------------------------- model
@dataclass
class SeirmSpec:
    """Constants of the SEIRM model (simulation_metadata + environment tables, config.yaml)."""
    sfs: tuple = (0.35, 0.69, 1.03, 1.03, 1.03, 1.03, 1.27, 1.52, 1.52)
    sfi: tuple = (0.0, 0.33, 0.72, 0.0, 0.0)
    lam_gamma: tuple = ()
    num_steps: int = 21
    exposed_to_infected_time: int = 5
    infected_to_recovered_time: int = 5
    infinity_time: Optional[int] = None      # default num_steps + 20 (seirm_progression/transition.py:28-30)

    def __post_init__(self):
        if self.infinity_time is None:
            self.infinity_time = self.num_steps + 20
        if not self.lam_gamma:
            self.lam_gamma = tuple(lam_gamma_integrals(5.5, 2.4, self.num_steps).tolist())
        if not 1 <= len(self.lam_gamma) <= nat.MAX_LAMGAMMA:
            raise ValueError(f"lam_gamma_integrals must have 1..{nat.MAX_LAMGAMMA} entries")
        if len(self.sfs) > nat.MAX_SFS or len(self.sfi) > nat.MAX_SFI:
            raise ValueError("SFSusceptibility / SFInfector table too long")
        # limits of the kernels, refused here rather than inside a launch (README "Limits"): the step index is stored in a byte,
        # the fused backward keeps one accumulator per step and lane
        if self.num_steps > 255:
            raise ValueError("episodes of more than 255 steps are not supported (recovered_step is a uint8)")

    def to_struct(self, i_bar, i_bar_layers=()):
        m = nat.SeirmModel()
        m.n_layers = len(i_bar_layers) if len(i_bar_layers) > 1 else 0
        for i, v in enumerate(i_bar_layers):
            m.i_bar_layers[i] = float(v)
        m.n_sfs, m.n_sfi, m.n_lam_gamma = len(self.sfs), len(self.sfi), len(self.lam_gamma)
        m.exposed_to_infected_time = int(self.exposed_to_infected_time)
        m.infected_to_recovered_time = int(self.infected_to_recovered_time)
        m.infinity_time = int(self.infinity_time)
        m.i_bar = float(i_bar)
        for i, v in enumerate(self.sfs):
            m.sfs[i] = float(v)
        for i, v in enumerate(self.sfi):
            m.sfi[i] = float(v)
        for i, v in enumerate(self.lam_gamma):
            m.lam_gamma[i] = float(v)
        return m


def lam_gamma_integrals(scale, rate, t):
    """Daily gamma-CDF differences, float32 (initialiser `get_lam_gamma_integrals`,
    models/covid/substeps/utils.py:14-27; a = scale/b, b = rate^2/scale evaluated in float32)."""
    from scipy.stats import gamma

    scale_t = torch.tensor([float(scale)], dtype=torch.float32)
    rate_t = torch.tensor([float(rate)], dtype=torch.float32)
    b = rate_t * rate_t / scale_t
    a = scale_t / b
    a64, b64 = float(a), float(b)
    vals = [gamma.cdf(i, a=a64, loc=0, scale=b64) - gamma.cdf(i - 1, a=a64, loc=0, scale=b64) for i in range(int(t))]
    return torch.tensor(vals, dtype=torch.float64).float()


# ----------------------------------------------------------------------------------------- state
@dataclass
class SeirmState:
    """Packed SoA agent state of one shard + the (T,) daily series."""
    stage: torch.Tensor             # (n,) float32 — integers, or k + (1-W) under soft isolation (Q4)
    next_stage_time: torch.Tensor   # (n,) int16 (or float32 in the reference layout)
    infected_time: torch.Tensor     # (n,) int16 (or float32)
    age: torch.Tensor               # (n,) uint8 (or float32)
    daily_infected: torch.Tensor    # (T,) float32
    daily_deaths: torch.Tensor      # (T,) float32

    @staticmethod
    def from_initial(stage0, age, num_steps, device, packed=True):
        """Initial timers as the reference's initialisers (utils.py:30-57): infected_time 500 / -1 / -3,
        next_stage_time 500 / 3 / 5 for stage 0 / 1 / 2."""
        stage = torch.as_tensor(stage0).reshape(-1).to(device=device, dtype=torch.float32).clone().contiguous()
        tdt = torch.int16 if packed else torch.float32
        itime = torch.full_like(stage, 500, dtype=tdt)
        nst = torch.full_like(stage, 500, dtype=tdt)
        itime[stage == 1] = -1
        itime[stage == 2] = -3
        nst[stage == 1] = 3
        nst[stage == 2] = 5
        age_t = torch.as_tensor(age).reshape(-1).to(device=device)
        age_t = age_t.to(torch.uint8 if packed else torch.float32).contiguous()
        z = torch.zeros(num_steps, dtype=torch.float32, device=device)
        return SeirmState(stage, nst, itime, age_t, z, z.clone())

    def clone(self):
        return SeirmState(*(getattr(self, f).clone() for f in
                            ("stage", "next_stage_time", "infected_time", "age", "daily_infected", "daily_deaths")))


@dataclass
class NoiseSpec:
    """Where the step's randomness comes from (stands in for torch.rand / torch.bernoulli)."""
    seed: int = 0
    isolation: str = "soft"          # 'soft' | 'hard' | 'none' | 'tensor'
    p_isolate: float = 0.3
    uniform: str = "hash"            # 'hash' | 'tensor'
    isolation_tensor: Optional[torch.Tensor] = None   # (n,) or (T,n) float32
    uniform_tensor: Optional[torch.Tensor] = None

    def iso_mode(self):
        return {"tensor": nat.ISO_TENSOR, "soft": nat.ISO_HASH_SOFT, "hard": nat.ISO_HASH_HARD,
                "none": nat.ISO_NONE}[self.isolation]

    def u_mode(self):
        return {"tensor": nat.U_TENSOR, "hash": nat.U_HASH}[self.uniform]


_TIMER_TAG = {torch.float32: nat.TIMER_F32, torch.int16: nat.TIMER_I16}
_AGE_TAG = {torch.float32: nat.AGE_F32, torch.uint8: nat.AGE_U8}


def _require_cuda(message, is_cuda):
    """No CPU path.  (One function so that the host-emulation tests, which drive this engine against the kernel source
    compiled for the CPU, can lift the check.)"""
    if not is_cuda:
        raise RuntimeError(message)


def _ptr(t, dtype=None, name="tensor"):
    if t is None:
        return None
    _require_cuda(f"{name} must be a CUDA tensor (agenttorch_b200 has no CPU path)", t.is_cuda)
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must have dtype {dtype}, got {t.dtype}")
    return t.data_ptr()


class SeirmEngine:
    def __init__(self, graph, spec: SeirmSpec, process_group=None, grid=0, code_len=None, block_size=None,
                 tasks_per_group=0, edges_per_chunk=None, warps_per_cta=0, hashed=None, hash_words=None, hash_min_warps=20, symmetric=None, sparse_exchange=True,
                 layout="auto", tile_rows=None, bank_aware=True, overlap_exchange=None):
        """graph: ContactGraph (plain CSR, re-blocked here) or a prebuilt BlockedGraph.  block_size: sources
        per column block (default: the largest that fits the shared-memory bitmap, i.e. as few blocks as
        possible); code_len: length of the code/bitmap arrays (multi-GPU: world_size equal slices).
        layout: "stream" (flat per-source-block streams of self-describing entries against the block's exact bitmap,
        fixed-point lambda sums, gather + update kernels: any number of sources), "bcsr" (row_start + col; one exact block
        fused in one kernel, a hashed filter, or 256-row column blocks; the only layout with edge weights) or "auto"
        (stream unless the graph has edge weights or a bcsr-only option is given)."""
        from .graph import (BlockedGraph, block_graph, block_graph_stream, choose_block_size, choose_edges_per_chunk,
                            choose_tile_rows)

        _require_cuda("SeirmEngine needs the contact graph on a CUDA device; there is no CPU fallback", graph.device.type == "cuda")
        self.lib = nat.lib()
        self.spec = spec
        self.device = graph.device
        self.model = spec.to_struct(graph.i_bar, getattr(graph, "i_bar_layers", ()))
        if float(spec.sfi[0]) != 0.0:
            raise NotImplementedError("SFInfector[0] must be 0 (agents with a fractional stage are not tracked as infectious)")
        self.pg = process_group
        self.grid = grid
        self.tasks_per_group = tasks_per_group
        self.warps_per_cta = warps_per_cta
        n_src = graph.n_sources if code_len is None else int(code_len)
        if code_len is not None and (code_len < graph.n_sources or code_len % 1024):
            raise ValueError("code_len must cover all sources and be a multiple of 1024")
        if process_group is not None and code_len is not None:
            import torch.distributed as dist

            world = dist.get_world_size(process_group)
            if graph.n_sources <= (world - 1) * (int(code_len) // world):
                # every rank sees the same numbers, so every rank raises (no rank is left waiting in a rendezvous)
                raise ValueError(f"{graph.n_sources} agents leave a rank without agents: ownership ranges are aligned to 1024, "
                                 f"so {world} ranks need more than {(world - 1) * (int(code_len) // world)} agents")
        def capacity(epc, warps):
            with torch.cuda.device(self.device):
                return int(self.lib.at_seirm_max_block_size(len(spec.sfs), len(spec.lam_gamma), int(epc), int(warps)))

        def stream_capacity(warps):
            with torch.cuda.device(self.device):
                return int(self.lib.at_seirm_stream_max_block_size(len(spec.sfs), len(spec.lam_gamma), int(warps)))

        if layout not in ("auto", "bcsr", "stream"):
            raise ValueError("layout must be 'auto', 'bcsr' or 'stream'")
        if layout == "auto" and not isinstance(graph, BlockedGraph):
            bcsr_only = (hashed or hash_words or edges_per_chunk or graph.edge_weight is not None) and graph.edge_layer is None
            layout = "bcsr" if bcsr_only else "stream"
        # Multi-GPU: which 32-source groups (bitmap words, code sectors) do the edges of THIS rank's rows refer to?  The halo
        # exchange pulls only those (include/at_seirm.h, halo_mask); from the plain CSR, before it is re-blocked.
        self.halo_groups = None
        if process_group is not None and not isinstance(graph, BlockedGraph) and graph.num_edges > 0:
            self.halo_groups = torch.unique(graph.col_idx[: graph.num_edges].to(torch.int64) >> 5)
        if isinstance(graph, BlockedGraph):
            self.graph = graph
        elif layout == "stream":
            if block_size is None:
                # the streamed units pass through L1 (whatever shared memory leaves of the SM's 256 KB), which has to hold
                # the ~96 KB of loads in flight: measured on B200, bitmaps of <= 88 KB per block stream at 5.3 TB/s, 175 KB
                # ones at 4.3 TB/s
                _, block_size = choose_block_size(n_src, min(stream_capacity(warps_per_cta or 0), 88 * 1024 * 8))
            nb = -(-n_src // int(block_size))
            if tile_rows is None:
                tile_rows = choose_tile_rows(graph.n_rows, nb)
            self.graph = block_graph_stream(graph, int(block_size), int(tile_rows), bank_aware=bool(bank_aware))
        else:
            # one column block if the whole bitmap fits next to >= 12 warps of TMA buffers, else equal blocks
            # sized for 16 warps
            # Default layout: ONE column block.  If all sources fit the shared-memory bitmap next to >= 12 warps of
            # TMA buffers the bitmap is exact; otherwise it is a hashed filter (bit = src mod M) whose false positives
            # are rejected by the code lookup.  block_size=... forces exact column blocks (gather + update kernels).
            epc = edges_per_chunk or choose_edges_per_chunk(graph.num_edges / max(graph.n_rows, 1) * 32)
            full = ((n_src + 1023) // 1024) * 1024
            if block_size is not None:
                hash_words = 0
            if block_size is None:
                fits = full <= capacity(epc, 12)
                if hash_words:
                    hashed = True
                forced_words = int(hash_words or 0)
                hash_words = 0
                if hashed is True or (hashed is None and not fits):
                    hash_words = 1 << 10
                    while hash_words * 2 * 32 <= min(capacity(epc, hash_min_warps), full * 2 - 1):
                        hash_words *= 2
                    hash_words = forced_words or hash_words
                    block_size = full
                elif fits:
                    block_size = full
                else:
                    _, block_size = choose_block_size(n_src, capacity(epc, 16))
            self.graph = block_graph(graph, int(block_size), edges_per_chunk=edges_per_chunk, hash_words=hash_words)
        cap4 = stream_capacity(self.warps_per_cta or 0) if self.graph.layout == 1 else capacity(self.graph.edges_per_chunk, 4)
        if (self.graph.hash_words * 32 or self.graph.block_size) > cap4:
            raise RuntimeError(f"block_size {self.graph.block_size} exceeds the shared-memory bitmap capacity")
        g = self.graph
        self.n_code = ((n_src + 1023) // 1024) * 1024
        if g.layout != 1:
            # (the stream gather zero-fills what the last block's bitmap lacks; the older kernels read whole blocks)
            self.n_code = max(self.n_code, g.n_blocks * g.block_size)
        dev = self.device
        self.halo_mask = None
        if self.halo_groups is not None:
            n_mask = self.n_code // 1024
            acc = torch.zeros(n_mask, dtype=torch.int64, device=dev)
            grp = self.halo_groups.to(dev)
            acc.index_add_(0, grp >> 5, torch.ones_like(grp) << (grp & 31))            # groups are unique: the sum is the OR
            acc = torch.where(acc >= (1 << 31), acc - (1 << 32), acc)                      # as the int32 bit pattern
            self.halo_mask = acc.to(torch.int32)
            self.halo_fraction = float(grp.numel()) / max(self.n_code // 32, 1)            # share of the population's groups this rank pulls from
        # Multi-GPU: the code arrays live in symmetric memory when available, so that the per-step halo exchange is a
        # device-side barrier + peer copies over NVLink that can be captured into the episode's CUDA graph (NCCL
        # collectives could not).  symmetric=None: try, fall back to the NCCL all-gather.
        self.code_hdl = None
        self.sparse_exchange = bool(sparse_exchange)
        # Overlapped halo exchange: lam_gamma_integrals[0] == 0 (the gamma CDF difference at day 0, utils.py:14-27) means an agent
        # exposed at step t contributes nothing at t+1, so the infectiousness of step t+2 is known when step t ends — every
        # E/I agent's code follows from (stage, timers) and the progression it will make at t+1.  The step kernels then encode
        # TWO steps ahead (AT_DO_LOOKAHEAD) into a third code buffer and the exchange of step t+2 runs on a side stream while
        # step t+1 computes.  Checked here; anything else keeps the synchronous exchange.
        can_overlap = (process_group is not None and self.sparse_exchange and float(spec.lam_gamma[0]) == 0.0
                       and int(spec.exposed_to_infected_time) >= 1)
        import os

        if overlap_exchange is None and os.environ.get("AGENTTORCH_B200_OVERLAP_EXCHANGE") in ("0", "1"):
            overlap_exchange = os.environ["AGENTTORCH_B200_OVERLAP_EXCHANGE"] == "1" and can_overlap     # developer switch
        if overlap_exchange and not can_overlap:
            raise ValueError("overlap_exchange needs several ranks, the sparse exchange and lam_gamma_integrals[0] == 0")
        # Measured on 2 and 4 B200s (profiles/r02_multigpu.md) the overlapped form is 4-6 % SLOWER than the synchronous exchange:
        # the exchange is not what bounds a step there (fixed launch / prologue costs are), and the lean kernels plus the second
        # encode launch cost more than the hidden rendezvous saves.  So it is opt-in.
        self.overlap = False if overlap_exchange is None else bool(overlap_exchange)
        n_buf = 3 if self.overlap else 2
        if process_group is not None and symmetric is not False:
            try:
                import torch.distributed._symmetric_memory as symm

                # one symmetric buffer per step parity: [code bytes][hot bitmap][arrival flags] (include/at_seirm.h,
                # at_seirm_exchange_sparse); the step kernel writes this rank's slice of the first two parts
                nbytes = int(self.lib.at_seirm_exchange_bytes(self.n_code))
                self.xbuf = [symm.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(n_buf)]
                self.code_hdl = [symm.rendezvous(c, process_group) for c in self.xbuf]
                for c in self.xbuf:
                    c.zero_()
                self.code = [c[: self.n_code] for c in self.xbuf]
                self.hot = [c[self.n_code: self.n_code + self.n_code // 8 + 16].view(torch.int32) for c in self.xbuf]
                self.xsync = [torch.zeros(32, dtype=torch.int32, device=dev) for _ in range(n_buf)]
                torch.cuda.synchronize(dev)
                import torch.distributed as dist

                dist.barrier(group=process_group)               # nobody signals into a buffer that is still being zeroed
            except Exception as exc:
                if symmetric is True:
                    raise
                import warnings

                warnings.warn(f"symmetric memory unavailable on this rank ({type(exc).__name__}: {exc}); NCCL all-gather exchange")
                self.code_hdl = None
            # every rank must take the same exchange path: a rank that fell back alone would wait in an all-gather while the
            # others spin in the rendezvous kernel
            import torch.distributed as dist

            ok = torch.tensor([1 if self.code_hdl is not None else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=process_group)
            if int(ok.item()) == 0:
                self.code_hdl = None
        if self.code_hdl is None:
            self.overlap = False                                  # the overlapped exchange is the symmetric-memory kernel
            self.code = [torch.zeros(self.n_code, dtype=torch.uint8, device=dev) for _ in range(2)]
            self.hot = [torch.zeros(self.n_code // 32 + 4, dtype=torch.int32, device=dev) for _ in range(2)]
        self.cur = 0
        # hashed mode on one GPU: the folded filter of each hot bitmap, produced by the launch that writes the bitmap
        self.filt = None
        if g.hash_words:
            self.filt = [torch.zeros(g.hash_words + 4, dtype=torch.int32, device=dev) for _ in range(3 if self.overlap else 2)]   # one per code buffer
        nt = g.n_tasks
        self.sus = torch.zeros(nt, dtype=torch.int32, device=dev)
        self.exp = torch.zeros(nt, dtype=torch.int32, device=dev)
        self.inf = torch.zeros(nt, dtype=torch.int32, device=dev)
        self.lam_part = self.dirty = self.rec = self.tile_count = None
        if g.layout == 1:
            # stream layout: per-tile lists of hit records (capacity = the tile's edges) and their fill counters, which are all
            # zero between steps (the update kernel clears what the gather kernel counted)
            self.rec = torch.empty(max(g.num_edges, 1), dtype=torch.int32, device=dev)
            self.tile_count = torch.zeros(g.n_rows_pad // g.tile_rows, dtype=torch.int32, device=dev)
        elif g.n_blocks > 1:
            # per-(column block, row) partial lambdas; only the rows flagged in `dirty` are ever written or read
            self.lam_part = torch.empty((g.n_blocks, g.n_rows_pad, 2), dtype=torch.float64, device=dev)
            self.dirty = torch.zeros((g.n_blocks, nt), dtype=torch.int32, device=dev)
        self.workspace = torch.zeros(self.lib.at_seirm_workspace_bytes() // 8 + 8, dtype=torch.float64, device=dev)
        self.error_flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self.seed_buf = torch.zeros(1, dtype=torch.int64, device=dev)
        self._seed_value = 0                 # what seed_buf holds (set_seed skips the fill when an episode repeats a seed)
        self.launch_count = 0
        self._graphs = {}
        self._saved = {}
        self._plans = {}
        self._recording = None
        # optional callable(t) issued on the same stream right after the launches of step t (e.g. the quarantine / testing
        # kernels of a covid config, which read the stage array the step just updated); episodes with a hook are not captured
        # into CUDA graphs (the hook's buffers are not part of the capture key)
        self.after_step = None
        self._side = None                    # side stream of the overlapped halo exchange
        # "the epidemic is over" (no agent of any rank in E or I): device word maintained by the launches of an episode (one GPU)
        # or by the halo exchange (several); the gather launches and, after one more round over the code buffers, the exchanges
        # then return at once.  Only consulted inside run_episode(), which zeroes it first: between direct launches the caller may
        # have edited the state.
        self.over = torch.zeros(4, dtype=torch.int32, device=dev)
        self._use_over = False

    # ------------------------------------------------------------------------------ raw launch
    def launch(self, state: SeirmState, *, flags, t, t_encode=None, code_in=None, code_out=None,
               noise: Optional[NoiseSpec] = None, isolation=None, uniform=None, r_value=None, r_index=0,
               m_value=None, dne_dr=None, sum_dne_dr=None, death_count=None, recovered_step=None,
               seed_ptr=None, check_range=True, xchg_slot=0, reserve_sms=0, dne_e_dr=None, step_skipped=None):
        g = self.graph
        a = nat.SeirmLaunch()
        a.abi_version = nat.ABI_VERSION
        a.flags = int(flags)
        a.t = int(t)
        a.t_encode = int(t + 1 if t_encode is None else t_encode)
        a.n_rows, a.n_rows_pad, a.n_sources, a.row_offset = g.n_rows, g.n_rows_pad, self.n_code, g.row_offset
        a.n_blocks, a.block_size = g.n_blocks, g.block_size
        a.rows_per_task, a.edges_per_chunk = g.rows_per_task, g.edges_per_chunk
        a.row_start = _ptr(g.row_start, torch.int16, "row_start")
        a.seg_info = _ptr(g.seg_info, torch.int32, "seg_info")
        a.col = _ptr(g.col, torch.int32, "col")
        a.edge_weight = _ptr(g.edge_weight, torch.float32, "edge_weight")
        if state.stage.shape[0] != g.n_rows:
            raise RuntimeError("state size does not match the graph shard")
        a.stage = _ptr(state.stage, torch.float32, "stage")
        if state.next_stage_time.dtype not in _TIMER_TAG or state.infected_time.dtype != state.next_stage_time.dtype:
            raise RuntimeError("timers must both be float32 or both int16")
        a.next_stage_time = _ptr(state.next_stage_time, None, "next_stage_time")
        a.infected_time = _ptr(state.infected_time, None, "infected_time")
        a.timer_dtype = _TIMER_TAG[state.next_stage_time.dtype]
        if state.age.dtype not in _AGE_TAG:
            raise RuntimeError("age must be float32 or uint8")
        a.age = _ptr(state.age, None, "age")
        a.age_dtype = _AGE_TAG[state.age.dtype]
        a.sus_bitmap, a.exp_bitmap, a.inf_bitmap = _ptr(self.sus), _ptr(self.exp), _ptr(self.inf)
        ci = self.cur if code_in is None else code_in
        co = 1 - ci if code_out is None else code_out
        a.code_in = _ptr(self.code[ci], torch.uint8)
        a.hot_bitmap_in = _ptr(self.hot[ci], torch.int32)
        a.code_out = _ptr(self.code[co], torch.uint8)
        a.hot_bitmap_out = _ptr(self.hot[co], torch.int32)
        if self.filt is not None:
            a.hot_filter_in = _ptr(self.filt[ci], torch.int32)
            # multi-GPU: the filter is rebuilt from the all-gathered codes (exchange), not from this rank's slice
            a.hot_filter_out = _ptr(self.filt[co], torch.int32) if self.pg is None else None
        noise = noise or NoiseSpec()
        a.iso_mode, a.u_mode = noise.iso_mode(), noise.u_mode()
        a.isolation = _ptr(isolation, torch.float32, "isolation")
        a.uniform = _ptr(uniform, torch.float32, "uniform")
        a.seed = int(noise.seed) & ((1 << 64) - 1)
        a.seed_ptr = _ptr(seed_ptr, torch.int64, "seed_ptr")
        a.p_isolate = float(noise.p_isolate)
        a.r_value = _ptr(r_value, torch.float32, "r_value")
        a.r_index = int(r_index)
        a.m_value = _ptr(m_value, torch.float32, "m_value")
        a.daily_infected = _ptr(state.daily_infected, torch.float32, "daily_infected")
        a.daily_deaths = _ptr(state.daily_deaths, torch.float32, "daily_deaths")
        a.dne_dr = _ptr(dne_dr, torch.float32, "dne_dr")
        a.dne_e_dr = _ptr(dne_e_dr, torch.float32, "dne_e_dr")
        a.sum_dne_dr = _ptr(sum_dne_dr, torch.float32, "sum_dne_dr")
        a.death_count = _ptr(death_count, torch.float32, "death_count")
        a.recovered_step = _ptr(recovered_step, torch.uint8, "recovered_step")
        a.lam_part, a.dirty = _ptr(self.lam_part), _ptr(self.dirty)
        a.layout, a.tile_rows = int(g.layout), int(g.tile_rows)
        if g.layout == 1:
            a.seg_info = None
            a.unit_tile = _ptr(g.seg_info, torch.int32, "unit_tile")
            a.block_units = _ptr(g.block_units, torch.int32, "block_units")
            a.tile_rec_base = _ptr(g.tile_rec_base, torch.int64, "tile_rec_base")
            a.tile_count = _ptr(self.tile_count, torch.int32, "tile_count")
            a.rec = _ptr(self.rec, torch.int32, "rec")
            a.lam_deg_bits = int(g.lam_deg_bits)
        a.workspace = _ptr(self.workspace, torch.float64)
        # check_range=False: the launch encodes the infectiousness of a step that is never simulated (the last step of an
        # episode), so an infection older than lam_gamma_integrals there is not an error — the reference never looks it up
        a.error_flag = _ptr(self.error_flag, torch.int32) if check_range else None
        a.grid = int(self.grid)
        a.tasks_per_group = int(self.tasks_per_group)
        a.warps_per_cta = int(self.warps_per_cta)
        a.hash_words = int(g.hash_words)
        a.xchg_slot = int(xchg_slot)
        a.reserve_sms = int(reserve_sms)
        if self._use_over:
            a.over_state = _ptr(self.over, torch.int32)
            a.over_local = 1 if self.pg is None else 0
            a.step_skipped = _ptr(step_skipped, torch.int32, "step_skipped")
        rec = getattr(self, "_recording", None)
        if rec is not None:
            rec.append((a, self._kernels_of(flags)))
        self._submit(a, self._kernels_of(flags))

    def _submit(self, a, n_kernels):
        stream = torch.cuda.current_stream(self.device).cuda_stream
        nat.check(self.lib.at_seirm_step(ctypes.byref(self.model), ctypes.byref(a), ctypes.c_void_p(stream)))
        self.launch_count += n_kernels

    def _kernels_of(self, flags):
        """CUDA kernels one at_seirm_step call launches (gather + update in blocked mode, + the filter fold)"""
        n = 2 if ((self.graph.n_blocks > 1 or self.graph.layout == 1) and (flags & nat.DO_TRANSMISSION)) else 1
        if self.filt is not None and self.pg is None and (flags & nat.DO_ENCODE):
            n += 1
        return n

    def kernels_per_episode(self, num_steps, backward=True):
        full = nat.DO_TRANSMISSION | nat.DO_PROGRESSION | nat.DO_ENCODE
        n = self._kernels_of(nat.DO_ENCODE | nat.DO_INIT) + num_steps * self._kernels_of(full) + (1 if backward else 0)
        if self.pg is not None:
            n += num_steps + 1               # one pull / rebuild kernel per halo exchange
            if self.overlap:
                n += 1                       # the second encode launch (two steps are encoded before step 0)
        return n

    # ------------------------------------------------------------------------ halo exchange
    # the overlapped exchange runs beside the step kernels of the same rank, which are launched in their register-lean form
    # (at_seirm_launch.reserve_sms): 8 K registers and 256 threads of every SM stay free, one exchange CTA per SM fits
    OVERLAP_SMS = 148

    def exchange(self, buf_index, slot=0, max_ctas=0):
        """Make every rank's freshly written slice of code/bitmap visible to all ranks (the per-step
        halo exchange of boundary-agent infectiousness).  No-op on one GPU.  slot: the workspace word (at_seirm_launch.xchg_slot)
        in which the launch that wrote the slice noted whether it holds an infectious agent."""
        if self.pg is None:
            return
        from .parallel import halo_exchange

        # the code bytes are the whole halo; the hot bitmap / hashed filter are rebuilt locally from them
        stream = torch.cuda.current_stream(self.device).cuda_stream
        filt = self.filt[buf_index] if self.filt is not None else None
        if self.code_hdl is not None:
            import torch.distributed as dist

            hdl = self.code_hdl[buf_index]
            world, rank = dist.get_world_size(self.pg), dist.get_rank(self.pg)
            ptrs = (ctypes.c_void_p * world)(*[int(p) for p in hdl.buffer_ptrs])
            if self.sparse_exchange:
                # one kernel: rendezvous (arrival flags in the peers' buffers) + bitmap-directed pull over NVLink
                # (P2P loads of the non-empty 32-byte code sectors only) + filter fold
                nat.check(self.lib.at_seirm_exchange_sparse(_ptr(self.xbuf[buf_index]), int(self.n_code), _ptr(filt),
                                                            int(self.graph.hash_words), ptrs, world, rank,
                                                            _ptr(self.xsync[buf_index]), _ptr(self.workspace) + 4 * int(slot),
                                                            _ptr(self.error_flag), int(max_ctas),
                                                            _ptr(self.over, torch.int32) if self._use_over else None,
                                                            len(self.xbuf), _ptr(self.halo_mask, torch.int32),
                                                            ctypes.c_void_p(stream)))
            else:
                hdl.barrier(channel=buf_index)                   # every rank has written its slice of this buffer
                # dense: pull the peers' whole slices (P2P loads) + rebuild bitmap/filter from the codes, one kernel
                nat.check(self.lib.at_seirm_pull_codes(_ptr(self.code[buf_index]), int(self.n_code), _ptr(self.hot[buf_index]),
                                                       _ptr(filt), int(self.graph.hash_words), ptrs, world, rank,
                                                       ctypes.c_void_p(stream)))
        else:
            halo_exchange(self.code[buf_index], None, self.n_code, self.pg)
            nat.check(self.lib.at_seirm_hot_from_code(_ptr(self.code[buf_index]), int(self.n_code), _ptr(self.hot[buf_index]),
                                                      _ptr(filt), int(self.graph.hash_words), ctypes.c_void_p(stream)))
        self.launch_count += 1

    # ----------------------------------------------------------------------------- substeps
    def set_seed(self, seed):
        """The noise seed the step kernels read from device memory (so that a captured episode serves any seed)."""
        v = _as_i64(seed)
        if v != self._seed_value:
            self.seed_buf.fill_(v)
            self._seed_value = v

    def encode(self, state, t):
        """(Re)derive the row bitmaps from `state.stage` and write the infectiousness code/bitmap of step `t`
        into the live buffer: the first launch on a state, or after the caller edited it."""
        self.launch(state, flags=nat.DO_ENCODE | nat.DO_INIT, t=0, t_encode=t, code_in=1 - self.cur, code_out=self.cur)
        self.exchange(self.cur)

    def step(self, state, t, noise, r_value, r_index, m_value, *, isolation=None, uniform=None, flags=None,
             dne_dr=None, sum_dne_dr=None, death_count=None, recovered_step=None, seed_ptr=None, check_range=True, dne_e_dr=None,
             step_skipped=None):
        """One fused step: transmission + progression + encode of step t+1; flips the live buffer."""
        flags = (nat.DO_TRANSMISSION | nat.DO_PROGRESSION | nat.DO_ENCODE) if flags is None else flags
        self.launch(state, flags=flags, t=t, t_encode=t + 1, noise=noise, isolation=isolation, uniform=uniform,
                    r_value=r_value, r_index=r_index, m_value=m_value, dne_dr=dne_dr, sum_dne_dr=sum_dne_dr,
                    death_count=death_count, recovered_step=recovered_step, seed_ptr=seed_ptr, check_range=check_range,
                    dne_e_dr=dne_e_dr, step_skipped=step_skipped)
        if flags & nat.DO_ENCODE:
            self.cur = 1 - self.cur
            self.exchange(self.cur)

    # ------------------------------------------------------------------------------ episode
    # gradients with respect to the SFSusceptibility / SFInfector tables (config.yaml declares them learnable-capable): the forward
    # additionally saves, per step and agent, the part of d ne / dR that came from sources in stage E
    want_table_grad = False

    def saved_buffers(self, num_steps, want_grad):
        key = (num_steps, want_grad)
        if key not in self._saved:
            n, dev = self.graph.n_rows, self.device
            ld = ((n + 31) // 32) * 32
            self._saved[key] = {
                "ld": ld,
                "dne_dr": torch.zeros((num_steps, ld), dtype=torch.float32, device=dev) if want_grad else None,
                "dne_e": None,
                # [t] = 1: the forward found the epidemic over at step t and left dne_dr[t] / dne_e[t] untouched (they mean zero)
                "step_skipped": torch.zeros(num_steps, dtype=torch.int32, device=dev),
                "recovered_step": torch.full((n,), 255, dtype=torch.uint8, device=dev),
                "sum_dne_dr": torch.zeros(num_steps, dtype=torch.float32, device=dev),
                "death_count": torch.zeros(num_steps, dtype=torch.float32, device=dev),
                "r_steps": torch.zeros(num_steps, dtype=torch.float32, device=dev),
                "m_value": torch.zeros(1, dtype=torch.float32, device=dev),
            }
        return self._saved[key]

    def _issue_episode(self, state, num_steps, noise, sv, want_grad, t0=0, zero_series=None):
        """Issue the 1 + num_steps launches of an episode.  The filled launch structs are cached per (buffers, modes),
        so that re-issuing an episode costs one ctypes call per step (matters when the episode cannot be replayed from
        a CUDA graph, i.e. with the NCCL halo exchange between the steps)."""
        key = ("plan", num_steps, want_grad, t0, self.cur, noise.isolation, noise.uniform, float(noise.p_isolate),
               state.stage.data_ptr(), state.next_stage_time.data_ptr(), state.infected_time.data_ptr(),
               state.age.data_ptr(), state.daily_infected.data_ptr(), state.daily_deaths.data_ptr(),
               None if noise.isolation_tensor is None else noise.isolation_tensor.data_ptr(),
               None if noise.uniform_tensor is None else noise.uniform_tensor.data_ptr())
        sv["recovered_step"].fill_(255)
        self.over.zero_()
        sv["step_skipped"].zero_()
        if zero_series is not None:
            zero_series.zero_()
        # (a hook may change stages behind the step kernels' back; the overlapped exchange updates the word while a step runs)
        self._use_over = self.after_step is None and not self.overlap
        try:
            if self.overlap:
                self._issue_episode_overlapped(state, num_steps, noise, sv, want_grad, t0)
            else:
                self._issue_episode_sync(state, num_steps, noise, sv, want_grad, t0, key)
        finally:
            self._use_over = False
        if want_grad and sv.get("itime_valid"):
            sv["itime_copy"].copy_(state.infected_time)       # for the backward pass (see episode_backward_packed)

    def _issue_episode_sync(self, state, num_steps, noise, sv, want_grad, t0, key):
        plan = self._plans.get(key)
        if plan is not None:
            for k, (a, nk, flips) in enumerate(plan):
                self._submit(a, nk)
                if flips:
                    self.cur = 1 - self.cur
                self.exchange(self.cur)
                if flips and self.after_step is not None:
                    self.after_step(t0 + k - 1)              # entry 0 is the encode launch
            return
        plan = []
        self._recording = []
        try:
            self.encode(state, t0)
            plan.append((*self._recording[-1], False))
            for k in range(num_steps):
                t = t0 + k
                iso = noise.isolation_tensor[k] if noise.isolation == "tensor" else None
                uni = noise.uniform_tensor[k] if noise.uniform == "tensor" else None
                self.step(state, t, noise, sv["r_steps"], k, sv["m_value"], isolation=iso, uniform=uni,
                          dne_dr=sv["dne_dr"][k] if want_grad else None, sum_dne_dr=sv["sum_dne_dr"],
                          death_count=sv["death_count"], recovered_step=sv["recovered_step"], seed_ptr=self.seed_buf,
                          check_range=k + 1 < num_steps,
                          dne_e_dr=sv["dne_e"][k] if (want_grad and self.want_table_grad and sv["dne_e"] is not None) else None,
                          step_skipped=sv["step_skipped"])
                plan.append((*self._recording[-1], True))
                if self.after_step is not None:
                    self.after_step(t)
        finally:
            self._recording = None
        if len(self._plans) > 8:
            self._plans.clear()
        self._plans[key] = plan

    def _issue_episode_overlapped(self, state, num_steps, noise, sv, want_grad, t0=0):
        """Episode with the halo exchange taken off the critical path: step t encodes the infectiousness of step t+2
        (AT_DO_LOOKAHEAD) into code buffer (t+2) % 3, whose exchange runs on a side stream while step t+1 computes; step t+2
        waits for it.  Same results as the synchronous path (tests/test_gpu_multi.py, emulated multi-rank tests)."""
        streams = self.device.type == "cuda"                  # (the host-emulation tests run the same sequence in program order)
        main = side = None
        if streams:
            main = torch.cuda.current_stream(self.device)
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
            side = self._side
        full = nat.DO_TRANSMISSION | nat.DO_PROGRESSION | nat.DO_ENCODE | nat.DO_LOOKAHEAD
        # code(t0) from the initial state, exchanged on the main stream; code(t0+1) predicted, exchanged beside step t0
        self.launch(state, flags=nat.DO_ENCODE | nat.DO_INIT, t=0, t_encode=t0, code_in=2, code_out=0, xchg_slot=0)
        self.exchange(0, slot=0)
        self.launch(state, flags=nat.DO_ENCODE | nat.DO_LOOKAHEAD, t=t0, t_encode=t0 + 1, code_in=2, code_out=1, xchg_slot=1)
        done = {}

        def exchange_on_side(k):
            if not streams:
                self.exchange(k % 3, slot=k % 3, max_ctas=self.OVERLAP_SMS)
                return
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
            with torch.cuda.stream(side):
                self.exchange(k % 3, slot=k % 3, max_ctas=self.OVERLAP_SMS)
                done[k] = torch.cuda.Event()
                done[k].record(side)

        if num_steps > 1:
            exchange_on_side(1)
        for k in range(num_steps):
            t = t0 + k
            if k in done:
                main.wait_event(done.pop(k))
            iso = noise.isolation_tensor[k] if noise.isolation == "tensor" else None
            uni = noise.uniform_tensor[k] if noise.uniform == "tensor" else None
            needed = k + 2 < num_steps                               # somebody will gather from code(t+2)
            self.launch(state, flags=full, t=t, t_encode=t + 2, code_in=k % 3, code_out=(k + 2) % 3, noise=noise, isolation=iso,
                        uniform=uni, r_value=sv["r_steps"], r_index=k, m_value=sv["m_value"],
                        dne_dr=sv["dne_dr"][k] if want_grad else None, sum_dne_dr=sv["sum_dne_dr"], death_count=sv["death_count"],
                        recovered_step=sv["recovered_step"], seed_ptr=self.seed_buf, check_range=needed, xchg_slot=(k + 2) % 3,
                        reserve_sms=1,
                        dne_e_dr=sv["dne_e"][k] if (want_grad and self.want_table_grad and sv["dne_e"] is not None) else None)
            if needed:
                exchange_on_side(k + 2)
            if self.after_step is not None:
                self.after_step(t)
        if streams:
            main.wait_stream(side)
        self.cur = 0

    def run_episode(self, state, r_steps, m_value, num_steps, noise: NoiseSpec, want_grad=False, cuda_graph=False, zero_series=None):
        """Advance `state` in place through `num_steps` steps (t = 0..num_steps-1, as Runner.step does).

        r_steps: (num_steps,) float32 device tensor with the per-step R (transition.py:139-143);
        m_value: (1,) float32.  Returns the saved-for-backward buffers (views owned by the engine).
        With cuda_graph=True the launch sequence is captured once per (state buffers, shape) and
        replayed; the seed, R and M are read from device memory so the graph is reusable.
        zero_series: a tensor (normally the storage behind state.daily_infected / daily_deaths) to be zeroed as part of the
        episode's launch sequence, so that the series come out as this episode's increments without a launch of the caller's."""
        if want_grad and num_steps > 64:
            raise ValueError("gradients through episodes of more than 64 steps are not supported (the fused backward kernel keeps "
                             "one accumulator per step and lane); run forward-only or split the episode")
        sv = self.saved_buffers(num_steps, want_grad)
        if want_grad and self.want_table_grad and sv["dne_e"] is None:
            if self.graph.layout != 1:
                raise NotImplementedError("gradients with respect to SFInfector need the stream layout")
            sv["dne_e"] = torch.zeros_like(sv["dne_dr"])
        # (every tensor method here is a dispatcher call of several microseconds on the host: the common case - float32 (T,) and
        # (1,) device tensors - goes through unchanged)
        if want_grad:
            it = state.infected_time
            pool = sv.setdefault("itime_pool", {})
            buf = pool.get((it.dtype, it.numel()))
            if buf is None:                        # (kept for the engine's lifetime: captured episodes hold its address)
                buf = pool[(it.dtype, it.numel())] = torch.empty_like(it)
            sv["itime_copy"] = buf
            sv["itime_valid"] = self.after_step is None and it.dtype in (torch.float32, torch.int16)
        src_r = r_steps.detach() if r_steps.requires_grad else r_steps
        src_m = m_value.detach() if m_value.requires_grad else m_value
        if src_r.dim() != 1 or src_r.numel() != num_steps:
            src_r = src_r.reshape(-1)[:num_steps]
        if src_m.dim() != 1 or src_m.numel() != 1:
            src_m = src_m.reshape(-1)[:1]
        if src_r.dtype == torch.float32 and src_m.dtype == torch.float32 and src_r.device == src_m.device == sv["r_steps"].device:
            torch._foreach_copy_([sv["r_steps"], sv["m_value"]], [src_r, src_m])          # one launch
        else:
            sv["r_steps"].copy_(src_r)
            sv["m_value"].copy_(src_m)
        self.set_seed(noise.seed)
        hook = self.after_step
        if hook is not None and hasattr(hook, "prepare"):
            hook.prepare()                   # the episode's noise seeds of the hook's launches, into the device words they read
        capturable_hook = hook is None or hasattr(hook, "capture_key")
        if not cuda_graph or not capturable_hook or self.device.type != "cuda" or (self.pg is not None and self.code_hdl is None):
            # (NCCL collectives are not captured: capturing them hung on this stack; without symmetric memory a
            # multi-GPU episode is issued launch by launch)
            self._issue_episode(state, num_steps, noise, sv, want_grad, zero_series=zero_series)
            return sv
        key = (num_steps, want_grad, None if zero_series is None else zero_series.data_ptr(), noise.isolation, noise.uniform, float(noise.p_isolate),
               state.stage.data_ptr(), state.next_stage_time.data_ptr(), state.infected_time.data_ptr(),
               state.age.data_ptr(), state.daily_infected.data_ptr(), state.daily_deaths.data_ptr(),
               None if noise.isolation_tensor is None else noise.isolation_tensor.data_ptr(),
               None if noise.uniform_tensor is None else noise.uniform_tensor.data_ptr(),
               None if hook is None else hook.capture_key)
        entry = self._graphs.get(key)
        if entry is None:
            cur0 = self.cur
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(self.device)
            # capture issues no work; the replay below is the execution
            with torch.cuda.graph(g):
                self._issue_episode(state, num_steps, noise, sv, want_grad, zero_series=zero_series)
            entry = (g, cur0, self.cur)
            if len(self._graphs) >= 4:               # state tensors were re-created (reset_state): drop old captures
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = entry
            self.cur = cur0
        g, cur0, cur1 = entry
        if self.cur != cur0:                 # the captured sequence starts from buffer cur0
            self.cur = cur0
        g.replay()
        self.cur = cur1
        return sv

    def episode_backward(self, sv, num_steps, g_daily_infected, g_daily_deaths, m_value):
        """grad wrt the per-step R (num_steps,) and M (1,) — at_seirm_episode_backward."""
        packed = self.episode_backward_packed(sv, num_steps, g_daily_infected, g_daily_deaths, m_value)
        return packed[:num_steps], packed[num_steps:]

    def episode_backward_packed(self, sv, num_steps, g_daily_infected, g_daily_deaths, m_value):
        """(num_steps + 1,): dL/d(per-step R) followed by dL/dM."""
        dev = self.device
        packed = torch.empty(num_steps + 1, dtype=torch.float32, device=dev)      # [grad_r | grad_m]: one buffer, one all-reduce
        grad_r, grad_m = packed[:num_steps], packed[num_steps:]
        def plain(t, n):         # float32, contiguous, exactly n elements: the usual case costs no tensor operation
            if t.dtype == torch.float32 and t.numel() == n and t.is_contiguous() and not t.requires_grad:
                return t
            return t.detach().reshape(-1)[:n].to(torch.float32).contiguous()

        gi, gd, mv = plain(g_daily_infected, num_steps), plain(g_daily_deaths, num_steps), plain(m_value, 1)
        # infected_time as the episode left it (copied by the episode's own launch sequence): an agent is an exposure candidate only
        # up to its exposure, so the kernel fetches no dne_dr entry of later steps; not with an intervention hook, which may make
        # an agent susceptible again
        itime = sv.get("itime_copy") if sv.get("itime_valid") else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        nat.check(self.lib.at_seirm_episode_backward(
            int(num_steps), int(self.graph.n_rows), int(sv["ld"]), _ptr(sv["dne_dr"]), _ptr(sv["recovered_step"]),
            _ptr(sv["sum_dne_dr"]), _ptr(sv["death_count"]), _ptr(gi), _ptr(gd), _ptr(mv), _ptr(grad_r), _ptr(grad_m),
            _ptr(self.workspace), 0, _ptr(sv["step_skipped"], torch.int32),
            None if itime is None else ctypes.c_void_p(itime.data_ptr()), nat.TIMER_F32 if (itime is not None and itime.dtype == torch.float32) else nat.TIMER_I16,
            ctypes.c_void_p(stream)))
        self.launch_count += 1
        if self.pg is not None:
            import torch.distributed as dist

            dist.all_reduce(packed, group=self.pg)
        return packed

    def table_gradients(self, sv, num_steps, g_daily_infected, g_daily_deaths, m_value, r_steps, age):
        """dL/dSFSusceptibility (n_sfs,) and dL/dSFInfector (n_sfi,) of an episode, from the buffers the forward saved — no per-edge
        intermediate, like the R / M gradients.  lambda_i is linear in S[age_i] and, source by source, in A[stage_j]
        (transition.py:63-64,78-83), so with c_i(t) = gI[t] + gD[recovery step of i] * M / 2 (the cotangent of ne_i(t), through the
        float stage to the deaths):
            dL/dS[a] = sum_t R_t / S[a] * sum_{i: age_i = a} c_i(t) * dne_dr[t, i]
            dL/dA[1] = sum_t R_t / A[1] * sum_i c_i(t) * dne_e[t, i]          (sources in stage E)
            dL/dA[2] = sum_t R_t / A[2] * sum_i c_i(t) * (dne_dr - dne_e)[t, i]
        dL/dA[0] (fractional-stage sources under soft isolation, whose A[0] = 0 contributes nothing to the forward) is not
        tracked and returned as 0, like dL/dA[3:].  A composition of torch reductions over the saved (T, N) buffers: an opt-in
        calibration path (both tables are `learnable: false` in the reference's config), not the hot path."""
        T, n, dev = int(num_steps), self.graph.n_rows, self.device
        if sv.get("dne_e") is None or sv.get("dne_dr") is None:
            raise RuntimeError("run the episode with want_grad=True and engine.want_table_grad = True first")
        gi = g_daily_infected.detach().reshape(-1)[:T].double()
        gd = g_daily_deaths.detach().reshape(-1)[:T].double()
        half_m = float(m_value.detach().reshape(-1)[0]) * 0.5
        rstep = sv["recovered_step"].long()
        c_rec = torch.where(rstep < T, gd[rstep.clamp(max=T - 1)] * half_m, torch.zeros((), dtype=torch.float64, device=dev))
        r = r_steps.detach().reshape(-1)[:T].double()
        age_l = age.reshape(-1).long().clamp(0, len(self.spec.sfs) - 1)
        grad_sfs = torch.zeros(len(self.spec.sfs), dtype=torch.float64, device=dev)
        ge = torch.zeros((), dtype=torch.float64, device=dev)
        gi2 = torch.zeros((), dtype=torch.float64, device=dev)
        skipped = sv["step_skipped"].cpu().tolist()                    # steps the forward skipped: their rows mean zero
        for t in range(T):                                             # one (N,) pass per step keeps the temporaries small
            if skipped[t]:
                continue
            w = (gi[t] + c_rec) * r[t]
            d = sv["dne_dr"][t, :n].double() * w
            e = sv["dne_e"][t, :n].double() * w
            grad_sfs.index_add_(0, age_l, d)
            ge += e.sum()
            gi2 += (d - e).sum()
        sfs = torch.tensor(self.spec.sfs, dtype=torch.float64, device=dev)
        grad_sfs = torch.where(sfs != 0, grad_sfs / sfs, torch.zeros_like(grad_sfs))
        grad_sfi = torch.zeros(len(self.spec.sfi), dtype=torch.float64, device=dev)
        if len(self.spec.sfi) > 1 and self.spec.sfi[1] != 0:
            grad_sfi[1] = ge / float(self.spec.sfi[1])
        if len(self.spec.sfi) > 2 and self.spec.sfi[2] != 0:
            grad_sfi[2] = gi2 / float(self.spec.sfi[2])
        if self.pg is not None:
            import torch.distributed as dist

            packed = torch.cat([grad_sfs, grad_sfi])
            dist.all_reduce(packed, group=self.pg)
            grad_sfs, grad_sfi = packed[: grad_sfs.numel()], packed[grad_sfs.numel():]
        return grad_sfs.float(), grad_sfi.float()

    def poll_errors(self):
        """Raise for an error flagged by an EARLIER episode without stalling the stream: the flag is copied to pinned host memory
        asynchronously after every call and inspected on the next one (check_errors() is the synchronous form)."""
        if self.device.type != "cuda":
            return self.check_errors()
        if getattr(self, "_err_host", None) is None:
            self._err_host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self._err_event = None
        if self._err_event is not None and self._err_event.query():
            code = int(self._err_host[0])
            if code != 0:
                self._err_host.zero_()
                self.error_flag.zero_()
                self._raise_for(code)
        self._err_host.copy_(self.error_flag, non_blocking=True)
        if self._err_event is None:
            self._err_event = torch.cuda.Event()
        self._err_event.record(torch.cuda.current_stream(self.device))

    @staticmethod
    def _raise_for(code):
        if code == 2:
            raise RuntimeError("halo exchange: a peer rank never arrived at the rendezvous (at_seirm_exchange_sparse timed out)")
        raise RuntimeError("lam_gamma_integrals index out of range during the episode (SURVEY.md A.2 Q10)")

    def check_errors(self):
        code = int(self.error_flag.item())
        if code == 2:
            raise RuntimeError("halo exchange: a peer rank never arrived at the rendezvous (at_seirm_exchange_sparse timed out)")
        if code != 0:
            raise RuntimeError("lam_gamma_integrals index out of range during the episode (SURVEY.md A.2 Q10)")


def _as_i64(v):
    v = int(v) & ((1 << 64) - 1)
    return v - (1 << 64) if v >= (1 << 63) else v


def noise_fill(out, seed, step, what, p_isolate=0.3, id_offset=0):
    """Fill `out` (float32 CUDA tensor) with the counter-hash noise of (seed, step): what =
    'uniform' (exposure U), 'isolation' (uniform W) or 'isolation_hard'."""
    code = {"uniform": 0, "isolation": 1, "isolation_hard": 2}[what]
    stream = torch.cuda.current_stream(out.device).cuda_stream
    nat.check(nat.lib().at_noise_fill(_ptr(out, torch.float32, "out"), out.numel(), int(id_offset),
                                      int(seed) & ((1 << 64) - 1), int(step), code, float(p_isolate),
                                      ctypes.c_void_p(stream)))
    return out
