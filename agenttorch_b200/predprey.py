"""Predator-prey grid model on one B200: engine over libat_predprey.so (include/at_predprey.h).

Reference: agent_torch/models/predator_prey/substeps/{move,eat,hunt,grow}.py driven by core/runner.py.  The reference
keeps positions as (n,2) int64 coordinates, the grass as (cells,1) stage + countdown tensors and the lattice as a
dense (cells x cells) adjacency matrix; the engine keeps (cell int32, energy float32) per agent, one bit + one
countdown per cell and an implicit lattice (or a CSR adjacency for other networks), and converts at the boundary.

No CPU fallback: everything here needs CUDA tensors and the built library.
"""
import ctypes
from dataclasses import dataclass

import torch

from . import _native_pp as nat

STD_OUTPUTS = ("prey_pos", "prey_energy", "pred_pos", "pred_energy")


def _require_cuda(what, is_cuda):
    """No CPU fallback: every tensor and the engine's device must be CUDA.  (One function so that the host-emulation tests,
    which drive this engine against the kernel source compiled for the CPU, can lift the check.)"""
    if not is_cuda:
        raise RuntimeError(f"agenttorch_b200 predator-prey {what} need CUDA (no CPU fallback)")


def _ptr(t, dtype=None):
    if t is None:
        return None
    _require_cuda("kernels: tensors", t.is_cuda)
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class GridLattice:
    """What this package's `grid` network helper returns instead of the reference's dense adjacency matrix
    (core/helpers/environment.py:6-9): nx.grid_graph(dim=shape) numbers its nodes so that node k touches
    k +- shape[0] and, inside a run of shape[0] nodes, k +- 1."""

    def __init__(self, shape):
        self.shape = tuple(int(s) for s in shape)
        self.width = self.shape[0]
        self.n_cells = 1
        for s in self.shape:
            self.n_cells *= s
        if len(self.shape) != 2:
            raise ValueError("GridLattice: only 2-D grids")

    def __len__(self):
        return self.n_cells

    def neighbors(self, cells):
        """(n,4) ascending neighbour cells (-1 = unused slot) and (n,) counts, as torch tensors on cells.device."""
        k = cells.to(torch.int64)
        i = k % self.width
        cand = torch.stack([k - self.width, k - 1, k + 1, k + self.width], dim=1)
        ok = torch.stack([k - self.width >= 0, i > 0, i < self.width - 1, k + self.width < self.n_cells], dim=1)
        order = torch.argsort((~ok).to(torch.int8), dim=1, stable=True)
        nbr = torch.where(torch.gather(ok, 1, order), torch.gather(cand, 1, order), torch.full_like(cand, -1))
        return nbr, ok.sum(1)

    def to_dense(self, device="cpu"):
        cells = torch.arange(self.n_cells, device=device)
        nbr, cnt = self.neighbors(cells)
        dense = torch.zeros(self.n_cells, self.n_cells, dtype=torch.int64, device=device)
        for j in range(4):
            sel = nbr[:, j] >= 0
            dense[cells[sel], nbr[sel, j]] = 1
        return dense


class CsrAdjacency:
    """Ascending neighbour lists of an arbitrary dense adjacency matrix (entries == 1, move.py:38-46)."""

    def __init__(self, dense, device):
        dense = torch.as_tensor(dense)
        self.n_cells = int(dense.shape[0])
        nz = torch.nonzero(dense == 1)                     # row-major order: neighbours ascending within a row
        counts = torch.bincount(nz[:, 0], minlength=self.n_cells)
        ptr = torch.zeros(self.n_cells + 1, dtype=torch.int64)
        ptr[1:] = torch.cumsum(counts.cpu(), 0)
        self.ptr = ptr.to(device)
        self.idx = nz[:, 1].to(torch.int32).contiguous().to(device)
        if self.idx.numel() == 0:
            self.idx = torch.zeros(1, dtype=torch.int32, device=device)


@dataclass
class PredPreySpec:
    """Model constants (reference config.yaml:37-220) + the lattice."""
    max_x: int
    max_y: int
    n_cells: int
    width: int
    pred_work: float = 1.0
    prey_work: float = 5.0
    prey_nutrition: float = 20.0
    grass_nutrition: float = 7.0
    regrowth_time: float = 100.0
    apply_pred_work: bool = True
    apply_prey_work: bool = True
    tiled: bool = False       # info bytes / countdowns stored in tiles of 8 rows x 16 columns (implicit lattice only)

    @property
    def tiles_per_row(self):
        return (int(self.width) + 15) // 16 if self.tiled else 0

    @property
    def n_storage(self):
        """Cells as stored: the lattice padded to whole tiles when tiled."""
        if not self.tiled:
            return int(self.n_cells)
        rows = -(-int(self.n_cells) // int(self.width))
        return -(-rows // 8) * self.tiles_per_row * 128

    def storage_index(self, device):
        """(n_cells,) int64: where cell k is stored (identity when untiled).  Cached per device."""
        if not self.tiled:
            return None
        key = (str(device), int(self.width), int(self.n_cells))
        hit = _STORAGE_INDEX.get(key)
        if hit is None:
            if len(_STORAGE_INDEX) > 4:
                _STORAGE_INDEX.clear()
            k = torch.arange(int(self.n_cells), device=device)
            row, col = k // int(self.width), k % int(self.width)
            hit = _STORAGE_INDEX[key] = (((row >> 3) * self.tiles_per_row + (col >> 4)) << 7) | ((row & 7) << 4) | (col & 15)
        return hit

    @staticmethod
    def work_flags(update_outputs):
        """Which new energies survive the transition's YAML output mapping (move.py:118-123, controller.py:73-81):
        the four results are keyed by output_variables in order, later duplicates overwrite earlier ones."""
        res = {}
        for key, what in zip(update_outputs, STD_OUTPUTS):
            res[key] = what
        return res.get("pred_energy") == "pred_energy", res.get("prey_energy") == "prey_energy"

    def c_model(self):
        return nat.PPModel(abi_version=nat.ABI_VERSION, width=int(self.width), max_y=int(self.max_y),
                           apply_pred_work=int(self.apply_pred_work), apply_prey_work=int(self.apply_prey_work),
                           tiles_per_row=int(self.tiles_per_row), n_storage=int(self.n_storage), n_cells=int(self.n_cells), pred_work=float(self.pred_work), prey_work=float(self.prey_work),
                           prey_nutrition=float(self.prey_nutrition), grass_nutrition=float(self.grass_nutrition),
                           regrowth_time=float(self.regrowth_time))


_STORAGE_INDEX = {}


def info_bytes(n_cells):
    return (int(n_cells) + 15) & ~15


def occupancy_bits(parity):
    """(predator bit, prey bit) of the info byte for occupancy parity 0 / 1 (include/at_predprey.h)."""
    return 4 << (2 * parity), 8 << (2 * parity)


def pack_cells(coords, max_y, n_cells):
    """(n,2) int64 [x, y] -> (n,) int32 cells; raises on coordinates outside the grid."""
    coords = coords.reshape(-1, 2).to(torch.int64).contiguous()
    n = coords.shape[0]
    cells = torch.empty(n, dtype=torch.int32, device=coords.device)
    flag = torch.zeros(1, dtype=torch.int32, device=coords.device)
    nat.check(nat.lib().at_pp_pack_cells(_ptr(coords), n, int(max_y), int(n_cells), _ptr(cells), _ptr(flag), _stream()))
    return cells, flag


def unpack_cells(cells, max_y, out=None):
    n = cells.shape[0]
    if out is None:
        out = torch.empty(n, 2, dtype=torch.int64, device=cells.device)
    nat.check(nat.lib().at_pp_unpack_cells(_ptr(cells, torch.int32), n, int(max_y), _ptr(out, torch.int64), _stream()))
    return out


def pack_stage(stage, n_cells):
    stage = stage.reshape(-1)
    if stage.dtype not in (torch.int64, torch.float32):
        stage = stage.to(torch.float32)
    stage = stage.contiguous()
    info = torch.empty(info_bytes(n_cells), dtype=torch.uint8, device=stage.device)
    nat.check(nat.lib().at_pp_pack_stage(_ptr(stage), 0 if stage.dtype == torch.int64 else 1, int(n_cells), _ptr(info), _stream()))
    return info


def unpack_stage(info, n_cells, dtype=torch.int64, out=None):
    if out is None:
        out = torch.empty(n_cells, 1, dtype=dtype, device=info.device)
    nat.check(nat.lib().at_pp_unpack_stage(_ptr(info, torch.uint8), int(n_cells), _ptr(out), 0 if out.dtype == torch.int64 else 1, _stream()))
    return out


class PredPreyState:
    """Device state in the engine's layout.  Agents may be kept in any slot order: `pred_id` / `prey_id` (None =
    identity) give the agent number of every slot.  `sort_by_cell()` orders the slots by cell, so that the random
    cell lookups of a warp fall into a few 128-byte lines; random walks destroy that order only slowly (a warp's
    agents spread over ~sqrt(steps) grid rows)."""

    AGENT_KEYS = ("pred_cell", "pred_energy", "prey_cell", "prey_energy")
    KEYS = AGENT_KEYS + ("info", "countdown")

    def __init__(self, pred_cell, pred_energy, prey_cell, prey_energy, info, countdown, pred_id=None, prey_id=None):
        self.pred_cell, self.pred_energy = pred_cell, pred_energy
        self.prey_cell, self.prey_energy = prey_cell, prey_energy
        self.info, self.countdown = info, countdown          # info: one byte per cell (bit 0 = grown; other bits are scratch of a step)
        self.pred_id, self.prey_id = pred_id, prey_id

    @classmethod
    def from_reference(cls, spec, pred_pos, pred_energy, prey_pos, prey_energy, growth_stage, growth_countdown, validate=True,
                       sort=False):
        """From tensors in the reference's shapes (copies; the inputs are not aliased)."""
        pc, f0 = pack_cells(pred_pos, spec.max_y, spec.n_cells)
        yc, f1 = pack_cells(prey_pos, spec.max_y, spec.n_cells)
        if validate and int((f0 | f1).item()):
            raise ValueError("agent coordinates outside the grid (cell = max_y * x + y must be in [0, n_cells))")
        cd = growth_countdown.reshape(-1)
        cd = cd.clone() if cd.dtype in (torch.float32, torch.float64) else cd.to(torch.float32)
        if growth_stage.numel() != spec.n_cells or cd.numel() != spec.n_cells:
            raise ValueError("grass tensors must have one entry per cell")
        f32 = lambda t: t.reshape(-1).to(torch.float32).clone().contiguous()          # noqa: E731
        info = pack_stage(growth_stage, spec.n_cells)
        if spec.tiled:                            # cell-number order -> tile order; the padding never grows
            sidx = spec.storage_index(cd.device)
            tiled_info = torch.zeros(info_bytes(spec.n_storage), dtype=torch.uint8, device=cd.device)
            tiled_info[sidx] = info[: spec.n_cells]
            tiled_cd = torch.full((spec.n_storage,), float("inf"), dtype=cd.dtype, device=cd.device)
            tiled_cd[sidx] = cd
            info, cd = tiled_info, tiled_cd
        st = cls(pc, f32(pred_energy), yc, f32(prey_energy), info, cd.contiguous())
        if sort:
            st.sort_by_cell(spec)
        return st

    def sort_by_cell(self, spec=None):
        """Reorder the slots of both species by (the storage address of) their cell, IN PLACE (buffer addresses — and
        captured graphs — stay valid)."""
        sidx = spec.storage_index(self.pred_cell.device) if spec is not None and spec.tiled else None
        for cell, energy, name in ((self.pred_cell, self.pred_energy, "pred_id"), (self.prey_cell, self.prey_energy, "prey_id")):
            if cell.numel() == 0:
                continue
            order = torch.argsort(cell if sidx is None else sidx[cell.to(torch.int64)])
            ids = getattr(self, name)
            if ids is None:
                setattr(self, name, order.to(torch.int32))
            else:
                ids.copy_(ids[order])
            cell.copy_(cell[order])
            energy.copy_(energy[order])

    def clone(self):
        c = lambda t: None if t is None else t.clone()                                 # noqa: E731
        return PredPreyState(*(c(getattr(self, k)) for k in self.KEYS), pred_id=c(self.pred_id), prey_id=c(self.prey_id))

    def copy_(self, other):
        for k in self.KEYS:
            getattr(self, k).copy_(getattr(other, k))
        for k in ("pred_id", "prey_id"):
            mine, theirs = getattr(self, k), getattr(other, k)
            if theirs is None:
                setattr(self, k, None)
            elif mine is None:
                setattr(self, k, theirs.clone())
            else:
                mine.copy_(theirs)

    def in_agent_order(self, key):
        """The per-agent array `key` indexed by agent number (undoes the slot order)."""
        v = getattr(self, key)
        ids = self.pred_id if key.startswith("pred") else self.prey_id
        if ids is None:
            return v
        out = torch.empty_like(v)
        out[ids.to(torch.int64)] = v
        return out

    def grass_in_cell_order(self, spec):
        """(info bytes, countdown) indexed by cell number (undoes the tiled storage order)."""
        if not spec.tiled:
            return self.info, self.countdown
        sidx = spec.storage_index(self.info.device)
        info = torch.zeros(info_bytes(spec.n_cells), dtype=torch.uint8, device=self.info.device)
        info[: spec.n_cells] = self.info[sidx]
        return info, self.countdown[sidx]

    def to_reference(self, spec):
        """Dict of tensors in the reference's shapes (agents in agent-number order, grass in cell-number order)."""
        info, cd = self.grass_in_cell_order(spec)
        return {"pred_pos": unpack_cells(self.in_agent_order("pred_cell"), spec.max_y),
                "pred_energy": self.in_agent_order("pred_energy").reshape(-1, 1).clone(),
                "prey_pos": unpack_cells(self.in_agent_order("prey_cell"), spec.max_y),
                "prey_energy": self.in_agent_order("prey_energy").reshape(-1, 1).clone(),
                "growth_stage": unpack_stage(info, spec.n_cells), "growth_countdown": cd.reshape(-1, 1).clone()}


class PredPreyEngine:
    """Issues the kernels of one or many steps.  Per simulated step the fused schedule launches TWO kernels:
    agents [EAT, HUNT of step t  +  MOVE of step t+1] and cells [eaten -> reset, GROW, clear]."""

    def __init__(self, spec, device, adjacency=None, grid=0):
        self.spec = spec
        self.device = torch.device(device)
        _require_cuda("engine: the device", self.device.type == "cuda")
        nat.lib()
        self.model = spec.c_model()
        self.csr = adjacency if isinstance(adjacency, CsrAdjacency) else None
        if spec.tiled and self.csr is not None:
            raise ValueError("tiled cell storage needs the implicit lattice (no CSR adjacency)")
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.t_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.grid = int(grid)
        self.launch_count = 0
        self._graphs = {}

    # ---------------------------------------------------------------------------------------- launches
    def _launch(self, st, flags, t_move=0, occ_in=0, occ_out=1, seed=None, use_dev_clock=False):
        a = nat.PPLaunch()
        a.abi_version, a.flags, a.t_move = nat.ABI_VERSION, int(flags), int(t_move)
        a.countdown_f64 = int(st.countdown.dtype == torch.float64)
        if use_dev_clock:
            a.seed_ptr, a.t_ptr = _ptr(self.seed_dev), _ptr(self.t_dev)
        else:
            a.seed = int(seed or 0) & 0xFFFFFFFFFFFFFFFF
        a.n_pred, a.n_prey = st.pred_cell.numel(), st.prey_cell.numel()
        a.pred_cell, a.pred_energy = _ptr(st.pred_cell, torch.int32), _ptr(st.pred_energy, torch.float32)
        a.prey_cell, a.prey_energy = _ptr(st.prey_cell, torch.int32), _ptr(st.prey_energy, torch.float32)
        a.info, a.countdown = _ptr(st.info, torch.uint8) if st.info is not None else None, _ptr(st.countdown)
        a.occ_in, a.occ_out = int(occ_in), -1 if occ_out is None else int(occ_out)
        if st.pred_id is not None:
            a.pred_id = _ptr(st.pred_id, torch.int32)
        if st.prey_id is not None:
            a.prey_id = _ptr(st.prey_id, torch.int32)
        if self.csr is not None:
            a.nbr_ptr, a.nbr_idx = _ptr(self.csr.ptr, torch.int64), _ptr(self.csr.idx, torch.int32)
        a.grid = self.grid
        nat.check(nat.lib().at_pp_step(ctypes.byref(self.model), ctypes.byref(a), _stream()))
        self.launch_count += self.kernels_of(flags)

    @staticmethod
    def kernels_of(flags):
        return (int(bool(flags & nat.OCCUPANCY)) + int(bool(flags & (nat.EAT | nat.HUNT | nat.MOVE)))
                + int(bool(flags & (nat.EAT_APPLY | nat.GROW | nat.CLEAR))))

    def kernels_per_run(self, num_steps, resort_every=0):
        chunks = -(-num_steps // resort_every) if resort_every else 1
        return chunks + 2 * num_steps

    # substep-level entry points (the Controller path and the parity tests drive these one by one)
    def move(self, st, t, seed):
        """Substep 0 on the engine state: positions, energies."""
        self._launch(st, nat.MOVE, t_move=t, seed=seed, occ_out=None)

    def eat(self, st):
        self._launch(st, nat.EAT | nat.EAT_APPLY)

    def hunt(self, st):
        self._launch(st, nat.OCCUPANCY | nat.HUNT | nat.CLEAR)

    def grow(self, st):
        self._launch(st, nat.GROW)

    # ------------------------------------------------------------------------------------- fused steps
    def _issue(self, st, num_steps, seed, t0, dev_clock, resort_every=0):
        kw = dict(seed=seed, use_dev_clock=dev_clock)
        base = 0 if dev_clock else t0
        done = 0
        while done < num_steps:
            chunk = min(resort_every, num_steps - done) if resort_every else num_steps
            if done and st.pred_id is not None:
                st.sort_by_cell(self.spec)             # between two steps: nothing is in flight, the order is free
            self._launch(st, nat.MOVE, t_move=base + done, occ_in=1, occ_out=0, **kw)
            for s in range(chunk):
                flags = nat.EAT | nat.HUNT | nat.EAT_APPLY | nat.GROW | nat.CLEAR | nat.EAT_IMPLICIT
                if s + 1 < chunk:
                    flags |= nat.MOVE
                self._launch(st, flags, t_move=base + done + s + 1, occ_in=s & 1, occ_out=(s + 1) & 1, **kw)
            done += chunk

    def run(self, st, num_steps, seed, t0=0, cuda_graph=True, resort_every=0):
        """Advance `st` in place by `num_steps` steps; the move draws are addressed by steps t0, t0+1, ...
        `resort_every` = k > 0 re-sorts a cell-sorted state (st.pred_id set) every k steps."""
        if num_steps <= 0:
            return
        st.info.bitwise_and_(nat.INFO_GROWN)                   # scratch bits of an interrupted earlier call
        if not cuda_graph:
            self._issue(st, num_steps, seed, t0, dev_clock=False, resort_every=resort_every)
            return
        self.seed_dev.fill_(int(seed) if int(seed) < 2 ** 63 else int(seed) - 2 ** 64)
        self.t_dev.fill_(int(t0))
        key = (num_steps, resort_every) + tuple(0 if t is None else t.data_ptr() for t in (
            st.pred_cell, st.pred_energy, st.prey_cell, st.prey_energy, st.info, st.countdown, st.pred_id, st.prey_id))
        graph = self._graphs.get(key)
        if graph is None:
            before = self.launch_count
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(graph):
                self._issue(st, num_steps, seed, t0, dev_clock=True, resort_every=resort_every)
            self.launch_count = before
            if len(self._graphs) >= 4:
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = graph
        graph.replay()
        self.launch_count += self.kernels_per_run(num_steps, resort_every)

    def census(self, st):
        """(predators alive, prey alive, grown cells) — counts the reference's scripts print (vmap_simulation.py:193-200)."""
        a = nat.PPLaunch()
        a.abi_version = nat.ABI_VERSION
        a.n_pred, a.n_prey = st.pred_cell.numel(), st.prey_cell.numel()
        a.pred_energy, a.prey_energy, a.info = _ptr(st.pred_energy), _ptr(st.prey_energy), _ptr(st.info, torch.uint8)
        out = torch.zeros(3, dtype=torch.int64, device=self.device)
        nat.check(nat.lib().at_pp_census(ctypes.byref(a), int(self.spec.n_storage), _ptr(out), _stream()))
        return tuple(int(v) for v in out.tolist())
