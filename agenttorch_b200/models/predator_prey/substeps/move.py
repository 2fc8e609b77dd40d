"""Substep 'Move' — same registry names, constructor and outputs as the reference's
models/predator_prey/substeps/move.py:49-123, computed by libat_predprey.so (AT_PP_MOVE)."""
import torch

from agenttorch_b200.core.registry import Registry
from agenttorch_b200.core.substep import SubstepAction, SubstepObservation, SubstepTransition
from agenttorch_b200.predprey import PredPreyState, pack_cells, unpack_cells

from ..context import engine_for, get_var, require_cuda, spec_from_bounds, species_of, topology


class NeighborLists:
    """`possible_neighbors` (move.py:66-70) without materialising a Python list per agent: behaves like the
    reference's list (len, indexing -> (k,2) int64 coordinates in ascending node order), and carries what
    DecideMovement needs to draw the moves on the device."""

    def __init__(self, positions, bounds, lattice, csr):
        self.positions, self.bounds, self.lattice, self.csr = positions, bounds, lattice, csr
        self.max_y = int(bounds.reshape(-1)[1])

    def __len__(self):
        return int(self.positions.shape[0])

    def __getitem__(self, i):
        pos = self.positions[i]
        node = int(self.max_y * pos[0] + pos[1])
        if self.lattice is not None:
            nbr, cnt = self.lattice.neighbors(torch.tensor([node], device=self.positions.device))
            idx = nbr[0, : int(cnt[0])]
        else:
            idx = self.csr.idx[int(self.csr.ptr[node]): int(self.csr.ptr[node + 1])].to(torch.int64)
        y = idx % self.max_y
        return torch.stack([(idx - y) // self.max_y, y], dim=1)

    def __iter__(self):
        return (self[i] for i in range(len(self)))


@Registry.register_substep("find_neighbors", "observation")
class FindNeighbors(SubstepObservation):
    def forward(self, state):
        iv = self.input_variables
        bounds, adj, positions = get_var(state, iv["bounds"]), get_var(state, iv["adj_grid"]), get_var(state, iv["positions"])
        require_cuda(positions)
        lattice, csr = topology(adj, positions.device)
        return {self.output_variables[0]: NeighborLists(positions, bounds, lattice, csr)}


@Registry.register_substep("decide_movement", "policy")
class DecideMovement(SubstepAction):
    """next position = a uniformly drawn neighbour if energy > 0, else the current one (move.py:82-92).  The draw of
    agent i at the k-th call is the counter hash (seed, k, species, i) instead of Python's `random.choice`."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.calls = 0
        meta = (self.config or {}).get("simulation_metadata", {})
        self.seed = int(meta.get("seed", 0))

    def forward(self, state, observations):
        iv = self.input_variables
        positions, energy = get_var(state, iv["positions"]), get_var(state, iv["energy"])
        require_cuda(positions, energy)
        nb = observations["possible_neighbors"]
        t, self.calls = self.calls, self.calls + 1
        species = species_of(iv["positions"])
        if not isinstance(nb, NeighborLists):
            return {self.output_variables[0]: self._from_lists(positions, energy, nb, t, species)}
        spec = spec_from_bounds(nb.bounds, n_cells=nb.lattice.n_cells if nb.lattice is not None else nb.csr.n_cells,
                                lattice=nb.lattice, apply_pred_work=False, apply_prey_work=False)
        eng = engine_for(spec, positions.device, nb.csr)
        cells, flag = pack_cells(positions, spec.max_y, spec.n_cells)
        e = energy.reshape(-1).to(torch.float32).clone().contiguous()
        none_i = torch.zeros(0, dtype=torch.int32, device=positions.device)
        none_f = torch.zeros(0, dtype=torch.float32, device=positions.device)
        st = PredPreyState(cells, e, none_i, none_f, None, none_f) if species == 0 else \
            PredPreyState(none_i, none_f, cells, e, None, none_f)
        eng.move(st, t, self.seed)
        return {self.output_variables[0]: unpack_cells(cells, spec.max_y)}

    def _from_lists(self, positions, energy, lists, t, species):
        """Explicit per-agent neighbour lists (as the reference's unit tests pass them): same draw, same pick."""
        from agenttorch_b200 import noise

        n = positions.shape[0]
        draw = noise.noise_u64(self.seed, t, torch.arange(n, device=positions.device) | (species << 39))
        draw = (draw >> 40) & 0xFFFFFF
        out = positions.clone()
        alive = (energy.reshape(-1) > 0).tolist()
        for i, nbrs in enumerate(lists):
            if alive[i] and len(nbrs) > 0:
                out[i] = torch.as_tensor(nbrs[(int(draw[i]) * len(nbrs)) >> 24]).to(out)
        return out


@Registry.register_substep("update_positions", "transition")
class UpdatePositions(SubstepTransition):
    def forward(self, state, action):
        iv = self.input_variables
        prey_energy, pred_energy = get_var(state, iv["prey_energy"]), get_var(state, iv["pred_energy"])
        prey_work, pred_work = get_var(state, iv["prey_work"]), get_var(state, iv["pred_work"])
        ov = self.output_variables
        # keyed by output_variables in this order; a name listed twice keeps its LAST value (config.yaml:281-285)
        return {ov[0]: action["prey"]["next_positions"], ov[1]: prey_energy + (-prey_work.reshape(-1)[0]),
                ov[2]: action["predator"]["next_positions"], ov[3]: pred_energy + (-pred_work.reshape(-1)[0])}
