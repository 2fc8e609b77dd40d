"""Substep 'Eat' — reference models/predator_prey/substeps/eat.py:24-110 on libat_predprey.so (AT_PP_EAT)."""
import torch

from agenttorch_b200.core.registry import Registry
from agenttorch_b200.core.substep import SubstepAction, SubstepTransition
from agenttorch_b200.predprey import PredPreyState, info_bytes, pack_cells, unpack_stage

from ..context import engine_for, get_var, require_cuda, spec_from_bounds


class MaskedPositions:
    """A list of positions selected by a device mask (the reference builds a Python list of row tensors,
    eat.py:37-45 / hunt.py:36-43): len / iteration / indexing materialise on demand."""

    def __init__(self, positions, mask):
        self.positions, self.mask = positions, mask
        self._rows = None

    def rows(self):
        if self._rows is None:
            self._rows = self.positions[self.mask]
        return self._rows

    def __len__(self):
        return int(self.rows().shape[0])

    def __getitem__(self, i):
        return self.rows()[i]

    def __iter__(self):
        return iter(self.rows())


def listed_cells(position_list, max_y, n_cells, device):
    """Cells named by an explicit list of [x, y] positions (int64 tensor)."""
    pos = torch.stack([torch.as_tensor(p) for p in position_list]).to(device=device, dtype=torch.int64)
    return pack_cells(pos, max_y, n_cells)[0].to(torch.int64)


@Registry.register_substep("find_eatable_grass", "policy")
class FindEatableGrass(SubstepAction):
    def forward(self, state, observations):
        iv = self.input_variables
        bounds, positions = get_var(state, iv["bounds"]), get_var(state, iv["positions"])
        growth = get_var(state, iv["grass_growth"])
        require_cuda(positions, growth)
        node = int(bounds.reshape(-1)[1]) * positions[:, 0] + positions[:, 1]            # eat.py:40
        return {self.output_variables[0]: MaskedPositions(positions, growth.reshape(-1)[node] == 1)}


@Registry.register_substep("eat_grass", "transition")
class EatGrass(SubstepTransition):
    def forward(self, state, action):
        iv = self.input_variables
        eatable = action["prey"]["eatable_grass_positions"]
        if len(eatable) < 1:                                                            # eat.py:66-67
            return {}
        bounds, prey_pos, energy = get_var(state, iv["bounds"]), get_var(state, iv["prey_pos"]), get_var(state, iv["energy"])
        growth, countdown = get_var(state, iv["grass_growth"]), get_var(state, iv["growth_countdown"])
        require_cuda(prey_pos, energy, growth, countdown)
        n_cells = growth.reshape(-1).shape[0]
        spec = spec_from_bounds(bounds, n_cells=n_cells, grass_nutrition=float(get_var(state, iv["nutrition"]).reshape(-1)[0]),
                                regrowth_time=float(get_var(state, iv["regrowth_time"]).reshape(-1)[0]))
        eng = engine_for(spec, energy.device)
        dev = energy.device
        none_i, none_f = torch.zeros(0, dtype=torch.int32, device=dev), torch.zeros(0, dtype=torch.float32, device=dev)
        st = PredPreyState.from_reference(spec, torch.zeros(0, 2, dtype=torch.int64, device=dev), none_f, prey_pos, energy, growth, countdown)
        st.pred_cell = none_i
        if not isinstance(eatable, MaskedPositions):
            # explicit position list (the reference's unit tests): exactly the listed cells are eaten (eat.py:80-99)
            true_info, cells = st.info, listed_cells(eatable, spec.max_y, n_cells, dev)
            st.info = torch.zeros(info_bytes(n_cells), dtype=torch.uint8, device=dev)
            st.info[cells] = 1
            eng.eat(st)
            st.info = true_info
            st.info[cells] = 0
        else:
            eng.eat(st)
        out_stage = unpack_stage(st.info, n_cells, dtype=growth.dtype if growth.dtype in (torch.int64, torch.float32) else torch.int64)
        ov = self.output_variables
        return {ov[0]: st.prey_energy.reshape(energy.shape).to(energy.dtype), ov[1]: out_stage.reshape(growth.shape),
                ov[2]: st.countdown.reshape(countdown.shape)}
