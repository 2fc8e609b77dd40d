"""Substep 'Grow' — reference models/predator_prey/substeps/grow.py:24-46 on libat_predprey.so (AT_PP_GROW)."""
import torch

from agenttorch_b200.core.registry import Registry
from agenttorch_b200.core.substep import SubstepTransition
from agenttorch_b200.predprey import PredPreyState, pack_stage, unpack_stage

from ..context import engine_for, get_var, require_cuda, spec_from_bounds


@Registry.register_substep("grow_grass", "transition")
class GrowGrass(SubstepTransition):
    def forward(self, state, action):
        iv = self.input_variables
        growth, countdown = get_var(state, iv["grass_growth"]), get_var(state, iv["growth_countdown"])
        require_cuda(growth, countdown)
        n_cells = growth.reshape(-1).shape[0]
        dev = growth.device
        spec = spec_from_bounds(torch.tensor([n_cells, 1]), n_cells=n_cells)
        eng = engine_for(spec, dev)
        cd = countdown.reshape(-1)
        cd = cd.clone() if cd.dtype in (torch.float32, torch.float64) else cd.to(torch.float32)     # int + full(-1) stays integral
        none_i, none_f = torch.zeros(0, dtype=torch.int32, device=dev), torch.zeros(0, dtype=torch.float32, device=dev)
        st = PredPreyState(none_i, none_f, none_i, none_f, pack_stage(growth, n_cells), cd.contiguous())
        eng.grow(st)
        ov = self.output_variables
        new_cd = st.countdown.reshape(countdown.shape)
        if not countdown.is_floating_point():
            new_cd = new_cd.to(countdown.dtype)
        return {ov[0]: unpack_stage(st.info, n_cells).view(-1, 1), ov[1]: new_cd}              # grow.py:43-46
