"""Substep 'Hunt' — reference models/predator_prey/substeps/hunt.py:24-92 on libat_predprey.so (AT_PP_HUNT)."""
import torch

from agenttorch_b200 import _native_pp as nat
from agenttorch_b200.core.registry import Registry
from agenttorch_b200.core.substep import SubstepAction, SubstepTransition
from agenttorch_b200.predprey import PredPreyState, info_bytes, occupancy_bits, pack_cells

from ..context import engine_for, get_var, require_cuda, spec_from_bounds
from .eat import MaskedPositions, listed_cells


def _grid_extent(*positions):
    """Bounds are not an input of the hunt substep (config.yaml:325-347): any injective cell numbering works, so
    use the bounding box of the positions."""
    mx = max(int(p[:, 0].max()) for p in positions if p.numel()) + 1
    my = max(int(p[:, 1].max()) for p in positions if p.numel()) + 1
    return mx, my


@Registry.register_substep("find_targets", "policy")
class FindTargets(SubstepAction):
    def forward(self, state, observations):
        iv = self.input_variables
        prey_pos, pred_pos = get_var(state, iv["prey_pos"]), get_var(state, iv["pred_pos"])
        require_cuda(prey_pos, pred_pos)
        mx, my = _grid_extent(prey_pos, pred_pos)
        has_prey = torch.zeros(mx * my, dtype=torch.bool, device=prey_pos.device)
        has_prey[my * prey_pos[:, 0] + prey_pos[:, 1]] = True
        return {self.output_variables[0]: MaskedPositions(pred_pos, has_prey[my * pred_pos[:, 0] + pred_pos[:, 1]])}   # hunt.py:38-40


@Registry.register_substep("hunt_prey", "transition")
class HuntPrey(SubstepTransition):
    def forward(self, state, action):
        iv = self.input_variables
        targets = action["predator"]["target_positions"]
        if len(targets) < 1:                                                            # hunt.py:60-61
            return {}
        prey_pos, prey_energy = get_var(state, iv["prey_pos"]), get_var(state, iv["prey_energy"])
        pred_pos, pred_energy = get_var(state, iv["pred_pos"]), get_var(state, iv["pred_energy"])
        require_cuda(prey_pos, prey_energy, pred_pos, pred_energy)
        mx, my = _grid_extent(prey_pos, pred_pos)
        dev = prey_energy.device
        spec = spec_from_bounds(torch.tensor([mx, my]), n_cells=mx * my,
                                prey_nutrition=float(get_var(state, iv["nutritional_value"]).reshape(-1)[0]))
        eng = engine_for(spec, dev)
        f32 = lambda t: t.reshape(-1).to(torch.float32).clone().contiguous()          # noqa: E731
        none_f = torch.zeros(0, dtype=torch.float32, device=dev)
        info = torch.zeros(info_bytes(mx * my), dtype=torch.uint8, device=dev)
        st = PredPreyState(pack_cells(pred_pos, my, mx * my)[0], f32(pred_energy), pack_cells(prey_pos, my, mx * my)[0],
                           f32(prey_energy), info, none_f)
        if isinstance(targets, MaskedPositions):
            eng.hunt(st)                          # target cells = cells holding a predator and a prey (hunt.py:36-40)
        else:
            # explicit target list (the reference's unit tests): prey / predators AT the listed cells (hunt.py:69-80)
            info[listed_cells(targets, my, mx * my, dev)] = sum(occupancy_bits(0))
            eng._launch(st, nat.HUNT)
        ov = self.output_variables
        return {ov[0]: st.prey_energy.reshape(prey_energy.shape).to(prey_energy.dtype),
                ov[1]: st.pred_energy.reshape(pred_energy.shape).to(pred_energy.dtype)}
