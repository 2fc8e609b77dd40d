"""SEIRMProgression — substep '1' transition, B200 path.

Drop-in for the reference class (models/covid/substeps/seirm_progression/transition.py:9-121): E->I->R
stage moves on `next_stage_time <= t`, timer rewrites on `== t` (Q5), daily deaths
`+= M * sum(stage * [stage==I and due] / 2)` with `calibrate_M` / `learnable_args["M"]`; one CUDA launch
(`agenttorch_b200::seirm_progression`) with registered autograd."""
import re

import torch

from agenttorch_b200 import ops
from agenttorch_b200.core.helpers import get_by_path
from agenttorch_b200.core.substep import SubstepTransition
from agenttorch_b200.models.covid.context import as_kernel_f32, get_context


class SEIRMProgression(SubstepTransition):
    fusable = True

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        meta = self.config["simulation_metadata"]
        self.device = torch.device(meta["device"])
        self.num_timesteps = meta["num_steps_per_episode"]
        self.calibration_mode = meta["calibration"]
        self.INFINITY_TIME = meta["num_steps_per_episode"] + 20
        self.INFECTED_TO_RECOVERED_TIME = meta["INFECTED_TO_RECOVERED_TIME"]

    def m_tensor(self):
        return self.calibrate_M if self.calibration_mode else self.learnable_args["M"]

    def forward(self, state, action):
        iv = self.input_variables
        t = int(state["current_step"])

        def var(name):
            return get_by_path(state, re.split("/", iv[name]))

        # the transmission substep registered the engine for this state's contact network
        net_vars = dict(self.config["substeps"]["0"]["transition"]["new_transmission"]["input_variables"])
        ctx = get_context(self.config, state, net_vars)
        stage = as_kernel_f32(var("disease_stage"))
        nst = as_kernel_f32(var("next_stage_time"))
        daily_deaths = var("daily_deaths")
        new_stage, new_nst, new_dd, _, _ = ops.seirm_progression(ctx.handle, stage, nst, daily_deaths.reshape(-1),
                                                                  self.m_tensor(), t)
        ov = self.output_variables
        # the reference's daily_deaths becomes (1,T) after the first call (Q6)
        return {ov[0]: new_stage, ov[1]: new_nst, ov[2]: new_dd.reshape(1, -1)}
