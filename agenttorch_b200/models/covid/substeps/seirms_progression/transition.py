"""SEIRMSProgression transition — B200 path.

Drop-in for the reference class (models/covid/substeps/seirms_progression/transition.py:10-181): E -> I on the timer,
due infected agents draw a Bernoulli(M) death, recovered agents return to susceptible; daily deaths
`+= M * sum(stage * [I and due] / 2)`.  The reference's quirks are reproduced (I5-I7 of oracle/interventions_port.py:
swapped dead/recovered paths, rows 0/1 of next_stage_time overwritten, doubled stage tangents).  M is `external_M`, a leaf
created from the learnable argument like the reference's (:38-40).  Two CUDA launches
(`agenttorch_b200::seirms_progression`, autograd registered)."""
import torch

from agenttorch_b200 import interventions as iv_ops
from agenttorch_b200.core.substep import SubstepTransition
from agenttorch_b200.models.covid.context import as_kernel_f32
from agenttorch_b200.models.covid.substeps.common import EpisodeNoise, require_cuda, state_var


class SEIRMSProgression(SubstepTransition, EpisodeNoise):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        meta = self.config["simulation_metadata"]
        self.device = torch.device(meta["device"])
        self.num_timesteps = meta["num_steps_per_episode"]
        for name, want in (("SUSCEPTIBLE_VAR", 0), ("EXPOSED_VAR", 1), ("INFECTED_VAR", 2), ("RECOVERED_VAR", 3)):
            if int(meta.get(name, want)) != want:
                raise NotImplementedError(f"simulation_metadata.{name} must be {want} for the SEIRMS kernel")
        self.INFINITY_TIME = meta["num_steps_per_episode"] + 20
        self.INFECTED_TO_RECOVERED_TIME = meta["INFECTED_TO_RECOVERED_TIME"]
        self.RECOVERED_TO_SUSCEPTIBLE_TIME = meta["RECOVERED_TO_SUSCEPTIBLE_TIME"]
        self.external_M = self.learnable_args["M"].detach().clone().requires_grad_(True)
        self.init_noise(meta)
        self.uniform = None               # optional callable t -> (N,) uniforms replacing the hash draw (cf. st_bernoulli)

    def forward(self, state, action):
        t = int(state["current_step"])
        iv = self.input_variables
        stage = as_kernel_f32(state_var(state, iv, "disease_stage"))
        require_cuda(stage, "SEIRMSProgression")
        nst = as_kernel_f32(state_var(state, iv, "next_stage_time"))
        daily_deaths = state_var(state, iv, "daily_deaths")
        m = self.external_M
        if m.device != stage.device:
            m = self.external_M = m.detach().to(stage.device).requires_grad_(True)
        u = self.uniform(t) if self.uniform is not None else None
        new_stage, new_nst, new_dd, _, _ = iv_ops.seirms_progression(
            stage, nst, daily_deaths.reshape(-1), m, t, self.episode_seed(t), self.id_offset, u,
            float(self.INFECTED_TO_RECOVERED_TIME), float(self.INFINITY_TIME), float(self.RECOVERED_TO_SUSCEPTIBLE_TIME))
        ov = self.output_variables
        return {ov[0]: new_stage, ov[1]: new_nst, ov[2]: new_dd.reshape(1, -1)}       # (1,T) like the reference (Q6)
