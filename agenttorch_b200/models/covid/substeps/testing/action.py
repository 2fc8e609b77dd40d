"""AcceptTest policy — B200 path.

Drop-in for the reference class (models/covid/substeps/testing/action.py:13-53): reads `is_quarantined`,
`disease_stage`, `test_compliance_prob` by path from the state, returns the float32 0/1 `(N,1)` acceptance.  The
reference's eligibility test `stage > S OR stage < R` holds for every stage (quirk I3), so the disease stage does not
enter: accepted = not quarantined and compliance draw.  One CUDA launch (`agenttorch_b200::testing_accept`)."""
import torch

from agenttorch_b200 import interventions as iv_ops
from agenttorch_b200.core.substep import SubstepAction
from agenttorch_b200.models.covid.substeps.common import EpisodeNoise, ScalarCache, require_cuda, state_var


class AcceptTest(SubstepAction, EpisodeNoise):
    def __init__(self, config, input_variables, output_variables, arguments):
        super().__init__(config, input_variables, output_variables, arguments)
        meta = self.config["simulation_metadata"]
        self.num_agents = meta["num_agents"]
        self.device = torch.device(meta["device"])
        self.init_noise(meta)
        self.uniform = None
        self._scalar = ScalarCache()

    def forward(self, state, observation):
        t = int(state["current_step"])
        is_q = state_var(state, self.input_variables, "is_quarantined")
        require_cuda(is_q, "AcceptTest")
        p = self._scalar("p", state_var(state, self.input_variables, "test_compliance_prob"))
        u = self.uniform(t) if self.uniform is not None else None
        return {self.output_variables[0]: iv_ops.testing_accept(is_q.detach(), p, t, self.episode_seed(t), self.id_offset, u)}
