"""UpdateTestStatus transition — B200 path.

Drop-in for the reference class (models/covid/substeps/testing/transition.py:18-175): agents whose result is due today
leave the queue and become ineligible for `test_ineligible_days`; willing agents that are neither cooling down nor
waiting enrol and get a result date `t + test_result_delay_days`.  The reference also draws true/false-positive
outcomes but its forward() drops them (quirk I4); they are not computed here.  One CUDA launch
(`agenttorch_b200::testing_update`)."""
import torch

from agenttorch_b200 import interventions as iv_ops
from agenttorch_b200.core.substep import SubstepTransition
from agenttorch_b200.models.covid.substeps.common import require_cuda, state_var


class UpdateTestStatus(SubstepTransition):
    def __init__(self, config, input_variables, output_variables, arguments):
        super().__init__(config, input_variables, output_variables, arguments)
        meta = self.config["simulation_metadata"]
        self.device = torch.device(meta["device"])
        self.num_agents = meta["num_agents"]
        self.test_ineligible_days = meta["test_ineligible_days"]
        self.test_result_delay_days = meta["test_result_delay_days"]
        self.INFINITY_TIME = meta["INFINITY_TIME"]

    def forward(self, state, action=None):
        t = int(state["current_step"])
        iv = self.input_variables
        awaiting = state_var(state, iv, "awaiting_test_result")
        require_cuda(awaiting, "UpdateTestStatus")
        new_aw, new_rd, new_re = iv_ops.testing_update(
            awaiting, state_var(state, iv, "test_result_date"), state_var(state, iv, "test_re_eligble_date"),
            action["citizens"]["test_acceptance_action"], t, float(self.test_ineligible_days), float(self.test_result_delay_days),
            float(self.INFINITY_TIME))
        ov = self.output_variables
        return {ov[0]: new_aw, ov[1]: new_rd, ov[2]: new_re}
