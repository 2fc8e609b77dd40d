"""UpdateQuarantineStatus transition — B200 path.

Drop-in for the reference class (models/covid/substeps/quarantine/transition.py:8-122): quarantines that have lasted
`quarantine_days` end, `start_compliance_action` starts one (start date = t), `break_compliance_action` breaks one; the
status is the reference's float32 counter (quirk I1 of oracle/interventions_port.py: it is decremented by "end" and
"break" independently).  One CUDA launch (`agenttorch_b200::quarantine_update`)."""
import torch

from agenttorch_b200 import interventions as iv_ops
from agenttorch_b200.core.substep import SubstepTransition
from agenttorch_b200.models.covid.substeps.common import require_cuda, state_var


class UpdateQuarantineStatus(SubstepTransition):
    def __init__(self, config, input_variables, output_variables, arguments):
        super().__init__(config, input_variables, output_variables, arguments)
        meta = self.config["simulation_metadata"]
        self.device = torch.device(meta["device"])
        self.num_agents = meta["num_agents"]
        self.quarantine_days = meta["quarantine_days"]
        self.INFINITY_TIME = meta["INFINITY_TIME"]

    def forward(self, state, action):
        t = int(state["current_step"])
        is_q = state_var(state, self.input_variables, "is_quarantined")
        start_date = state_var(state, self.input_variables, "quarantine_start_date")
        require_cuda(is_q, "UpdateQuarantineStatus")
        acts = action["citizens"]
        new_q, new_date = iv_ops.quarantine_update(is_q, start_date, acts["start_compliance_action"], acts["break_compliance_action"],
                                                   t, float(self.quarantine_days), float(self.INFINITY_TIME))
        return {self.output_variables[0]: new_q, self.output_variables[1]: new_date}
