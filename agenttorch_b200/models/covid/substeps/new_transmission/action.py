"""MakeIsolationDecision policy (reference: models/covid/substeps/new_transmission/action.py:19-67).
heuristic mode, or no behavior attached: W ~ U[0,1) per agent; otherwise W = behavior(observation)."""
import torch

from agenttorch_b200.core.decorators import with_behavior
from agenttorch_b200.core.substep import SubstepAction


@with_behavior
class MakeIsolationDecision(SubstepAction):
    fusable = True      # the Runner may replace torch.rand by the in-kernel counter hash (same law)

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        meta = self.config["simulation_metadata"]
        self.device = torch.device(meta["device"])
        self.mode = meta["EXECUTION_MODE"]
        self.num_agents = meta["num_agents"]

    def draws_uniform(self):
        return self.mode == "heuristic" or self.behavior is None

    def forward(self, state, observation):
        if self.draws_uniform():
            will_isolate = torch.rand(self.num_agents, 1, device=self.device)
        else:
            will_isolate = self.behavior(observation)
        key = self.output_variables[0] if self.output_variables else "isolation_decision"
        return {key: will_isolate}
