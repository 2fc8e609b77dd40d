"""StartCompliance / BreakCompliance policies — B200 path.

Drop-ins for the reference classes (models/covid/substeps/quarantine/action.py:16-68): same constructor, same
observation keys (`quarantine_start_prob`, `quarantine_break_prob`, `is_quarantined`, `disease_stage`), same bool
`(N,1)` outputs.  The reference samples with `discrete_sample` (Gumbel-softmax, tau 0.1, hard: "yes" with probability
p); here the draw is `u < p` on the counter hash (seed, step, agent; quarantine stream: start = high 24 bits, break =
bits 8..31), or on `self.uniform(t)` when a test injects it.  One CUDA launch each (`agenttorch_b200::quarantine_decide`)."""
import torch

from agenttorch_b200 import interventions as iv_ops
from agenttorch_b200.core.substep import SubstepAction
from agenttorch_b200.models.covid.substeps.common import EpisodeNoise, ScalarCache, require_cuda


class _Compliance(SubstepAction, EpisodeNoise):
    which = None

    def __init__(self, config, input_variables, output_variables, arguments):
        super().__init__(config, input_variables, output_variables, arguments)
        meta = self.config["simulation_metadata"]
        self.num_agents = meta["num_agents"]
        self.device = torch.device(meta["device"])
        self.init_noise(meta)
        self.uniform = None               # optional callable t -> (N,) uniforms replacing the hash draw
        self._scalar = ScalarCache()

    def forward(self, state, observation):
        t = int(state["current_step"])
        is_q, stage = observation["is_quarantined"], observation["disease_stage"]
        require_cuda(is_q, "quarantine policies")
        start = self.which == "start"
        p = self._scalar("p", observation["quarantine_start_prob" if start else "quarantine_break_prob"])
        u = self.uniform(t) if self.uniform is not None else None
        out = iv_ops.quarantine_decide(stage.detach(), is_q.detach(), p if start else 0.0, 0.0 if start else p, t,
                                       self.episode_seed(t), self.id_offset, u if start else None, None if start else u)
        return {self.output_variables[0]: out[0] if start else out[1]}


class StartCompliance(_Compliance):
    """Exposed or infected agents that are not quarantined start a quarantine with probability quarantine_start_prob."""
    which = "start"


class BreakCompliance(_Compliance):
    """Quarantined agents break their quarantine with probability quarantine_break_prob."""
    which = "break"
