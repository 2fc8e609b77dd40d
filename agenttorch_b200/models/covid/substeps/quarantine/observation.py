"""GetFromState observation (reference: models/covid/substeps/quarantine/observation.py:9-22, testing/observation.py):
returns every input variable read from the state, keyed by its name."""
import re

from agenttorch_b200.core.helpers import get_by_path
from agenttorch_b200.core.substep import SubstepObservation


class GetFromState(SubstepObservation):
    def forward(self, state):
        iv = self.input_variables or {}
        return {k: get_by_path(state, re.split("/", iv[k])) for k in iv.keys()}
