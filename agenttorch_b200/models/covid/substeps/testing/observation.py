"""GetFromState observation of the testing substep (reference: models/covid/substeps/testing/observation.py:9-22)."""
from agenttorch_b200.models.covid.substeps.quarantine.observation import GetFromState  # noqa: F401
