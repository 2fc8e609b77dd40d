"""covid (SEIRM) model: registry wiring identical to the reference's (models/covid/__init__.py:4-50) —
same names, same keys — with the B200 substep classes behind them."""
from agenttorch_b200.core.registry import Registry


def get_registry():
    reg = Registry()
    from .substeps.new_transmission.action import MakeIsolationDecision
    from .substeps.new_transmission.transition import NewTransmission
    from .substeps.seirm_progression.transition import SEIRMProgression
    from .substeps import utils

    reg.register(MakeIsolationDecision, "make_isolation_decision", key="policy")
    reg.register(NewTransmission, "new_transmission", key="transition")
    reg.register(SEIRMProgression, "seirm_progression", key="transition")
    # classes the reference ships without registering them (substeps/{quarantine,testing,seirms_progression}/)
    from .substeps.quarantine.action import BreakCompliance, StartCompliance
    from .substeps.quarantine.observation import GetFromState
    from .substeps.quarantine.transition import UpdateQuarantineStatus
    from .substeps.seirms_progression.transition import SEIRMSProgression
    from .substeps.testing.action import AcceptTest
    from .substeps.testing.transition import UpdateTestStatus

    reg.register(GetFromState, "get_from_state", key="observation")
    reg.register(StartCompliance, "start_compliance", key="policy")
    reg.register(BreakCompliance, "break_compliance", key="policy")
    reg.register(UpdateQuarantineStatus, "update_quarantine_status", key="transition")
    reg.register(AcceptTest, "accept_test", key="policy")
    reg.register(UpdateTestStatus, "update_test_status", key="transition")
    reg.register(SEIRMSProgression, "seirms_progression", key="transition")
    reg.register(utils.network_from_file, "network_from_file", key="network")
    for name in ("read_from_file", "get_lam_gamma_integrals", "get_mean_agent_interactions", "get_infected_time",
                 "get_next_stage_time", "load_population_attribute", "initialize_id"):
        reg.register(getattr(utils, name), name, key="initialization")
    from . import synthetic  # noqa: F401  (registers the synthetic_* helpers)
    return reg


registry = get_registry()


def get_runner(config, registry=None):
    from agenttorch_b200.core.runner import Runner

    return Runner(config, get_registry())          # re-assert this model's helpers in the shared tables (see Executor)
