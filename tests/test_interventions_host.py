"""Host side of the quarantine / testing / SEIRMS path without a GPU: the C-ABI library loads and exports what
include/at_interventions.h declares, argument errors come back as status codes, the registry and config helpers wire the
substeps the way the reference's Initializer parses them, the op backward formula equals autograd of the oracle, and
nothing falls back to the CPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from agenttorch_b200 import _native_iv
from oracle import noise as onoise

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    _native_iv.build_library()
    return _native_iv.lib()


def test_header_symbols_are_exported(built):
    header = open(os.path.join(ROOT, "include", "at_interventions.h")).read()
    declared = set(re.findall(r"^\s*(?:int|int64_t|const char\*)\s+(at_[a-z0-9_]+)\s*\(", header, flags=re.M))
    assert declared == set(_native_iv.EXPORTS), declared ^ set(_native_iv.EXPORTS)
    handle = ctypes.CDLL(_native_iv.LIB_PATH)
    for name in declared:
        assert getattr(handle, name) is not None
    assert f"#define AT_IV_ABI_VERSION {_native_iv.ABI_VERSION}" in header
    assert built.at_iv_sizeof_quarantine_args() == ctypes.sizeof(_native_iv.QuarantineArgs)
    assert built.at_iv_sizeof_testing_args() == ctypes.sizeof(_native_iv.TestingArgs)
    assert built.at_iv_sizeof_seirms_args() == ctypes.sizeof(_native_iv.SeirmsArgs)


def test_bad_arguments_return_errors_not_exceptions(built):
    assert built.at_iv_quarantine(None, None) != 0 and b"null" in built.at_iv_last_error()
    q = _native_iv.QuarantineArgs(abi_version=_native_iv.ABI_VERSION + 1)
    assert built.at_iv_quarantine(ctypes.byref(q), None) != 0 and b"ABI" in built.at_iv_last_error()
    q = _native_iv.QuarantineArgs(abi_version=_native_iv.ABI_VERSION, flags=0, n=8)
    assert built.at_iv_quarantine(ctypes.byref(q), None) != 0 and b"flags" in built.at_iv_last_error()
    q = _native_iv.QuarantineArgs(abi_version=_native_iv.ABI_VERSION, flags=_native_iv.APPLY, n=8)
    assert built.at_iv_quarantine(ctypes.byref(q), None) != 0 and b"is_q_in" in built.at_iv_last_error()
    q = _native_iv.QuarantineArgs(abi_version=_native_iv.ABI_VERSION, flags=_native_iv.APPLY, n=0)
    assert built.at_iv_quarantine(ctypes.byref(q), None) == 0                   # empty population: nothing to do
    t = _native_iv.TestingArgs(abi_version=_native_iv.ABI_VERSION, flags=_native_iv.DECIDE, n=8)
    assert built.at_iv_testing(ctypes.byref(t), None) != 0 and b"is_q" in built.at_iv_last_error()
    s = _native_iv.SeirmsArgs(abi_version=_native_iv.ABI_VERSION, n=1)
    assert built.at_iv_seirms_progression(ctypes.byref(s), None) != 0 and b"rows 0 and 1" in built.at_iv_last_error()
    s = _native_iv.SeirmsArgs(abi_version=_native_iv.ABI_VERSION, n=1 << 37)
    assert built.at_iv_seirms_progression(ctypes.byref(s), None) != 0


def test_registry_and_config_helpers():
    from agenttorch_b200.models import covid
    from agenttorch_b200.models.covid.interventions import add_intervention_substeps, use_seirms_progression
    from agenttorch_b200.models.covid.substeps.quarantine.action import BreakCompliance, StartCompliance
    from agenttorch_b200.models.covid.substeps.seirms_progression.transition import SEIRMSProgression
    from agenttorch_b200.models.covid.synthetic import synthetic_config

    reg = covid.get_registry()
    assert reg.helpers["policy"]["start_compliance"] is StartCompliance
    assert reg.helpers["policy"]["break_compliance"] is BreakCompliance
    assert reg.helpers["transition"]["seirms_progression"] is SEIRMSProgression
    for key, names in (("observation", ["get_from_state"]), ("policy", ["accept_test"]),
                       ("transition", ["update_quarantine_status", "update_test_status"])):
        assert all(n in reg.helpers[key] for n in names)

    base = synthetic_config(256, 4.0, device="cpu")
    props = base["state"]["agents"]["citizens"]["properties"]
    for k in ("is_quarantined", "quarantine_start_date", "awaiting_test_result", "test_result_date", "test_re_eligble_date"):
        assert k in props                                               # reference config.yaml:79-203
    assert base["simulation_metadata"]["quarantine_days"] == 12 and base["simulation_metadata"]["RECOVERED_TO_SUSCEPTIBLE_TIME"] == 100
    cfg = add_intervention_substeps(base)
    assert list(cfg["substeps"].keys()) == ["0", "1", "2", "3"] and cfg["simulation_metadata"]["num_substeps_per_step"] == 4
    assert list(base["substeps"].keys()) == ["0", "1"]                  # the input config is not modified
    assert list(cfg["substeps"]["2"]["policy"]["citizens"].keys()) == ["start_compliance", "break_compliance"]
    only_q = add_intervention_substeps(base, testing=False)
    assert list(only_q["substeps"].keys()) == ["0", "1", "2"]
    sm = use_seirms_progression(base)
    assert list(sm["substeps"]["1"]["transition"].keys()) == ["seirms_progression"]
    assert sm["substeps"]["1"]["transition"]["seirms_progression"]["arguments"]["M"]["learnable"] is True


def test_runner_builds_on_cpu_and_substeps_refuse_cpu_tensors():
    from agenttorch_b200.core.runner import Runner
    from agenttorch_b200.models import covid
    from agenttorch_b200.models.covid.interventions import add_intervention_substeps, use_seirms_progression
    from agenttorch_b200.models.covid.synthetic import synthetic_config

    cfg = use_seirms_progression(add_intervention_substeps(synthetic_config(256, 4.0, device="cpu")))
    runner = Runner(cfg, covid.get_registry())
    runner.init()
    cit = runner.state["agents"]["citizens"]
    assert cit["is_quarantined"].dtype == torch.float32 and float(cit["quarantine_start_date"][0]) == 130.0
    pg = runner.initializer.transition_function["1"]["seirms_progression"]
    assert pg.external_M.requires_grad and float(pg.external_M) == pytest.approx(0.12)
    state = dict(runner.state, current_step=0)
    obs = runner.initializer.observation_function["2"]["citizens"]["get_from_state"](state)
    assert set(obs) == {"quarantine_start_prob", "quarantine_break_prob", "is_quarantined", "disease_stage"}
    for module, args in ((runner.initializer.policy_function["2"]["citizens"]["start_compliance"], (state, obs)),
                         (runner.initializer.policy_function["3"]["citizens"]["accept_test"], (state, None)),
                         (pg, (state, None))):
        with pytest.raises(RuntimeError, match="CUDA"):
            module(*args)


@pytest.mark.parametrize("seed", [0, 1])
def test_seirms_backward_formula_equals_autograd_of_the_oracle(seed):
    """One SEIRMS progression call: (d/d stage, d/d M) of a random cotangent, closed form vs torch autograd of the port."""
    from agenttorch_b200 import _native_iv as nat
    from agenttorch_b200.interventions import seirms_backward_formula
    from oracle.interventions_port import SeirmsPort
    from oracle.seirm_port import SeirmState

    rng = np.random.default_rng(seed)
    n, t, T = 400, 7, 21
    stage0 = rng.choice([0.0, 1.0, 2.0, 3.0, 4.0, 0.37], size=n, p=[0.2, 0.2, 0.3, 0.15, 0.05, 0.1])
    nst = rng.choice([3.0, 6.0, 7.0, 9.0, 41.0], size=n).astype(np.float32)
    port = SeirmsPort(np.zeros(n), stage0, np.array([[0, 1]]))
    stage = torch.tensor(stage0, dtype=torch.float64, requires_grad=True)
    m = torch.tensor([0.4], dtype=torch.float32, requires_grad=True)
    u = onoise.stream_uniforms(5, t, np.arange(n), onoise.STREAM_MORTALITY)[0]
    st = SeirmState(stage, torch.from_numpy(nst), torch.zeros(n), torch.zeros(T), torch.zeros(T))
    out = port.progression_seirms(st, t, m, u)
    g_stage = torch.from_numpy(rng.standard_normal(n))
    g_dd = torch.from_numpy(rng.standard_normal(T).astype(np.float32))
    ((out.stage * g_stage).sum() + (out.daily_deaths * g_dd).sum()).backward()
    # move codes as the kernel writes them
    s_np, due = stage0, nst <= t
    died = u < np.float32(0.4)
    code = np.zeros(n, dtype=np.uint8)
    code[(s_np == 1) & due] = nat.MOVE_E_TO_I
    code[(s_np == 3) & due] = nat.MOVE_R_TO_S
    code[(s_np == 2) & due & ~died] = nat.MOVE_SURVIVED
    code[(s_np == 2) & due & died] = nat.MOVE_DIED
    counts = torch.tensor([int(((s_np == 2) & due).sum()), 0, 0, 0])
    gs, gm = seirms_backward_formula(torch.from_numpy(code), counts, m.detach(), t, g_stage, g_dd, stage.shape)
    assert torch.allclose(gs, stage.grad, rtol=1e-6, atol=1e-6)
    assert torch.allclose(gm.to(torch.float64), m.grad.to(torch.float64).reshape(()), rtol=1e-5, atol=1e-5)
    # the new stages the codes stand for
    expect = s_np.copy()
    expect[code == nat.MOVE_E_TO_I], expect[code == nat.MOVE_R_TO_S] = 2, 0
    expect[code == nat.MOVE_SURVIVED], expect[code == nat.MOVE_DIED] = 4, 3
    assert np.array_equal(out.stage.detach().numpy(), expect)


def test_scalar_cache_follows_replacement_and_in_place_edits():
    from agenttorch_b200.models.covid.substeps.common import EpisodeNoise, ScalarCache

    cache = ScalarCache()
    p = torch.tensor([0.25])
    assert cache("p", p) == 0.25 and cache("p", p) == 0.25
    p.fill_(0.5)                                   # in-place edit bumps the version
    assert cache("p", p) == 0.5
    assert cache("p", torch.tensor([0.75])) == 0.75 and cache("p", 0.1) == 0.1

    class N(EpisodeNoise):
        pass

    n = N()
    n.init_noise({"noise_seed": 5})
    s1 = n.episode_seed(0)
    assert n.episode_seed(3) == s1 and n.episode_seed(0) != s1          # one stream per episode
    n.advance_seed = False
    assert n.episode_seed(0) == 5 and -(1 << 63) <= s1 < (1 << 63)
