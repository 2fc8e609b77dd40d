"""The DEVICE code of csrc/at_interventions.cu, compiled for the host with a small shim (tests/host_emulation/), against
the reference's outputs and the oracle: the kernel source itself is checked on a machine without a GPU.  (The -m gpu
tests run the same comparisons on the real library.)"""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from agenttorch_b200 import _native_iv as nat
from oracle import noise as onoise
from oracle.interventions_port import SeirmsPort
from tests.helpers import INTERVENTION_CASES, SEIRMS_CASES, load_interventions_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emu") / "iv_emulation.so")
    cmd = ["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "host_emulation", "iv_emulation.cpp"), "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    lib = ctypes.CDLL(out)
    for name, struct in (("emu_quarantine", nat.QuarantineArgs), ("emu_testing", nat.TestingArgs), ("emu_seirms", nat.SeirmsArgs)):
        getattr(lib, name).argtypes = [ctypes.POINTER(struct), ctypes.c_int, ctypes.c_int]
        getattr(lib, name).restype = None
    return lib


def aligned(n, dtype, fill=0, offset=0):
    """(n,) array whose first element sits `offset` elements after a 16-byte boundary."""
    item = np.dtype(dtype).itemsize
    raw = np.zeros((n + offset) * item + 64, dtype=np.uint8)
    start = (-raw.ctypes.data) % 16
    arr = raw[start:start + (n + offset) * item].view(dtype)[offset:]
    arr[:] = fill
    return arr


def ptr(a):
    return None if a is None else a.ctypes.data


def make_quarantine_args(g, t, flags, stage, q_in, d_in, q_out, d_out, start=None, brk=None, us=None, ub=None):
    ov = g["overrides"]
    return nat.QuarantineArgs(
        abi_version=nat.ABI_VERSION, flags=flags, t=t, stage_f64=int(stage.dtype == np.float64), n=len(q_in), id_offset=0,
        seed=g["seed"], seed_ptr=None, p_start=ov["env"]["quarantine_start_prob"], p_break=ov["env"]["quarantine_break_prob"],
        quarantine_days=ov["meta"]["quarantine_days"], infinity_time=ov["meta"]["INFINITY_TIME"], stage=ptr(stage), u_start=ptr(us),
        u_break=ptr(ub), is_q_in=ptr(q_in), start_date_in=ptr(d_in), start_action=ptr(start), break_action=ptr(brk),
        is_q_out=ptr(q_out), start_date_out=ptr(d_out))


def make_testing_args(g, t, flags, is_q, aw, rd, re, aw_o, rd_o, re_o, accept=None, uc=None):
    ov = g["overrides"]
    return nat.TestingArgs(
        abi_version=nat.ABI_VERSION, flags=flags, t=t, reserved=0, n=len(aw), id_offset=0, seed=g["seed"], seed_ptr=None,
        p_comply=ov["env"]["test_compliance_prob"], ineligible_days=ov["meta"]["test_ineligible_days"],
        result_delay_days=ov["meta"]["test_result_delay_days"], infinity_time=ov["meta"]["INFINITY_TIME"], u_comply=ptr(uc),
        is_q=ptr(is_q), accept_action=ptr(accept), awaiting_in=ptr(aw), result_date_in=ptr(rd), re_eligible_in=ptr(re),
        awaiting_out=ptr(aw_o), result_date_out=ptr(rd_o), re_eligible_out=ptr(re_o))


@pytest.mark.parametrize("name", INTERVENTION_CASES)
@pytest.mark.parametrize("variant", ["fused_inplace_vec", "split_scalar_f64stage", "split_injected_unaligned"])
def test_quarantine_and_testing_kernels_match_reference(emu, name, variant):
    g = load_interventions_golden(name)
    n, T = len(g["age"]), g["num_steps"]
    ov = g["overrides"]
    ini, inf_t = ov["initial"], ov["meta"]["INFINITY_TIME"]
    off = 1 if variant == "split_injected_unaligned" else 0
    vec = int(variant == "fused_inplace_vec")
    f = np.float32
    q = aligned(n, f, ini.get("is_quarantined", 0), off)
    d = aligned(n, f, ini.get("quarantine_start_date", inf_t), off)
    aw = aligned(n, f, ini.get("awaiting_test_result", 0), off)
    rd = aligned(n, f, ini.get("test_result_date", inf_t), off)
    re = aligned(n, f, ini.get("test_re_eligble_date", inf_t), off)
    ids = np.arange(n)
    for t in range(T):
        stage32 = aligned(n, f, 0, off)
        stage32[:] = g["ref_disease_stage_history"][t]
        if variant == "fused_inplace_vec":
            emu.emu_quarantine(make_quarantine_args(g, t, nat.DECIDE | nat.APPLY, stage32, q, d, q, d), vec, 3)
            emu.emu_testing(make_testing_args(g, t, nat.DECIDE | nat.APPLY, q, aw, rd, re, aw, rd, re), vec, 2)
        else:
            stage = stage32.astype(np.float64) if variant == "split_scalar_f64stage" else stage32
            us = ub = uc = None
            if variant == "split_injected_unaligned":
                u0, u1 = onoise.stream_uniforms(g["seed"], t, ids, onoise.STREAM_QUARANTINE)
                us, ub = aligned(n, f, 0, off), aligned(n, f, 0, off)
                us[:], ub[:] = u0, u1
                uc = aligned(n, f, 0, off)
                uc[:] = onoise.stream_uniforms(g["seed"], t, ids, onoise.STREAM_TESTING)[0]
            start, brk = aligned(n, np.uint8, 7, off), aligned(n, np.uint8, 7, off)
            # the two policies are separate launches, each writing its own action
            emu.emu_quarantine(make_quarantine_args(g, t, nat.DECIDE, stage, q, None, None, None, start=start, us=us, ub=ub), vec, 2)
            emu.emu_quarantine(make_quarantine_args(g, t, nat.DECIDE, stage, q, None, None, None, brk=brk, us=us, ub=ub), vec, 2)
            assert set(np.unique(start)) <= {0, 1} and set(np.unique(brk)) <= {0, 1}
            q2, d2 = aligned(n, f, -77, off), aligned(n, f, -77, off)
            emu.emu_quarantine(make_quarantine_args(g, t, nat.APPLY, stage, q, d, q2, d2, start=start, brk=brk), vec, 5)
            q, d = q2, d2
            accept = aligned(n, f, -5, off)
            emu.emu_testing(make_testing_args(g, t, nat.DECIDE, q, aw, None, None, None, None, None, accept=accept, uc=uc), vec, 1)
            o = [aligned(n, f, -77, off) for _ in range(3)]
            emu.emu_testing(make_testing_args(g, t, nat.APPLY, None, aw, rd, re, *o, accept=accept), vec, 4)
            aw, rd, re = o
        for key, arr in (("is_quarantined", q), ("quarantine_start_date", d), ("awaiting_test_result", aw), ("test_result_date", rd),
                         ("test_re_eligble_date", re)):
            assert np.array_equal(arr.astype(np.float64), g["ref_" + key + "_history"][t].astype(np.float64)), (key, t)


@pytest.mark.parametrize("name", SEIRMS_CASES)
@pytest.mark.parametrize("variant", ["inplace_vec", "scalar_injected_unaligned"])
def test_seirms_kernel_matches_oracle_and_reference(emu, name, variant):
    g = load_interventions_golden(name)
    n, T = len(g["age"]), g["num_steps"]
    meta = g["overrides"]["meta"]
    port = SeirmsPort(g["age"], g["stage0"], g["edges"], recovered_to_susceptible_time=meta["RECOVERED_TO_SUSCEPTIBLE_TIME"])
    r2, m = torch.tensor(g["r2"]), torch.tensor(g["m"])
    st = port.initial_state()
    ids = np.arange(n)
    off = 0 if variant == "inplace_vec" else 3
    f = np.float32
    dd = aligned(T, f)
    m_dev = aligned(1, f, float(g["m"][0]))
    for t in range(T):
        w = torch.from_numpy(onoise.isolation_decision(g["seed"], t, ids, g["mode"], float(g["p_isolate"])))
        u = torch.from_numpy(onoise.exposure_uniform(g["seed"], t, ids))
        st, _ = port.transmission(st, t, w, u, r2)
        s_in, nt_in = aligned(n, f, 0, off), aligned(n, f, 0, off)
        s_in[:], nt_in[:] = st.stage.numpy(), st.next_stage_time.numpy()
        ud = None
        if off:
            ud = aligned(n, f, 0, off)
            ud[:] = onoise.stream_uniforms(g["seed"], t, ids, onoise.STREAM_MORTALITY)[0]
            s_out, nt_out = aligned(n, f, -9, off), aligned(n, f, -9, off)
        else:
            s_out, nt_out = s_in, nt_in
        code = aligned(n, np.uint8, 9, off)
        counts = aligned(4, np.int64)
        args = nat.SeirmsArgs(
            abi_version=nat.ABI_VERSION, t=t, n=n, id_offset=0, seed=g["seed"], seed_ptr=None, infected_to_recovered_time=5.0,
            infinity_time=T + 20.0, recovered_to_susceptible_time=meta["RECOVERED_TO_SUSCEPTIBLE_TIME"], num_steps=T, m_value=ptr(m_dev),
            u_death=ptr(ud), stage_in=ptr(s_in), next_stage_time_in=ptr(nt_in), stage_out=ptr(s_out), next_stage_time_out=ptr(nt_out),
            move_code=ptr(code), counts=ptr(counts), daily_deaths=ptr(dd))
        before = st.stage.numpy().copy()
        emu.emu_seirms(args, int(off == 0), 3)
        st = port.progression_seirms(st, t, m, onoise.stream_uniforms(g["seed"], t, ids, onoise.STREAM_MORTALITY)[0])
        assert np.array_equal(s_out.astype(np.float64), st.stage.numpy()), t
        assert np.array_equal(nt_out, st.next_stage_time.numpy()), t
        assert np.array_equal(s_out.astype(np.float64), g["ref_disease_stage_history"][t].astype(np.float64)), t
        assert np.array_equal(nt_out.astype(np.float64), g["ref_next_stage_time_history"][t].astype(np.float64)), t
        due = (before == 2) & (code >= nat.MOVE_SURVIVED)
        assert counts[0] == due.sum() == counts[1] + counts[2] and counts[1] == (code == nat.MOVE_DIED).sum()
        assert np.array_equal(code == nat.MOVE_E_TO_I, (before == 1) & (s_out == 2))
        assert np.array_equal(code == nat.MOVE_R_TO_S, (before == 3) & (s_out == 0))
    assert np.allclose(dd, g["ref_daily_deaths"], rtol=1e-6, atol=0)
    assert np.array_equal(dd, st.daily_deaths.numpy())


@pytest.mark.parametrize("n", [1, 2, 3, 5, 63, 1025])
@pytest.mark.parametrize("off", [0, 1])
def test_ragged_sizes_and_every_status_combination_against_the_oracle(emu, n, off):
    """Tail groups (n not a multiple of 4), unaligned arrays, negative / fractional statuses, dates at and around t."""
    from oracle.interventions_port import quarantine_step, testing_step as oracle_testing_step

    rng = np.random.default_rng(1000 * n + off)
    f = np.float32
    g = {"seed": 77, "overrides": {"env": {"quarantine_start_prob": 0.6, "quarantine_break_prob": 0.35, "test_compliance_prob": 0.7},
                                   "meta": {"quarantine_days": 2, "INFINITY_TIME": 130, "test_ineligible_days": 1,
                                            "test_result_delay_days": 2}}}
    stage = aligned(n, f, 0, off)
    stage[:] = rng.choice([0, 1, 2, 3, 4, 0.5], size=n)
    q, d = aligned(n, f, 0, off), aligned(n, f, 0, off)
    q[:] = rng.choice([-2, -1, 0, 1, 0.5], size=n)
    d[:] = rng.choice([0, 1, 2, 3, 130], size=n)
    aw, rd, re = aligned(n, f, 0, off), aligned(n, f, 0, off), aligned(n, f, 0, off)
    aw[:], rd[:], re[:] = rng.choice([0, 1, -1], size=n), rng.choice([3, 4, 5, 130], size=n), rng.choice([0, 3, 4, 5, 6], size=n)
    ids = np.arange(n)
    for t in (3, 4, 5):
        us, ub = onoise.stream_uniforms(77, t, ids, onoise.STREAM_QUARANTINE)
        uc = onoise.stream_uniforms(77, t, ids, onoise.STREAM_TESTING)[0]
        eq, ed = quarantine_step(t, stage, q, d, us, ub, 0.6, 0.35, 2, 130)
        ea, er, ee = oracle_testing_step(t, eq, aw, rd, re, uc, 0.7, 1, 2, 130)
        emu.emu_quarantine(make_quarantine_args(g, t, nat.DECIDE | nat.APPLY, stage, q, d, q, d), int(off == 0), 2)
        emu.emu_testing(make_testing_args(g, t, nat.DECIDE | nat.APPLY, q, aw, rd, re, aw, rd, re), int(off == 0), 3)
        for got, want in ((q, eq), (d, ed), (aw, ea), (rd, er), (re, ee)):
            assert np.array_equal(got, want), t


# ------------------------------------------------------------------------------------------------------------------
# The torch custom ops (agenttorch_b200/interventions.py) and the substep classes against the same host-compiled source,
# through its C-ABI wrappers: argument structs, flags, action dtypes and the registered autograd on a machine without a GPU.
@pytest.fixture()
def ops_on_host(emu, monkeypatch):
    from agenttorch_b200 import interventions as ops

    monkeypatch.setattr(nat, "LIB_PATH", emu._name)
    monkeypatch.setattr(nat, "_lib", None)
    monkeypatch.setattr(ops, "_require_cuda", lambda name, t: None)
    monkeypatch.setattr(ops, "_stream", lambda t: None)
    yield ops
    monkeypatch.setattr(nat, "_lib", None)


@pytest.mark.parametrize("name", INTERVENTION_CASES)
@pytest.mark.parametrize("variant", ["fused_inplace", "split"])
def test_custom_ops_match_reference_histories(ops_on_host, name, variant):
    ops = ops_on_host
    g = load_interventions_golden(name)
    n, T = len(g["age"]), g["num_steps"]
    meta, env, ini = g["overrides"]["meta"], g["overrides"]["env"], g["overrides"]["initial"]
    inf_t = float(meta["INFINITY_TIME"])
    default = {"is_quarantined": 0, "quarantine_start_date": inf_t, "awaiting_test_result": 0, "test_result_date": inf_t,
               "test_re_eligble_date": inf_t}
    s = {k: torch.full((n, 1), float(ini.get(k, v)), dtype=torch.float32) for k, v in default.items()}
    for t in range(T):
        stage = torch.from_numpy(g["ref_disease_stage_history"][t].astype(np.float64)).reshape(n, 1)
        if variant == "fused_inplace":
            ops.quarantine_step(stage.float(), s["is_quarantined"], s["quarantine_start_date"], env["quarantine_start_prob"],
                                env["quarantine_break_prob"], t, g["seed"], 0, float(meta["quarantine_days"]), inf_t)
            ops.testing_step(s["is_quarantined"], s["awaiting_test_result"], s["test_result_date"], s["test_re_eligble_date"],
                             env["test_compliance_prob"], t, g["seed"], 0, float(meta["test_ineligible_days"]),
                             float(meta["test_result_delay_days"]), inf_t)
        else:
            start, _ = ops.quarantine_decide(stage, s["is_quarantined"], env["quarantine_start_prob"], 0.0, t, g["seed"], 0, None, None)
            _, brk = ops.quarantine_decide(stage, s["is_quarantined"], 0.0, env["quarantine_break_prob"], t, g["seed"], 0, None, None)
            assert start.dtype == torch.bool and start.shape == (n, 1)
            s["is_quarantined"], s["quarantine_start_date"] = ops.quarantine_update(
                s["is_quarantined"], s["quarantine_start_date"], start, brk.long(), t, float(meta["quarantine_days"]), inf_t)
            accept = ops.testing_accept(s["is_quarantined"], env["test_compliance_prob"], t, g["seed"], 0, None)
            s["awaiting_test_result"], s["test_result_date"], s["test_re_eligble_date"] = ops.testing_update(
                s["awaiting_test_result"], s["test_result_date"], s["test_re_eligble_date"], accept, t,
                float(meta["test_ineligible_days"]), float(meta["test_result_delay_days"]), inf_t)
        for k in default:
            assert np.array_equal(s[k].numpy().reshape(-1).astype(np.float64), g["ref_" + k + "_history"][t].astype(np.float64)), (k, t)


@pytest.mark.parametrize("name", SEIRMS_CASES)
def test_seirms_op_and_its_autograd_against_the_oracle(ops_on_host, name):
    """Step by step: the op (kernel + epilogue + registered backward) against torch autograd of the oracle's progression,
    for a random cotangent on (stage, daily_deaths)."""
    ops = ops_on_host
    g = load_interventions_golden(name)
    n, T = len(g["age"]), g["num_steps"]
    meta = g["overrides"]["meta"]
    port = SeirmsPort(g["age"], g["stage0"], g["edges"], recovered_to_susceptible_time=meta["RECOVERED_TO_SUSCEPTIBLE_TIME"])
    r2 = torch.tensor(g["r2"])
    st = port.initial_state()
    ids = np.arange(n)
    rng = np.random.default_rng(5)
    dd_op = torch.zeros(T)
    for t in range(T):
        w = torch.from_numpy(onoise.isolation_decision(g["seed"], t, ids, g["mode"], float(g["p_isolate"])))
        u = torch.from_numpy(onoise.exposure_uniform(g["seed"], t, ids))
        st, _ = port.transmission(st, t, w, u, r2)
        st.stage = st.stage.detach().requires_grad_(True)
        m_ref = torch.tensor(g["m"], requires_grad=True)
        out = port.progression_seirms(st, t, m_ref, onoise.stream_uniforms(g["seed"], t, ids, onoise.STREAM_MORTALITY)[0])
        stage_in = st.stage.detach().float().reshape(n, 1).requires_grad_(True)
        m_op = torch.tensor(g["m"], requires_grad=True)
        s2, n2, dd2, code, counts = ops.seirms_progression(stage_in, st.next_stage_time.reshape(n, 1), dd_op, m_op, t, g["seed"], 0, None,
                                                           5.0, T + 20.0, float(meta["RECOVERED_TO_SUSCEPTIBLE_TIME"]))
        assert np.array_equal(s2.detach().numpy().reshape(-1).astype(np.float64), out.stage.detach().numpy())
        assert np.array_equal(n2.numpy().reshape(-1), out.next_stage_time.numpy())
        assert torch.equal(dd2.detach(), out.daily_deaths.detach())
        g_stage = torch.from_numpy(rng.standard_normal(n).astype(np.float32))
        g_dd = torch.from_numpy(rng.standard_normal(T).astype(np.float32))
        ((out.stage * g_stage.double()).sum() + (out.daily_deaths * g_dd).sum()).backward()
        ((s2.reshape(-1) * g_stage).sum() + (dd2 * g_dd).sum()).backward()
        assert torch.allclose(stage_in.grad.reshape(-1).double(), st.stage.grad, rtol=1e-5, atol=1e-6), t
        assert torch.allclose(m_op.grad, m_ref.grad, rtol=1e-4, atol=1e-4), t
        dd_op = dd2.detach()
        st = type(st)(out.stage.detach(), out.next_stage_time, out.infected_time, out.daily_infected.detach(), out.daily_deaths.detach())


def test_substep_classes_on_host(ops_on_host, monkeypatch):
    """StartCompliance / BreakCompliance / UpdateQuarantineStatus / AcceptTest / UpdateTestStatus through the Controller."""
    from agenttorch_b200.core.runner import Runner
    from agenttorch_b200.models import covid
    from agenttorch_b200.models.covid.interventions import add_intervention_substeps
    from agenttorch_b200.models.covid.substeps.quarantine import action as q_action, transition as q_transition
    from agenttorch_b200.models.covid.substeps.testing import action as t_action, transition as t_transition
    from agenttorch_b200.models.covid.synthetic import synthetic_config
    from oracle.interventions_port import quarantine_step, testing_step as oracle_testing_step

    cfg = add_intervention_substeps(synthetic_config(500, 4.0, device="cpu", noise_seed=31))
    cfg["state"]["agents"]["citizens"]["properties"]["test_re_eligble_date"]["value"] = 0
    runner = Runner(cfg, covid.get_registry())
    runner.init()
    for module in (q_action, q_transition, t_action, t_transition):
        monkeypatch.setattr(module, "require_cuda", lambda t, what: None)
    state = runner.state
    cit = state["agents"]["citizens"]
    cit["disease_stage"] = torch.from_numpy(np.random.default_rng(2).choice([0.0, 1.0, 2.0, 3.0], size=(500, 1)))
    pf, tf, of = runner.initializer.policy_function, runner.initializer.transition_function, runner.initializer.observation_function
    for mod in (pf["2"]["citizens"]["start_compliance"], pf["2"]["citizens"]["break_compliance"], pf["3"]["citizens"]["accept_test"]):
        mod.advance_seed = False
    ids = np.arange(500)
    q, d = np.zeros(500, np.float32), np.full(500, 130, np.float32)
    aw, rd, re = np.zeros(500, np.float32), np.full(500, 130, np.float32), np.zeros(500, np.float32)
    for t in range(6):
        state["current_step"] = t
        for sub in ("2", "3"):
            state["current_substep"] = sub
            obs = runner.controller.observe(state, of, "citizens")
            act = {"citizens": runner.controller.act(state, obs, pf, "citizens")}
            state = runner.controller.progress(state, act, tf)
        us, ub = onoise.stream_uniforms(31, t, ids, onoise.STREAM_QUARANTINE)
        q, d = quarantine_step(t, cit["disease_stage"].numpy().reshape(-1), q, d, us, ub, 0.7, 0.1, 12, 130)
        aw, rd, re = oracle_testing_step(t, q, aw, rd, re, onoise.stream_uniforms(31, t, ids, onoise.STREAM_TESTING)[0], 0.95, 2, 3, 130)
        c = state["agents"]["citizens"]
        for key, want in (("is_quarantined", q), ("quarantine_start_date", d), ("awaiting_test_result", aw), ("test_result_date", rd),
                          ("test_re_eligble_date", re)):
            assert np.array_equal(c[key].numpy().reshape(-1), want), (key, t)
    assert (q != 0).any() and (aw != 0).any()
