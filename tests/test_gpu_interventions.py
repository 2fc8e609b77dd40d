"""GPU parity of the quarantine / testing / SEIRMS-progression path (libat_interventions.so through the torch custom ops
and through the Runner's substep API) against the real reference's outputs (tests/golden/interventions_*.npz,
seirms_*.npz) and the oracle, plus size-independent properties at millions of agents."""
import numpy as np
import pytest
import torch

from oracle import noise as onoise
from oracle.interventions_port import quarantine_step as oracle_quarantine_step
from oracle.interventions_port import testing_step as oracle_testing_step
from tests.helpers import INTERVENTION_CASES, SEIRMS_CASES, load_interventions_golden, make_noise, rel_err

pytestmark = pytest.mark.gpu

STATE_KEYS = ("is_quarantined", "quarantine_start_date", "awaiting_test_result", "test_result_date", "test_re_eligble_date")


def _initial(g, dev, n):
    ini, inf_t = g["overrides"]["initial"], g["overrides"]["meta"]["INFINITY_TIME"]
    default = {"is_quarantined": 0, "quarantine_start_date": inf_t, "awaiting_test_result": 0, "test_result_date": inf_t,
               "test_re_eligble_date": inf_t}
    return {k: torch.full((n, 1), float(ini.get(k, default[k])), dtype=torch.float32, device=dev) for k in STATE_KEYS}


# ------------------------------------------------------------------------------------------- custom ops
@pytest.mark.parametrize("name", INTERVENTION_CASES)
@pytest.mark.parametrize("variant", ["fused_inplace", "split_f64stage", "split_injected"])
def test_ops_match_reference_histories(name, variant):
    from agenttorch_b200 import interventions as ops

    g = load_interventions_golden(name)
    dev = torch.device("cuda", 0)
    n, T = len(g["age"]), g["num_steps"]
    meta, env = g["overrides"]["meta"], g["overrides"]["env"]
    s = _initial(g, dev, n)
    ids = np.arange(n)
    for t in range(T):
        stage = torch.from_numpy(g["ref_disease_stage_history"][t].astype(np.float32)).to(dev).reshape(n, 1)
        if variant == "fused_inplace":
            ops.quarantine_step(stage, s["is_quarantined"], s["quarantine_start_date"], env["quarantine_start_prob"],
                                env["quarantine_break_prob"], t, g["seed"], 0, float(meta["quarantine_days"]), float(meta["INFINITY_TIME"]))
            ops.testing_step(s["is_quarantined"], s["awaiting_test_result"], s["test_result_date"], s["test_re_eligble_date"],
                             env["test_compliance_prob"], t, g["seed"], 0, float(meta["test_ineligible_days"]),
                             float(meta["test_result_delay_days"]), float(meta["INFINITY_TIME"]))
        else:
            us = ub = uc = None
            if variant == "split_injected":
                u0, u1 = onoise.stream_uniforms(g["seed"], t, ids, onoise.STREAM_QUARANTINE)
                us, ub = torch.from_numpy(u0).to(dev), torch.from_numpy(u1).to(dev)
                uc = torch.from_numpy(onoise.stream_uniforms(g["seed"], t, ids, onoise.STREAM_TESTING)[0]).to(dev)
                seed = 12345                                         # must not matter when every draw is injected
            else:
                stage, seed = stage.double(), g["seed"]
            start, _ = ops.quarantine_decide(stage, s["is_quarantined"], env["quarantine_start_prob"], 0.0, t, seed, 0, us, None)
            _, brk = ops.quarantine_decide(stage, s["is_quarantined"], 0.0, env["quarantine_break_prob"], t, seed, 0, None, ub)
            assert start.dtype == torch.bool and start.shape == (n, 1)
            s["is_quarantined"], s["quarantine_start_date"] = ops.quarantine_update(
                s["is_quarantined"], s["quarantine_start_date"], start, brk.float(), t, float(meta["quarantine_days"]),
                float(meta["INFINITY_TIME"]))
            accept = ops.testing_accept(s["is_quarantined"], env["test_compliance_prob"], t, seed, 0, uc)
            s["awaiting_test_result"], s["test_result_date"], s["test_re_eligble_date"] = ops.testing_update(
                s["awaiting_test_result"], s["test_result_date"], s["test_re_eligble_date"], accept, t,
                float(meta["test_ineligible_days"]), float(meta["test_result_delay_days"]), float(meta["INFINITY_TIME"]))
        for k in STATE_KEYS:
            got = s[k].cpu().numpy().reshape(-1).astype(np.float64)
            assert np.array_equal(got, g["ref_" + k + "_history"][t].astype(np.float64)), (k, t)


def test_ops_refuse_cpu_tensors():
    from agenttorch_b200 import interventions as ops

    z = torch.zeros(8, 1)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.quarantine_decide(z, z, 0.5, 0.5, 0, 1, 0, None, None)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.testing_accept(z, 0.5, 0, 1, 0, None)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.seirms_progression(z, z, torch.zeros(21), torch.tensor([0.1]), 0, 1, 0, None, 5.0, 41.0, 100.0)


# ---------------------------------------------------------------------------------- through the Runner
def _runner(g, tmp_path, seirms=False, interventions=False):
    from agenttorch_b200.core.dataloader import DataLoader, LoadPopulation
    from agenttorch_b200.core.runner import Runner
    from agenttorch_b200.models import covid
    from agenttorch_b200.models.covid.interventions import add_intervention_substeps, use_seirms_progression
    from agenttorch_b200.populations import write_population_folder

    folder = write_population_folder(str(tmp_path / "pop"), g["age"], g["stage0"], g["edges"])
    loader = DataLoader(covid, LoadPopulation(folder))
    ov = g["overrides"]
    for k, v in ov["meta"].items():
        loader.set_config_attribute(k, v)
    for k, v in ov["initial"].items():
        loader.config["state"]["agents"]["citizens"]["properties"][k]["value"] = v
    for k, v in ov["env"].items():
        loader.config["state"]["environment"][k]["value"] = v
    cfg = loader.get_config()
    if interventions:
        cfg = add_intervention_substeps(cfg)
    if seirms:
        cfg = use_seirms_progression(cfg)
    runner = Runner(cfg, covid.get_registry())
    runner.init()
    dev = runner.initializer.device
    assert dev.type == "cuda"
    tf, pf = runner.initializer.transition_function, runner.initializer.policy_function
    tr = tf["0"]["new_transmission"]
    pg = tf["1"]["seirms_progression" if seirms else "seirm_progression"]
    r2 = torch.tensor(g["r2"], dtype=torch.float32, device=dev, requires_grad=True)
    params = {"initializer.transition_function.0.new_transmission.learnable_args.R2": r2}
    if seirms:
        pg.external_M = torch.tensor(g["m"], dtype=torch.float32, device=dev, requires_grad=True)
        m = pg.external_M
    else:
        m = torch.tensor(g["m"], dtype=torch.float32, device=dev, requires_grad=True)
        params["initializer.transition_function.1.seirm_progression.learnable_args.M"] = m
    runner._set_parameters(params)
    T = g["num_steps"]
    U, W = make_noise(g["seed"], g["mode"], len(g["age"]), T, float(g["p_isolate"]))
    Ut, Wt = torch.from_numpy(U).to(dev), torch.from_numpy(W).to(dev)
    pol = pf["0"]["citizens"]["make_isolation_decision"]
    pol.mode = "injected"
    pol.behavior = lambda obs: Wt[int(runner.state["current_step"])].reshape(-1, 1)
    tr.exposure_uniform = lambda t: Ut[t]
    noisy = [pg] if seirms else []
    if interventions:
        noisy += [pf["2"]["citizens"]["start_compliance"], pf["2"]["citizens"]["break_compliance"], pf["3"]["citizens"]["accept_test"]]
    for mod in noisy:
        mod.noise_seed, mod.advance_seed = g["seed"], False
    return runner, r2, m


@pytest.mark.parametrize("name", INTERVENTION_CASES)
def test_runner_with_quarantine_and_testing_substeps(name, tmp_path):
    g = load_interventions_golden(name)
    runner, r2, m = _runner(g, tmp_path, interventions=True)
    history = {k: [] for k in STATE_KEYS + ("disease_stage", "next_stage_time")}

    def record(module, args, kwargs, output):                           # last transition of every step
        cit = kwargs["state"]["agents"]["citizens"]
        for k in history:
            v = output.get(k, cit[k])
            history[k].append(v.detach().cpu().numpy().reshape(-1).astype(np.float64))

    runner.initializer.transition_function["3"]["update_test_status"].register_forward_hook(record, with_kwargs=True)
    runner.step(g["num_steps"])
    assert len(runner.state_trajectory) == g["num_steps"] + 1 and len(runner.state_trajectory[-1]) == 4
    for k, rows in history.items():
        assert np.array_equal(np.stack(rows), g["ref_" + k + "_history"].astype(np.float64)), k
    cit = runner.state["agents"]["citizens"]
    for k in STATE_KEYS:
        assert cit[k].dtype == torch.float32 and cit[k].shape == (len(g["age"]), 1)
    env = runner.state["environment"]
    assert rel_err(g["ref_daily_infected"], env["daily_infected"].detach().cpu().numpy()) <= 1e-5
    assert rel_err(g["ref_daily_deaths"], env["daily_deaths"].detach().cpu().numpy()) <= 1e-5
    dev = env["daily_infected"].device
    loss = (torch.from_numpy(g["gI"]).to(dev) * env["daily_infected"].reshape(-1)).sum() + \
           (torch.from_numpy(g["gD"]).to(dev) * env["daily_deaths"].reshape(-1)).sum()
    loss.backward()
    assert rel_err(g["ref_grad_R2"], r2.grad.cpu().numpy()) <= 1e-4
    assert rel_err(g["ref_grad_M"], m.grad.cpu().numpy()) <= 1e-4


@pytest.mark.parametrize("name", SEIRMS_CASES)
def test_runner_with_seirms_progression(name, tmp_path):
    g = load_interventions_golden(name)
    runner, r2, m = _runner(g, tmp_path, seirms=True)
    history = {"disease_stage": [], "next_stage_time": []}

    def record(module, args, kwargs, output):
        for k in history:
            history[k].append(output[k].detach().cpu().numpy().reshape(-1).astype(np.float64))

    runner.initializer.transition_function["1"]["seirms_progression"].register_forward_hook(record, with_kwargs=True)
    runner.step(g["num_steps"])
    assert runner.perf_stats["fused_episodes"] == 0
    for k, rows in history.items():
        assert np.array_equal(np.stack(rows), g["ref_" + k + "_history"].astype(np.float64)), k
    env = runner.state["environment"]
    di, dd = env["daily_infected"], env["daily_deaths"]
    assert tuple(dd.shape) == (1, g["num_steps"])
    assert rel_err(g["ref_daily_infected"], di.detach().cpu().numpy()) <= 1e-5
    assert rel_err(g["ref_daily_deaths"], dd.detach().cpu().numpy()) <= 1e-5
    dev = di.device
    loss = (torch.from_numpy(g["gI"]).to(dev) * di.reshape(-1)).sum() + (torch.from_numpy(g["gD"]).to(dev) * dd.reshape(-1)).sum()
    loss.backward()
    assert abs(float(loss) - float(g["ref_loss"])) <= 1e-5 * abs(float(g["ref_loss"]))
    assert rel_err(g["ref_grad_R2"], r2.grad.cpu().numpy()) <= 1e-4
    assert rel_err(g["ref_grad_M"], m.grad.cpu().numpy()) <= 1e-4


# ------------------------------------------------------------------------- properties at larger sizes
def _random_state(n, dev, seed):
    gen = torch.Generator(device="cpu").manual_seed(seed)
    stage = torch.randint(0, 5, (n,), generator=gen).float()
    is_q = torch.randint(-2, 2, (n,), generator=gen).float()
    sd = torch.where(is_q != 0, torch.randint(0, 6, (n,), generator=gen).float(), torch.full((n,), 130.0))
    aw = torch.randint(0, 2, (n,), generator=gen).float()
    rd = torch.where(aw != 0, torch.randint(3, 8, (n,), generator=gen).float(), torch.full((n,), 130.0))
    re = torch.randint(0, 9, (n,), generator=gen).float()
    return [x.to(dev) for x in (stage, is_q, sd, aw, rd, re)]


@pytest.mark.parametrize("n", [1, 3, 1000003, 4 * 1024 * 1024])
def test_fused_equals_split_equals_oracle_on_random_states(n):
    """Ragged sizes (tail groups), an odd-offset view (scalar path), every status/date combination; three steps."""
    from agenttorch_b200 import interventions as ops

    dev = torch.device("cuda", 0)
    stage, is_q, sd, aw, rd, re = _random_state(n + 1, dev, seed=n)
    views = [x[1:] for x in (stage, is_q, sd, aw, rd, re)]                          # 4-byte offset: not 16-byte aligned
    stage, is_q, sd, aw, rd, re = [v.clone() for v in views]                       # aligned twins
    u_stage, u_q, u_sd, u_aw, u_rd, u_re = views
    seed, ids = 99, np.arange(n)
    for t in (4, 5, 6):
        # oracle on the current state
        us, ub = onoise.stream_uniforms(seed, t, ids, onoise.STREAM_QUARANTINE)
        uc = onoise.stream_uniforms(seed, t, ids, onoise.STREAM_TESTING)[0]
        if n <= 1000003:
            eq, ed = oracle_quarantine_step(t, stage.cpu().numpy(), is_q.cpu().numpy(), sd.cpu().numpy(), us, ub, 0.6, 0.3, 3, 130)
            ea, er, ee = oracle_testing_step(t, eq, aw.cpu().numpy(), rd.cpu().numpy(), re.cpu().numpy(), uc, 0.8, 2, 3, 130)
        # split path on the aligned tensors
        start, brk = ops.quarantine_decide(stage, is_q, 0.6, 0.3, t, seed, 0, None, None)
        q2, d2 = ops.quarantine_update(is_q, sd, start, brk, t, 3.0, 130.0)
        acc = ops.testing_accept(q2, 0.8, t, seed, 0, None)
        a2, r2_, e2 = ops.testing_update(aw, rd, re, acc, t, 2.0, 3.0, 130.0)
        # fused in-place path on the unaligned views
        ops.quarantine_step(u_stage, u_q, u_sd, 0.6, 0.3, t, seed, 0, 3.0, 130.0)
        ops.testing_step(u_q, u_aw, u_rd, u_re, 0.8, t, seed, 0, 2.0, 3.0, 130.0)
        for got, want in ((u_q, q2), (u_sd, d2), (u_aw, a2), (u_rd, r2_), (u_re, e2)):
            assert torch.equal(got, want)
        if n <= 1000003:
            for got, want in ((q2, eq), (d2, ed), (a2, ea), (r2_, er), (e2, ee)):
                assert np.array_equal(got.cpu().numpy(), want)
        # invariants: a start needs an exposed/infected, non-quarantined agent; dates are t, INFINITY or unchanged
        assert not bool((start & ~((stage == 1) | (stage == 2))).any()) and not bool((start & (is_q.trunc() != 0)).any())
        assert not bool((brk & (is_q == 0)).any())
        changed = d2 != sd
        assert bool(((d2[changed] == t) | (d2[changed] == 130.0)).all())
        is_q, sd, aw, rd, re = q2, d2, a2, r2_, e2


@pytest.mark.parametrize("n", [2, 5, 2 * 1024 * 1024 + 1])
def test_seirms_kernel_on_random_states(n):
    from agenttorch_b200 import _native_iv as nat
    from agenttorch_b200 import interventions as ops

    dev = torch.device("cuda", 0)
    gen = torch.Generator().manual_seed(n)
    stage = torch.tensor([0.0, 1.0, 2.0, 3.0, 4.0, 0.25])[torch.randint(0, 6, (n,), generator=gen)].to(dev)
    nst = torch.randint(5, 10, (n,), generator=gen).float().to(dev)
    dd = torch.zeros(21, device=dev)
    m = torch.tensor([0.35], device=dev, requires_grad=True)
    t, seed = 7, 4242
    stage_in = stage.clone().requires_grad_(True)
    s2, n2, dd2, code, counts = ops.seirms_progression(stage_in, nst, dd, m, t, seed, 0, None, 5.0, 41.0, 100.0)
    u = torch.from_numpy(onoise.stream_uniforms(seed, t, np.arange(n), onoise.STREAM_MORTALITY)[0]).to(dev)
    due = nst <= t
    died = (stage == 2) & due & (u < 0.35)
    survived = (stage == 2) & due & ~died
    want = stage.clone()
    want[(stage == 1) & due], want[(stage == 3) & due], want[died], want[survived] = 2.0, 0.0, 3.0, 4.0
    assert torch.equal(s2.detach(), want)
    assert counts.tolist() == [int(died.sum() + survived.sum()), int(died.sum()), int(survived.sum()), 0]
    wn = nst.clone()
    wn[(stage == 1) & (nst == t)] = t + 5.0
    if n > 1:                                                   # quirk I6: rows 0 and 1
        wn[0] = 100.0 if int(died.sum()) < n else 41.0
        if int(died.sum()) > 0:
            wn[1] = 100.0
        elif int(survived.sum()) > 0:
            wn[1] = 41.0
    assert torch.equal(n2, wn)
    assert float(dd2[t]) == pytest.approx(float(counts[0]) * 0.35, rel=1e-6) and float(dd2.sum()) == pytest.approx(float(dd2[t]))
    assert (code == nat.MOVE_DIED).sum() == died.sum() and (code == nat.MOVE_E_TO_I).sum() == ((stage == 1) & due).sum()
    # injected uniforms reproduce the hash draws; autograd: d/dM of daily_deaths[t] = due count, stage tangents per quirk I7
    s3, _, _, code3, _ = ops.seirms_progression(stage, nst, dd, m.detach(), t, 1, 0, u, 5.0, 41.0, 100.0)
    assert torch.equal(s3, s2.detach()) and torch.equal(code3, code)
    (dd2[t] + s2.sum()).backward()
    assert float(m.grad) == pytest.approx(float(counts[0]) - float(counts[0]), abs=1e-3)      # +count (deaths) - count (ST draw)
    lut = torch.tensor([1.0, 2.0, 0.0, 3.0, 2.5], device=dev)
    want_g = lut[code.long()] + 0.5 * 0.35 * (code >= nat.MOVE_SURVIVED).float()
    assert torch.allclose(stage_in.grad, want_g, rtol=1e-6, atol=1e-6)
