"""Generate tests/golden/interventions_*.npz and seirms_*.npz by running the real reference (/root/reference) on CPU.

TEST INFRASTRUCTURE.  Build container only:  python -m oracle.ref_harness.make_golden_interventions
Inputs (age, initial stages, edge list), the noise address (seed, isolation mode), the configuration overrides and
the reference's per-step histories of every agent property the substeps write are stored.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_harness.make_golden import GOLDEN, T, synthetic_population  # noqa: E402
from oracle.ref_harness.run_interventions_reference import run_interventions_reference  # noqa: E402

META_DEFAULT = {"quarantine_days": 12, "test_ineligible_days": 2, "test_result_delay_days": 3, "INFINITY_TIME": 130,
                "RECOVERED_TO_SUSCEPTIBLE_TIME": 100}
ENV_DEFAULT = {"quarantine_start_prob": 0.7, "quarantine_break_prob": 0.1, "test_compliance_prob": 0.95}
SMALL = {"is_quarantined": np.int8, "awaiting_test_result": np.int8}


def save_case(name, variant, age, stage0, edges, seed, mode, r2, m, meta=None, initial=None, env=None, keep_steps=None,
              store_inputs=True):
    gI = np.linspace(0.5, 1.5, T).astype(np.float32)
    gD = np.linspace(2.0, 1.0, T).astype(np.float32)
    meta_all = dict(META_DEFAULT, **(meta or {}))
    env_all = dict(ENV_DEFAULT, **(env or {}))
    out = run_interventions_reference(age, stage0, edges, seed, mode=mode, num_steps=T, variant=variant, meta=meta_all,
                                      initial=initial, env=env_all, r2=r2, m=m, gI=gI, gD=gD)
    payload = dict(
        seed=np.int64(seed), mode=np.array(mode), p_isolate=np.float32(0.3), num_steps=np.int64(T), variant=np.array(variant),
        gI=gI, gD=gD, r2=np.asarray(r2, dtype=np.float32), m=np.asarray(m, dtype=np.float32),
        overrides=np.array(json.dumps({"meta": meta_all, "initial": initial or {}, "env": env_all})),
        ref_daily_infected=out["daily_infected"].astype(np.float32), ref_daily_deaths=out["daily_deaths"].astype(np.float32),
        ref_grad_R2=out["grad_R2"].astype(np.float32), ref_grad_M=out["grad_M"].astype(np.float32), ref_loss=np.float64(out["loss"]),
    )
    if store_inputs:
        payload.update(age=age.astype(np.uint8), stage0=stage0.astype(np.uint8), edges=edges.astype(np.int32))
    if keep_steps is not None:
        payload["kept_steps"] = np.asarray(keep_steps, dtype=np.int64)
    keys = ["disease_stage", "next_stage_time"]
    if variant == "interventions":
        keys += ["is_quarantined", "quarantine_start_date", "awaiting_test_result", "test_result_date", "test_re_eligble_date"]
    for k in keys:
        h = out[k + "_history"]
        if keep_steps is not None:
            h = h[list(keep_steps)]
        dt = np.float32 if k == "disease_stage" else SMALL.get(k, np.int16)
        payload["ref_" + k + "_history"] = h.astype(dt)
        assert np.array_equal(payload["ref_" + k + "_history"].astype(np.float64), h), k
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **payload)
    print(name, "%.0f KB" % (os.path.getsize(path) / 1024), "deaths", out["daily_deaths"][:8], "grad_R2", out["grad_R2"].reshape(-1),
          "grad_M", out["grad_M"], {k: (float(out[k + "_history"].min()), float(out[k + "_history"].max())) for k in keys})


def astoria():
    """The bundled astoria population (inputs: tests/golden/astoria_inputs.npz): states kept at steps 0, 10 and 20."""
    from oracle.ref_harness.run_reference import load_bundled_population

    age, stage0, edges = load_bundled_population("astoria")
    r2 = np.full((3, 1), 4.75, dtype=np.float32)
    save_case("interventions_astoria", "interventions", age, stage0, edges, seed=57, mode="hard", r2=r2, m=np.array([0.12], dtype=np.float32),
              meta={"quarantine_days": 5}, initial={"test_re_eligble_date": 0}, keep_steps=(0, 10, 20), store_inputs=False)
    save_case("seirms_astoria", "seirms", age, stage0, edges, seed=58, mode="hard", r2=r2, m=np.array([0.4], dtype=np.float32),
              keep_steps=(0, 10, 20), store_inputs=False)


def main():
    r2 = np.array([[1.5], [1.1], [0.8]], dtype=np.float32)
    a, s, e = synthetic_population(1500, 6.0, seed=8, frac_e=0.05, frac_i=0.05)
    # short quarantines and an open testing queue, so that end / start / break and result / enrol paths all fire
    save_case("interventions_1500_hard", "interventions", a, s, e, seed=51, mode="hard", r2=r2, m=np.array([0.3], dtype=np.float32),
              meta={"quarantine_days": 4}, initial={"test_re_eligble_date": 0}, env={"quarantine_break_prob": 0.2})
    # the shipped defaults: 12-day quarantines, testing never opens (re-eligibility date = INFINITY_TIME), soft isolation
    save_case("interventions_1500_default", "interventions", a, s, e, seed=52, mode="soft", r2=r2, m=np.array([0.12], dtype=np.float32))
    # a population that starts quarantined / awaiting results
    save_case("interventions_1500_busy", "interventions", a, s, e, seed=53, mode="hard", r2=r2, m=np.array([0.2], dtype=np.float32),
              meta={"quarantine_days": 3, "test_ineligible_days": 1, "test_result_delay_days": 2},
              initial={"is_quarantined": 1, "quarantine_start_date": 1, "awaiting_test_result": 1, "test_result_date": 2,
                       "test_re_eligble_date": 4},
              env={"quarantine_start_prob": 0.5, "quarantine_break_prob": 0.4, "test_compliance_prob": 0.6})
    # reference-default soft isolation with an open testing queue: the case the fused Runner path is checked on
    save_case("interventions_1500_softopen", "interventions", a, s, e, seed=59, mode="soft", r2=np.array([[4.0], [3.0], [2.0]], dtype=np.float32),
              m=np.array([0.25], dtype=np.float32), meta={"quarantine_days": 4}, initial={"test_re_eligble_date": 0},
              env={"quarantine_break_prob": 0.25, "test_compliance_prob": 0.6})
    save_case("seirms_1500_hard", "seirms", a, s, e, seed=54, mode="hard", r2=r2, m=np.array([0.3], dtype=np.float32))
    save_case("seirms_1500_soft", "seirms", a, s, e, seed=55, mode="soft", r2=np.array([[4.0], [3.0], [2.0]], dtype=np.float32),
              m=np.array([0.6], dtype=np.float32))
    a, s, e = synthetic_population(64, 4.0, seed=9, frac_e=0.2, frac_i=0.3)
    save_case("seirms_64_hard", "seirms", a, s, e, seed=56, mode="hard", r2=r2, m=np.array([0.5], dtype=np.float32))
    astoria()


if __name__ == "__main__":
    main()
