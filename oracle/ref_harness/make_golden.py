"""Generate tests/golden/*.npz by running the real reference (/root/reference) on CPU.

TEST INFRASTRUCTURE.  Run in the build container only:
    python -m oracle.ref_harness.make_golden
Each fixture stores the inputs (age, initial stages, edge list), the noise
address (seed, isolation mode) — the uniforms themselves are regenerated from
oracle/noise.py — the loss weights, and the reference's outputs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import noise  # noqa: E402
from oracle.ref_harness.run_reference import load_bundled_population, run_reference  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
T = 21


def make_noise(seed, mode, n, p_iso=0.3, steps=T):
    ids = np.arange(n)
    U = np.stack([noise.exposure_uniform(seed, t, ids) for t in range(steps)])
    W = np.stack([noise.isolation_decision(seed, t, ids, mode, p_iso) for t in range(steps)])
    return U, W


def synthetic_population(n, avg_deg, seed, frac_e=0.02, frac_i=0.01):
    rng = np.random.default_rng(seed)
    probs = np.array([6723, 11573, 5724, 4647, 3910, 4941], dtype=np.float64)
    age = rng.choice(6, size=n, p=probs / probs.sum())
    stage0 = np.zeros(n)
    r = rng.random(n)
    stage0[r < frac_e] = 1.0
    stage0[(r >= frac_e) & (r < frac_e + frac_i)] = 2.0
    e = int(n * avg_deg)
    # directed, unsorted, duplicates and self-loops kept
    edges = np.stack([rng.integers(0, n, e), rng.integers(0, n, e)], axis=1)
    return age.astype(np.int64), stage0, edges.astype(np.int64)


def save_case(name, age, stage0, edges, seed, mode, r2=None, m=None, store_inputs=True, edge_type=None, table_grads=False):
    n = len(age)
    U, W = make_noise(seed, mode, n)
    gI = np.linspace(0.5, 1.5, T).astype(np.float32)
    gD = np.linspace(2.0, 1.0, T).astype(np.float32)
    out = run_reference(age, stage0, edges, U, W, num_steps=T, r2=r2, m=m, gI=gI, gD=gD, edge_type=edge_type, table_grads=table_grads)
    payload = dict(
        seed=np.int64(seed), mode=np.array(mode), p_isolate=np.float32(0.3), num_steps=np.int64(T),
        gI=gI, gD=gD,
        r2=np.asarray(r2 if r2 is not None else np.full((3, 1), 4.75), dtype=np.float32),
        m=np.asarray(m if m is not None else [0.12], dtype=np.float32),
        ref_stage=out["stage"].astype(np.float64),
        ref_next_stage_time=out["next_stage_time"].astype(np.int16),
        ref_infected_time=out["infected_time"].astype(np.int16),
        ref_stage_history=out["stage_history"].astype(np.float32),
        ref_daily_infected=out["daily_infected"].astype(np.float32),
        ref_daily_deaths=out["daily_deaths"].astype(np.float32),
        ref_grad_R2=out["grad_R2"].astype(np.float32),
        ref_grad_M=out["grad_M"].astype(np.float32),
        ref_loss=np.float64(out["loss"]),
        ref_lam_gamma_integrals=out["lam_gamma_integrals"].astype(np.float32),
    )
    assert np.array_equal(payload["ref_stage_history"].astype(np.float64), out["stage_history"])
    assert np.array_equal(payload["ref_next_stage_time"].astype(np.float32), out["next_stage_time"])
    assert np.array_equal(payload["ref_infected_time"].astype(np.float32), out["infected_time"])
    if edge_type is not None:
        payload["edge_type"] = np.asarray(edge_type, dtype=np.int32)
    if table_grads:
        payload["ref_grad_SFS"] = out["grad_SFS"].astype(np.float32)
        payload["ref_grad_SFI"] = out["grad_SFI"].astype(np.float32)
    if store_inputs:
        payload.update(age=age.astype(np.uint8), stage0=stage0.astype(np.uint8), edges=edges.astype(np.int32))
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **payload)
    print(name, "N=%d E=%d" % (n, len(edges)), "%.0f KB" % (os.path.getsize(path) / 1024),
          "daily_infected[:6]=", out["daily_infected"][:6], "grad_R2=", out["grad_R2"].reshape(-1), "grad_M=", out["grad_M"])


def layered_edge_types(age, edges, seed):
    """Per-edge network types for a three-layer contact network: the reference gathers I_bar = mean_interactions[dst[type]]
    (transition.py:71-73), so the three types are numbers of edges whose destinations have 2, 3 and 4 mean interactions."""
    mi = np.where(age < 1, 2.0, np.where(age > 4, 4.0, 3.0))
    dst = edges[:, 1]
    ks = np.array([int(np.where(mi[dst] == v)[0][0]) for v in (2.0, 3.0, 4.0)])
    return ks[np.random.default_rng(seed).integers(0, 3, len(edges))]


def main_round2():
    """Round-2 fixtures: multi-layer contact networks and the gradients w.r.t. SFSusceptibility / SFInfector."""
    a, s, e = synthetic_population(3000, 8.0, seed=4242)
    save_case("synth3k_layers_hard", a, s, e, seed=77, mode="hard", edge_type=layered_edge_types(a, e, 5), table_grads=True)
    save_case("synth3k_tables_soft", a, s, e, seed=78, mode="soft", table_grads=True,
              r2=np.array([[1.2], [0.4], [0.8]], dtype=np.float32), m=np.array([0.25], dtype=np.float32), store_inputs=False)
    a, s, e = synthetic_population(400, 5.0, seed=9, frac_e=0.05, frac_i=0.05)
    save_case("layers400_none", a, s, e, seed=79, mode="none", edge_type=layered_edge_types(a, e, 6), table_grads=True)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    age, stage0, edges = load_bundled_population("astoria")
    np.savez_compressed(os.path.join(GOLDEN, "astoria_inputs.npz"), age=age.astype(np.uint8),
                        stage0=stage0.astype(np.uint8), edges=edges.astype(np.int32))
    save_case("astoria_hard", age, stage0, edges, seed=11, mode="hard", store_inputs=False)
    save_case("astoria_soft", age, stage0, edges, seed=12, mode="soft", store_inputs=False)
    a, s, e = synthetic_population(3000, 8.0, seed=5)
    save_case("synth3k_hard", a, s, e, seed=21, mode="hard",
              r2=np.array([[0.9], [0.7], [0.5]], dtype=np.float32), m=np.array([0.3], dtype=np.float32))
    save_case("synth3k_soft", a, s, e, seed=22, mode="soft",
              r2=np.array([[1.2], [0.4], [0.8]], dtype=np.float32), m=np.array([0.25], dtype=np.float32))
    # a user-injected (NUM_WEEKS,) R2 (reference example.py:73-79): R = R2[week], per-week gradients differ
    save_case("synth3k_r2vec", a, s, e, seed=23, mode="hard",
              r2=np.array([1.4, 0.6, 0.9], dtype=np.float32), m=np.array([0.2], dtype=np.float32), store_inputs=False)
    a, s, e = synthetic_population(96, 4.0, seed=6, frac_e=0.1, frac_i=0.1)
    save_case("tiny96_hard", a, s, e, seed=31, mode="hard",
              r2=np.array([[1.5], [1.5], [1.5]], dtype=np.float32))
    a, s, e = synthetic_population(500, 3.0, seed=7, frac_e=0.05, frac_i=0.05)
    save_case("noiso500", a, s, e, seed=41, mode="none")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "round2":
        main_round2()
    else:
        main()
