"""Shared helpers for the parity tests (test infrastructure)."""
import os

import numpy as np

from oracle import noise

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

GOLDEN_CASES = ["astoria_hard", "astoria_soft", "synth3k_hard", "synth3k_soft", "synth3k_r2vec",
                "tiny96_hard", "noiso500"]
_INPUTS_FROM = {"astoria_hard": "astoria_inputs", "astoria_soft": "astoria_inputs", "synth3k_r2vec": "synth3k_hard"}


def load_golden(name):
    """Returns a dict: inputs (age, stage0, edges), noise address, reference outputs."""
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    if "age" not in g:
        src = np.load(os.path.join(GOLDEN_DIR, _INPUTS_FROM[name] + ".npz"))
        for k in ("age", "stage0", "edges"):
            g[k] = src[k]
    g["mode"] = str(g["mode"])
    g["seed"] = int(g["seed"])
    g["num_steps"] = int(g["num_steps"])
    return g


def make_noise(seed, mode, n, steps, p_iso=0.3, id_offset=0):
    ids = np.arange(n) + id_offset
    U = np.stack([noise.exposure_uniform(seed, t, ids) for t in range(steps)])
    W = np.stack([noise.isolation_decision(seed, t, ids, mode, p_iso) for t in range(steps)])
    return U, W


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.abs(a - b).max() / max(np.abs(a).max(), 1e-30))
