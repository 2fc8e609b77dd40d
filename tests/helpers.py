"""Shared helpers for the parity tests (test infrastructure)."""
import os

import numpy as np

from oracle import noise

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

GOLDEN_CASES = ["astoria_hard", "astoria_soft", "synth3k_hard", "synth3k_soft", "synth3k_r2vec",
                "tiny96_hard", "noiso500"]
_INPUTS_FROM = {"astoria_hard": "astoria_inputs", "astoria_soft": "astoria_inputs", "synth3k_r2vec": "synth3k_hard",
                "synth3k_tables_soft": "synth3k_layers_hard"}
# round-2 fixtures of the real reference: per-edge network types (multi-layer contact networks) and the gradients with respect
# to SFSusceptibility / SFInfector (oracle/ref_harness/make_golden.py round2)
TABLE_CASES = ["synth3k_layers_hard", "synth3k_tables_soft", "layers400_none"]


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


# ------------------------------------------------------------------------------------ predator-prey
PREDPREY_CASES = ["predprey_18x25", "predprey_18x25_work", "predprey_24x28", "predprey_18x25_starve", "predprey_40x40_work"]


def load_predprey_golden(name):
    """Fixture written by oracle/ref_harness/make_golden_predprey.py (outputs of the real reference)."""
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    g["seed"], g["steps"], g["n_cells"] = int(g["seed"]), int(g["steps"]), int(g["n_cells"])
    g["adjacency"] = np.unpackbits(g["adjacency"], axis=1)[:, : g["n_cells"]]
    g["update_outputs"] = [str(s) for s in g["update_outputs"]]
    return g


def predprey_port_from_golden(g, use_adjacency=False):
    from oracle.predprey_port import PredPreyConstants, PredPreyPort

    pw, yw, pn, gn, rt = (float(v) for v in g["consts"])
    consts = PredPreyConstants(int(g["bounds"][0]), int(g["bounds"][1]), pw, yw, pn, gn, rt, g["update_outputs"])
    return PredPreyPort(consts, g["init_pred_pos"], g["init_pred_energy"], g["init_prey_pos"], g["init_prey_energy"],
                        g["init_growth_stage"], g["init_growth_countdown"], adjacency=g["adjacency"] if use_adjacency else None)


PREDPREY_KEYS = ("pred_pos", "pred_energy", "prey_pos", "prey_energy", "growth_stage", "growth_countdown")


# ------------------------------------------------------------------- quarantine / testing / SEIRMS progression
INTERVENTION_CASES = ["interventions_1500_hard", "interventions_1500_default", "interventions_1500_busy", "interventions_1500_softopen"]
SEIRMS_CASES = ["seirms_1500_hard", "seirms_1500_soft", "seirms_64_hard"]


def load_interventions_golden(name):
    """Fixture written by oracle/ref_harness/make_golden_interventions.py (outputs of the real reference)."""
    import json

    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    g["mode"], g["variant"] = str(g["mode"]), str(g["variant"])
    g["seed"], g["num_steps"] = int(g["seed"]), int(g["num_steps"])
    g["overrides"] = json.loads(str(g["overrides"]))
    return g
