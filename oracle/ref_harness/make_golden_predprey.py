"""Generate tests/golden/predprey_*.npz by running the UNMODIFIED reference predator-prey model here.

TEST INFRASTRUCTURE — build container only (needs /root/reference).  Each fixture holds the initial state the
reference's Initializer produced from the bundled CSVs (+ its random_float energies), the model constants, the
injected-noise seed, the dense adjacency matrix of `grid_network`, and the full state after every step.

    python oracle/ref_harness/make_golden_predprey.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_harness.run_predprey_reference import run_reference_predprey      # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
STD_OUTPUTS = ["prey_pos", "prey_energy", "pred_pos", "pred_energy"]

CASES = [
    # name, grid, steps, seed, energy_seed, overrides
    ("predprey_18x25", "18x25", 30, 11, 0, None),                                      # config.yaml as shipped (prey keep their energy)
    ("predprey_18x25_work", "18x25", 30, 12, 1,                                        # config_vmap.yaml:283 output mapping
     {"substeps/0/transition/update_positions/output_variables": STD_OUTPUTS}),
    ("predprey_24x28", "24x28", 24, 13, 2, None),
    ("predprey_18x25_starve", "18x25", 16, 15, 4,                                      # predators run out of energy too
     {"substeps/0/transition/update_positions/output_variables": STD_OUTPUTS,
      "state/agents/predator/properties/stride_work/value": 17, "state/agents/prey/properties/nutritional_value/value": 9,
      "state/objects/grass/properties/regrowth_time/value": 6.3}),
    ("predprey_40x40_work", "40x40", 24, 14, 3,
     {"substeps/0/transition/update_positions/output_variables": STD_OUTPUTS}),
]


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    for name, grid, steps, seed, eseed, overrides in CASES:
        init, traj, c, adj = run_reference_predprey(grid=grid, num_steps=steps, seed=seed, energy_seed=eseed, overrides=overrides)
        out = {f"init_{k}": v for k, v in init.items()}
        for k in traj[0]:
            out[f"traj_{k}"] = np.stack([s[k] for s in traj])
        out.update(bounds=c["bounds"], seed=np.int64(seed), steps=np.int64(steps),
                   consts=np.array([c["pred_work"], c["prey_work"], c["prey_nutrition"], c["grass_nutrition"], c["regrowth_time"]], dtype=np.float64),
                   update_outputs=np.array(c["update_outputs"]), adjacency=np.packbits(adj.astype(np.uint8), axis=1),
                   n_cells=np.int64(adj.shape[0]))
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
        last = traj[-1]
        print(name, "prey alive", int((last["prey_energy"] > 0).sum()), "pred alive", int((last["pred_energy"] > 0).sum()),
              "grown", int(last["growth_stage"].sum()))


if __name__ == "__main__":
    main()
