"""Population folders in the layout the covid config reads (age.pickle, disease_stages.csv,
mobility_networks/0.csv — reference: populations/astoria/, models/covid/yamls/config.yaml:22,31-32)."""
import os

import numpy as np
import pandas as pd


def write_population_folder(folder, age, stage0, edges):
    """age (N,) ints 0..5, stage0 (N,) in {0,1,2}, edges (E,2) [src,dst] -> files under `folder`."""
    os.makedirs(os.path.join(folder, "mobility_networks"), exist_ok=True)
    pd.Series(np.asarray(age, dtype=np.int64)).to_pickle(os.path.join(folder, "age.pickle"))
    np.savetxt(os.path.join(folder, "disease_stages.csv"), np.asarray(stage0, dtype=np.float64), fmt="%.1f")
    pd.DataFrame(np.asarray(edges, dtype=np.int64)).to_csv(
        os.path.join(folder, "mobility_networks", "0.csv"), header=False, index=False)
    return folder
