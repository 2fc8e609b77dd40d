"""State initialisers and the network loader of the covid model, with the reference's names, signatures
and values (agent_torch/models/covid/substeps/utils.py:14-148).  Differences: the disease-stage column is
kept in float32 (every value the model produces is exactly representable) so that the
CUDA kernels read it without a conversion pass."""
import hashlib
import os

import numpy as np
import pandas as pd
import torch

from agenttorch_b200.engine import lam_gamma_integrals as _lam_gamma


def get_lam_gamma_integrals(shape, params):
    scale, rate, t = params["scale"], params["rate"], params["t"]
    return _lam_gamma(float(torch.as_tensor(scale).reshape(-1)[0]), float(torch.as_tensor(rate).reshape(-1)[0]),
                      int(torch.as_tensor(t).reshape(-1)[0].item()))


def read_from_file(shape, params):
    path = params["file_path"]
    if path.endswith("csv"):
        df = pd.read_csv(path, header=None)
        values = df.apply(pd.to_numeric, errors="coerce").to_numpy()
    else:
        values = np.asarray(pd.read_pickle(path).values)
    if values.ndim == 1:
        values = values.reshape(-1, 1)
    assert values.shape == tuple(shape), f"read_from_file: shape mismatch {values.shape} vs {tuple(shape)} for {path}"
    return torch.from_numpy(np.ascontiguousarray(values, dtype=np.float32))


def _timers(stages, default, at_exposed, at_infected):
    out = torch.full_like(stages, float(default))
    out[stages == 1] = float(at_exposed)
    out[stages == 2] = float(at_infected)
    return out.float()


def get_infected_time(shape, params):
    # 500 = "never"; previously exposed agents count as infected one day ago, infected ones three days ago
    return _timers(read_from_file(shape, params), 500, -1, -3)


def get_next_stage_time(shape, params, exposed_to_infected_times=3, infected_to_recovered_times=5):
    return _timers(read_from_file(shape, params), 500, exposed_to_infected_times, infected_to_recovered_times)


def load_population_attribute(shape, params):
    df = pd.read_pickle(params["file_path"])
    return torch.from_numpy(np.array(df.values, copy=True)).float().unsqueeze(1)


def get_mean_agent_interactions(shape, params):
    ages = load_population_attribute(shape, params).view(-1)
    out = torch.zeros(size=tuple(shape))
    out[ages < 1, 0] = 2.0                      # children
    out[(ages >= 1) & (ages <= 4), 0] = 3.0     # adults ('20t29' .. '50t64')
    out[ages > 4, 0] = 4.0                      # elderly
    return out


def initialize_id(shape, params):
    return torch.arange(0, shape[0]).reshape(-1, 1).float()


def _edge_cache_path(path):
    """Opt-in cache of parsed edge lists: set AGENTTORCH_B200_CACHE_DIR.  Keyed by path, size and mtime of the CSV."""
    root = os.environ.get("AGENTTORCH_B200_CACHE_DIR")
    if not root:
        return None
    st = os.stat(path)
    key = hashlib.sha1(f"{os.path.abspath(path)}|{st.st_size}|{st.st_mtime_ns}".encode()).hexdigest()[:20]
    return os.path.join(root, f"edges_{key}.npy")


def _parse_edge_list(path):
    """(2, E) int64 edge_index of an edge-list CSV.  Native parser (libat_io.so, include/at_io.h: memory-mapped, one thread
    per core, same row semantics) when it has been built, else the reference's pandas sequence."""
    from agenttorch_b200 import _native_io

    if _native_io.available():
        return _native_io.read_edge_list(path)[0]
    df = pd.read_csv(path, header=None)
    if df.iloc[0].apply(lambda v: isinstance(v, str)).any():
        df = pd.read_csv(path, header=0)
    df = df.iloc[:, :2].apply(pd.to_numeric, errors="coerce").dropna().astype("int64")
    return torch.tensor(df.values.T, dtype=torch.long).contiguous()


def network_from_file(params):
    """Edge-list CSV -> (None, (edge_index (2,E) int64, edge_attr (2,E) float32 = [type==1; weight==1])).
    (reference utils.py:129-148).  pandas needs minutes for 168 M text edges; the native parser reads them at several
    hundred MB/s, and with AGENTTORCH_B200_CACHE_DIR set the parsed (E,2) int32 array is also kept as .npy and
    memory-mapped on the next run."""
    path = params["file_path"]
    cache = _edge_cache_path(path)
    if cache is not None and os.path.exists(cache):
        pairs = np.load(cache, mmap_mode="r")
        edge_index = torch.from_numpy(np.array(pairs.T, dtype=np.int64, order="C"))
    else:
        edge_index = _parse_edge_list(path)
        if cache is not None and edge_index.numel() and int(edge_index.max()) < 2**31 and int(edge_index.min()) >= -2**31:
            os.makedirs(os.path.dirname(cache), exist_ok=True)
            tmp = cache + f".{os.getpid()}.tmp.npy"
            np.save(tmp, edge_index.t().to(torch.int32).numpy())
            os.replace(tmp, cache)
    e = edge_index.shape[1]
    edge_attr = torch.vstack((torch.ones(e, dtype=torch.float32), torch.ones(e, dtype=torch.float32)))
    return None, (edge_index, edge_attr)
