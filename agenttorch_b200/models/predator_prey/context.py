"""Glue between the reference-shaped state dict and the predator-prey engine."""
import re

import torch

from agenttorch_b200.core.helpers import get_by_path
from agenttorch_b200.predprey import CsrAdjacency, GridLattice, PredPreyEngine, PredPreySpec

SPECIES = {"predator": 0, "prey": 1}


def get_var(state, var):
    return get_by_path(state, re.split("/", var))


def require_cuda(*tensors):
    for t in tensors:
        if torch.is_tensor(t) and not t.is_cuda:
            raise RuntimeError("agenttorch_b200 predator-prey substeps need CUDA tensors (simulation_metadata.device: "
                               "cuda); there is no CPU fallback")


def species_of(path):
    """'agents/prey/coordinates' -> 1.  The species number addresses the move draws (oracle/noise.py: move_draw)."""
    name = re.split("/", path)[1]
    if name not in SPECIES:
        raise KeyError(f"predator-prey substeps expect agent types {sorted(SPECIES)}, got '{name}'")
    return SPECIES[name]


_ADJ_CACHE = {}


def topology(adjacency, device):
    """(GridLattice | None, CsrAdjacency | None) from what state['network'][...]['adjacency_matrix'] holds."""
    if isinstance(adjacency, GridLattice):
        return adjacency, None
    if torch.is_tensor(adjacency):
        key = (adjacency.data_ptr(), tuple(adjacency.shape), adjacency._version)
        hit = _ADJ_CACHE.get(key)
        if hit is None:
            if len(_ADJ_CACHE) > 8:
                _ADJ_CACHE.clear()
            hit = _ADJ_CACHE[key] = CsrAdjacency(adjacency, device)
        return None, hit
    raise TypeError(f"unsupported adjacency payload {type(adjacency).__name__}")


def spec_from_bounds(bounds, n_cells=None, lattice=None, **consts):
    mx, my = (int(v) for v in bounds.reshape(-1).tolist())
    n_cells = int(n_cells if n_cells is not None else (lattice.n_cells if lattice is not None else mx * my))
    return PredPreySpec(max_x=mx, max_y=my, n_cells=n_cells, width=int(lattice.width) if lattice is not None else mx, **consts)


_ENGINES = {}


def engine_for(spec, device, csr=None):
    """One engine (bitmap buffers) per distinct model on a device; constants are part of the key."""
    key = (str(device), id(csr), tuple(sorted(vars(spec).items())))
    eng = _ENGINES.get(key)
    if eng is None:
        if len(_ENGINES) > 16:
            _ENGINES.clear()
        eng = _ENGINES[key] = PredPreyEngine(spec, device, adjacency=csr)
    return eng
