"""Per-simulation engine context shared by the covid substeps and the Runner's fused episode path."""
import re
import weakref

import torch

from agenttorch_b200 import ops
from agenttorch_b200.core.helpers import get_by_path
from agenttorch_b200.engine import SeirmEngine, SeirmSpec
from agenttorch_b200.graph import ContactGraph, build_contact_graph

_CACHE = {}
_MAX_CACHED = 4


def _check_constants(meta):
    want = {"SUSCEPTIBLE_VAR": 0, "EXPOSED_VAR": 1, "INFECTED_VAR": 2, "RECOVERED_VAR": 3}
    for k, v in want.items():
        if int(meta.get(k, v)) != v:
            raise NotImplementedError(f"simulation_metadata.{k} must be {v} for the fused SEIRM kernels")


def spec_from_state(config, state, input_variables):
    meta = config["simulation_metadata"]
    _check_constants(meta)

    def env(name, default_path):
        return get_by_path(state, re.split("/", input_variables.get(name, default_path)))

    sfs = env("SFSusceptibility", "environment/SFSusceptibility").detach().reshape(-1).cpu().tolist()
    sfi = env("SFInfector", "environment/SFInfector").detach().reshape(-1).cpu().tolist()
    lg = env("lam_gamma_integrals", "environment/lam_gamma_integrals").detach().reshape(-1).cpu().tolist()
    return SeirmSpec(sfs=tuple(sfs), sfi=tuple(sfi), lam_gamma=tuple(lg), num_steps=int(meta["num_steps_per_episode"]),
                     exposed_to_infected_time=int(meta["EXPOSED_TO_INFECTED_TIME"]),
                     infected_to_recovered_time=int(meta["INFECTED_TO_RECOVERED_TIME"]))


class SeirmContext:
    def __init__(self, engine):
        self.engine = engine
        self.handle = ops.register_engine(engine)
        self.edge_ref = lambda: None


def _require_cuda(is_cuda):
    """No CPU fallback.  (One function so that the host-emulation tests can lift the check.)"""
    if not is_cuda:
        raise RuntimeError("agenttorch_b200 SEIRM substeps need CUDA tensors (simulation_metadata.device: cuda); "
                           "there is no CPU fallback")


def get_context(config, state, input_variables):
    """Engine for the contact network this substep reads; built on first use and cached by the identity
    of the adjacency payload.  Raises on CPU tensors: the kernels are the only implementation."""
    adjacency = get_by_path(state, re.split("/", input_variables["adjacency_matrix"]))
    edge_index, edge_attr = adjacency
    _require_cuda(edge_index.is_cuda)
    # The key names the adjacency tensor (address, size, in-place version) AND everything the engine bakes in at construction:
    # the model constants, the agent count and I_bar's source.  A re-initialised runner normally gets the freed tensor's
    # address back from the caching allocator, so the address alone would hand a new graph or a changed config the old engine;
    # the weak reference drops the entry as soon as the tensor it was built from dies.
    meta = config["simulation_metadata"]
    mean_int = get_by_path(state, re.split("/", input_variables["mean_interactions"]))
    n = int(meta["num_agents"])

    def ident(name, default_path):       # identity + in-place version of a table tensor (no device read on the hit path)
        t = get_by_path(state, re.split("/", input_variables.get(name, default_path)))
        return (t.data_ptr(), t._version, tuple(t.shape))

    key = (edge_index.data_ptr(), edge_index.shape[1], str(edge_index.device), edge_index._version, n,
           ident("SFSusceptibility", "environment/SFSusceptibility"), ident("SFInfector", "environment/SFInfector"),
           ident("lam_gamma_integrals", "environment/lam_gamma_integrals"), int(meta["num_steps_per_episode"]),
           int(meta["EXPOSED_TO_INFECTED_TIME"]), int(meta["INFECTED_TO_RECOVERED_TIME"]), mean_int.data_ptr(), mean_int._version)
    ctx = _CACHE.get(key)
    if ctx is not None and ctx.edge_ref() is not edge_index:
        ops.release_engine(_CACHE.pop(key).handle)          # same address, different tensor: the graph was re-created
        ctx = None
    if ctx is None:
        net = get_by_path(state, re.split("/", input_variables["adjacency_matrix"])[:-1])
        prebuilt = net.get("graph") if isinstance(net, dict) else None
        if isinstance(prebuilt, ContactGraph) and prebuilt.num_edges == edge_index.shape[1]:
            graph = prebuilt.to(edge_index.device)      # network helper already produced the dst-sorted CSR
        else:
            graph = build_contact_graph(edge_index, n, edge_attr=edge_attr, mean_interactions=mean_int, device=edge_index.device)
        spec = spec_from_state(config, state, input_variables)
        ctx = SeirmContext(SeirmEngine(graph, spec))
        ctx.table_fingerprint = (spec.sfs, spec.sfi, spec.lam_gamma)
        ctx.edge_ref = weakref.ref(edge_index)
        for old in [k for k, c in _CACHE.items() if c.edge_ref() is None]:       # engines whose graph tensor died
            ops.release_engine(_CACHE.pop(old).handle)
        if len(_CACHE) >= _MAX_CACHED:
            old = next(iter(_CACHE))
            ops.release_engine(_CACHE.pop(old).handle)
        _CACHE[key] = ctx
    return ctx


def as_kernel_f32(t):
    """State tensors are float32 in this package; accept float64 inputs (the reference's CSV path) by value."""
    return t if t.dtype == torch.float32 and t.is_contiguous() else t.to(torch.float32).contiguous()
