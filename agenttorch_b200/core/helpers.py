"""Path, copy and config helpers with the reference's names and behaviour
(agent_torch/core/helpers/general.py:11-129), re-implemented without omegaconf."""
import copy
import re

import torch
import yaml
from torch import nn

from .soft import *  # noqa: F401,F403  (the reference's core.helpers re-exports soft.py and initializer.py the same way)


def _walk(root, keys):
    node = root
    for k in keys:
        node = node[k]
    return node


def get_by_path(root, items):
    """Nested lookup; a plain nn.Module leaf is called, a ModuleDict is returned (general.py:11-21)."""
    obj = _walk(root, items)
    if isinstance(obj, nn.ModuleDict):
        return obj
    if isinstance(obj, nn.Module):
        return obj()
    return obj


def get_var(state, var):
    return get_by_path(state, re.split("/", var))


def set_by_path(root, items, value):
    parent = get_by_path(root, items[:-1])
    if isinstance(parent, nn.ModuleDict):
        parent[items[-1]].param.data.copy_(value)
        parent[items[-1]].param.requires_grad = value.requires_grad
        return None
    parent[items[-1]] = value
    return root


def del_by_path(root, items):
    del get_by_path(root, items[:-1])[items[-1]]


def copy_module(tree):
    """New dict tree with every tensor cloned (general.py:49-65).  Unlike the reference, tuples/lists of
    tensors (the `(edge_index, edge_attr)` adjacency payload) are shared, not deep-copied: they are
    immutable inputs and copying ~4 GB of edges per substep is the reference's quirk Q8."""
    out = {}
    for k, v in tree.items():
        if torch.is_tensor(v):
            out[k] = v.clone()
        elif isinstance(v, dict):
            out[k] = copy_module(v)
        elif isinstance(v, (tuple, list)) and all(torch.is_tensor(x) for x in v):
            out[k] = v
        else:
            out[k] = copy.deepcopy(v)
    return out


def to_cpu(tree):
    out = {}
    for k, v in tree.items():
        if torch.is_tensor(v):
            out[k] = v.clone().cpu()
        elif isinstance(v, dict):
            out[k] = to_cpu(v)
        else:
            out[k] = v
    return out


def process_shape(config, s):
    return get_by_path(config, re.split("/", s)) if isinstance(s, str) else s


# ------------------------------------------------------------------------------------------ config
_RESOLVERS = {"sum": lambda x, y: x + y, "multiply": lambda x, y: x * y, "divide": lambda x, y: x // y}
_INTERP = re.compile(r"\$\{([^${}]+)\}")


def register_resolver(name, resolver):
    _RESOLVERS[name] = resolver


def _lookup_dotted(root, dotted):
    node = root
    for part in dotted.split("."):
        node = node[int(part)] if isinstance(node, list) else node[part]
    return node


def _scalar(tok):
    try:
        return yaml.safe_load(tok)
    except yaml.YAMLError:
        return tok


def _resolve(root, value):
    """`${a.b.c}` interpolation and `${name:x,y}` resolvers, applied until a fixed point."""
    for _ in range(64):
        if not isinstance(value, str):
            return value
        m = _INTERP.search(value)
        if m is None:
            return value
        expr = m.group(1).strip()
        head = expr.split(":", 1)[0]
        if ":" in expr and head in _RESOLVERS:
            args = [_scalar(_resolve(root, a.strip())) if isinstance(a, str) else a for a in expr.split(":", 1)[1].split(",")]
            args = [_scalar(a) if isinstance(a, str) else a for a in args]
            sub = _RESOLVERS[head](*args)
        else:
            sub = _resolve(root, _lookup_dotted(root, expr))
        if m.span() == (0, len(value)):
            value = sub
        else:
            value = value[: m.start()] + str(sub) + value[m.end():]
    raise ValueError("config interpolation did not converge")


def _resolve_tree(root, node):
    if isinstance(node, dict):
        return {k: _resolve_tree(root, v) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve_tree(root, v) for v in node]
    return _resolve(root, node)


def resolve_config(raw):
    """Plain-dict config with every interpolation resolved (OmegaConf.to_object equivalent)."""
    return _resolve_tree(raw, raw)


def read_config(config_file, register_resolvers=True):
    """Load a YAML config and resolve interpolations (general.py:101-129)."""
    if str(config_file)[-5:] != ".yaml":
        raise ValueError("Config file type should be yaml")
    try:
        with open(config_file, "r", encoding="utf-8") as f:
            raw = yaml.safe_load(f)
        return resolve_config(raw)
    except Exception as e:  # same error surface as the reference
        raise ValueError(f"Could not load config file. Please check path and file type. Error message is {str(e)}")


def read_from_file(shape, params):
    """CSV (with a header row) -> tensor of the file's own dtype (general.py:132-143)."""
    import pandas as pd

    file_path = params["file_path"]
    if not str(file_path).endswith("csv"):
        raise ValueError(f"read_from_file: only csv files are supported, got {file_path}")
    values = pd.read_csv(file_path).values
    assert values.shape == tuple(shape), f"read_from_file: shape mismatch {values.shape} vs {tuple(shape)} for {file_path}"
    return torch.from_numpy(values.copy())


def grid_network(params):
    """`grid` network helper (core/helpers/environment.py:6-9).  The reference materialises the dense
    (cells x cells) adjacency matrix of nx.grid_graph(dim=shape); that is 2.8e14 entries for a 4096x4096 grid, so
    this returns the same lattice in implicit form (agenttorch_b200.predprey.GridLattice; `.to_dense()` gives the
    reference's matrix for small grids)."""
    from agenttorch_b200.predprey import GridLattice

    return None, GridLattice(params["shape"])


def _nx_network(graph):
    """networkx graph -> (graph, (edge_index (2,E) int64, edge_attr (E,1) ones)): the payload of the reference's random-graph
    helpers (core/helpers/environment.py:12-41), each edge listed once in networkx's order."""
    edges = list(graph.edges())
    edge_index = torch.tensor(edges, dtype=torch.long).reshape(-1, 2).t().contiguous()
    return graph, (edge_index, torch.ones(edge_index.shape[1], 1))


def small_world(params):
    """Watts-Strogatz graph: n (or num_agents, default 1000), k = 10 neighbours, rewiring probability p = 0.1."""
    import networkx as nx

    return _nx_network(nx.watts_strogatz_graph(params.get("n", params.get("num_agents", 1000)), params.get("k", 10),
                                               params.get("p", 0.1)))


def erdos_renyi(params):
    """G(n, p) graph: n (or num_agents, default 1000), p = 0.05."""
    import networkx as nx

    return _nx_network(nx.erdos_renyi_graph(params.get("n", params.get("num_agents", 1000)), params.get("p", 0.05)))


def scale_free(params):
    """Barabasi-Albert graph: n (or num_agents, default 1000), m = 3 edges per new node."""
    import networkx as nx

    return _nx_network(nx.barabasi_albert_graph(params.get("n", params.get("num_agents", 1000)), params.get("m", 3)))


def memory_checkpoint(name):
    allocated, reserved = torch.cuda.memory_allocated(), torch.cuda.memory_reserved()
    print("Checkpoint: ", name, " Allocated: ", allocated, " Reserved: ", reserved)
    return allocated, reserved
