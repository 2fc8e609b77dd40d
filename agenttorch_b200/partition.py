"""Contact-graph partitioning for the multi-GPU path: a re-ordering of the agents after which contiguous id ranges — what
`graph.shard_bounds` hands to the ranks — are communities of the contact network, so that a rank's edges refer to few of its peers'
agents and the halo mask of the per-step exchange (`at_seirm_exchange_sparse`, include/at_seirm.h) leaves most of them out.

The reference has no counterpart (its distributed runner slices agent tensors by position and never exchanges anything,
core/distributed_runner.py:126-144, :218-228).  This is a pre-processing step of the POPULATION, applied before `runner.init()`:
the noise streams of the kernels are keyed by agent id, so a relabelled population is a different (isomorphic) population whose
trajectories are the same at every world size, not the trajectories of the original labelling.

Method: balanced label propagation (Ugander & Backstrom 2013; Spinner; PuLP): every agent looks at the parts its contacts
(in either direction) belong to and asks for the one holding most of them; requests are granted as pairwise SWAPS between parts,
best gains first, so the part sizes — the ranks' ownership ranges, aligned to 1024 agents — never change.  A handful of
rounds of dense torch operations (CPU or GPU); no multilevel coarsening, so it finds the community structure a population has
(households, neighbourhoods, the `comm` family of `graph.synthetic_contact_graph`) and does nothing for a uniformly random graph.
"""
import torch

from .graph import shard_bounds


def part_sizes(num_agents, parts):
    """Sizes of the ranks' ownership ranges (`graph.shard_bounds`: equal 1024-aligned ranges, the last one shorter)."""
    return [hi - lo for lo, hi in shard_bounds(num_agents, parts)]


def edge_cut_fraction(src, dst, labels):
    """Share of the edges whose two ends lie in different parts."""
    if src.numel() == 0:
        return 0.0
    return float((labels[src] != labels[dst]).double().mean())


def halo_fraction(src, dst, labels, parts):
    """Mean over the parts of: 32-agent groups of OTHER parts that the in-edges of the part's agents refer to, as a share of all
    groups outside the part — what the halo mask lets through, for agents numbered part by part (`order_from_labels`)."""
    new_id = new_ids_from_order(order_from_labels(labels, torch.stack([src, dst])))
    s, d = new_id[src], new_id[dst]
    n = labels.numel()
    bounds = shard_bounds(n, parts)
    total = 0.0
    for lo, hi in bounds:
        mine = (d >= lo) & (d < hi)
        srcs = s[mine]
        outside = srcs[(srcs < lo) | (srcs >= hi)]
        groups_outside = max(-(-n // 32) - -(-(hi - lo) // 32), 1)
        total += float(torch.unique(outside >> 5).numel()) / groups_outside
    return total / len(bounds)


def order_from_labels(labels, edge_index=None):
    """Permutation `order` with order[k] = old id of the agent that gets new id k: parts in sequence; inside a part the old order,
    or — given the edges — [interior | boundary]: agents that are the source of an edge into ANOTHER part come last, so that the
    32-agent groups a peer refers to are few and adjacent (the halo mask works on groups)."""
    if edge_index is None:
        return torch.argsort(labels, stable=True)
    src, dst = edge_index[0].long(), edge_index[1].long()
    boundary = torch.zeros_like(labels)
    boundary[src[labels[src] != labels[dst]]] = 1
    return torch.argsort(labels * 2 + boundary, stable=True)


def new_ids_from_order(order):
    new_id = torch.empty_like(order)
    new_id[order] = torch.arange(order.numel(), device=order.device, dtype=order.dtype)
    return new_id


def label_propagation_partition(edge_index, num_agents, parts, rounds=10, labels=None):
    """Balanced label propagation.  edge_index: (2, E) int64 [source; destination]; returns labels (num_agents,) int64 in
    [0, parts) with exactly `part_sizes(num_agents, parts)` agents per part.  Start: `labels` or the current contiguous ranges."""
    src, dst = edge_index[0].long(), edge_index[1].long()
    n, dev = int(num_agents), edge_index.device
    sizes = part_sizes(n, parts)
    if labels is None:
        labels = torch.repeat_interleave(torch.arange(parts, device=dev), torch.tensor(sizes, device=dev))
    labels = labels.clone().long()
    if parts < 2 or src.numel() == 0:
        return labels
    ar = torch.arange(n, device=dev)
    for _ in range(int(rounds)):
        # counts[i, p] = contacts of agent i (either direction) that are in part p
        counts = torch.zeros(n * parts, dtype=torch.int32, device=dev)
        one = torch.ones(src.numel(), dtype=torch.int32, device=dev)
        counts.index_add_(0, dst * parts + labels[src], one)
        counts.index_add_(0, src * parts + labels[dst], one)
        counts = counts.view(n, parts)
        here = counts[ar, labels]
        best_count, best = counts.max(dim=1)
        gain = best_count - here
        wants = (best != labels) & (gain > 0)
        if not bool(wants.any()):
            break
        moved = 0
        for a in range(parts):
            for b in range(a + 1, parts):
                ab = torch.nonzero(wants & (labels == a) & (best == b)).view(-1)
                ba = torch.nonzero(wants & (labels == b) & (best == a)).view(-1)
                m = min(ab.numel(), ba.numel())
                if m == 0:
                    continue
                ab = ab[torch.argsort(gain[ab], descending=True, stable=True)[:m]]
                ba = ba[torch.argsort(gain[ba], descending=True, stable=True)[:m]]
                labels[ab] = b
                labels[ba] = a
                wants[ab] = False
                wants[ba] = False
                moved += 2 * m
        if moved == 0:
            break
    return labels


def reorder_population(edge_index, per_agent, labels):
    """Relabel a population part by part.  per_agent: dict of tensors whose first dimension is the agent; returns
    (edge_index', per_agent', order) with order[k] = old id of new agent k."""
    order = order_from_labels(labels, edge_index)
    new_id = new_ids_from_order(order)
    new_edges = torch.stack([new_id[edge_index[0].long()], new_id[edge_index[1].long()]]).to(edge_index.dtype)
    return new_edges, {k: v[order.to(v.device)] for k, v in per_agent.items()}, order
