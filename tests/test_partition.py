"""Contact-graph partitioner (agenttorch_b200/partition.py): balanced label propagation + [interior | boundary] numbering."""
import torch

from agenttorch_b200 import partition as P
from agenttorch_b200.graph import shard_bounds, synthetic_contact_graph


def _shuffled_community_graph(n=8192, deg=8.0, block=256, p_local=0.995, seed=3):
    g = synthetic_contact_graph(n, deg, seed, family="comm", block=block, p_local=p_local, device="cpu")
    rp = g.row_ptr.long()
    dst = torch.repeat_interleave(torch.arange(n), rp[1:] - rp[:-1])
    src = g.col_idx[: g.num_edges].long()
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(seed + 1))        # hide the structure: random agent ids
    return torch.stack([perm[src], perm[dst]]), n


def test_label_propagation_keeps_part_sizes_and_finds_the_communities():
    edges, n = _shuffled_community_graph()
    for parts in (2, 4):
        start = P.label_propagation_partition(edges, n, parts, rounds=0)
        labels = P.label_propagation_partition(edges, n, parts, rounds=20)
        assert torch.bincount(labels, minlength=parts).tolist() == [hi - lo for lo, hi in shard_bounds(n, parts)]
        cut0, cut1 = P.edge_cut_fraction(edges[0], edges[1], start), P.edge_cut_fraction(edges[0], edges[1], labels)
        assert cut0 > 0.45 and cut1 < 0.25 * cut0                      # random ids cut (parts-1)/parts of the edges
        halo0, halo1 = P.halo_fraction(edges[0], edges[1], start, parts), P.halo_fraction(edges[0], edges[1], labels, parts)
        assert halo0 > 0.9 and halo1 < 0.5 * halo0                     # what the exchange's halo mask lets through


def test_uniformly_random_graph_is_left_balanced_and_valid():
    n = 4096
    g = synthetic_contact_graph(n, 6.0, 5, family="rand", device="cpu")
    rp = g.row_ptr.long()
    edges = torch.stack([g.col_idx[: g.num_edges].long(), torch.repeat_interleave(torch.arange(n), rp[1:] - rp[:-1])])
    labels = P.label_propagation_partition(edges, n, 4, rounds=5)
    assert torch.bincount(labels, minlength=4).tolist() == P.part_sizes(n, 4)
    assert P.edge_cut_fraction(edges[0], edges[1], labels) > 0.5       # nothing to find


def test_reorder_population_is_a_relabelling():
    edges, n = _shuffled_community_graph(n=4096)
    age = torch.arange(n) % 6
    labels = P.label_propagation_partition(edges, n, 2, rounds=10)
    new_edges, per_agent, order = P.reorder_population(edges, {"age": age}, labels)
    assert torch.equal(torch.sort(order).values, torch.arange(n))      # a permutation
    assert torch.equal(per_agent["age"], age[order])
    assert torch.equal(order[new_edges[0]], edges[0]) and torch.equal(order[new_edges[1]], edges[1])     # same contacts, new names
    # parts are contiguous ranges of the new ids (what graph.shard_bounds hands to the ranks), boundary agents last inside a part
    new_labels = labels[order]
    assert torch.equal(new_labels, torch.sort(new_labels).values)
    lo, hi = shard_bounds(n, 2)[0]
    is_boundary = torch.zeros(n, dtype=torch.bool)
    cut = new_labels[new_edges[0]] != new_labels[new_edges[1]]
    is_boundary[new_edges[0][cut]] = True
    first = is_boundary[lo:hi]
    assert not bool(first[: int((~first).sum())].any())                # [interior | boundary]
    assert P.edge_cut_fraction(edges[0], edges[1], labels) == P.edge_cut_fraction(new_edges[0], new_edges[1], new_labels)


def test_degenerate_inputs():
    empty = torch.zeros((2, 0), dtype=torch.long)
    assert P.label_propagation_partition(empty, 2048, 2).tolist() == [0] * 1024 + [1] * 1024
    assert P.edge_cut_fraction(empty[0], empty[1], torch.zeros(2048, dtype=torch.long)) == 0.0
    one = P.label_propagation_partition(torch.tensor([[0, 1], [1, 0]]), 1500, 1)
    assert one.tolist() == [0] * 1500
