"""Host-side data preparation: CSR builder, column-blocked layout, synthetic generators, counter-hash noise."""
import numpy as np
import pytest
import torch

from agenttorch_b200 import noise as tnoise
from agenttorch_b200.graph import (block_graph, build_contact_graph, choose_block_size, mean_interactions_from_age,
                                   shard_bounds, shard_graph, synthetic_contact_graph, synthetic_population)
from oracle import noise as onoise
from tests.helpers import load_golden


def _decode_rows(bg):
    """BlockedGraph -> list of per-row source lists (block order, then original edge order inside a block)"""
    rpt = bg.rows_per_task
    ngt = bg.n_rows_pad // rpt
    rs = bg.row_start.int() & 0xFFFF
    rows = [[] for _ in range(bg.n_rows)]
    col = bg.col.tolist()
    info = bg.seg_info.tolist()
    for b in range(bg.n_blocks):
        for gt in range(ngt):
            base, ln = info[b * ngt + gt][0] * 4, info[b * ngt + gt][1]
            for l in range(rpt):
                r = gt * rpt + l
                if r >= bg.n_rows:
                    continue
                s0 = int(rs[b, r])
                e0 = int(rs[b, r + 1]) if l < rpt - 1 else ln
                for enc in col[base + s0: base + e0]:
                    c = enc if bg.hash_words else (((enc >> 7) << 5) | (enc & 31))
                    rows[r].append(c + b * bg.block_size)
    return rows


def test_csr_keeps_original_edge_order_duplicates_and_self_loops():
    # edges as written: duplicates (2->0 twice), a self loop (1->1), unsorted destinations (Q9)
    ei = torch.tensor([[2, 3, 2, 1, 0, 2], [0, 1, 0, 1, 3, 3]])
    mi = torch.tensor([2.0, 3.0, 4.0, 3.0])
    g = build_contact_graph(ei, 4, mean_interactions=mi)
    assert g.row_ptr.tolist() == [0, 2, 4, 4, 6]
    assert g.col_idx[:6].tolist() == [2, 2, 3, 1, 0, 2]
    # Q1: I_bar = mean_interactions[dst of ORIGINAL edge #1] = mean_interactions[1]
    assert g.i_bar_agent == 1 and g.i_bar == 3.0
    assert g.col_idx.numel() >= g.num_edges + 4 and g.col_idx.data_ptr() % 16 == 0


def test_i_bar_follows_edge_type_and_mixed_types_become_layers():
    ei = torch.tensor([[0, 1, 2], [1, 2, 0]])
    mi = torch.tensor([2.0, 3.0, 4.0])
    attr = torch.tensor([[2.0, 2.0, 2.0], [1.0, 1.0, 1.0]])
    assert build_contact_graph(ei, 3, edge_attr=attr, mean_interactions=mi).i_bar == 2.0     # dst[2] = 0
    # mixed types: one I_bar per distinct type k, mean_interactions[dst[k]] (the reference's per-edge gather, transition.py:71-73)
    lay = build_contact_graph(ei, 3, edge_attr=torch.tensor([[1.0, 2.0, 1.0], [1.0, 1.0, 1.0]]), mean_interactions=mi)
    assert lay.i_bar_layers == (4.0, 2.0)                                                     # types 1, 2 -> dst[1] = 2, dst[2] = 0
    assert lay.edge_layer[:3].tolist() == [0, 0, 1]                                           # dst-sorted edge order: (2->0), (0->1), (1->2)
    with pytest.raises(NotImplementedError):                                                 # layers + weights: not supported
        build_contact_graph(ei, 3, edge_attr=torch.tensor([[1.0, 2.0, 1.0], [1.0, 0.5, 1.0]]), mean_interactions=mi)
    w = build_contact_graph(ei, 3, edge_attr=torch.tensor([[1.0, 1.0, 1.0], [1.0, 0.5, 1.0]]), mean_interactions=mi)
    assert w.edge_weight is not None and w.edge_weight[:3].tolist() == [1.0, 1.0, 0.5]       # dst-sorted order


@pytest.mark.parametrize("block_size,hash_words", [(8192, 0), (2048, 0), (1024, 0), (8192, 64)])
def test_blocked_layout_round_trips(block_size, hash_words):
    g = synthetic_contact_graph(5000, 6.0, 3)
    bg = block_graph(g, block_size, hash_words=hash_words)
    rec = _decode_rows(bg)
    rp = g.row_ptr.tolist()
    for r in range(g.n_rows):
        ref = g.col_idx[rp[r]: rp[r + 1]].tolist()
        if bg.n_blocks == 1:
            assert rec[r] == ref                       # original edge order inside a row
        else:
            assert sorted(rec[r]) == sorted(ref)
            for b in range(bg.n_blocks):                # order preserved inside every block
                assert [x for x in rec[r] if x // block_size == b] == [x for x in ref if x // block_size == b]
    info = bg.seg_info
    assert int(info[:, 1].sum()) == g.num_edges and bool((info[:, 1] <= 65535).all())
    assert bg.col.numel() >= int(info[-1, 0]) * 4 + int(info[-1, 1]) + 1024


def test_empty_and_ragged_graphs():
    g = build_contact_graph(torch.zeros((2, 0), dtype=torch.long), 70)
    bg = block_graph(g, 1024)
    assert bg.num_edges == 0 and int(bg.seg_info[:, 1].sum()) == 0 and bg.n_rows_pad % 256 == 0
    # one hub row with many in-edges, the rest empty
    ei = torch.stack((torch.arange(3000) % 97, torch.full((3000,), 5)))
    hub = block_graph(build_contact_graph(ei, 100), 1024)
    assert _decode_rows(hub)[5] == (torch.arange(3000) % 97).tolist()
    with pytest.raises(ValueError):
        build_contact_graph(torch.tensor([[0], [7]]), 4)


def test_astoria_fixture_statistics():
    g = load_golden("astoria_hard")
    n = len(g["age"])
    graph = build_contact_graph(torch.from_numpy(g["edges"].astype(np.int64)).t().contiguous(), n,
                                mean_interactions=mean_interactions_from_age(torch.from_numpy(g["age"].astype(np.int64))))
    assert (n, graph.num_edges) == (37518, 136378)
    deg = graph.row_ptr[1:] - graph.row_ptr[:-1]
    assert int(deg.max()) == 36 and int((deg == 0).sum()) == 6481
    assert graph.i_bar_agent == 5488 and graph.i_bar == 2.0


def test_synthetic_graph_is_seeded_and_shards_agree_with_the_full_graph():
    a = synthetic_contact_graph(6000, 5.0, 11, family="comm", block=512)
    b = synthetic_contact_graph(6000, 5.0, 11, family="comm", block=512)
    assert torch.equal(a.col_idx, b.col_idx) and torch.equal(a.row_ptr, b.row_ptr)
    assert abs(a.avg_degree() - 5.0) < 0.2
    for r0, r1 in shard_bounds(6000, 2, align=1024):
        part = synthetic_contact_graph(6000, 5.0, 11, family="comm", block=512, row_begin=r0, row_end=r1)
        ref = shard_graph(a, r0, r1)
        assert part.row_offset == r0 and torch.equal(part.row_ptr, ref.row_ptr)
        assert torch.equal(part.col_idx[: part.num_edges], ref.col_idx[: ref.num_edges])
    age, st = synthetic_population(20000, 5)
    assert set(age.unique().tolist()) <= set(range(6)) and 0.005 < float(st.mean()) < 0.015


def test_shard_bounds_are_word_aligned_and_cover():
    b = shard_bounds(8_400_000, 8)
    assert b[0][0] == 0 and b[-1][1] == 8_400_000 and all(lo % 1024 == 0 for lo, _ in b)
    assert len({hi - lo for lo, hi in b[:-1]}) == 1
    assert choose_block_size(8_400_000, 1_100_000)[0] == 8


def test_noise_twins_agree_bitwise():
    ids = torch.arange(50000)
    for seed in (0, 7, 2**63 + 5, 2**64 - 1):
        for t in (0, 3, 20):
            assert np.array_equal(tnoise.exposure_uniform(seed, t, ids).numpy(), onoise.exposure_uniform(seed, t, ids.numpy()))
            for mode in ("soft", "hard", "none"):
                assert np.array_equal(tnoise.isolation_decision(seed, t, ids, mode).numpy(),
                                      onoise.isolation_decision(seed, t, ids.numpy(), mode))
    assert int(onoise.noise_u64(0, 0, [0])[0]) == 0xE220A8397B1DCDAF        # SplitMix64 known answer (seed 0, first output)
    u = onoise.exposure_uniform(3, 1, np.arange(200000))
    assert u.min() >= 0.0 and u.max() < 1.0 and abs(u.mean() - 0.5) < 5e-3


def test_contact_graph_file_round_trip_and_edge_list_cache(tmp_path, monkeypatch):
    import pandas as pd

    from agenttorch_b200.graph import build_contact_graph, load_contact_graph, save_contact_graph
    from agenttorch_b200.models.covid.substeps import utils

    rng = np.random.default_rng(3)
    n, e = 500, 4000
    edges = np.stack([rng.integers(0, n, e), rng.integers(0, n, e)], 1)
    g = build_contact_graph(torch.from_numpy(edges.T.copy()), n, mean_interactions=torch.full((n,), 3.0))
    path = str(tmp_path / "graph.atg")
    save_contact_graph(g, path)
    h = load_contact_graph(path)
    assert (h.n_rows, h.n_sources, h.num_edges, h.i_bar, h.i_bar_agent) == (g.n_rows, g.n_sources, g.num_edges, g.i_bar, g.i_bar_agent)
    assert torch.equal(h.row_ptr, g.row_ptr) and torch.equal(h.col_idx, g.col_idx) and h.edge_weight is None
    torch.save({"version": 99}, path)
    with pytest.raises(ValueError, match="version"):
        load_contact_graph(path)

    csv = tmp_path / "0.csv"
    pd.DataFrame(edges).to_csv(csv, header=False, index=False)
    monkeypatch.setenv("AGENTTORCH_B200_CACHE_DIR", str(tmp_path / "cache"))
    _, (ei1, attr1) = utils.network_from_file({"file_path": str(csv)})
    assert np.array_equal(ei1.numpy(), edges.T) and attr1.shape == (2, e) and bool((attr1 == 1).all())
    monkeypatch.setattr(utils.pd, "read_csv", lambda *a, **k: (_ for _ in ()).throw(AssertionError("the cache should have been used")))
    _, (ei2, _) = utils.network_from_file({"file_path": str(csv)})
    assert torch.equal(ei1, ei2) and ei2.dtype == torch.int64
    monkeypatch.undo()
    csv.write_text("1,2\n3,4\n")                                   # the CSV changed: the stale cache entry is not used
    monkeypatch.setenv("AGENTTORCH_B200_CACHE_DIR", str(tmp_path / "cache"))
    _, (ei3, _) = utils.network_from_file({"file_path": str(csv)})
    assert ei3.tolist() == [[1, 3], [2, 4]]
