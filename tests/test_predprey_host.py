"""Host-side logic of the predator-prey path that needs no GPU: config / Initializer / registry surface, the implicit
lattice against networkx and the oracle, the YAML output mapping, and the loud failure on CPU tensors."""
import numpy as np
import pytest
import torch

from agenttorch_b200.predprey import CsrAdjacency, GridLattice, PredPreySpec
from oracle.predprey_port import grid_neighbors


@pytest.mark.parametrize("shape", [(3, 3), (18, 25), (5, 2), (1, 6)])
def test_grid_lattice_matches_networkx_and_oracle(shape):
    nx = pytest.importorskip("networkx")
    lat = GridLattice(shape)
    dense_ref = torch.tensor(np.asarray(nx.adjacency_matrix(nx.grid_graph(dim=tuple(shape))).todense()))     # environment.py:6-9
    assert torch.equal(lat.to_dense(), dense_ref.to(torch.int64))
    cells = torch.arange(lat.n_cells)
    nbr, cnt = lat.neighbors(cells)
    n2, c2 = grid_neighbors(cells.numpy(), shape[0], lat.n_cells)
    assert np.array_equal(nbr.numpy(), n2) and np.array_equal(cnt.numpy(), c2)
    csr = CsrAdjacency(dense_ref, "cpu")
    for c in range(lat.n_cells):
        assert csr.idx[csr.ptr[c]: csr.ptr[c + 1]].tolist() == nbr[c, : cnt[c]].tolist()


def test_output_mapping_decides_which_work_is_charged():
    # reference config.yaml:281-285 lists pred_energy twice (prey keep their energy); config_vmap.yaml:283 does not
    assert PredPreySpec.work_flags(["prey_pos", "pred_energy", "pred_pos", "pred_energy"]) == (True, False)
    assert PredPreySpec.work_flags(["prey_pos", "prey_energy", "pred_pos", "pred_energy"]) == (True, True)


def test_packaged_config_initialises_with_the_reference_schema():
    from agenttorch_b200.models import predator_prey as pp
    from agenttorch_b200.models.predator_prey.synthetic import synthetic_config

    cfg = synthetic_config(12, 20, 30, 70, seed=3, device="cpu")
    assert list(cfg["substeps"].keys()) == ["0", "1", "2", "3"]
    assert [cfg["substeps"][s]["name"] for s in "0123"] == ["Move", "Eat", "Hunt", "Grow"]
    r = pp.get_runner(cfg, pp.registry)
    r.init()
    s = r.state
    assert s["agents"]["predator"]["coordinates"].shape == (30, 2) and s["agents"]["prey"]["energy"].shape == (70, 1)
    assert s["objects"]["grass"]["growth_stage"].shape == (240, 1) and s["environment"]["bounds"].tolist() == [12, 20]
    stage, cd = s["objects"]["grass"]["growth_stage"], s["objects"]["grass"]["growth_countdown"]
    assert torch.equal(stage == 1, cd <= 0)
    assert isinstance(s["network"]["agent_agent"]["predator_prey"]["adjacency_matrix"], GridLattice)
    for name, key in (("find_neighbors", "observation"), ("decide_movement", "policy"), ("update_positions", "transition"),
                      ("find_eatable_grass", "policy"), ("eat_grass", "transition"), ("find_targets", "policy"),
                      ("hunt_prey", "transition"), ("grow_grass", "transition")):
        assert name in pp.registry.helpers[key]
    # same seed -> same population; reset() rebuilds it
    a = s["agents"]["prey"]["coordinates"].clone()
    r.reset()
    assert torch.equal(r.state["agents"]["prey"]["coordinates"], a)
    # no CPU fallback: stepping a CPU state fails loudly instead of running a slow path
    with pytest.raises(RuntimeError, match="CUDA"):
        r.step(1)


def test_reference_file_helpers(tmp_path):
    from agenttorch_b200.core.helpers import grid_network, read_from_file

    p = tmp_path / "coords.csv"
    p.write_text("x, y\n3, 12\n1, 5\n")
    t = read_from_file([2, 2], {"file_path": str(p)})                      # header row skipped, dtype of the file kept
    assert t.dtype == torch.int64 and t.tolist() == [[3, 12], [1, 5]]
    g, lat = grid_network({"shape": [4, 6]})
    assert g is None and lat.n_cells == 24 and lat.width == 4


@pytest.mark.parametrize("shape", [(18, 25), (16, 8), (1, 9), (33, 31), (64, 64)])
def test_tiled_storage_index_is_an_injection_into_whole_tiles(shape):
    w, h = shape
    spec = PredPreySpec(w, h, w * h, w, tiled=True)
    sidx = spec.storage_index("cpu")
    assert sidx.shape == (w * h,) and int(sidx.min()) >= 0 and int(sidx.max()) < spec.n_storage
    assert torch.unique(sidx).numel() == w * h and spec.n_storage % 128 == 0
    # the four lattice neighbours of an interior cell share its 128-byte tile unless it sits on the tile's border
    k = torch.arange(w * h)
    row, col = k // w, k % w
    inner = (row % 8 > 0) & (row % 8 < 7) & (col % 16 > 0) & (col % 16 < 15) & (row + 1 < h) & (col + 1 < w)
    for d in (-w, -1, 1, w):
        kk = k[inner] + d
        assert bool(((sidx[kk] >> 7) == (sidx[k[inner]] >> 7)).all())
    assert PredPreySpec(w, h, w * h, w).storage_index("cpu") is None and PredPreySpec(w, h, w * h, w).n_storage == w * h
