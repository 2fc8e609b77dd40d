"""Pin the CPU restatement of the predator-prey step (oracle/predprey_port.py) against

  * outputs of the REAL reference run in the build container (tests/golden/predprey_*.npz, written by
    oracle/ref_harness/make_golden_predprey.py: Runner + substeps/{move,eat,hunt,grow}.py + grid_network), and
  * the known-answer vectors of the reference's own tests (tests/test_vmap_substeps.py:129-492).
"""
import numpy as np
import pytest

from oracle.predprey_port import PredPreyConstants, PredPreyPort, adjacency_neighbors, grid_neighbors
from tests.helpers import PREDPREY_CASES, PREDPREY_KEYS, load_predprey_golden, predprey_port_from_golden


@pytest.mark.parametrize("case", PREDPREY_CASES)
@pytest.mark.parametrize("use_adjacency", [False, True])
def test_port_matches_reference_trajectory(case, use_adjacency):
    g = load_predprey_golden(case)
    port = predprey_port_from_golden(g, use_adjacency)
    for t in range(g["steps"]):
        port.step(t, g["seed"])
        snap = port.snapshot()
        for k in PREDPREY_KEYS:                       # positions, energies (float32), grass: all bit-exact
            assert snap[k].dtype == g["traj_" + k].dtype, k
            assert np.array_equal(snap[k], g["traj_" + k][t]), (case, t, k)


def test_golden_covers_the_interesting_events():
    g = load_predprey_golden("predprey_18x25_starve")
    assert (g["traj_pred_energy"][-1] <= 0).all()                      # predators starve (and stop moving)
    dead = g["traj_pred_energy"][5].reshape(-1) <= 0
    assert dead.any() and np.array_equal(g["traj_pred_pos"][6][dead], g["traj_pred_pos"][5][dead])
    g = load_predprey_golden("predprey_18x25")
    e0 = g["traj_prey_energy"][0]
    assert ((e0 >= g["init_prey_energy"]) | (e0 == 0)).all()            # shipped config: moving costs prey nothing
    assert (g["traj_prey_energy"][-1] == 0).any()                      # hunted prey
    cd = g["traj_growth_countdown"]
    assert cd.dtype == np.float64 and (cd != np.round(cd)).any()       # float64 countdowns with fractional parts
    assert g["update_outputs"].count("pred_energy") == 2


@pytest.mark.parametrize("shape", [(3, 3), (2, 5), (5, 2), (18, 25), (7, 1)])
def test_grid_neighbors_match_networkx(shape):
    nx = pytest.importorskip("networkx")
    G = nx.grid_graph(dim=tuple(shape))                                 # core/helpers/environment.py:6-9
    A = np.asarray(nx.adjacency_matrix(G).todense())
    cells = np.arange(shape[0] * shape[1])
    n1, c1 = grid_neighbors(cells, shape[0], len(cells))
    n2, c2 = adjacency_neighbors(cells, A)
    assert np.array_equal(c1, c2)
    for i in cells:
        assert np.array_equal(n1[i, : c1[i]], n2[i, : c2[i]])


# ------------------------------------------------------------------ reference KATs (tests/test_vmap_substeps.py)
def _kat_port(pred_pos, pred_e, prey_pos, prey_e, stage, countdown, **consts):
    return PredPreyPort(PredPreyConstants(3, 3, **consts), pred_pos, pred_e, prey_pos, prey_e, stage, countdown)


def test_kat_update_positions_energy():           # test_vmap_substeps.py:196-226
    port = _kat_port([[0, 0], [1, 1], [2, 2]], [10.0, 5.0, 0.0], [[0, 1], [1, 0], [2, 1]], [8.0, 6.0, 4.0],
                     np.zeros(9), np.ones(9), pred_work=1.0, prey_work=1.0)
    port.move(0, 42)
    assert np.array_equal(port.prey_energy.reshape(-1), [7.0, 5.0, 3.0])
    assert np.array_equal(port.pred_energy.reshape(-1), [9.0, 4.0, 0.0 - 1.0])      # the KAT's dead predator had 0 -> stays put
    assert np.array_equal(port.pred_pos[2], [2, 2])                                  # :193 no energy, no move
    assert not np.array_equal(port.pred_pos[0], [0, 0]) and not np.array_equal(port.pred_pos[1], [1, 1])
    nbr, cnt = grid_neighbors([0], 3, 9)                                              # :141-147 cell (0,0) -> (0,1), (1,0)
    assert cnt[0] == 2 and nbr[0, :2].tolist() == [1, 3]


def test_kat_eat_grass():                         # test_vmap_substeps.py:229-345
    port = _kat_port(np.zeros((0, 2)), np.zeros(0), [[0, 0], [1, 1], [2, 2]], [10.0, 5.0, 2.0],
                     [1, 0, 1, 0, 1, 0, 1, 0, 1], np.array([0, 5, 0, 10, 0, 15, 0, 8, 0], dtype=np.float32),
                     grass_nutrition=5.0, regrowth_time=20.0)
    port.eat()
    assert np.array_equal(port.prey_energy.reshape(-1), [15.0, 10.0, 7.0])
    assert np.array_equal(port.growth_stage.reshape(-1), [0, 0, 1, 0, 0, 0, 1, 0, 0])
    assert np.array_equal(port.growth_countdown.reshape(-1), [20, 5, 0, 10, 20, 15, 0, 8, 20])


def test_kat_hunt_prey():                         # test_vmap_substeps.py:348-437
    port = _kat_port([[0, 0], [1, 1], [2, 2]], [10.0, 5.0, 2.0], [[0, 0], [1, 1], [0, 1], [2, 1]], [8.0, 6.0, 4.0, 3.0],
                     np.zeros(9), np.ones(9), prey_nutrition=15.0)
    port.hunt()
    assert np.array_equal(port.prey_energy.reshape(-1), [0.0, 0.0, 4.0, 3.0])
    assert np.array_equal(port.pred_energy.reshape(-1), [25.0, 20.0, 2.0])


def test_kat_grow_grass():                        # test_vmap_substeps.py:440-488
    port = _kat_port(np.zeros((0, 2)), np.zeros(0), np.zeros((0, 2)), np.zeros(0), [0, 0, 0, 0, 1, 0, 1, 0, 1],
                     np.array([5, 3, 1, 0, 0, 10, 0, 7, 0], dtype=np.float32))
    port.grow()
    assert np.array_equal(port.growth_countdown.reshape(-1), [4, 2, 0, -1, -1, 9, -1, 6, -1])
    assert np.array_equal(port.growth_stage.reshape(-1), [0, 0, 1, 1, 1, 0, 1, 0, 1])
