"""CPU restatement (numpy) of the reference predator-prey step.

TEST INFRASTRUCTURE — imported only by tests/, __graft_entry__.smoke() and bench/probe CPU baselines; the product
(agenttorch_b200/) never imports it.  Pinned against the UNMODIFIED reference run in this container
(oracle/ref_harness/run_predprey_reference.py -> tests/golden/predprey_*.npz, checked by tests/test_oracle_predprey.py)
and against the known-answer tests of the reference's tests/test_vmap_substeps.py:129-492.

One step = four substeps (reference models/predator_prey/config.yaml:222-372):
  0 move   substeps/move.py:49-123   neighbours of the agent's cell from the adjacency matrix (:24-46, ascending
           node order), `random.choice` among them if energy > 0 (:88-90), then energy -= stride_work for every
           agent (:112-116).  Which energies are actually written back follows the transition's YAML
           `output_variables` (controller.py:73-81): the shipped config.yaml:281-285 lists pred_energy twice, so
           prey energy is NOT reduced there; config_vmap.yaml:283 lists prey_energy.
  1 eat    substeps/eat.py:24-110    every prey (dead ones too, :81 TODO) standing on a cell with growth_stage == 1
           gains the grass nutrition once (:93-97 OR-ed masks); that cell's stage drops to 0 and its countdown
           becomes c + float32(regrowth_time - c) (:71-72 float32 masks, :99,:105).
  2 hunt   substeps/hunt.py:24-92    a cell holding >= 1 predator and >= 1 prey (dead or alive) is a target: its
           prey get energy 0 (:81), its predators gain the prey nutrition once (:83-86).
  3 grow   substeps/grow.py:24-46    countdown -= 1; growth_stage = (countdown <= 0).

Grid topology (core/helpers/environment.py:6-9): `nx.grid_graph(dim=(max_x, max_y))` orders its nodes so that node k
is adjacent to k +- max_x and, inside a run of max_x consecutive nodes, to k +- 1; the substeps number cells as
node = max_y * x + y (move.py:36).  For square grids both agree; for the bundled 18x25 data the lattice has seams.
Both are reproduced as they are (`grid_neighbors` is checked against networkx in the tests).

Noise: the draw of agent i of a species at step t is oracle.noise.move_draw(seed, t, species, i), a 24-bit integer;
the chosen neighbour is number (draw * n_neighbours) >> 24 of the ascending list.
"""
import numpy as np

from . import noise


def grid_neighbors(cells, width, n_cells):
    """Ascending neighbour lists of `cells` on the lattice nx.grid_graph(dim=(width, n_cells // width)) builds.
    Returns (nbr (n,4) int64, count (n,) int64); unused slots hold -1."""
    k = np.asarray(cells, dtype=np.int64)
    i = k % width
    cand = np.stack([k - width, k - 1, k + 1, k + width], axis=1)
    ok = np.stack([k - width >= 0, i > 0, i < width - 1, k + width < n_cells], axis=1)
    order = np.argsort(~ok, axis=1, kind="stable")                 # valid candidates first, ascending order kept
    nbr = np.where(np.take_along_axis(ok, order, 1), np.take_along_axis(cand, order, 1), -1)
    return nbr, ok.sum(1)


def adjacency_neighbors(cells, adjacency):
    """Same from a dense adjacency matrix (move.py:38-46): row `cell`, entries == 1, ascending."""
    adjacency = np.asarray(adjacency)
    cells = np.asarray(cells, dtype=np.int64)
    nbr = np.full((len(cells), int((adjacency == 1).sum(1).max(initial=0))), -1, dtype=np.int64)
    cnt = np.zeros(len(cells), dtype=np.int64)
    for n, c in enumerate(cells):
        idx = np.nonzero(adjacency[c] == 1)[0]
        nbr[n, : len(idx)] = idx
        cnt[n] = len(idx)
    return nbr, cnt


class PredPreyConstants:
    def __init__(self, max_x, max_y, pred_work=1.0, prey_work=5.0, prey_nutrition=20.0, grass_nutrition=7.0,
                 regrowth_time=100.0, update_outputs=("prey_pos", "prey_energy", "pred_pos", "pred_energy")):
        self.max_x, self.max_y = int(max_x), int(max_y)
        self.pred_work, self.prey_work = np.float32(pred_work), np.float32(prey_work)
        self.prey_nutrition, self.grass_nutrition = np.float32(prey_nutrition), np.float32(grass_nutrition)
        self.regrowth_time = np.float32(regrowth_time)
        self.update_outputs = tuple(update_outputs)


class PredPreyPort:
    """State arrays follow the reference's shapes: positions (n,2) int64 [x, y], energies (n,1) float32,
    growth_stage (cells,1), growth_countdown (cells,1) float32/float64."""

    def __init__(self, consts, pred_pos, pred_energy, prey_pos, prey_energy, growth_stage, growth_countdown,
                 adjacency=None):
        self.c = consts
        self.pred_pos = np.array(pred_pos, dtype=np.int64).reshape(-1, 2)
        self.prey_pos = np.array(prey_pos, dtype=np.int64).reshape(-1, 2)
        self.pred_energy = np.array(pred_energy, dtype=np.float32).reshape(-1, 1)
        self.prey_energy = np.array(prey_energy, dtype=np.float32).reshape(-1, 1)
        self.growth_stage = np.array(growth_stage).reshape(-1, 1).astype(np.int64)
        gc = np.array(growth_countdown).reshape(-1, 1)
        self.growth_countdown = gc.astype(np.float32 if gc.dtype == np.float32 else np.float64)
        self.n_cells = self.growth_stage.shape[0]
        self.adjacency = adjacency

    # ------------------------------------------------------------------------------------------ helpers
    def node(self, pos):
        return self.c.max_y * pos[:, 0] + pos[:, 1]                                   # move.py:36, eat.py:40

    def coords(self, idx):
        y = idx % self.c.max_y                                                        # move.py:43-44
        return np.stack([(idx - y) // self.c.max_y, y], axis=1)

    def neighbors(self, cells):
        if self.adjacency is not None:
            return adjacency_neighbors(cells, self.adjacency)
        return grid_neighbors(cells, self.c.max_x, self.n_cells)

    # ------------------------------------------------------------------------------------------ substeps
    def move(self, t, seed):
        nxt = {}
        for species, pos, energy in ((0, self.pred_pos, self.pred_energy), (1, self.prey_pos, self.prey_energy)):
            nbr, cnt = self.neighbors(self.node(pos))
            draw = noise.move_draw(seed, t, species, np.arange(len(pos)))
            pick = (draw * cnt) >> 24
            target = nbr[np.arange(len(pos)), np.minimum(pick, np.maximum(cnt - 1, 0))]
            alive = energy.reshape(-1) > 0                                            # move.py:89
            nxt[species] = np.where(alive[:, None], self.coords(target), pos)
        results = {}                                                                  # move.py:118-123 + controller.py:73-81
        for key, val in zip(self.c.update_outputs, (nxt[1], self.prey_energy + (-self.c.prey_work),
                                                    nxt[0], self.pred_energy + (-self.c.pred_work))):
            results[key] = val
        for key, val in results.items():
            setattr(self, {"prey_pos": "prey_pos", "prey_energy": "prey_energy", "pred_pos": "pred_pos",
                           "pred_energy": "pred_energy"}[key], val)

    def eat(self):
        node = self.node(self.prey_pos)
        grown = self.growth_stage.reshape(-1)[node] == 1                              # eat.py:41
        if not grown.any():
            return
        cells = np.unique(node[grown])
        self.prey_energy = np.where(grown[:, None], self.prey_energy + self.c.grass_nutrition, self.prey_energy)
        self.growth_stage[cells, 0] += -1                                             # eat.py:98,104
        cd = self.growth_countdown
        reset = (self.c.regrowth_time.astype(cd.dtype) - cd[cells, 0]).astype(np.float32)   # float32 mask, eat.py:72,99
        cd[cells, 0] = cd[cells, 0] + reset.astype(cd.dtype)

    def hunt(self):
        prey_node, pred_node = self.node(self.prey_pos), self.node(self.pred_pos)
        size = max(self.n_cells, int(prey_node.max(initial=0)) + 1, int(pred_node.max(initial=0)) + 1)
        has_prey = np.zeros(size, dtype=bool)
        has_prey[prey_node] = True
        has_pred = np.zeros(size, dtype=bool)
        has_pred[pred_node] = True
        target = has_prey & has_pred                                                  # hunt.py:38-40
        self.prey_energy = np.where(target[prey_node][:, None], np.float32(0), self.prey_energy)
        self.pred_energy = np.where(target[pred_node][:, None], self.pred_energy + self.c.prey_nutrition, self.pred_energy)

    def grow(self):
        self.growth_countdown = self.growth_countdown + self.growth_countdown.dtype.type(-1)     # grow.py:34-35
        self.growth_stage = (self.growth_countdown <= 0).astype(np.int64)             # grow.py:39-40

    def step(self, t, seed):
        self.move(t, seed)
        self.eat()
        self.hunt()
        self.grow()

    def snapshot(self):
        return {k: getattr(self, k).copy() for k in
                ("pred_pos", "pred_energy", "prey_pos", "prey_energy", "growth_stage", "growth_countdown")}
