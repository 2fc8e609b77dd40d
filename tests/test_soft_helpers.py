"""core.soft: the reference's straight-through helpers (agent_torch/core/helpers/soft.py:6-76) — hard values forward,
surrogate gradients backward — and its initialisation helpers (core/helpers/initializer.py:8-22).  Known answers worked out
by hand from the reference's formulas (checked bitwise against the reference functions in the build container)."""
import math

import pytest
import torch

from agenttorch_b200.core import helpers as h
from agenttorch_b200.core import substep


def test_compare_max_min_values_and_gradients():
    a = torch.tensor([2.0, -1.0, 0.5], requires_grad=True)
    b = torch.tensor([1.0, 3.0, 0.5], requires_grad=True)
    c = h.compare(a, b)
    assert torch.allclose(c, torch.tensor([1.0, 0.0, 0.0]), atol=1e-6)                 # a > b, ties are 0
    c.sum().backward()
    sig = torch.sigmoid(0.8 * (a - b)).detach()
    assert torch.allclose(a.grad, 0.8 * sig * (1 - sig)) and torch.allclose(b.grad, -0.8 * sig * (1 - sig))
    assert torch.allclose(h.max(a, b).detach(), torch.tensor([2.0, 3.0, 0.0]), atol=1e-5)   # a tie yields 0 (reference quirk)
    assert torch.allclose(h.min(a, b).detach(), torch.tensor([1.0, -1.0, 0.0]), atol=1e-5)


def test_logic_truncates_to_integers_and_passes_surrogate_gradients():
    a = torch.tensor([0.0, 1.0, 0.4, 2.0, -1.0], requires_grad=True)
    b = torch.tensor([1.0, 1.0, 0.9, 0.0, 0.0], requires_grad=True)
    # truth values are int(x) != 0: 0.4 and 0.9 count as false
    assert h.logical_not(a).tolist() == pytest.approx([1.0, 0.0, 1.0, 0.0, 0.0])
    assert h.logical_or(a, b).tolist() == pytest.approx([1.0, 1.0, 0.0, 1.0, 1.0])
    assert h.logical_and(a, b).tolist() == pytest.approx([0.0, 1.0, 0.0, 0.0, 0.0])
    h.logical_and(a, b).sum().backward()
    assert torch.equal(a.grad, b.detach()) and torch.equal(b.grad, a.detach())           # d(a*b)
    a.grad = None
    h.logical_not(a).sum().backward()
    assert torch.equal(a.grad, -torch.ones(5))                                            # d(1 - a)
    # grad=False: plain bool tensors
    assert h.logical_or(a, b, grad=False).dtype == torch.bool and h.logical_not(torch.tensor([0, 3])).tolist() == [1, 0]
    assert h.logical_not(torch.tensor([0, 3])).dtype == torch.int64                       # long in, long out (bool + long zeros)


def test_discrete_sample_law_shape_and_gradient():
    torch.manual_seed(1)
    p = torch.tensor([0.3], requires_grad=True)
    x = h.discrete_sample(p, 20000)
    assert x.shape == (20000,) and set(x.detach().unique().tolist()) <= {0.0, 1.0}
    assert abs(float(x.mean()) - 0.3) < 4 * math.sqrt(0.3 * 0.7 / 20000)
    x.sum().backward()
    assert p.grad is not None and torch.isfinite(p.grad).all()
    assert float(h.discrete_sample(torch.tensor([1.0 - 1e-7]), 100).sum()) == 100.0


def test_reexports_and_initialisation_helpers():
    for name in ("compare", "logical_and", "logical_or", "logical_not", "discrete_sample", "get_by_path", "read_config"):
        assert hasattr(substep, name), name                       # `from agent_torch.core.substep import *` habits keep working
    z = h.zeros([2, 3], {"dtype": "int"})
    assert z.dtype == torch.int32 and z.shape == (2, 3) and h.zeros([1], {"dtype": "float"}).dtype == torch.float32
    assert h.constant([2, 1], {"value": 2.5}).tolist() == [[2.5], [2.5]]
    assert h.random_normal([4], {"mu": 1.5, "sigma": 0.0}).tolist() == [1.5] * 4


def test_random_graph_helpers_return_the_reference_payload():
    import random

    for fn, params, edges in ((h.small_world, {"n": 30, "k": 4, "p": 0.0}, 60), (h.scale_free, {"num_agents": 20, "m": 2}, None),
                              (h.erdos_renyi, {"n": 12, "p": 1.0}, 66)):
        random.seed(0)
        graph, (edge_index, edge_attr) = fn(params)
        assert edge_index.dtype == torch.int64 and edge_index.shape[0] == 2 and edge_attr.shape == (edge_index.shape[1], 1)
        assert graph.number_of_edges() == edge_index.shape[1] and (edges is None or edge_index.shape[1] == edges)
        assert bool((edge_attr == 1).all())
