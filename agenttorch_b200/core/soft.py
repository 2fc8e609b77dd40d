"""Straight-through ("soft") comparison, logic and sampling helpers with the reference's names and semantics
(agent_torch/core/helpers/soft.py:6-76), plus its small initialisation helpers (core/helpers/initializer.py:8-22).

User-written substeps import these from `core.helpers` / `core.substep` (the reference's quarantine and testing classes do);
they are plain elementwise torch expressions on whatever device their inputs live on.  Each function returns the HARD value in
the forward pass and the gradient of a smooth surrogate in the backward pass: `hard + (soft - soft.detach())`.  (The covid
substeps shipped in this package do not call them: their draws and logic run inside the CUDA kernels.)"""
import torch

__all__ = ["compare", "max", "min", "logical_not", "logical_or", "logical_and", "discrete_sample", "zeros", "constant",
           "random_normal", "dtype_dict"]


def _straight_through(hard, soft):
    return hard + (soft - soft.detach())


def compare(a, b):
    """1.0 where a > b; gradient of sigmoid(0.8 (a - b)).  The reference sums left to right, `(hard + soft) - soft.detach()`
    (soft.py:13-14), so the value carries one float32 rounding of `hard + soft`; kept for bitwise-equal results."""
    soft = torch.sigmoid(0.8 * (a - b))
    return ((a > b).float() + soft) - soft.detach()


def max(a, b):  # noqa: A001  (the reference shadows the builtin as well)
    return a * compare(a, b) + b * compare(b, a)


def min(a, b):  # noqa: A001
    return a * compare(b, a) + b * compare(a, b)


def logical_not(a, grad=True):
    """not int(a); gradient of 1 - a.  The truth value is taken after truncation to an integer, as the reference does."""
    hard = torch.logical_not(a.long())
    return _straight_through(hard, 1 - a) if grad else hard


def logical_or(a, b, grad=True):
    """int(a) or int(b); gradient of a + b."""
    hard = torch.logical_or(a.long(), b.long())
    return _straight_through(hard, a + b) if grad else hard


def logical_and(a, b, grad=True):
    """int(a) and int(b); gradient of a * b."""
    hard = torch.logical_and(a.long(), b.long())
    return _straight_through(hard, a * b) if grad else hard


def discrete_sample(sample_prob, size, device="cpu", hard=True):
    """Bernoulli(sample_prob) draws of shape `size` through a two-class Gumbel-softmax at temperature 0.1: the value is 1.0
    with probability sample_prob, the gradient flows to sample_prob (soft.py:70-76)."""
    probs = sample_prob * torch.ones(size).to(device)
    logits = torch.vstack((probs, 1.0 - probs)).transpose(0, 1).log()
    return torch.nn.functional.gumbel_softmax(logits, tau=0.1, hard=hard)[:, 0]


# --------------------------------------------------------------------------------- initialisation helpers
dtype_dict = {"int": torch.int32, "float": torch.float32}


def zeros(shape, params):
    return torch.zeros(size=list(shape), dtype=dtype_dict[params["dtype"]])


def constant(shape, params):
    return params["value"] * torch.ones(size=list(shape))


def random_normal(shape, params):
    return params["mu"] + params["sigma"] * torch.randn(size=list(shape))
