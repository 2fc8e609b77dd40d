"""StraightThroughBernoulli (reference: agent_torch/core/distributions/distributions.py:8-24).
Forward samples torch.bernoulli(p); backward is the identity.  Unlike the reference the backward
stays on p's device (the reference multiplies by a CPU ones tensor, which breaks CUDA — SURVEY Q7)."""
import torch


class StraightThroughBernoulli(torch.autograd.Function):
    generate_vmap_rule = True

    @staticmethod
    def forward(p):
        return torch.bernoulli(p)

    @staticmethod
    def setup_context(ctx, inputs, output):
        pass

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output
