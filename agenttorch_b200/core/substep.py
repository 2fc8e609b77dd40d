"""Substep base classes (reference: agent_torch/core/substep.py:13-116).

Constructor contract (called by Initializer.substeps): cls(config, input_variables, output_variables,
arguments={"learnable": {...}, "fixed": {...}}).  Learnable arguments become an nn.ParameterDict and,
when simulation_metadata.calibration is true, leaf tensors `calibrate_<name>` (substep.py:26-30)."""
from abc import ABC, abstractmethod

import torch
import torch.nn as nn

from .helpers import *  # noqa: F401,F403  (the reference re-exports its helpers here)
from .vmap import sample_grid, vmap  # noqa: F401


class _SubstepBase(nn.Module, ABC):
    def __init__(self, config, input_variables, output_variables, arguments):
        super().__init__()
        self.config = config
        self.input_variables = input_variables
        self.output_variables = output_variables
        arguments = arguments or {"learnable": {}, "fixed": {}}
        self.learnable_args, self.fixed_args = arguments["learnable"], arguments["fixed"]
        if self.learnable_args:
            self.learnable_args = nn.ParameterDict(self.learnable_args)
        # config may be None when a substep is built stand-alone (the reference's tests/test_vmap_substeps.py do so)
        if self.config is not None and self.config["simulation_metadata"]["calibration"] == True:  # noqa: E712
            for key, value in self.learnable_args.items():
                setattr(self, f"calibrate_{key}", value.detach().clone().requires_grad_(True))
        self.args = {**self.fixed_args, **self.learnable_args}


class SubstepObservation(_SubstepBase):
    def __init__(self, config, input_variables, output_variables, arguments):
        super().__init__(config, input_variables, output_variables, arguments)
        self.custom_observation_network = None

    @abstractmethod
    def forward(self, state):
        pass


class SubstepAction(_SubstepBase):
    def __init__(self, config, input_variables, output_variables, arguments):
        super().__init__(config, input_variables, output_variables, arguments)
        self.custom_action_network = None

    @abstractmethod
    def forward(self, state, observation):
        pass


class SubstepTransition(_SubstepBase):
    def __init__(self, config, input_variables, output_variables, arguments):
        super().__init__(config, input_variables, output_variables, arguments)
        self.custom_transition_network = None

    @abstractmethod
    def forward(self, state, action):
        pass


class SubstepTransitionMessagePassing(_SubstepBase):
    """Message-passing transition with aggr="add" (substep.py:95-116).  The reference inherits PyG's
    MessagePassing; this class provides the same `propagate` contract natively: arguments named `x_i`
    are gathered at edge_index[1], `x_j` at edge_index[0], `message()` results are summed onto
    edge_index[1].  (The SEIRM transmission does not use it: its gather runs in the fused CUDA kernel.)"""

    def __init__(self, config, input_variables, output_variables, arguments):
        super().__init__(config, input_variables, output_variables, arguments)
        self.aggr = "add"

    def propagate(self, edge_index, size=None, **kwargs):
        import inspect

        src, dst = edge_index[0], edge_index[1]
        call, num_nodes = {}, None
        for name in inspect.signature(self.message).parameters:
            if name.endswith("_i") or name.endswith("_j"):
                base = kwargs[name[:-2]]
                num_nodes = base.size(0)
                call[name] = base.index_select(0, dst if name.endswith("_i") else src)
            else:
                call[name] = kwargs[name]
        msg = self.message(**call)
        out = torch.zeros((num_nodes,) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device)
        return out.index_add(0, dst, msg)

    def message(self, x_j):
        return x_j

    @abstractmethod
    def forward(self, state, action):
        pass
