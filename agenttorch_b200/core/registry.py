"""Registry — the plugin surface (reference: agent_torch/core/registry.py:7-43).

Tables are class-level and shared by every instance, keyed by
transition | observation | policy | initialization | network."""
import json

import torch.nn as nn

_KEYS = ("transition", "observation", "policy", "initialization", "network")


class Registry(nn.Module):
    helpers = {k: {} for k in _KEYS}

    def __init__(self):
        super().__init__()
        self.initialization_helpers = self.helpers["initialization"]
        self.observation_helpers = self.helpers["observation"]
        self.policy_helpers = self.helpers["policy"]
        self.transition_helpers = self.helpers["transition"]
        self.network_helpers = self.helpers["network"]

    def register(self, obj_source, name, key):
        if key not in self.helpers:
            raise KeyError(f"unknown registry key '{key}' (expected one of {_KEYS})")
        self.helpers[key][name] = obj_source

    def view(self):
        return json.dumps({k: {n: getattr(o, "__name__", str(o)) for n, o in v.items()} for k, v in self.helpers.items()}, indent=2)

    def forward(self):
        print("Invoke registry.register(class_obj, key)")

    @classmethod
    def register_helper(cls, name, key):
        def decorator(obj):
            cls.helpers[key][name] = obj
            return obj

        return decorator

    register_substep = register_helper
