"""Registry — the plugin surface (reference: agent_torch/core/registry.py:7-43).

One table per kind of plug-in (transition | observation | policy | initialization | network), stored on the CLASS: every
instance, and every model package imported into the process, sees and edits the same tables — exactly the reference's
behaviour, which is why `Executor` re-asserts a model's registrations before it builds a runner."""
import json

import torch.nn as nn

KINDS = ("transition", "observation", "policy", "initialization", "network")
_SUFFIX = "_helpers"


class Registry(nn.Module):
    helpers = dict((kind, {}) for kind in KINDS)

    @classmethod
    def _table(cls, key):
        try:
            return cls.helpers[key]
        except KeyError:
            raise KeyError(f"unknown registry key '{key}' (expected one of {KINDS})") from None

    def __getattr__(self, name):
        # registry.transition_helpers, registry.policy_helpers, ...: the shared tables (registry.py:17-21)
        if name.endswith(_SUFFIX) and name[: -len(_SUFFIX)] in Registry.helpers:
            return Registry.helpers[name[: -len(_SUFFIX)]]
        return super().__getattr__(name)

    def register(self, obj_source, name, key):
        """registry.register(SubstepClass_or_function, "yaml_name", key="transition")"""
        self._table(key)[name] = obj_source

    @classmethod
    def register_helper(cls, name, key):
        """@Registry.register_helper("yaml_name", "initialization") / @Registry.register_substep("yaml_name", "policy")"""
        table = cls._table(key)

        def bind(obj):
            table[name] = obj
            return obj

        return bind

    register_substep = register_helper

    def view(self):
        """The registry as JSON text: names of the registered objects per kind."""
        names = {kind: {n: getattr(o, "__name__", str(o)) for n, o in table.items()} for kind, table in self.helpers.items()}
        return json.dumps(names, indent=2)

    def forward(self):
        print("Invoke registry.register(class_obj, key)")
