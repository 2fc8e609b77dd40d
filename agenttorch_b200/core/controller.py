"""Controller: observe / act / progress (reference: agent_torch/core/controller.py:17-103).

`progress` gives transitions a shallow working copy of the state tree (new dicts, SAME tensors):
transitions return new tensors which are written back by path, so nothing aliases — the reference
instead clones every tensor and deep-copies the edge list each substep (copy_module, Q8)."""
import re

import torch.nn as nn

from .helpers import set_by_path


def _shallow_tree(tree):
    return {k: _shallow_tree(v) if isinstance(v, dict) else v for k, v in tree.items()}


class Controller(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.returns = []

    def _keys(self, substep, kind, agent_type):
        block = self.config["substeps"].get(substep, {}).get(kind)
        if not isinstance(block, dict):
            return None
        agent_map = block.get(agent_type) or {}
        return list(agent_map.keys()) if isinstance(agent_map, dict) else []

    def observe(self, state, observation_function, agent_type):
        substep = state["current_substep"]
        keys = self._keys(substep, "observation", agent_type)
        if not keys:
            return None
        out = {}
        for k in keys:
            out.update(observation_function[substep][agent_type][k](state))
        return out

    def act(self, state, observation, policy_function, agent_type):
        substep = state["current_substep"]
        keys = self._keys(substep, "policy", agent_type)
        if not keys:
            return None
        out = {}
        for k in keys:
            out.update(policy_function[substep][agent_type][k](state, observation))
        return out

    def _apply(self, state, action, transition_function):
        substep = state["current_substep"]
        n_sub = self.config["simulation_metadata"]["num_substeps_per_step"]
        state["current_substep"] = str((int(substep) + 1) % n_sub)
        spec = self.config["substeps"][substep]["transition"]
        for name in spec.keys():
            updated = transition_function[substep][name](state=state, action=action)
            for var_name, value in updated.items():
                set_by_path(state, re.split("/", spec[name]["input_variables"][var_name]), value)
        return state

    def progress(self, state, action, transition_function):
        return self._apply(_shallow_tree(state), action, transition_function)

    def progress_inplace(self, state, action, transition_function):
        return self._apply(state, action, transition_function)
