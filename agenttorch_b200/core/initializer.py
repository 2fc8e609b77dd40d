"""Initializer: builds the nested state dict and the substep ModuleDicts from a resolved config
(reference: agent_torch/core/initializer.py:9-444).  Same YAML schema, same registry look-ups:
initialization helpers `fn(shape, params) -> Tensor`, network helpers `fn(params) -> (graph, adjacency)`,
substep classes keyed by the YAML dict key (initializer.py:347,368,385)."""
import torch
import torch.nn as nn


class Initializer(nn.Module):
    def __init__(self, config, registry):
        super().__init__()
        self.config = config
        self.registry = registry
        want = str(config["simulation_metadata"].get("device", "auto")).lower()
        if want == "auto":
            want = "cuda" if torch.cuda.is_available() else "cpu"
        self.device = torch.device(want)
        self.is_cuda = self.device.type == "cuda"
        self.config["simulation_metadata"]["device"] = self.device.type     # initializer.py:24
        self.state = {}
        self.environment, self.agents, self.objects, self.networks = {}, {}, {}, {}
        self.fixed_parameters, self.learnable_parameters = {}, {}
        self.observation_function, self.policy_function = nn.ModuleDict(), nn.ModuleDict()
        self.transition_function, self.reward_function = nn.ModuleDict(), nn.ModuleDict()

    # ---------------------------------------------------------------------------- values
    def _to_device(self, t):
        if not torch.is_tensor(t):
            return t
        if self.is_cuda and t.device.type == "cpu":
            return t.contiguous().pin_memory().to(self.device, non_blocking=True)
        return t.to(self.device)

    def _initialize_from_default(self, src_val, shape):
        """Literal `value:` -> tensor; float32 whatever the YAML dtype says (initializer.py:78-116)."""
        if type(src_val) == str:
            return src_val
        if isinstance(shape, int):            # predator_prey/config.yaml:31 writes `shape: 2`
            shape = [shape]
        if type(src_val) == list:
            src = torch.tensor(src_val)
            if shape and list(src.shape) != list(shape):
                if src.dim() == 1 and len(shape) >= 2 and src.shape[0] == shape[-1]:
                    val = src.unsqueeze(0).expand(shape)
                else:
                    val = torch.zeros(size=shape).fill_(src.flatten()[0].item())
            else:
                val = src
        else:
            val = src_val * torch.ones(size=shape)
        return self._to_device(val)

    def _initialize_from_generator(self, spec, shape, name_root):
        params = {}
        for arg, obj in spec["arguments"].items():
            if obj["initialization_function"] is not None:
                raise NotImplementedError("Dynamic argument initialization is not supported yet.")
            val = self._initialize_from_default(obj["value"], obj["shape"])
            params[arg] = val
            (self.learnable_parameters if obj["learnable"] else self.fixed_parameters)[f"{name_root}_{arg}"] = val
        return self._to_device(self.registry.initialization_helpers[spec["generator"]](shape, params))

    def _initialize_property(self, obj, key):
        init = obj["initialization_function"]
        if init is None:
            val = self._initialize_from_default(obj["value"], obj["shape"])
        else:
            val = self._initialize_from_generator(init, obj["shape"], key)
        (self.learnable_parameters if obj["learnable"] else self.fixed_parameters)[key] = val
        return val, obj["learnable"]

    # ----------------------------------------------------------------------------- state
    def init_environment(self, key="environment"):
        if self.config["state"][key] is None:
            return
        for prop, obj in self.config["state"][key].items():
            self.environment[prop], _ = self._initialize_property(obj, f"{key}_{prop}")

    def _init_instances(self, key, target):
        if self.config["state"][key] is None:
            return
        for inst, spec in self.config["state"][key].items():
            if inst == "metadata":
                continue
            target[inst] = {}
            props = spec["properties"]
            if props is None:
                continue
            for prop, obj in props.items():
                target[inst][prop], _ = self._initialize_property(obj, f"{key}_{inst}_{prop}")

    def init_agents(self, key="agents"):
        self._init_instances(key, self.agents)

    def init_objects(self, key="objects"):
        self._init_instances(key, self.objects)

    def init_network(self, key="network"):
        if self.config["state"][key] is None:
            return
        for itype, nets in self.config["state"][key].items():
            self.networks[itype] = {}
            if nets is None:
                continue
            for name, spec in nets.items():
                graph, adjacency = self.registry.network_helpers[spec["type"]](spec["arguments"])
                if isinstance(adjacency, (tuple, list)) and len(adjacency) == 2:
                    adjacency = (self._to_device(adjacency[0]), self._to_device(adjacency[1]))
                elif torch.is_tensor(adjacency):
                    adjacency = self._to_device(adjacency)
                self.networks[itype][name] = {"graph": graph, "adjacency_matrix": adjacency}

    def simulator(self):
        self.init_environment()
        self.init_agents()
        self.init_objects()
        self.init_network()
        self.parameters_dict = nn.ParameterDict(
            {k: v for k, v in self.learnable_parameters.items() if torch.is_tensor(v) and v.is_floating_point()})

    # -------------------------------------------------------------------------- substeps
    def _parse_function(self, fn_obj, name_root):
        learnable, fixed = {}, {}
        for arg, obj in (fn_obj["arguments"] or {}).items():
            if obj["initialization_function"] is None:
                val = self._initialize_from_default(obj["value"], obj["shape"])
            else:
                val = self._initialize_from_generator(obj["initialization_function"], obj["shape"], f"{name_root}_{arg}")
            if obj["learnable"]:
                self.learnable_parameters[f"{name_root}_{arg}"] = val
                learnable[arg] = val
            else:
                self.fixed_parameters[f"{name_root}_{arg}"] = val
                fixed[arg] = val
        return fn_obj["input_variables"], fn_obj["output_variables"], {"learnable": learnable, "fixed": fixed}

    def substeps(self):
        for substep, spec in self.config["substeps"].items():
            self.observation_function[substep] = nn.ModuleDict()
            self.policy_function[substep] = nn.ModuleDict()
            self.transition_function[substep] = nn.ModuleDict()
            for agent_type in spec["active_agents"]:
                for kind, table, helpers in (("observation", self.observation_function, self.registry.observation_helpers),
                                             ("policy", self.policy_function, self.registry.policy_helpers)):
                    table[substep][agent_type] = nn.ModuleDict()
                    fns = (spec.get(kind) or {}).get(agent_type)
                    if fns is None:
                        continue
                    for name, fn_obj in fns.items():
                        iv, ov, args = self._parse_function(fn_obj, f"{agent_type}_{kind}_{name}")
                        table[substep][agent_type][name] = helpers[name](self.config, iv, ov, args)
            for name, fn_obj in spec["transition"].items():
                iv, ov, args = self._parse_function(fn_obj, f"_transition_{name}")
                self.transition_function[substep][name] = self.registry.transition_helpers[name](self.config, iv, ov, args)

    # ------------------------------------------------------------------------ entry points
    def _publish(self):
        self.state["environment"] = self.environment
        self.state["network"] = self.networks
        self.state["agents"] = self.agents
        self.state["objects"] = self.objects
        self.state["parameters"] = self.parameters_dict

    def initialize(self):
        self.state["current_step"] = 0
        self.state["current_substep"] = "0"
        self.simulator()
        self.substeps()
        self._publish()

    def reset_state(self):
        """Fresh state tensors, substep modules (and their learnable parameters) kept (initializer.py:405-428)."""
        self.state["current_step"] = 0
        self.state["current_substep"] = "0"
        self.environment, self.agents, self.objects, self.networks = {}, {}, {}, {}
        self.fixed_parameters, self.learnable_parameters = {}, {}
        self.simulator()
        self._publish()

    def forward(self):
        self.initialize()

    def __getstate__(self):
        d = self.state.copy()
        params = getattr(self, "parameters_dict", None)
        if isinstance(params, nn.ParameterDict):
            d["parameters"] = params.state_dict()
        return d

    def __setstate__(self, state):
        nn.Module.__init__(self)
        self.parameters_dict = nn.ParameterDict(state.get("parameters", {}))
        self.state = state
        self.state["parameters"] = self.parameters_dict
