"""Shared pieces of the intervention substeps (quarantine, testing, SEIRMS progression)."""
import re

import torch

from agenttorch_b200.core.helpers import get_by_path


class EpisodeNoise:
    """Counter-hash noise address of a substep: `simulation_metadata.noise_seed` (or a random seed), advanced once per
    episode like NewTransmission.episode_seed, unless `advance_seed` is switched off (parity tests)."""

    def init_noise(self, meta):
        seed = meta.get("noise_seed")
        self.noise_seed = int(seed) if seed is not None else int(torch.randint(0, 2**62, (1,)).item())
        self.episode_index = 0
        self.advance_seed = True
        self.id_offset = 0                # global id of this process's agent 0

    def episode_seed(self, t):
        if t == 0:
            self.episode_index += 1
        v = self.noise_seed if not self.advance_seed else self.noise_seed + 0x9E3779B97F4A7C15 * self.episode_index
        v &= (1 << 64) - 1
        return v - (1 << 64) if v >= (1 << 63) else v


class ScalarCache:
    """float() of a one-element device tensor costs a device sync: read it once per (tensor object, version).  The tensor is
    kept referenced so that its storage address cannot be reused by another tensor while the entry is alive."""

    def __init__(self):
        self._seen = {}

    def __call__(self, name, tensor):
        if not torch.is_tensor(tensor):
            return float(tensor)
        hit = self._seen.get(name)
        if hit is None or hit[0] is not tensor or hit[1] != tensor._version:
            hit = (tensor, tensor._version, float(tensor.detach().reshape(-1)[0]))
            self._seen[name] = hit
        return hit[2]


def state_var(state, input_variables, name):
    return get_by_path(state, re.split("/", input_variables[name]))


def require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"agenttorch_b200 {what} needs CUDA tensors (simulation_metadata.device: cuda); there is no CPU fallback")
