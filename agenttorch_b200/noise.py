"""Counter-hash noise, torch (host or device) restatement of `at_noise_u64` in csrc/at_seirm.cu.

The step kernel regenerates these uniforms in registers (iso_mode / u_mode HASH); this module
materialises the same values as tensors for the tensor-mode API (e.g. feeding an identical
`isolation_decision` to a user-written substep).  int64 arithmetic wraps like uint64.
"""
import torch

_GOLD = -7046029254386353131          # 0x9E3779B97F4A7C15 as int64
_M1 = -4658895280553007687            # 0xBF58476D1CE4E5B9
_M2 = -7723592293110705685            # 0x94D049BB133111EB


def _lsr(z, k):
    """logical shift right on int64 tensors"""
    return (z >> k) & ((1 << (64 - k)) - 1)


def _as_i64(v):
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


def noise_u64(seed, step, agent_ids):
    ids = agent_ids.to(torch.int64)
    z = ((int(step) << 40) | ids) + 1
    z = z * _GOLD + _as_i64(int(seed))
    z = (z ^ _lsr(z, 30)) * _M1
    z = (z ^ _lsr(z, 27)) * _M2
    return z ^ _lsr(z, 31)


def exposure_uniform(seed, step, agent_ids):
    return _lsr(noise_u64(seed, step, agent_ids), 40).to(torch.float32) * (1.0 / 16777216.0)


def isolation_uniform(seed, step, agent_ids):
    return (_lsr(noise_u64(seed, step, agent_ids), 8) & 0xFFFFFF).to(torch.float32) * (1.0 / 16777216.0)


def isolation_decision(seed, step, agent_ids, mode="soft", p_isolate=0.3):
    if mode == "none":
        return torch.zeros(agent_ids.shape[0], dtype=torch.float32, device=agent_ids.device)
    w = isolation_uniform(seed, step, agent_ids)
    if mode == "soft":
        return w
    if mode == "hard":
        return (w < torch.tensor(p_isolate, dtype=torch.float32)).to(torch.float32)
    raise ValueError(mode)


STREAM_QUARANTINE, STREAM_TESTING, STREAM_MORTALITY = 1, 2, 3


def stream_uniforms(seed, step, agent_ids, stream):
    """(top 24 bits, bits 8..31) of the hash addressed (seed, step, agent | stream << 36): the draws of the
    intervention kernels (csrc/at_interventions.cu) as tensors."""
    z = noise_u64(seed, step, agent_ids.to(torch.int64) | (int(stream) << 36))
    return (_lsr(z, 40).to(torch.float32) * (1.0 / 16777216.0),
            (_lsr(z, 8) & 0xFFFFFF).to(torch.float32) * (1.0 / 16777216.0))
