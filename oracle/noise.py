"""Counter-based noise shared by the parity tests (numpy restatement).

TEST INFRASTRUCTURE.  The reference draws its randomness from torch's global
generator (`torch.rand(N,1)` in new_transmission/action.py:54 and
`torch.bernoulli` in core/distributions/distributions.py:16); that stream cannot
be reproduced by a fused kernel, so parity runs INJECT the same uniforms into
both paths.  The uniforms are SplitMix64 outputs addressed by
(seed, step, agent): the CUDA kernels regenerate them in registers
(agenttorch_b200/csrc/at_seirm.cu, `at_noise_u64`), this module regenerates them
with numpy, and the reference harness feeds them through the reference's own
`behavior` / `st_bernoulli` hooks.
"""
import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_INV24 = np.float32(1.0 / 16777216.0)


def noise_u64(seed, step, agent_ids):
    """SplitMix64 output number ((step << 40) | agent) of the stream seeded `seed`."""
    ids = np.asarray(agent_ids, dtype=np.uint64)
    with np.errstate(over="ignore"):
        ctr = (np.uint64(step) << np.uint64(40)) | ids
        z = (ctr + np.uint64(1)) * _GOLD + np.uint64(seed)
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return z


def exposure_uniform(seed, step, agent_ids):
    """U in [0,1): top 24 bits of the hash, as float32 (exact k * 2^-24)."""
    z = noise_u64(seed, step, agent_ids)
    return (z >> np.uint64(40)).astype(np.float32) * _INV24


def isolation_uniform(seed, step, agent_ids):
    """Second uniform from bits 8..31 of the same hash."""
    z = noise_u64(seed, step, agent_ids)
    return ((z >> np.uint64(8)) & np.uint64(0xFFFFFF)).astype(np.float32) * _INV24


def isolation_decision(seed, step, agent_ids, mode="soft", p_isolate=0.3):
    """W: 'soft' = uniform in [0,1) (reference default, action.py:54);
    'hard' = 1.0 with probability p_isolate else 0.0; 'none' = 0."""
    if mode == "none":
        return np.zeros(len(np.asarray(agent_ids)), dtype=np.float32)
    w = isolation_uniform(seed, step, agent_ids)
    if mode == "soft":
        return w
    if mode == "hard":
        return (w < np.float32(p_isolate)).astype(np.float32)
    raise ValueError(mode)


def move_draw(seed, step, species, agent_ids):
    """24-bit draw that picks a predator-prey agent's move: neighbour number (draw * n_neighbours) >> 24.
    Species 0 = predator, 1 = prey; the species is bit 39 of the counter's agent field."""
    ids = np.asarray(agent_ids, dtype=np.uint64) | (np.uint64(species) << np.uint64(39))
    return (noise_u64(seed, step, ids) >> np.uint64(40)).astype(np.int64)


# ----------------------------------------------------------------- intervention substeps
# Streams live in bits 36..38 of the counter's agent field (agent ids stay below 2^36):
STREAM_QUARANTINE, STREAM_TESTING, STREAM_MORTALITY = 1, 2, 3


def stream_uniforms(seed, step, agent_ids, stream):
    """Two uniforms per (seed, step, agent, stream): (top 24 bits, bits 8..31) of one hash, float32 k * 2^-24.
    quarantine: (start draw, break draw); testing: (compliance draw, unused); mortality: (death draw, unused)."""
    ids = np.asarray(agent_ids, dtype=np.uint64) | (np.uint64(stream) << np.uint64(36))
    z = noise_u64(seed, step, ids)
    return ((z >> np.uint64(40)).astype(np.float32) * _INV24,
            ((z >> np.uint64(8)) & np.uint64(0xFFFFFF)).astype(np.float32) * _INV24)
