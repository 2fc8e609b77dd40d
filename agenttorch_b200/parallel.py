"""Multi-GPU plumbing of the SEIRM path: one process per GPU (torchrun), contiguous agent ownership.

The only per-step cross-agent data is the infectiousness of the source agents for the NEXT step: one code byte and
one "hot" bit per agent.  Each rank writes the slice of its own agents (the step kernel does) and `halo_exchange`
makes all slices visible everywhere with in-place all-gathers.  The reference has no equivalent: its
`_handle_cross_gpu_interactions` is a stub (agent_torch/core/distributed_runner.py:218-228)."""
import torch
import torch.distributed as dist


def halo_exchange(code, hot_words, n_code, group=None):
    """code: (n_code,) uint8; hot_words: (>= n_code/32,) int32 or None (the engine rebuilds the bitmap from the
    gathered codes on the device) — this rank's slice already written in place.
    n_code must be world_size equal slices (graph.shard_bounds).  No-op without a process group."""
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    rank = dist.get_rank(group)
    if n_code % (world * 32):
        raise ValueError("n_code must split into world_size slices of whole bitmap words")
    per = n_code // world
    dist.all_gather_into_tensor(code[:n_code], code[rank * per:(rank + 1) * per], group=group)
    if hot_words is None:
        return
    words = hot_words[: n_code // 32]
    wper = per // 32
    dist.all_gather_into_tensor(words, words[rank * wper:(rank + 1) * wper], group=group)


def reduce_episode_outputs(tensors, group=None):
    """Per-episode all-reduce of the daily series / gradients (partial sums over the rank's agents)."""
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return tensors
    if dist.get_world_size(group) > 1:
        for t in tensors:
            dist.all_reduce(t, group=group)
    return tensors


def reduce_daily_series(state, group=None):
    """(daily_infected, daily_deaths) summed over the ranks with ONE small all-reduce (the two series packed)."""
    di, dd = state.daily_infected, state.daily_deaths
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return di.clone(), dd.clone()
    packed = torch.cat([di.reshape(-1), dd.reshape(-1)])
    if dist.get_world_size(group) > 1:
        dist.all_reduce(packed, group=group)
    n = di.numel()
    return packed[:n].reshape(di.shape), packed[n:].reshape(dd.shape)          # views of the fresh packed tensor
