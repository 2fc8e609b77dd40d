"""agenttorch_b200 — B200 (sm_100a) implementation of AgentTorch's SEIRM agent-update hot path.

Public surface mirrors the reference for this path: `agenttorch_b200.core` (Registry, Substep*,
Runner, envs, Executor, LoadPopulation) and `agenttorch_b200.models.covid`.  The arithmetic runs in
hand-written CUDA kernels (csrc/at_seirm.cu) behind a C ABI (include/at_seirm.h); there is no CPU
fallback.
"""
__version__ = "0.1.0"
