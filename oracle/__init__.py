"""TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's SEIRM hot path (AgentTorch
`agent_torch/models/covid`), used as the parity checker by `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py`.  Nothing under `agenttorch_b200/` imports this package; the
product path is the CUDA extension and fails loudly without it.
"""
