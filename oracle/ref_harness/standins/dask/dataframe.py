class DataFrame:  # stand-in: agent_torch/core/executor.py:4 only does isinstance checks
    pass


class Series:
    pass
