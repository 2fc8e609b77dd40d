def to_networkx(*args, **kwargs):
    raise NotImplementedError("stand-in: not on the SEIRM path")
