"""`with_behavior` class decorator (reference: agent_torch/core/decorators.py:6-40): a class-level
default behavior settable with `cls.set_behavior(b)` before construction, overridable per instance
through the `behavior` property."""
from functools import wraps


def with_behavior(cls):
    plain_init = cls.__init__

    @wraps(plain_init)
    def init(self, *args, **kwargs):
        plain_init(self, *args, **kwargs)
        self._behavior = getattr(cls, "_class_behavior", None)

    def _get(self):
        return self._behavior if hasattr(self, "_behavior") else getattr(type(self), "_class_behavior", None)

    def _set(self, behavior):
        self._behavior = behavior

    cls.__init__ = init
    cls.set_behavior = classmethod(lambda c, behavior: setattr(c, "_class_behavior", behavior))
    cls.behavior = property(_get, _set)
    return cls
