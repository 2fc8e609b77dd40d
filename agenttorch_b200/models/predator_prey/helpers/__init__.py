from .random import random_float, random_int  # noqa: F401
