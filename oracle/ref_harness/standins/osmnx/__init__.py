"""Stand-in for `osmnx` (absent here, no network).  TEST INFRASTRUCTURE.
The reference imports it at module level in models/predator_prey/helpers/map.py:7 for the street-map
network variant; the grid variant exercised by the parity harness never calls it."""


def graph_from_point(*args, **kwargs):
    raise RuntimeError("osmnx stand-in: street-map networks need network access")
