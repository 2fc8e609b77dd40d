"""TEST INFRASTRUCTURE — run predator-prey episodes (plain and tiled + sorted storage) on the emulated library given on the
command line and check them against the reference trajectories; started under ThreadSanitizer by tests/test_racecheck.py.

    python racecheck_predprey.py <libat_predprey_emulated.so>
"""
import dataclasses
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from agenttorch_b200 import _native_pp, predprey  # noqa: E402
from tests.helpers import PREDPREY_KEYS, load_predprey_golden  # noqa: E402
from tests.test_predprey_emulated import _assert_equal, _spec, _state  # noqa: E402


def main(lib_path):
    _native_pp.LIB_PATH, _native_pp._lib = lib_path, None
    predprey._require_cuda = lambda what, is_cuda: None
    predprey._stream = lambda: None
    for case, tiled in (("predprey_18x25_work", False), ("predprey_24x28", True)):
        g = load_predprey_golden(case)
        spec = _spec(g)
        if tiled:
            spec = dataclasses.replace(spec, tiled=True)
        eng = predprey.PredPreyEngine(spec, "cpu")
        st = _state(spec, g)
        if tiled:
            st.sort_by_cell(spec)
        eng.run(st, g["steps"], g["seed"], cuda_graph=False, resort_every=5 if tiled else 0)
        _assert_equal(st, spec, {k: g["traj_" + k][g["steps"] - 1] for k in PREDPREY_KEYS}, case)
        print("episode ok:", case, "tiled" if tiled else "plain", flush=True)


if __name__ == "__main__":
    main(sys.argv[1])
