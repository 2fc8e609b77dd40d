"""TEST INFRASTRUCTURE — quarantine / testing / SEIRMS-progression kernels on the emulated library given on the command
line, a few steps on a random population against the oracle; started under ThreadSanitizer by tests/test_racecheck.py.

    python racecheck_interventions.py <libat_interventions_emulated.so>
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from agenttorch_b200 import _native_iv  # noqa: E402
from agenttorch_b200 import interventions as ops  # noqa: E402
from oracle import noise as onoise  # noqa: E402
from oracle.interventions_port import quarantine_step, testing_step  # noqa: E402


def main(lib_path):
    _native_iv.LIB_PATH, _native_iv._lib = lib_path, None
    ops._require_cuda = lambda name, t: None
    ops._stream = lambda t: None
    n, seed = 5003, 41
    rng = np.random.default_rng(1)
    stage = torch.from_numpy(rng.choice([0.0, 1.0, 2.0, 3.0], size=n).astype(np.float32))
    q, d = torch.zeros(n), torch.full((n,), 130.0)
    aw, rd, re = torch.zeros(n), torch.full((n,), 130.0), torch.zeros(n)
    ids = np.arange(n)
    for t in range(5):
        us, ub = onoise.stream_uniforms(seed, t, ids, onoise.STREAM_QUARANTINE)
        eq, ed = quarantine_step(t, stage.numpy(), q.numpy(), d.numpy(), us, ub, 0.7, 0.3, 2, 130)
        ea, er, ee = testing_step(t, eq, aw.numpy(), rd.numpy(), re.numpy(), onoise.stream_uniforms(seed, t, ids, onoise.STREAM_TESTING)[0],
                                  0.8, 1, 2, 130)
        ops.quarantine_step(stage, q, d, 0.7, 0.3, t, seed, 0, 2.0, 130.0)
        ops.testing_step(q, aw, rd, re, 0.8, t, seed, 0, 1.0, 2.0, 130.0)
        for got, want in ((q, eq), (d, ed), (aw, ea), (rd, er), (re, ee)):
            assert np.array_equal(got.numpy(), want), t
    nst = torch.from_numpy(rng.integers(0, 8, size=n).astype(np.float32))
    s2, n2, dd2, code, counts = ops.seirms_progression(stage, nst, torch.zeros(21), torch.tensor([0.4]), 4, seed, 0, None, 5.0, 41.0, 100.0)
    due = (stage == 2) & (nst <= 4)
    assert int(counts[0]) == int(due.sum()) == int(counts[1] + counts[2]) and float(dd2[4]) > 0
    print("episode ok: interventions", flush=True)


if __name__ == "__main__":
    main(sys.argv[1])
