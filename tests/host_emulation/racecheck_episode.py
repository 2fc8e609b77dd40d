"""TEST INFRASTRUCTURE — run SEIRM episodes on the emulated library given on the command line, in the layouts given, and check
them against the reference golden.  Started by tests/test_racecheck.py under ThreadSanitizer (LD_PRELOAD=libtsan.so with
the emulated library built with -fsanitize=thread): every CUDA thread of the emulator is an OS thread, so TSan sees the
kernel's shared- and global-memory accesses with the happens-before edges of the barriers, warp rendezvous and atomics.

    python racecheck_episode.py <libat_seirm_emulated.so> <golden case> <layout> [<layout> ...]      layouts: fused blocked hashed
"""
import contextlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from agenttorch_b200 import _native, engine  # noqa: E402
from tests.helpers import load_golden, rel_err  # noqa: E402

LAYOUTS = {"fused": dict(), "blocked": dict(block_size=1024), "hashed": dict(hashed=True, hash_words=32)}


def main(lib_path, case, layouts):
    _native.LIB_PATH, _native._lib = lib_path, None
    engine._require_cuda = lambda message, is_cuda: None
    torch.cuda.device = lambda device: contextlib.nullcontext()
    torch.cuda.current_stream = lambda device=None: types.SimpleNamespace(cuda_stream=0)
    torch.cuda.synchronize = lambda device=None: None
    from agenttorch_b200.engine import NoiseSpec, SeirmEngine, SeirmSpec, SeirmState
    from agenttorch_b200.graph import build_contact_graph, mean_interactions_from_age

    g = load_golden(case)
    T, n = g["num_steps"], len(g["age"])
    age = torch.from_numpy(g["age"].astype(np.int64))
    graph = build_contact_graph(torch.from_numpy(g["edges"].astype(np.int64)).t().contiguous(), n,
                                mean_interactions=mean_interactions_from_age(age))
    for layout in layouts:
        eng = SeirmEngine(graph, SeirmSpec(num_steps=T), warps_per_cta=2, **LAYOUTS[layout])
        st = SeirmState.from_initial(g["stage0"].astype(np.float32), g["age"], T, torch.device("cpu"))
        r2 = torch.tensor(g["r2"], dtype=torch.float32)
        eye = torch.eye(3)
        r_steps = torch.stack([(r2 * eye[int(t / 7)]).sum() for t in range(T)])
        m = torch.tensor(g["m"], dtype=torch.float32)
        sv = eng.run_episode(st, r_steps, m, T, NoiseSpec(seed=g["seed"], isolation=g["mode"], p_isolate=float(g["p_isolate"])),
                             want_grad=True)
        _, gm = eng.episode_backward(sv, T, torch.from_numpy(g["gI"]), torch.from_numpy(g["gD"]), m)
        eng.check_errors()
        assert np.array_equal(st.stage.numpy().astype(np.float64), g["ref_stage"]), layout
        assert rel_err(g["ref_daily_infected"], st.daily_infected.numpy()) <= 1e-5 and rel_err(g["ref_grad_M"], gm.numpy()) <= 1e-4
        print("episode ok:", case, layout, flush=True)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3:])
