"""ctypes binding of libat_interventions.so (C ABI in include/at_interventions.h): quarantine, testing and SEIRMS
progression of the covid model.

Built in-tree by `build_library()` (nvcc, sm_100a) into agenttorch_b200/lib/.  No CPU fallback: a missing library or a
failing call raises RuntimeError.
"""
import ctypes
import os
import subprocess
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "lib", "libat_interventions.so")
SRC_PATH = os.path.join(_PKG, "csrc", "at_interventions.cu")
INCLUDE_DIR = os.path.join(_ROOT, "include")

ABI_VERSION = 1
DECIDE, APPLY = 1, 2
MOVE_NONE, MOVE_E_TO_I, MOVE_R_TO_S, MOVE_SURVIVED, MOVE_DIED = 0, 1, 2, 3, 4

EXPORTS = [
    "at_iv_abi_version", "at_iv_last_error", "at_iv_sizeof_quarantine_args", "at_iv_sizeof_testing_args",
    "at_iv_sizeof_seirms_args", "at_iv_quarantine", "at_iv_testing", "at_iv_seirms_progression",
]

_vp, _i32, _i64, _u64, _f = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64, ctypes.c_float


class QuarantineArgs(ctypes.Structure):
    """Mirror of `at_iv_quarantine_args` (field order and types must match include/at_interventions.h)."""
    _fields_ = [
        ("abi_version", _i32), ("flags", _i32), ("t", _i32), ("stage_f64", _i32), ("n", _i64), ("id_offset", _i64),
        ("seed", _u64), ("seed_ptr", _vp),
        ("p_start", _f), ("p_break", _f), ("quarantine_days", _f), ("infinity_time", _f),
        ("stage", _vp), ("u_start", _vp), ("u_break", _vp), ("is_q_in", _vp), ("start_date_in", _vp),
        ("start_action", _vp), ("break_action", _vp), ("is_q_out", _vp), ("start_date_out", _vp),
    ]


class TestingArgs(ctypes.Structure):
    """Mirror of `at_iv_testing_args`."""
    _fields_ = [
        ("abi_version", _i32), ("flags", _i32), ("t", _i32), ("reserved", _i32), ("n", _i64), ("id_offset", _i64),
        ("seed", _u64), ("seed_ptr", _vp),
        ("p_comply", _f), ("ineligible_days", _f), ("result_delay_days", _f), ("infinity_time", _f),
        ("u_comply", _vp), ("is_q", _vp), ("accept_action", _vp),
        ("awaiting_in", _vp), ("result_date_in", _vp), ("re_eligible_in", _vp),
        ("awaiting_out", _vp), ("result_date_out", _vp), ("re_eligible_out", _vp),
    ]


class SeirmsArgs(ctypes.Structure):
    """Mirror of `at_iv_seirms_args`."""
    _fields_ = [
        ("abi_version", _i32), ("t", _i32), ("n", _i64), ("id_offset", _i64), ("seed", _u64), ("seed_ptr", _vp),
        ("infected_to_recovered_time", _f), ("infinity_time", _f), ("recovered_to_susceptible_time", _f), ("num_steps", _i32),
        ("m_value", _vp), ("u_death", _vp), ("stage_in", _vp), ("next_stage_time_in", _vp),
        ("stage_out", _vp), ("next_stage_time_out", _vp), ("move_code", _vp), ("counts", _vp), ("daily_deaths", _vp),
    ]


def nvcc_command(out_path=LIB_PATH, extra=()):
    return ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
            "-shared", "-Xcompiler", "-fPIC", "-I", INCLUDE_DIR, *extra, SRC_PATH, "-o", out_path]


def build_library(force=False, verbose=False):
    """Compile csrc/at_interventions.cu for sm_100a into agenttorch_b200/lib/ (nvcc cross-compiles without a GPU)."""
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    deps = [SRC_PATH, os.path.join(INCLUDE_DIR, "at_interventions.h"), os.path.join(_PKG, "csrc", "at_noise.cuh")]
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps)):
        return LIB_PATH
    cmd = nvcc_command()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB_PATH


_lib = None
_lock = threading.Lock()


def lib():
    """Load the shared library (once).  Raises if it has not been built — no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(agenttorch_b200 has no CPU or PyTorch fallback for the quarantine / testing / SEIRMS kernels)")
        L = ctypes.CDLL(LIB_PATH)
        L.at_iv_abi_version.restype = ctypes.c_int
        L.at_iv_last_error.restype = ctypes.c_char_p
        for name in ("at_iv_sizeof_quarantine_args", "at_iv_sizeof_testing_args", "at_iv_sizeof_seirms_args"):
            getattr(L, name).restype = _i64
        for name, struct in (("at_iv_quarantine", QuarantineArgs), ("at_iv_testing", TestingArgs),
                             ("at_iv_seirms_progression", SeirmsArgs)):
            fn = getattr(L, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [ctypes.POINTER(struct), _vp]
        if L.at_iv_abi_version() != ABI_VERSION:
            raise RuntimeError(f"libat_interventions.so has ABI {L.at_iv_abi_version()}, the binding expects {ABI_VERSION}: rebuild")
        if (L.at_iv_sizeof_quarantine_args() != ctypes.sizeof(QuarantineArgs) or L.at_iv_sizeof_testing_args() != ctypes.sizeof(TestingArgs)
                or L.at_iv_sizeof_seirms_args() != ctypes.sizeof(SeirmsArgs)):
            raise RuntimeError("argument struct layout mismatch between include/at_interventions.h and _native_iv.py")
        _lib = L
        return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("libat_interventions: " + lib().at_iv_last_error().decode())
