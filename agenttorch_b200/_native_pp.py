"""ctypes binding of libat_predprey.so (C ABI in include/at_predprey.h).

Built in-tree by `build_library()` (nvcc, sm_100a) into agenttorch_b200/lib/.  No CPU fallback: a missing library or a
failing call raises RuntimeError.
"""
import ctypes
import os
import subprocess
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "lib", "libat_predprey.so")
SRC_PATH = os.path.join(_PKG, "csrc", "at_predprey.cu")
INCLUDE_DIR = os.path.join(_ROOT, "include")

ABI_VERSION = 4
OCCUPANCY, EAT, HUNT, MOVE, EAT_APPLY, GROW, CLEAR, EAT_IMPLICIT = 1, 2, 4, 8, 16, 32, 64, 128
INFO_GROWN, INFO_EATEN = 1, 2

EXPORTS = [
    "at_pp_abi_version", "at_pp_last_error", "at_pp_sizeof_model", "at_pp_sizeof_launch", "at_pp_info_bytes",
    "at_pp_storage_cells",
    "at_pp_step", "at_pp_pack_cells", "at_pp_unpack_cells", "at_pp_pack_stage", "at_pp_unpack_stage", "at_pp_census",
]


class PPModel(ctypes.Structure):
    """Mirror of `at_pp_model`."""
    _fields_ = [
        ("abi_version", ctypes.c_int32), ("width", ctypes.c_int32), ("max_y", ctypes.c_int32),
        ("apply_pred_work", ctypes.c_int32), ("apply_prey_work", ctypes.c_int32), ("tiles_per_row", ctypes.c_int32),
        ("n_cells", ctypes.c_int64),
        ("pred_work", ctypes.c_float), ("prey_work", ctypes.c_float), ("prey_nutrition", ctypes.c_float),
        ("grass_nutrition", ctypes.c_float), ("regrowth_time", ctypes.c_float), ("reserved_f", ctypes.c_float),
        ("n_storage", ctypes.c_int64),
    ]


class PPLaunch(ctypes.Structure):
    """Mirror of `at_pp_launch` (field order and types must match include/at_predprey.h)."""
    _fields_ = [
        ("abi_version", ctypes.c_int32), ("flags", ctypes.c_int32), ("t_move", ctypes.c_int32), ("countdown_f64", ctypes.c_int32),
        ("seed", ctypes.c_uint64), ("seed_ptr", ctypes.c_void_p), ("t_ptr", ctypes.c_void_p),
        ("n_pred", ctypes.c_int64), ("n_prey", ctypes.c_int64),
        ("pred_cell", ctypes.c_void_p), ("pred_energy", ctypes.c_void_p),
        ("prey_cell", ctypes.c_void_p), ("prey_energy", ctypes.c_void_p),
        ("info", ctypes.c_void_p), ("countdown", ctypes.c_void_p),
        ("occ_in", ctypes.c_int32), ("occ_out", ctypes.c_int32),
        ("pred_id", ctypes.c_void_p), ("prey_id", ctypes.c_void_p),
        ("nbr_ptr", ctypes.c_void_p), ("nbr_idx", ctypes.c_void_p),
        ("grid", ctypes.c_int32), ("reserved", ctypes.c_int32),
    ]


def nvcc_command(out_path=LIB_PATH, extra=()):
    return ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
            "-shared", "-Xcompiler", "-fPIC", "-I", INCLUDE_DIR, *extra, SRC_PATH, "-o", out_path]


def build_library(force=False, verbose=False):
    """Compile csrc/at_predprey.cu for sm_100a into agenttorch_b200/lib/ (nvcc cross-compiles without a GPU)."""
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    deps = [SRC_PATH, os.path.join(INCLUDE_DIR, "at_predprey.h"), os.path.join(_PKG, "csrc", "at_noise.cuh")]
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
                "(agenttorch_b200 has no CPU or PyTorch fallback for the predator-prey kernels)")
        L = ctypes.CDLL(LIB_PATH)
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
        L.at_pp_abi_version.restype = ctypes.c_int
        L.at_pp_last_error.restype = ctypes.c_char_p
        L.at_pp_sizeof_model.restype = i64
        L.at_pp_sizeof_launch.restype = i64
        L.at_pp_info_bytes.restype = i64
        L.at_pp_info_bytes.argtypes = [i64]
        L.at_pp_storage_cells.restype = i64
        L.at_pp_storage_cells.argtypes = [i32, i64, i32]
        L.at_pp_step.restype = ctypes.c_int
        L.at_pp_step.argtypes = [ctypes.POINTER(PPModel), ctypes.POINTER(PPLaunch), vp]
        L.at_pp_pack_cells.restype = ctypes.c_int
        L.at_pp_pack_cells.argtypes = [vp, i64, i32, i64, vp, vp, vp]
        L.at_pp_unpack_cells.restype = ctypes.c_int
        L.at_pp_unpack_cells.argtypes = [vp, i64, i32, vp, vp]
        L.at_pp_pack_stage.restype = ctypes.c_int
        L.at_pp_pack_stage.argtypes = [vp, i32, i64, vp, vp]
        L.at_pp_unpack_stage.restype = ctypes.c_int
        L.at_pp_unpack_stage.argtypes = [vp, i64, vp, i32, vp]
        L.at_pp_census.restype = ctypes.c_int
        L.at_pp_census.argtypes = [ctypes.POINTER(PPLaunch), i64, vp, vp]
        if L.at_pp_abi_version() != ABI_VERSION:
            raise RuntimeError(f"libat_predprey.so has ABI {L.at_pp_abi_version()}, the binding expects {ABI_VERSION}: rebuild")
        if L.at_pp_sizeof_launch() != ctypes.sizeof(PPLaunch) or L.at_pp_sizeof_model() != ctypes.sizeof(PPModel):
            raise RuntimeError("at_pp_launch / at_pp_model layout mismatch between include/at_predprey.h and _native_pp.py")
        _lib = L
        return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("libat_predprey: " + lib().at_pp_last_error().decode())
