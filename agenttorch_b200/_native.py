"""ctypes binding of libat_seirm.so (C ABI in include/at_seirm.h).

The library is built in-tree by `build_library()` (nvcc, sm_100a) and loaded from
agenttorch_b200/lib/.  There is no CPU fallback: if the library is missing or a call
fails, a RuntimeError is raised.
"""
import ctypes
import os
import subprocess
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "lib", "libat_seirm.so")
SRC_PATH = os.path.join(_PKG, "csrc", "at_seirm.cu")
INCLUDE_DIR = os.path.join(_ROOT, "include")

ABI_VERSION = 10
MAX_SFS, MAX_SFI, MAX_LAMGAMMA = 16, 8, 84
TIMER_F32, TIMER_I16 = 0, 1
AGE_F32, AGE_U8 = 0, 1
ISO_TENSOR, ISO_HASH_SOFT, ISO_HASH_HARD, ISO_NONE = 0, 1, 2, 3
U_TENSOR, U_HASH = 0, 1
DO_TRANSMISSION, DO_PROGRESSION, DO_ENCODE, DO_INIT, DO_LOOKAHEAD = 1, 2, 4, 8, 16
MAX_BLOCKS = 64
MAX_LAYERS = 8
LAYOUT_BCSR, LAYOUT_STREAM = 0, 1
STREAM_UNIT, STREAM_MAX_TILE_ROWS, STREAM_MAX_BLOCK_SIZE = 256, 2048, (1 << 21) - 1024

EXPORTS = [
    "at_seirm_abi_version", "at_last_error", "at_seirm_sizeof_launch", "at_seirm_sizeof_model",
    "at_seirm_workspace_bytes", "at_seirm_max_block_size", "at_seirm_stream_max_block_size", "at_seirm_step",
    "at_seirm_episode_backward", "at_seirm_hot_from_code", "at_seirm_pull_codes", "at_seirm_exchange_bytes",
    "at_seirm_exchange_sparse", "at_noise_fill",
]


class SeirmModel(ctypes.Structure):
    """Mirror of `at_seirm_model`."""
    _fields_ = [
        ("n_sfs", ctypes.c_int32), ("n_sfi", ctypes.c_int32), ("n_lam_gamma", ctypes.c_int32),
        ("exposed_to_infected_time", ctypes.c_int32), ("infected_to_recovered_time", ctypes.c_int32),
        ("infinity_time", ctypes.c_int32), ("i_bar", ctypes.c_float),
        ("sfs", ctypes.c_float * MAX_SFS), ("sfi", ctypes.c_float * MAX_SFI),
        ("lam_gamma", ctypes.c_float * MAX_LAMGAMMA), ("n_layers", ctypes.c_int32), ("i_bar_layers", ctypes.c_float * MAX_LAYERS),
    ]


class SeirmLaunch(ctypes.Structure):
    """Mirror of `at_seirm_launch` (field order and types must match include/at_seirm.h)."""
    _fields_ = [
        ("abi_version", ctypes.c_int32), ("flags", ctypes.c_int32), ("t", ctypes.c_int32), ("t_encode", ctypes.c_int32),
        ("n_rows", ctypes.c_int64), ("n_rows_pad", ctypes.c_int64), ("n_sources", ctypes.c_int64), ("row_offset", ctypes.c_int64),
        ("n_blocks", ctypes.c_int32), ("block_size", ctypes.c_int32),
        ("rows_per_task", ctypes.c_int32), ("edges_per_chunk", ctypes.c_int32),
        ("row_start", ctypes.c_void_p), ("seg_info", ctypes.c_void_p), ("col", ctypes.c_void_p), ("edge_weight", ctypes.c_void_p),
        ("stage", ctypes.c_void_p), ("next_stage_time", ctypes.c_void_p), ("infected_time", ctypes.c_void_p),
        ("age", ctypes.c_void_p), ("timer_dtype", ctypes.c_int32), ("age_dtype", ctypes.c_int32),
        ("sus_bitmap", ctypes.c_void_p), ("exp_bitmap", ctypes.c_void_p), ("inf_bitmap", ctypes.c_void_p),
        ("code_in", ctypes.c_void_p), ("hot_bitmap_in", ctypes.c_void_p), ("code_out", ctypes.c_void_p),
        ("hot_bitmap_out", ctypes.c_void_p), ("hot_filter_in", ctypes.c_void_p), ("hot_filter_out", ctypes.c_void_p),
        ("iso_mode", ctypes.c_int32), ("u_mode", ctypes.c_int32),
        ("isolation", ctypes.c_void_p), ("uniform", ctypes.c_void_p),
        ("seed", ctypes.c_uint64), ("seed_ptr", ctypes.c_void_p), ("p_isolate", ctypes.c_float), ("_pad0", ctypes.c_int32),
        ("r_value", ctypes.c_void_p), ("r_index", ctypes.c_int64), ("m_value", ctypes.c_void_p),
        ("daily_infected", ctypes.c_void_p), ("daily_deaths", ctypes.c_void_p), ("dne_dr", ctypes.c_void_p),
        ("sum_dne_dr", ctypes.c_void_p), ("death_count", ctypes.c_void_p), ("recovered_step", ctypes.c_void_p),
        ("lam_part", ctypes.c_void_p), ("layout", ctypes.c_int32), ("tile_rows", ctypes.c_int32), ("dirty", ctypes.c_void_p),
        ("workspace", ctypes.c_void_p), ("error_flag", ctypes.c_void_p),
        ("grid", ctypes.c_int32), ("tasks_per_group", ctypes.c_int32),
        ("warps_per_cta", ctypes.c_int32), ("hash_words", ctypes.c_int32),
        ("unit_tile", ctypes.c_void_p), ("block_units", ctypes.c_void_p), ("tile_rec_base", ctypes.c_void_p),
        ("tile_count", ctypes.c_void_p), ("rec", ctypes.c_void_p),
        ("lam_deg_bits", ctypes.c_int32), ("xchg_slot", ctypes.c_int32),
        ("dne_e_dr", ctypes.c_void_p), ("over_state", ctypes.c_void_p), ("over_local", ctypes.c_int32), ("reserve_sms", ctypes.c_int32),
        ("step_skipped", ctypes.c_void_p),
    ]


def nvcc_command(out_path=LIB_PATH, extra=()):
    return ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
            "-shared", "-Xcompiler", "-fPIC", "-I", INCLUDE_DIR, *extra, SRC_PATH, "-o", out_path]


def build_library(force=False, verbose=False):
    """Compile csrc/at_seirm.cu for sm_100a into agenttorch_b200/lib/ (nvcc cross-compiles without a GPU)."""
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    deps = [SRC_PATH, os.path.join(INCLUDE_DIR, "at_seirm.h"), os.path.join(_PKG, "csrc", "at_noise.cuh")]
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
                "(agenttorch_b200 has no CPU or PyTorch fallback for the SEIRM kernels)")
        L = ctypes.CDLL(LIB_PATH)
        L.at_seirm_abi_version.restype = ctypes.c_int
        L.at_last_error.restype = ctypes.c_char_p
        for name in ("at_seirm_sizeof_launch", "at_seirm_sizeof_model", "at_seirm_workspace_bytes"):
            getattr(L, name).restype = ctypes.c_int64
        L.at_seirm_max_block_size.restype = ctypes.c_int64
        L.at_seirm_max_block_size.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]
        L.at_seirm_stream_max_block_size.restype = ctypes.c_int64
        L.at_seirm_stream_max_block_size.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]
        L.at_seirm_step.restype = ctypes.c_int
        L.at_seirm_step.argtypes = [ctypes.POINTER(SeirmModel), ctypes.POINTER(SeirmLaunch), ctypes.c_void_p]
        L.at_seirm_episode_backward.restype = ctypes.c_int
        L.at_seirm_episode_backward.argtypes = [
            ctypes.c_int32, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]
        L.at_seirm_hot_from_code.restype = ctypes.c_int
        L.at_seirm_hot_from_code.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_int32, ctypes.c_void_p]
        L.at_seirm_pull_codes.restype = ctypes.c_int
        L.at_seirm_pull_codes.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32,
                                          ctypes.POINTER(ctypes.c_void_p), ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]
        L.at_seirm_exchange_bytes.restype = ctypes.c_int64
        L.at_seirm_exchange_bytes.argtypes = [ctypes.c_int64]
        L.at_seirm_exchange_sparse.restype = ctypes.c_int
        L.at_seirm_exchange_sparse.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int32,
                                               ctypes.POINTER(ctypes.c_void_p), ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32,
                                               ctypes.c_void_p, ctypes.c_void_p]
        L.at_noise_fill.restype = ctypes.c_int
        L.at_noise_fill.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int32,
                                    ctypes.c_int32, ctypes.c_float, ctypes.c_void_p]
        if L.at_seirm_abi_version() != ABI_VERSION:
            raise RuntimeError("libat_seirm.so ABI version mismatch: rebuild the library")
        if L.at_seirm_sizeof_launch() != ctypes.sizeof(SeirmLaunch) or L.at_seirm_sizeof_model() != ctypes.sizeof(SeirmModel):
            raise RuntimeError("ctypes mirror of at_seirm_launch / at_seirm_model is out of sync with the header")
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise RuntimeError("libat_seirm: " + lib().at_last_error().decode())
