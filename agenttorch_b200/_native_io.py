"""ctypes binding of libat_io.so (C ABI in include/at_io.h): the native edge-list parser of the input pipeline.

Host code (plain C, pthreads), built in-tree by `build_library()` with gcc.  `read_edge_list(path)` returns the (2, E) int64
`edge_index` of an edge-list CSV with the row semantics of the reference's `network_from_file` (utils.py:129-148)."""
import ctypes
import os
import subprocess
import threading

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "lib", "libat_io.so")
SRC_PATH = os.path.join(_PKG, "csrc", "at_io.c")
INCLUDE_DIR = os.path.join(_ROOT, "include")

ABI_VERSION = 1
EXPORTS = ["at_io_abi_version", "at_io_last_error", "at_io_edge_list_count", "at_io_edge_list_parse"]


def gcc_command(out_path=LIB_PATH):
    return ["gcc", "-O3", "-Wall", "-shared", "-fPIC", "-pthread", "-I", INCLUDE_DIR, SRC_PATH, "-o", out_path, "-lm"]


def build_library(force=False, verbose=False):
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    deps = [SRC_PATH, os.path.join(INCLUDE_DIR, "at_io.h")]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    res = subprocess.run(gcc_command(), capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed:\n" + " ".join(gcc_command()) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB_PATH


_lib = None
_lock = threading.Lock()


def lib():
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        L = ctypes.CDLL(LIB_PATH)
        i32, i64p = ctypes.c_int32, ctypes.POINTER(ctypes.c_int64)
        L.at_io_abi_version.restype = ctypes.c_int
        L.at_io_last_error.restype = ctypes.c_char_p
        L.at_io_edge_list_count.restype = ctypes.c_int
        L.at_io_edge_list_count.argtypes = [ctypes.c_char_p, i32, i64p]
        L.at_io_edge_list_parse.restype = ctypes.c_int
        L.at_io_edge_list_parse.argtypes = [ctypes.c_char_p, i32, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, i64p, i64p]
        if L.at_io_abi_version() != ABI_VERSION:
            raise RuntimeError(f"libat_io.so has ABI {L.at_io_abi_version()}, the binding expects {ABI_VERSION}: rebuild")
        _lib = L
        return _lib


def available():
    return os.path.exists(LIB_PATH)


def read_edge_list(path, threads=0, pin_memory=False):
    """(edge_index (2, E) int64 on the host, largest id).  Two passes over the memory-mapped text: count, then fill."""
    L = lib()
    raw = os.fsencode(path)
    n = ctypes.c_int64(0)
    if L.at_io_edge_list_count(raw, int(threads), ctypes.byref(n)) != 0:
        raise OSError(L.at_io_last_error().decode())
    edge_index = torch.empty((2, n.value), dtype=torch.int64, pin_memory=bool(pin_memory and n.value))
    written, max_id = ctypes.c_int64(0), ctypes.c_int64(-1)
    if n.value:
        rc = L.at_io_edge_list_parse(raw, int(threads), n.value, edge_index[0].data_ptr(), edge_index[1].data_ptr(),
                                     ctypes.byref(written), ctypes.byref(max_id))
        if rc != 0:
            raise OSError(L.at_io_last_error().decode())
        if written.value != n.value:
            raise OSError(f"{path} changed while it was being read ({n.value} then {written.value} edges)")
    return edge_index, max_id.value
