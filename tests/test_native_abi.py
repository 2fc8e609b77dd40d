"""The C-ABI library loads on a CPU-only box and exports every symbol include/at_seirm.h declares
(no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

from agenttorch_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    _native.build_library()
    return _native.lib()


def test_header_symbols_are_exported(built):
    header = open(os.path.join(ROOT, "include", "at_seirm.h")).read()
    declared = set(re.findall(r"^\s*(?:int|int64_t|const char\*)\s+(at_[a-z0-9_]+)\s*\(", header, flags=re.M))
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    handle = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert getattr(handle, name) is not None


def test_struct_mirrors_match_the_header(built):
    assert built.at_seirm_abi_version() == _native.ABI_VERSION
    assert built.at_seirm_sizeof_launch() == ctypes.sizeof(_native.SeirmLaunch)
    assert built.at_seirm_sizeof_model() == ctypes.sizeof(_native.SeirmModel)
    assert built.at_seirm_workspace_bytes() > 0
    header = open(os.path.join(ROOT, "include", "at_seirm.h")).read()
    assert f"#define AT_SEIRM_ABI_VERSION {_native.ABI_VERSION}" in header


def test_bad_arguments_return_errors_not_exceptions(built):
    assert built.at_seirm_step(None, None, None) != 0
    assert b"null" in built.at_last_error()
    launch, model = _native.SeirmLaunch(), _native.SeirmModel()
    launch.abi_version = _native.ABI_VERSION + 1
    assert built.at_seirm_step(ctypes.byref(model), ctypes.byref(launch), None) != 0
    assert b"abi_version" in built.at_last_error()


def test_no_cpu_fallback():
    import torch

    from agenttorch_b200.engine import SeirmEngine, SeirmSpec
    from agenttorch_b200.graph import synthetic_contact_graph

    g = synthetic_contact_graph(512, 4.0, 1, device="cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        SeirmEngine(g, SeirmSpec())
    assert not torch.cuda.is_available() or True


# ------------------------------------------------------------------------------- libat_predprey.so
@pytest.fixture(scope="module")
def built_pp():
    from agenttorch_b200 import _native_pp

    _native_pp.build_library()
    return _native_pp.lib()


def test_predprey_header_symbols_are_exported(built_pp):
    from agenttorch_b200 import _native_pp

    header = open(os.path.join(ROOT, "include", "at_predprey.h")).read()
    declared = set(re.findall(r"^\s*(?:int|int64_t|const char\*)\s+(at_[a-z0-9_]+)\s*\(", header, flags=re.M))
    assert declared == set(_native_pp.EXPORTS), declared ^ set(_native_pp.EXPORTS)
    handle = ctypes.CDLL(_native_pp.LIB_PATH)
    for name in declared:
        assert getattr(handle, name) is not None
    assert built_pp.at_pp_sizeof_launch() == ctypes.sizeof(_native_pp.PPLaunch)
    assert built_pp.at_pp_sizeof_model() == ctypes.sizeof(_native_pp.PPModel)
    assert f"#define AT_PP_ABI_VERSION {_native_pp.ABI_VERSION}" in header
    assert built_pp.at_pp_info_bytes(450) == 464 and built_pp.at_pp_info_bytes(4096 * 4096) == 4096 * 4096


def test_predprey_bad_arguments_return_errors(built_pp):
    from agenttorch_b200 import _native_pp

    assert built_pp.at_pp_step(None, None, None) != 0
    assert b"null" in built_pp.at_pp_last_error()
    m, a = _native_pp.PPModel(), _native_pp.PPLaunch()
    m.abi_version, a.abi_version = _native_pp.ABI_VERSION, _native_pp.ABI_VERSION + 1
    assert built_pp.at_pp_step(ctypes.byref(m), ctypes.byref(a), None) != 0
    assert b"ABI" in built_pp.at_pp_last_error()
    a.abi_version = _native_pp.ABI_VERSION
    assert built_pp.at_pp_step(ctypes.byref(m), ctypes.byref(a), None) != 0           # n_cells == 0
    assert b"n_cells" in built_pp.at_pp_last_error()


def test_predprey_no_cpu_fallback():
    import torch

    from agenttorch_b200.predprey import PredPreyEngine, PredPreySpec, pack_cells

    with pytest.raises(RuntimeError, match="CUDA"):
        PredPreyEngine(PredPreySpec(4, 4, 16, 4), "cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        pack_cells(torch.zeros(3, 2, dtype=torch.int64), 4, 16)
