"""libat_io.so (include/at_io.h): the native edge-list parser against the reference's own pandas sequence
(agent_torch/models/covid/substeps/utils.py:129-148, restated in `reference_edge_index`) on well-formed and malformed
files, any thread count; `network_from_file` returns the reference's payload."""
import ctypes
import os
import re

import numpy as np
import pandas as pd
import pytest
import torch

from agenttorch_b200 import _native_io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    _native_io.build_library()
    return _native_io.lib()


def reference_edge_index(path):
    df = pd.read_csv(path, header=None)
    if df.iloc[0].apply(lambda v: isinstance(v, str)).any():
        df = pd.read_csv(path, header=0)
    df = df.iloc[:, :2]
    df = df.apply(pd.to_numeric, errors="coerce").dropna().astype("int64")
    return torch.tensor(df.values.T, dtype=torch.long).reshape(2, -1)


CASES = {
    "plain": "0,1\n2,3\n",
    "header": "src,dst\n0,1\n2,3\n",
    "header_three_columns": "a,b,c\n1,2,3\n4,5,6\n",
    "header_only": "a,b\n",
    "text_in_a_later_row": "0,1\na,2\n3,4\n",                 # pandas reads column 0 as strings -> row 0 becomes the header
    "text_in_a_later_column": "0,1,x\n2,3,y\n",
    "third_column_numeric": "0,1,5\n2,3,6\n",
    "floats_truncate": "0.0,1.0\n2.5,3.9\n",
    "blank_rows": "0,1\n\n2,3\n\n",
    "leading_blank_row": "\n0,1\n2,3\n",
    "crlf": "0,1\r\n2,3\r\n",
    "blanks_around_fields": " 0, 1\n2 ,3\n",
    "missing_field": "0,\n2,3\n",
    "negative": "-1,2\n3,4\n",
    "no_trailing_newline": "0,1\n2,3",
    "nan_token": "0,1\nnan,3\n4,5\n",
    "na_tokens": "1,2\nNA,3\n4,null\n5,6\n",
    "exponent": "1e2,3\n4,5\n",
    "plus_sign": "+1,2\n3,4\n",
    "quoted": '"0","1"\n2,3\n',
    "wide_ids": "123456789012,5\n7,8\n",
    "self_loops_and_duplicates": "4,4\n1,2\n1,2\n",
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("threads", [1, 3])
def test_same_rows_as_the_reference_loader(tmp_path, name, threads):
    path = str(tmp_path / (name + ".csv"))
    with open(path, "w", newline="") as f:
        f.write(CASES[name])
    want = reference_edge_index(path)
    got, max_id = _native_io.read_edge_list(path, threads=threads)
    assert got.dtype == torch.int64 and got.shape == want.shape and torch.equal(got, want)
    assert max_id == (int(want.max()) if want.numel() else -1)


@pytest.mark.parametrize("threads", [1, 2, 7, 0])
def test_large_file_keeps_row_order_for_any_thread_count(tmp_path, threads):
    rng = np.random.default_rng(3)
    pairs = rng.integers(0, 5_000_000, size=(400_000, 2))
    path = str(tmp_path / "edges.csv")
    pd.DataFrame(pairs).to_csv(path, header=False, index=False)
    got, max_id = _native_io.read_edge_list(path, threads=threads)
    assert torch.equal(got, torch.from_numpy(pairs.T)) and max_id == int(pairs.max())


def test_empty_file_missing_file_and_capacity(tmp_path, built):
    empty = str(tmp_path / "empty.csv")
    open(empty, "w").close()
    got, max_id = _native_io.read_edge_list(empty)
    assert got.shape == (2, 0) and max_id == -1
    with pytest.raises(OSError, match="cannot open"):
        _native_io.read_edge_list(str(tmp_path / "absent.csv"))
    path = str(tmp_path / "two.csv")
    open(path, "w").write("0,1\n2,3\n")
    out = torch.zeros((2, 1), dtype=torch.int64)
    n, mx = ctypes.c_int64(0), ctypes.c_int64(0)
    rc = built.at_io_edge_list_parse(os.fsencode(path), 1, 1, out[0].data_ptr(), out[1].data_ptr(), ctypes.byref(n), ctypes.byref(mx))
    assert rc != 0 and b"capacity" in built.at_io_last_error()
    assert built.at_io_edge_list_count(None, 1, ctypes.byref(n)) != 0


def test_header_symbols_are_exported(built):
    header = open(os.path.join(ROOT, "include", "at_io.h")).read()
    declared = set(re.findall(r"^\s*(?:int|int64_t|const char\*)\s+(at_[a-z0-9_]+)\s*\(", header, flags=re.M))
    assert declared == set(_native_io.EXPORTS), declared ^ set(_native_io.EXPORTS)
    assert f"#define AT_IO_ABI_VERSION {_native_io.ABI_VERSION}" in header and built.at_io_abi_version() == _native_io.ABI_VERSION


def test_network_from_file_payload(tmp_path, monkeypatch):
    from agenttorch_b200.models.covid.substeps.utils import network_from_file

    monkeypatch.delenv("AGENTTORCH_B200_CACHE_DIR", raising=False)
    path = str(tmp_path / "0.csv")
    open(path, "w").write("src,dst\n5,1\n2,3\n2,3\n")
    graph, (edge_index, edge_attr) = network_from_file({"file_path": path})
    assert graph is None and edge_index.tolist() == [[5, 2, 2], [1, 3, 3]] and edge_index.dtype == torch.int64
    assert edge_attr.dtype == torch.float32 and edge_attr.tolist() == [[1.0] * 3, [1.0] * 3]           # [type == 1; weight == 1]
    # without the native library the pandas sequence gives the same payload
    monkeypatch.setattr(_native_io, "available", lambda: False)
    _, (edge_index2, _) = network_from_file({"file_path": path})
    assert torch.equal(edge_index, edge_index2)
