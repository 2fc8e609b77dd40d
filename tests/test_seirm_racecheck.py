"""Race check of the SEIRM kernels without a GPU: csrc/at_seirm.cu compiled for the host on the SIMT emulator WITH
ThreadSanitizer, episodes of the three layouts (fused, column-blocked gather + update, hashed filter + fold) and the backward
kernel run in a subprocess under libtsan.  Every CUDA thread is an OS thread, __syncthreads / warp intrinsics / mbarrier waits /
atomics give the happens-before edges, so an unsynchronised shared- or global-memory access pair is reported.  The only races
allowed are the two the source marks as benign (several threads storing the same value 1 to a flag)."""
import glob
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCE = os.path.join(ROOT, "agenttorch_b200", "csrc", "at_seirm.cu")


def _libtsan():
    res = subprocess.run(["gcc", "-print-file-name=libtsan.so"], capture_output=True, text=True)
    path = res.stdout.strip()
    return path if res.returncode == 0 and os.path.isabs(path) and os.path.exists(path) else None


@pytest.mark.skipif(_libtsan() is None, reason="libtsan.so not available")
def test_only_the_documented_benign_races(tmp_path):
    lib = str(tmp_path / "libat_seirm_emulated_tsan.so")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-fsanitize=thread", "-shared", "-fPIC", "-pthread",
           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "host_emulation", "seirm_emulation.cpp"), "-o", lib]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    log = str(tmp_path / "tsan")
    env = dict(os.environ, LD_PRELOAD=_libtsan(), TSAN_OPTIONS=f"report_signal_unsafe=0 exitcode=0 history_size=2 log_path={log}")
    run = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "host_emulation", "racecheck_episode.py"), lib, "tiny96_hard",
                          "fused", "blocked", "hashed"], capture_output=True, text=True, env=env, timeout=1200)
    assert run.returncode == 0, run.stdout + run.stderr
    assert run.stdout.count("episode ok:") == 3
    source = open(SOURCE).read().splitlines()
    reports = "".join(open(p).read() for p in glob.glob(log + "*"))
    races = re.findall(r"SUMMARY: ThreadSanitizer: data race \S*at_seirm\.cu:(\d+)", reports)
    other = re.findall(r"SUMMARY: ThreadSanitizer: (?!data race \S*at_seirm\.cu)(.*)", reports)
    assert not other, other                                  # nothing outside the kernel source (e.g. in the shim) either
    not_benign = sorted({int(ln) for ln in races if "benign race" not in source[int(ln) - 1]})
    assert not not_benign, [(ln, source[ln - 1].strip()) for ln in not_benign]
    assert races, "ThreadSanitizer saw no thread interaction at all: the check did not run"
