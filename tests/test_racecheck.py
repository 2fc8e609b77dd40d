"""Race check of the kernels without a GPU.  SEIRM: csrc/at_seirm.cu compiled for the host on the SIMT emulator WITH
ThreadSanitizer, episodes of the three layouts (fused, column-blocked gather + update, hashed filter + fold) and the backward
kernel run in a subprocess under libtsan.  Every CUDA thread is an OS thread, __syncthreads / warp intrinsics / mbarrier waits /
atomics give the happens-before edges, so an unsynchronised shared- or global-memory access pair is reported.  The only races
allowed are the two the source marks as benign (several threads storing the same value 1 to a flag).
Predator-prey: the same on csrc/at_predprey.cu; the only interaction between the threads of a launch is the occupancy marking
(atomicOr into a word whose other bytes are read by `cell_info`), marked benign in the source — which is also what entitles
the sequential shim of tests/host_emulation/cuda_shim.h to run the threads one after the other."""
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


def _first_source_frames(reports, source_name):
    """For every TSan report: the source lines (in `source_name`) of the innermost frames of the two racing accesses."""
    out = []
    for report in reports.split("WARNING: ThreadSanitizer: data race")[1:]:
        lines = []
        for access in re.split(r"\n  (?:Previous )?(?:atomic )?(?:[Ww]rite|[Rr]ead) of size", report)[1:3]:
            m = re.search(re.escape(source_name) + r":(\d+)", access)
            lines.append(int(m.group(1)) if m else None)
        out.append(tuple(lines))
    return out


@pytest.mark.skipif(_libtsan() is None, reason="libtsan.so not available")
def test_predator_prey_threads_only_meet_in_the_occupancy_marks(tmp_path):
    lib = str(tmp_path / "libat_predprey_emulated_tsan.so")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-shared", "-fPIC", "-pthread", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "host_emulation", "pp_simt_emulation.cpp"), "-o", lib]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    log = str(tmp_path / "tsan")
    env = dict(os.environ, LD_PRELOAD=_libtsan(), TSAN_OPTIONS=f"report_signal_unsafe=0 exitcode=0 history_size=2 log_path={log}")
    run = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "host_emulation", "racecheck_predprey.py"), lib],
                         capture_output=True, text=True, env=env, timeout=1200)
    assert run.returncode == 0, run.stdout + run.stderr
    assert run.stdout.count("episode ok:") == 2
    source = open(os.path.join(ROOT, "agenttorch_b200", "csrc", "at_predprey.cu")).read().splitlines()
    reports = "".join(open(p).read() for p in glob.glob(log + "*"))
    pairs = _first_source_frames(reports, "at_predprey.cu")
    assert pairs, "ThreadSanitizer saw no thread interaction at all: the check did not run"
    bad = sorted({ln for pair in pairs for ln in pair if ln is None or "benign race" not in source[ln - 1]}, key=str)
    assert not bad, [(ln, source[ln - 1].strip() if ln else "frame outside the kernel source") for ln in bad]


@pytest.mark.skipif(_libtsan() is None, reason="libtsan.so not available")
def test_intervention_kernels_are_race_free(tmp_path):
    """csrc/at_interventions.cu: a thread owns its 4 agents, the SEIRMS counts go through warp shuffles + atomics, the
    epilogue runs after the launch boundary — ThreadSanitizer must report nothing at all."""
    lib = str(tmp_path / "libat_interventions_emulated_tsan.so")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-shared", "-fPIC", "-pthread", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "host_emulation", "iv_simt_emulation.cpp"), "-o", lib]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    symbols = subprocess.run(["nm", "-D", "--undefined-only", lib], capture_output=True, text=True).stdout
    assert "__tsan_" in symbols                                  # the library really is instrumented
    log = str(tmp_path / "tsan")
    env = dict(os.environ, LD_PRELOAD=_libtsan(), TSAN_OPTIONS=f"report_signal_unsafe=0 exitcode=0 history_size=2 log_path={log}")
    run = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "host_emulation", "racecheck_interventions.py"), lib],
                         capture_output=True, text=True, env=env, timeout=1200)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "episode ok: interventions" in run.stdout
    reports = "".join(open(p).read() for p in glob.glob(log + "*"))
    assert "WARNING: ThreadSanitizer" not in reports, reports[:2000]
