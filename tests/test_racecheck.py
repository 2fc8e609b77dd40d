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


def _first_source_frames(reports, source_name):
    """For every TSan report that touches `source_name`: the source lines of the innermost frames (in that file) of the two
    racing accesses."""
    out = []
    for report in reports.split("WARNING: ThreadSanitizer: data race")[1:]:
        if source_name not in report:
            continue
        lines = []
        for access in re.split(r"\n  (?:Previous )?(?:atomic )?(?:[Ww]rite|[Rr]ead) of size", report)[1:3]:
            m = re.search(re.escape(source_name) + r":(\d+)", access)
            lines.append(int(m.group(1)) if m else None)
        out.append(tuple(lines))
    return out


WORKLOADS = {   # kernel source -> (emulation translation unit, extra compiler flags, driver script and its arguments)
    "at_seirm.cu": ("seirm_emulation.cpp", ["-ffp-contract=off"], ["racecheck_episode.py", "{lib}", "tiny96_hard", "fused", "blocked", "hashed"]),
    "at_predprey.cu": ("pp_simt_emulation.cpp", [], ["racecheck_predprey.py", "{lib}"]),
    "at_interventions.cu": ("iv_simt_emulation.cpp", [], ["racecheck_interventions.py", "{lib}"]),
}


@pytest.fixture(scope="module")
def tsan_reports(tmp_path_factory):
    """Builds the three emulated libraries with -fsanitize=thread and runs each workload of tests/host_emulation/racecheck_*.py
    under libtsan in its own subprocess (all three at the same time); returns {source: (stdout, reports)}."""
    if _libtsan() is None:
        pytest.skip("libtsan.so not available")
    tmp = tmp_path_factory.mktemp("racecheck")
    emu = os.path.join(ROOT, "tests", "host_emulation")
    builds = {}
    for source, (unit, extra, _) in WORKLOADS.items():
        lib = str(tmp / (unit.replace(".cpp", "") + "_tsan.so"))
        cmd = ["g++", "-O1", "-g", "-std=c++17", *extra, "-fsanitize=thread", "-Wno-tsan", "-shared", "-fPIC", "-pthread",
               "-I", os.path.join(ROOT, "include"), os.path.join(emu, unit), "-o", lib]
        builds[source] = (lib, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    runs = {}
    for source, (lib, proc) in builds.items():
        out, _ = proc.communicate()
        assert proc.returncode == 0, out
        assert "__tsan_" in subprocess.run(["nm", "-D", "--undefined-only", lib], capture_output=True, text=True).stdout
        log = str(tmp / ("tsan_" + source))
        env = dict(os.environ, LD_PRELOAD=_libtsan(), TSAN_OPTIONS=f"report_signal_unsafe=0 exitcode=0 history_size=2 log_path={log}")
        script, *args = WORKLOADS[source][2]
        runs[source] = (log, subprocess.Popen([sys.executable, os.path.join(emu, script), *[a.format(lib=lib) for a in args]],
                                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env))
    results = {}
    for source, (log, proc) in runs.items():
        out, _ = proc.communicate(timeout=1800)
        assert proc.returncode == 0, out
        reports = "".join(open(p).read() for p in glob.glob(log + "*"))
        other = re.findall(r"SUMMARY: ThreadSanitizer: (?!data race)(.*)", reports)
        assert not other, other
        results[source] = (out, reports)
    return results


def _not_benign(reports, source_name):
    source = open(os.path.join(ROOT, "agenttorch_b200", "csrc", source_name)).read().splitlines()
    pairs = _first_source_frames(reports, source_name)
    bad = sorted({ln for pair in pairs for ln in pair if ln is None or "benign race" not in source[ln - 1]}, key=str)
    return pairs, [(ln, source[ln - 1].strip() if ln else "frame outside the kernel source") for ln in bad]


def test_seirm_only_the_documented_benign_races(tsan_reports):
    stdout, reports = tsan_reports["at_seirm.cu"]
    assert stdout.count("episode ok: tiny96_hard") == 3                      # fused, blocked, hashed (+ backward each)
    pairs, bad = _not_benign(reports, "at_seirm.cu")
    assert pairs, "ThreadSanitizer saw no thread interaction at all: the check did not run"
    assert not bad, bad


def test_predator_prey_threads_only_meet_in_the_occupancy_marks(tsan_reports):
    stdout, reports = tsan_reports["at_predprey.cu"]
    assert stdout.count("episode ok: predprey") == 2
    pairs, bad = _not_benign(reports, "at_predprey.cu")
    assert pairs, "ThreadSanitizer saw no thread interaction at all: the check did not run"
    assert not bad, bad


def test_intervention_kernels_are_race_free(tsan_reports):
    """A thread owns its 4 agents, the SEIRMS counts go through warp shuffles + atomics, the epilogue runs after the launch
    boundary: ThreadSanitizer must report nothing at all."""
    stdout, reports = tsan_reports["at_interventions.cu"]
    assert "episode ok: interventions" in stdout
    assert "WARNING: ThreadSanitizer" not in reports, reports[:2000]
