#!/usr/bin/env python
"""Put the UNMODIFIED reference where `bench.py --impl reference` finds it on the GPU box: baseline/_ref/ (git-ignored, but
it travels with the gpurun snapshot; /root/reference does not exist there).

1. the contract's recipe: `pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref`
   from a writable copy of /root/reference.  In this image it fails: the reference's build backend (`pdm.backend`,
   pyproject.toml) is not installed and there is no network.
2. fallback: the reference is a pure-Python package, so what pip would have installed is the `agent_torch/` tree itself;
   it is copied as is (notebooks and caches left out).  Nothing is modified; nothing enters the git history.

Run in the build container:  python baseline/install_reference.py
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("AGENT_TORCH_REFERENCE", "/root/reference")


def main():
    if not os.path.isdir(os.path.join(SOURCE, "agent_torch")):
        print(f"{SOURCE} holds no agent_torch package: nothing installed")
        return 1
    shutil.rmtree(TARGET, ignore_errors=True)
    work = tempfile.mkdtemp(prefix="ref_src_")
    copy = os.path.join(work, "reference")
    shutil.copytree(SOURCE, copy, ignore=shutil.ignore_patterns(".git", "docs", "*.ipynb", "__pycache__"))
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse",
           "--target", TARGET, copy]
    res = subprocess.run(cmd, capture_output=True, text=True)
    how = "pip"
    if res.returncode != 0 or not os.path.isdir(os.path.join(TARGET, "agent_torch")):
        how = "copy (pip failed: %s)" % (res.stderr.strip().splitlines() or ["?"])[-1]
        shutil.rmtree(TARGET, ignore_errors=True)
        shutil.copytree(os.path.join(copy, "agent_torch"), os.path.join(TARGET, "agent_torch"))
    shutil.rmtree(work, ignore_errors=True)
    with open(os.path.join(TARGET, "INSTALLED_BY"), "w") as f:
        f.write(how + "\n")
    print("reference installed into", TARGET, "by", how)
    return 0


if __name__ == "__main__":
    sys.exit(main())
