"""Stage the UNMODIFIED reference's ``spatial_engine`` package under ``oracle/_ref/`` (git-ignored, NOT gpurun-ignored).

TEST INFRASTRUCTURE.  The reference is pure Python: there is nothing to compile, "building" it means putting its
sources where the GPU box can import them, because ``/root/reference`` exists only in the build container.  The tree is
copied byte for byte (a manifest with SHA-256 sums is written next to it so a run can prove it used the stock code);
nothing under ``oracle/_ref`` is ever committed, and no product module imports it — only ``bench.py --impl reference``
(through ``oracle/ref_shim.py`` / ``oracle/ref_arm.py``) and the live-reference tests do.

    python oracle/build_ref.py            # called by __graft_entry__.build() when /root/reference is present
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("MSP_REFERENCE_SRC", "/root/reference")
DST_ROOT = os.path.join(HERE, "_ref")


def build(verbose: bool = True) -> bool:
    src = os.path.join(SRC_ROOT, "spatial_engine")
    if not os.path.isdir(src):
        if verbose:
            print(f"[build_ref] {src} not found: keeping whatever is already under {DST_ROOT}")
        return os.path.isdir(os.path.join(DST_ROOT, "spatial_engine"))
    dst = os.path.join(DST_ROOT, "spatial_engine")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(DST_ROOT, exist_ok=True)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    manifest = {}
    for root, _, files in os.walk(dst):
        for fn in sorted(files):
            p = os.path.join(root, fn)
            with open(p, "rb") as f:
                manifest[os.path.relpath(p, DST_ROOT)] = hashlib.sha256(f.read()).hexdigest()
    head = None
    try:
        with open(os.path.join(SRC_ROOT, ".git", "HEAD")) as f:
            head = f.read().strip()
    except OSError:
        pass
    with open(os.path.join(DST_ROOT, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC_ROOT, "git_head": head, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"[build_ref] copied {len(manifest)} files of the unmodified reference to {dst}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
