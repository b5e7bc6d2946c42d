"""Import the UNMODIFIED reference from /root/reference (build container only).

TEST INFRASTRUCTURE.  The reference needs ``mmengine`` (absent here); a stub
with the handful of names it touches is planted in ``sys.modules`` and the
read-only checkout — or, on the GPU box, the byte-for-byte copy that
``oracle/build_ref.py`` staged under ``oracle/_ref`` — is put on ``sys.path``.  Used by ``tests/golden/make_golden.py``
to produce golden vectors and by the live-reference tests, which skip when the
checkout is missing (the GPU box has no /root/reference).
"""
from __future__ import annotations

import contextlib
import os
import pickle
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")   # oracle/build_ref.py (travels to the GPU box)
REFERENCE_ROOT = os.environ.get("MSP_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isdir("/root/reference/spatial_engine") else _STAGED)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "spatial_engine"))


def _install_mmengine_stub():
    if "mmengine" in sys.modules:
        return
    mm = types.ModuleType("mmengine")

    def load(path, *a, **k):
        with open(path, "rb") as f:
            return pickle.load(f)

    def dump(obj, path, *a, **k):
        with open(path, "wb") as f:
            pickle.dump(obj, f)

    def mkdir_or_exist(d, mode=0o777):
        if d:
            os.makedirs(os.path.expanduser(d), mode=mode, exist_ok=True)

    def list_from_file(path, *a, **k):
        with open(path) as f:
            return [ln.rstrip("\n") for ln in f]

    mm.load, mm.dump, mm.mkdir_or_exist = load, dump, mkdir_or_exist
    mm.list_from_file, mm.exists = list_from_file, os.path.exists

    class TimeCounter:
        def __init__(self, *a, **k):
            pass

        def __call__(self, fn):
            return fn

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

    utils = types.ModuleType("mmengine.utils")
    dl = types.ModuleType("mmengine.utils.dl_utils")
    dl.TimeCounter = TimeCounter
    utils.dl_utils = dl
    mm.utils = utils
    sys.modules["mmengine"] = mm
    sys.modules["mmengine.utils"] = utils
    sys.modules["mmengine.utils.dl_utils"] = dl


def import_reference():
    """Returns a namespace with the reference modules the hot path lives in."""
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True          # the mount is read-only
    _install_mmengine_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    with contextlib.redirect_stdout(open(os.devnull, "w")):
        from spatial_engine.utils.scannet_utils.handler import info_handler
        from spatial_engine.camera_movement import calculate_frames_relations as cfr
        from spatial_engine.utils.scannet_utils import make_visibility_info as mvi
    ns.info_handler, ns.cfr, ns.mvi = info_handler, cfr, mvi
    return ns
