"""No-GPU checks of the boundary: the library loads, exports what include/msp.h declares, host helpers work."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "msp.h")).read()
    return sorted(set(re.findall(r"MSP_API\s+[\w\s\*]+?\b(msp_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from multi_spatialmllm_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(_lib.lib, name), f"{name} declared in msp.h but not exported by libmsp.so"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared          # the ctypes table covers the whole header
    assert _lib.version() == 101
    assert _lib.lib.msp_last_error() is not None


def test_record_layout_matches_header():
    from multi_spatialmllm_b200 import engine
    dt = engine.PAIR_RECORD_DTYPE
    assert dt.itemsize == 128
    assert dt.fields["overlap"][1] == 8 and dt.fields["rel_disp"][1] == 40 and dt.fields["point_id"][1] == 64
    assert dt.fields["permille"][1] == 72 and dt.fields["depth_mm"][1] == 88 and dt.fields["pix"][1] == 96
    assert ctypes.sizeof(engine.MspCamera) == 24


def test_argument_validation_needs_no_gpu():
    """EINVAL paths return before any CUDA call."""
    from multi_spatialmllm_b200 import _lib
    rc = _lib.lib.msp_project_visibility(None, 10, 7, None, None, 1, None, None, None, 4, None, None, None)
    assert rc == -1 and b"pts_dtype" in _lib.lib.msp_last_error()
    rc = _lib.lib.msp_pair_annotate(None, 10, 4, None, 1, None, 5, None, 0, None, None, None, None, None, None, None, None)
    assert rc == -1 and b"sel_mode" in _lib.lib.msp_last_error()
    assert _lib.lib.msp_pose_prepare(None, None, 0, None, None, None, None, None) == 0     # empty input is fine


def test_mask_stride_and_synth_determinism():
    from multi_spatialmllm_b200 import engine, synth
    assert engine.mask_stride_words(1) == 4 and engine.mask_stride_words(128) == 4 and engine.mask_stride_words(129) == 8
    assert engine.mask_stride_words(200_000) == 6252
    a = synth.make_scene(3, 5000, 4)
    b = synth.make_scene(3, 5000, 4)
    np.testing.assert_array_equal(a.points, b.points)
    np.testing.assert_array_equal(a.c2w, b.c2w)
    assert a.points.dtype == np.float32 and a.n_points == 5000 and a.c2w.shape == (4, 4, 4)
    assert np.array_equal(np.round(a.c2w, 6), a.c2w)           # %f-style pose text
    v = synth.make_video(1, n_frames=10, n_tracks=16)
    assert v.tracks_xyz.shape == (10, 16, 3) and v.tracks_xyz.dtype == np.float32


def test_engine_refuses_to_run_without_cuda():
    import pytest
    import torch
    from multi_spatialmllm_b200 import engine, synth
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    s = synth.make_scene(0, 100, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.DeviceScene(s.points, s.K, s.c2w)


def test_package_command_table_cpu(capsys):
    """`python -m multi_spatialmllm_b200 <reference script>`: every reference script name resolves to a callable entry point."""
    import importlib
    from multi_spatialmllm_b200 import __main__ as cli
    assert cli.main([]) == 0 and "calculate_frames_relations" in capsys.readouterr().out
    assert cli.main(["no_such_script"]) == 2
    for name, (module, fn, lead, ref) in cli.COMMANDS.items():
        mod = importlib.import_module(f"multi_spatialmllm_b200.{module}")
        assert callable(getattr(mod, fn)), name
    with pytest.raises(SystemExit) as e:                      # argparse --help of a dispatched command exits 0
        cli.main(["visual_correspondence_qa_engine_coor_2_coor.py", "--help"])
    assert e.value.code == 0


def test_bench_byte_model_follows_the_kernel_tile_size():
    """bench.py's compulsory-bytes model (roofline.frac) uses the tile size and the prepared-tile record size of
    msp_cull9.cuh; keep them in step with the header (CullTilePrep: re-centred points, boxes, header, id table, point slots)."""
    import importlib.util
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "multi_spatialmllm_b200", "csrc", "msp_cull9.cuh")).read()
    threads = int(re.search(r"#define MSP_C9_THREADS (\d+)", src).group(1))
    sg_per_warp = int(re.search(r"constexpr int kC9SgPerWarp = (\d+);", src).group(1))
    slots = int(re.search(r"constexpr int kC9Slots = (\d+);", src).group(1))
    sgs = threads // 32 * sg_per_warp
    points = sgs * 64
    prep = sgs * 3 * 32 * 8 + sgs * 2 * 16 + (3 + 3 + 2) * 4 + slots * 2 + sgs + points
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.TILE_POINTS == points
    assert bench.TILE_PREP_BYTES == prep and prep % 16 == 0
