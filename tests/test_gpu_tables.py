"""Device-formatted tables (msp_visibility_json_*, tables.visibility_arrow_table / relations_arrow_table):
the text must be byte-identical to json.dumps of the reference's lists (make_visibility_info.py:59-65), the row order the
reference's, and the parquet files readable by the reference's readers.  Parity bar: exact bytes."""
import json
import types

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu


def _fake_scene(mask_bool, valid=None, device="cuda:0"):
    """A DeviceScene stand-in that carries only what the formatter reads: bit masks, validity, sizes."""
    import torch
    from multi_spatialmllm_b200.engine import mask_stride_words
    F, n = mask_bool.shape
    stride = mask_stride_words(n)
    bits = np.zeros((F, stride * 32), np.uint8)
    bits[:, :n] = mask_bool
    words = np.packbits(bits, axis=1, bitorder="little").view(np.uint32)
    sc = types.SimpleNamespace()
    sc.device = torch.device(device)
    sc.mask = torch.from_numpy(words.view(np.int32).copy()).to(sc.device)
    sc.valid = torch.from_numpy(np.ones(F, np.int32) if valid is None else np.asarray(valid, np.int32)).to(sc.device)
    sc.F, sc.n, sc.stride, sc.perm = F, n, stride, None
    return sc


def _expected_rows(scene_id, mask_bool, image_ids, valid):
    rows = []
    vf = [f for f in range(len(image_ids)) if valid[f]]
    for f in vf:
        rows.append((f"{scene_id}:image_to_points:{image_ids[f]}", json.dumps(np.flatnonzero(mask_bool[f]).tolist())))
    for p in range(mask_bool.shape[1]):
        imgs = sorted({image_ids[f] for f in vf if mask_bool[f, p]})          # MVI:116-118
        rows.append((f"{scene_id}:point_to_images:{p}", json.dumps(imgs)))
    return rows


@pytest.mark.parametrize("n,F,density,seed", [(1, 1, 1.0, 0), (31, 3, 0.5, 1), (32, 2, 0.0, 2), (33, 5, 0.9, 3),
                                               (1000, 7, 0.3, 4), (10_050, 9, 0.05, 5), (10_050, 4, 1.0, 6),
                                               (100_100, 6, 0.02, 7), (8193, 300, 0.1, 8)])
def test_json_text_is_byte_identical_to_json_dumps(n, F, density, seed):
    from multi_spatialmllm_b200 import tables
    rng = np.random.default_rng(seed)
    m = rng.random((F, n)) < density
    if F > 2:
        m[1] = False                                                         # a frame that sees nothing -> "[]"
    valid = np.ones(F, bool)
    if F > 3:
        valid[2] = False                                                     # an invalid pose: no row, not listed (IH:184-189)
    # ids of different lengths, NOT in sorted order, so that point_to_images must re-sort them as strings
    image_ids = [("%05d" % (7 * (F - f))) if f % 3 else ("f%d" % (F - f)) for f in range(F)]
    sc = _fake_scene(m, valid)
    t = tables.visibility_arrow_table(sc, "scene0000_00", image_ids)
    got = list(zip(t["key"].to_pylist(), t["values"].to_pylist()))
    want = _expected_rows("scene0000_00", m, image_ids, valid)
    assert len(got) == len(want)
    assert got == want


def test_json_formatter_rejects_ids_json_would_escape():
    from multi_spatialmllm_b200 import tables
    sc = _fake_scene(np.ones((2, 40), bool))
    with pytest.raises(ValueError):
        tables.visibility_arrow_table(sc, "s", ['a"b', "c"])
    with pytest.raises(ValueError):
        tables.visibility_arrow_table(sc, "s", ["é", "c"])
    with pytest.raises(ValueError):
        tables.visibility_arrow_table(sc, "s", ["same", "same"])


def test_empty_scene_tables():
    from multi_spatialmllm_b200 import tables
    sc = _fake_scene(np.zeros((3, 10), bool), valid=[0, 0, 0])               # no valid frame at all
    t = tables.visibility_arrow_table(sc, "s", ["0", "1", "2"])
    assert t["key"].to_pylist() == [f"s:point_to_images:{p}" for p in range(10)]
    assert t["values"].to_pylist() == ["[]"] * 10


def test_golden_scene_tables_match_reference_rows(golden_scannet, tmp_path):
    """The arrow tables of the golden scene equal the rows built from the reference's own outputs, and the parquet
    written through the streaming writer reads back identical to the pandas writer the reference uses."""
    from multi_spatialmllm_b200 import engine, tables
    from test_gpu_parity import _scene_from_golden
    g = golden_scannet
    sc = _scene_from_golden(g)
    ids = [str(s) for s in g["image_ids"]]
    t = tables.visibility_arrow_table(sc, "scene0011_00", ids)
    valid = [str(v) for v in g["valid_ids"]]
    off = g["ref_i2p_off"]
    vals = t["values"].to_pylist()
    keys = t["key"].to_pylist()
    for k, v in enumerate(valid):                                             # reference lists (make_visibility_info.process_scene)
        assert keys[k] == f"scene0011_00:image_to_points:{v}"
        assert vals[k] == json.dumps(g["ref_i2p_flat"][off[k]:off[k + 1]].tolist())
    for p in range(sc.n):
        want = [valid[k] for k in np.flatnonzero(g["ref_mask"][:, p])]
        assert vals[len(valid) + p] == json.dumps(sorted(want))
    # parquet: streaming writer == pandas writer of the dict path
    tab = engine.process_scene_visibility(sc, ids)
    a, b = str(tmp_path / "dict.parquet"), str(tmp_path / "stream.parquet")
    tables.save_visibility_parquet({"scene0011_00": tab}, a)
    w = tables._ParquetStream(b)
    w.write(t)
    assert w.close() == len(vals)
    pd.testing.assert_frame_equal(pd.read_parquet(a), pd.read_parquet(b))
    # relations
    rel = engine.process_scene_relations(sc, ids)
    rt = tables.relations_arrow_table(sc, "scene0011_00", ids).to_pandas()
    want = tables.overlap_dataframe({"scene0011_00": rel})
    pd.testing.assert_frame_equal(rt, want, check_exact=True)


def test_formatter_throughput_smoke():
    """Config-2-sized scene (200k points x 64 frames, 15 % visible): the formatter must produce the same total size as
    the arithmetic says and finish quickly; prints the rate for the profile notes."""
    import time
    import torch
    from multi_spatialmllm_b200 import tables
    rng = np.random.default_rng(0)
    n, F = 200_000, 64
    m = rng.random((F, n)) < 0.15
    sc = _fake_scene(m)
    ids = ["%05d" % (5 * f) for f in range(F)]
    tables.visibility_arrow_table(sc, "scene0000_00", ids)                   # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t = tables.visibility_arrow_table(sc, "scene0000_00", ids)
    dt = time.perf_counter() - t0
    idx = np.flatnonzero(m.ravel()) % n
    digits = np.char.str_len(idx.astype(str)).sum() if len(idx) < 5_000_000 else None
    a_bytes = int(digits + 2 * len(idx)) if digits is not None else None
    b_bytes = int(m.sum() * 9) + int((~m.any(0)).sum()) * 2
    total = sum(len(v) for v in t["values"].to_pylist()[:F])
    assert a_bytes is None or total == a_bytes
    assert t["values"].nbytes >= a_bytes + b_bytes
    print(f"\nvisibility_arrow_table: {F} frames x {n} points, {t['values'].nbytes / 1e6:.1f} MB of JSON text in {dt * 1e3:.1f} ms")
