"""Drop-in L2 script modules (multi_spatialmllm_b200.make_visibility_info / calculate_frames_relations /
compute_object_visibility: the reference's function names and arguments) against the reference's own workers run on a
synthetic disk scene that reaches the warning branches — tests/golden/make_golden_l2_scripts.py.
Parity bar: identical lists / counts / overlap values and warning lines; yaw / pitch within 1e-12 (device atan2 / asin)."""
import gzip
import hashlib
import importlib.util
import json
import os
import re

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _gen():
    spec = importlib.util.spec_from_file_location("_golden_l2", os.path.join(ROOT, "tests", "golden", "make_golden_l2_scripts.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def golden():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "l2_scripts.json.gz"), "rt") as f:
        return json.load(f)


@pytest.fixture(scope="module")
def disk(tmp_path_factory):
    from multi_spatialmllm_b200.handler.info_handler import SceneInfoHandler
    gen = _gen()
    s = gen.scene()
    root = str(tmp_path_factory.mktemp("l2") / "data" / "scannet")
    info = gen.write(s, root)
    h = SceneInfoHandler(info, posed_images_root=os.path.join(root, "posed_images"),
                         instance_data_root=os.path.join(root, "scannet_instance_data"))
    return s, h, info


def _digest(obj):
    return hashlib.sha256(json.dumps(obj, sort_keys=True).encode()).hexdigest()


def _numbers(text):
    """Warning text with every float pulled out: (text with placeholders, [floats])."""
    pat = re.compile(r"np\.float64\(([^)]*)\)")
    return pat.sub("#", text), [float(x) for x in pat.findall(text)]


def test_make_visibility_info_process_scene(disk, golden, tmp_path):
    from multi_spatialmllm_b200 import make_visibility_info as mvi
    s, h, _ = disk
    w = str(tmp_path / "w.txt")
    sid, res = mvi.process_scene(s.scene_id, h, w)
    g = golden["visibility"]
    assert sid == s.scene_id and list(res["image_to_points"]) == golden["valid_ids"]
    assert {i: len(v) for i, v in res["image_to_points"].items()} == g["visible_counts"]
    assert _digest(res["image_to_points"]) == g["image_to_points_sha256"]
    assert _digest({str(k): v for k, v in res["point_to_images"].items()}) == g["point_to_images_sha256"]
    assert open(w).read() == g["warnings"] and g["warnings"].count("no in-bound points") == 2
    # run_split: same table through the device-formatted parquet, same warning lines
    out, w2 = str(tmp_path / "vis.parquet"), str(tmp_path / "w2.txt")
    mvi.run_split(h, out, w2)
    df = pd.read_parquet(out)
    i2p = {k.split(":")[2]: json.loads(v) for k, v in zip(df["key"], df["values"]) if ":image_to_points:" in k}
    assert _digest(i2p) == g["image_to_points_sha256"] and open(w2).read() == g["warnings"]
    import pickle
    pkl = str(tmp_path / "v.pkl")
    with open(pkl, "wb") as f:
        pickle.dump({sid: res}, f)
    mvi.convert_pkl_to_parquet(pkl)
    pd.testing.assert_frame_equal(pd.read_parquet(pkl.replace(".pkl", ".parquet")), df)


def test_calculate_frames_relations_process_scene(disk, golden, tmp_path):
    from multi_spatialmllm_b200 import calculate_frames_relations as cfr
    s, h, _ = disk
    w = str(tmp_path / "w.txt")
    sid, info = cfr.process_scene(s.scene_id, h, w)
    want = golden["relations"]["pairs"]
    assert [list(k) for k in info] == [p[:2] for p in want]
    for (a, b, ov, dist, yaw, pitch), v in zip(want, info.values()):
        assert (v["overlap"] == ov) or (np.isnan(ov) and np.isnan(v["overlap"]))
        assert v["distance"] == dist
        np.testing.assert_allclose([v["yaw"], v["pitch"]], [yaw, pitch], rtol=0, atol=1e-12)
    got_text, got_nums = _numbers(open(w).read())
    want_text, want_nums = _numbers(golden["relations"]["warnings"])
    assert got_text == want_text and "has something wrong" in want_text and want_text.count("no in bound points") == 2
    np.testing.assert_allclose(got_nums, want_nums, rtol=0, atol=1e-12)
    # run_split writes both files and the same warning lines
    out, w2 = str(tmp_path / "cam.parquet"), str(tmp_path / "w2.txt")
    res = cfr.run_split(h, out, w2)
    df, nz = pd.read_parquet(out), pd.read_parquet(out.replace(".parquet", "_nonzero.parquet"))
    assert res["rows"] == len(df) == len(want) and len(nz) == sum(1 for p in want if p[2] != 0.0)
    assert _numbers(open(w2).read())[0] == want_text
    cfr.save_overlap_info({sid: info}, str(tmp_path / "a.parquet"))
    cfr.save_overlap_info_nonzero({sid: info}, str(tmp_path / "a_nonzero.parquet"))
    a = pd.read_parquet(tmp_path / "a.parquet")
    assert list(a.columns) == ["scene_id", "image_id1", "image_id2", "overlap", "distance", "yaw", "pitch"]
    np.testing.assert_array_equal(a["overlap"].to_numpy(), df["overlap"].to_numpy())
    assert len(pd.read_parquet(tmp_path / "a_nonzero.parquet")) == len(nz)


def test_cfr_helpers(disk, golden):
    from multi_spatialmllm_b200 import calculate_frames_relations as cfr, make_visibility_info as mvi
    s, h, _ = disk
    for image_id, yaw, pitch in golden["yaw_pitch"]:
        E = h.get_extrinsic_matrix_align(s.scene_id, image_id)
        np.testing.assert_allclose(cfr.extract_yaw_pitch(E), (yaw, pitch), rtol=0, atol=1e-12)
        np.testing.assert_allclose(cfr.extract_yaw_pitch(E[:3, :3]), (yaw, pitch), rtol=0, atol=1e-12)
    _, res = mvi.process_scene(s.scene_id, h, os.devnull)
    masks = {i: np.isin(np.arange(s.n_points), v) for i, v in res["image_to_points"].items()}
    assert len(golden["overlap_calls"]) == 4
    for a, b, want in golden["overlap_calls"]:
        got = cfr.calculate_camera_overlap(masks, a, b)
        assert got == want or (np.isnan(want) and np.isnan(got))
    assert any(np.isnan(w) for _, _, w in golden["overlap_calls"]) and any(w == 0.0 for _, _, w in golden["overlap_calls"])
    with pytest.raises(ValueError):
        cfr.calculate_camera_overlap({"a": np.ones(4, bool), "b": np.ones(5, bool)}, "a", "b")


def test_compute_object_visibility_process_scene(disk, golden, tmp_path):
    from multi_spatialmllm_b200 import compute_object_visibility as cov
    s, h, info = disk
    sid, res, warnings = cov.process_scene(s.scene_id, h, {})
    g = golden["object_visibility"]
    assert json.loads(json.dumps(res)) == g["result"] and warnings == g["warnings"]
    assert "1" not in g["result"]["object_to_images"]                         # the wall is skipped (COV:58)
    _, empty, w = cov.process_scene("scene9999_00", h, {})
    assert empty == {"object_to_images": {}, "image_to_objects": {}} and "not found" in w[0]
    cov.process_split("val", h, None, str(tmp_path / "op"))
    import pickle
    with open(tmp_path / "op" / "object_visibility.pkl", "rb") as f:
        assert json.loads(json.dumps(pickle.load(f)[sid])) == g["result"]
    assert open(tmp_path / "op" / "warning.txt").read() == "".join(x + "\n" for x in g["warnings"])
