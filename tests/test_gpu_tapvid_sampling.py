"""TAPVid-3D per-video sample generation (multi_spatialmllm_b200.tapvid) against the reference's own
generate_qa_training_single_scene (OM-coord:406-575) on seeded synthetic clips — tests/golden/make_golden_tapvid_sampling.py.
Parity bar: the same samples in the same order (ids, integers, float vectors bit-identical, conversation text) and the
same position of Python's RNG stream afterwards."""
import gzip
import json
import os
import random

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _golden():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "tapvid_sampling.json.gz"), "rt") as f:
        return json.load(f)


def _clip(seed, T, N, G):
    from multi_spatialmllm_b200 import synth
    v = synth.make_video(seed, n_frames=T, n_tracks=N, n_groups=G, image_hw=(36, 64))
    v.tracks_xyz[:, : N // 6] = v.tracks_xyz[:1, : N // 6]
    return v


def test_pair_sequence_behaves_like_the_reference_lists_cpu():
    """random.choice / random.sample on the array-backed sequence draw exactly what they draw on a list of tuples."""
    from multi_spatialmllm_b200.tapvid import _Pairs
    d, a, b = np.arange(50) * 0.5, np.arange(50), np.arange(50) + 1
    seq, lst = _Pairs(d, a, b), list(zip(d, a, b))
    for k in (1, 3, 7, 30, 50):
        random.seed(k)
        got = random.sample(seq, k)
        random.seed(k)
        assert got == random.sample(lst, k)
    random.seed(5)
    got = random.choice(seq)
    random.seed(5)
    assert got == random.choice(lst)
    assert list(seq[3:5]) == lst[3:5] and len(seq[10:]) == 40


@pytest.mark.gpu
@pytest.mark.parametrize("case_index", [0, 1])
def test_video_samples_match_reference(case_index, tmp_path):
    from multi_spatialmllm_b200 import streams, tapvid
    g = _golden()
    case = g["cases"][case_index]
    vs, T, N, G, qt, npg, ratio, rs = case["case"]
    v = _clip(vs, T, N, G)
    tp = streams.Templates({f"object_movement:{qt}": g["templates"]})
    video = {"images_jpeg_bytes": None, "fx_fy_cx_cy": v.fxfycxcy, "tracks_XYZ": v.tracks_xyz, "visibility": v.visibility,
             "extrinsics_w2c": v.w2c}
    random.seed(rs)
    out = tapvid.generate_video_samples(video, v.video_id, qt, npoints_per_group=npg, npairs_per_bin=1e8, image_hw=(36, 64),
                                        augment=True, augment_ratio=ratio, templates=tp, rnd=random)
    after = random.random()
    want = case["samples"]
    assert [e["id"] for e in out] == [w["id"] for w in want]
    for e, w in zip(out, want):
        assert e["gt_value"] == w["gt_value"], e["id"]                       # int mm, or the float64 vector bit for bit
        assert list(e["p1"]) == w["p1"] and list(e["p2"]) == w["p2"]
        assert (e["point_moving"], e["cam_moving"]) == (w["point_moving"], w["cam_moving"])
        assert e["image"] == w["image"] and [e["height_list"][0], e["width_list"][0]] == w["hw"]
        assert e["conversations"] == w["conversations"], e["id"]
        assert e["question_type"] == qt
    assert after == case["random_after"]                                      # same number and kind of draws
    assert sum(w["point_moving"] == 0 for w in want) > 0 and json.dumps(out[0])


@pytest.mark.gpu
def test_all_pair_distances_numpy_order():
    """msp_track_all_pair_distances == np.linalg.norm(p2 - p1, axis=1) bit for bit, pairs in i-major order, ragged tracks."""
    from multi_spatialmllm_b200 import tapvid
    from multi_spatialmllm_b200.engine import DeviceVideo
    v = _clip(3, 40, 32, 3)
    vid = DeviceVideo(v.tracks_xyz, v.w2c, v.fxfycxcy, (36, 64)).project()
    world = vid.world.cpu().numpy()
    rng = np.random.default_rng(0)
    tracks = [5, 0, 17, 31]
    vis = [np.sort(rng.choice(40, k, replace=False)) for k in (40, 2, 13, 0)]
    res = tapvid.all_pair_distances(vid, tracks, vis)
    assert len(res) == 4
    for t, vf, (d, f1, f2) in zip(tracks, vis, res):
        pairs = [(i, j) for i in range(len(vf)) for j in range(i + 1, len(vf))]          # OM-coord:485-486
        assert len(d) == len(pairs)
        if pairs:
            fp = np.array(pairs)
            np.testing.assert_array_equal(f1, vf[fp[:, 0]])
            np.testing.assert_array_equal(f2, vf[fp[:, 1]])
            np.testing.assert_array_equal(d, np.linalg.norm(world[f2, t] - world[f1, t], axis=1))
    assert tapvid.all_pair_distances(vid, [], []) == []


@pytest.mark.gpu
def test_generate_split_writes_jsonl(tmp_path):
    import cv2
    from multi_spatialmllm_b200 import tapvid
    files = []
    for seed in (21, 22):
        v = _clip(seed, 16, 36, 2)
        ok, jpg = cv2.imencode(".jpg", np.zeros((36, 64, 3), np.uint8))
        p = str(tmp_path / f"{v.video_id}.npz")
        np.savez(p, images_jpeg_bytes=np.array([jpg.tobytes()] * 16, dtype=object), fx_fy_cx_cy=v.fxfycxcy,
                 tracks_XYZ=v.tracks_xyz, visibility=v.visibility, extrinsics_w2c=v.w2c)
        files.append(p)
    out = str(tmp_path / "om.jsonl")
    n = tapvid.generate_split(files, out, img_output_dir=str(tmp_path / "img"), max_samples=150)
    rows = [json.loads(l) for l in open(out)]
    assert n == len(rows) == 150
    assert os.path.exists(tmp_path / "img" / "video00021" / "00015.jpg")
    assert all(r["height_list"] == [36, 36] and r["width_list"] == [64, 64] for r in rows)
    np.savez(str(tmp_path / "noext.npz"), images_jpeg_bytes=np.array([b""], dtype=object), fx_fy_cx_cy=np.ones(4),
             tracks_XYZ=np.zeros((1, 1, 3)), visibility=np.ones((1, 1), bool))
    with pytest.raises(NotImplementedError):
        tapvid.load_video(str(tmp_path / "noext.npz"))


@pytest.mark.gpu
def test_dot_variant_matches_reference(tmp_path):
    """single_object_movement_engine_dot.py: same triples, `_ann` ids, vector gt (its question_type test never matches),
    one circle per first-seen (frame, point) with the colour drawn from the RNG; files written with the reference's names."""
    import cv2
    from multi_spatialmllm_b200 import streams, tapvid
    g = _golden()
    case = g["dot_cases"][0]
    vs, T, N, G, qt, npg, ratio, rs = case["case"]
    v = _clip(vs, T, N, G)
    ok, jpg = cv2.imencode(".jpg", np.full((36, 64, 3), 90, np.uint8))
    video = {"images_jpeg_bytes": np.array([jpg.tobytes()] * T, dtype=object), "fx_fy_cx_cy": v.fxfycxcy,
             "tracks_XYZ": v.tracks_xyz, "visibility": v.visibility, "extrinsics_w2c": v.w2c}
    tp = streams.Templates({f"object_movement:{qt}": g["templates"]})
    # marks first (no rendering), to compare every circle with the reference's cv2.circle calls
    from multi_spatialmllm_b200.engine import DeviceVideo
    random.seed(rs)
    vid = DeviceVideo(v.tracks_xyz, v.w2c, v.fxfycxcy, (36, 64)).project()
    groups, _ = vid.rigid_body_groups()
    pairs = tapvid.sample_video_pairs(vid, v.visibility, groups, npg, 1e8, True, ratio, rnd=random)
    tri = [(int(p["frame1"]), int(p["frame2"]), int(p["point_index"])) for p in pairs]
    samples, marks, plan = streams.object_movement_dot_samples(vid, v.video_id, tri, qt, templates=tp, rng=random)
    assert random.random() == case["random_after"]
    want = case["samples"]
    assert [e["id"] for e in samples] == [w["id"] for w in want]
    for e, w in zip(samples, want):
        assert e["gt_value"] == w["gt_value"] and e["image"] == w["image"] and e["conversations"] == w["conversations"]
        assert list(e["p1"]) == w["p1"] and list(e["p2"]) == w["p2"]
        assert (e["point_moving"], e["cam_moving"]) == (w["point_moving"], w["cam_moving"])
    got_circles = [[int(m["x"]), int(m["y"]), int(m["radius"]), [int(c) for c in m["bgr"]], -1] for m in marks]
    assert got_circles == case["circles"]                                     # centre, radius W//100, colour, filled
    assert len(plan) == len(case["circles"])
    # end to end with rendering: the reference's file set, decodable images with the dot where it belongs
    random.seed(rs)
    out = tapvid.generate_video_samples(video, v.video_id, qt, npoints_per_group=npg, image_hw=(36, 64), augment=True,
                                        augment_ratio=ratio, templates=tp, rnd=random, dot=True,
                                        img_output_dir=str(tmp_path / "ann"))
    assert [e["id"] for e in out] == [w["id"] for w in want]
    assert sorted(os.listdir(tmp_path / "ann" / v.video_id)) == case["files"]
    m0 = marks[0]
    img = cv2.imread(str(tmp_path / "ann" / plan[0][1]))
    assert img.shape == (36, 64, 3)
    if 0 <= m0["x"] < 64 and 0 <= m0["y"] < 36 and m0["radius"] >= 0:
        # radius W // 100 = 0: a single pixel, blurred by the JPEG round trip but clearly pulled towards the colour
        px = img[m0["y"], m0["x"]].astype(int)
        assert np.abs(px - m0["bgr"].astype(int)).sum() < np.abs(np.array([90, 90, 90]) - m0["bgr"].astype(int)).sum() or \
            np.abs(np.array([90, 90, 90]) - m0["bgr"].astype(int)).sum() < 60


@pytest.mark.gpu
def test_reference_named_engine_classes(tmp_path, golden_tapvid):
    """TwoFrameVideoQAEngine / TwoFrameVideoQAEngineDot: the reference's class surface (constructor, project_point,
    generate_qa_training_single_scene, generate_qa_training_data / generate_qa_eval_data, __main__ layout)."""
    import cv2
    from multi_spatialmllm_b200 import tapvid
    from oracle import geometry as og
    g = golden_tapvid
    eng = tapvid.TwoFrameVideoQAEngine("tapvid3d_total_distance", "adt")
    H, W = (int(x) for x in g["image_hw"])
    n_none = 0
    for t in range(0, g["tracks_xyz"].shape[0], 3):
        for n in range(0, g["tracks_xyz"].shape[1], 5):
            want = og.track_project_point(g["tracks_xyz"][t, n], g["fxfycxcy"], H, W)
            got = eng.project_point(g["tracks_xyz"][t, n], g["fxfycxcy"], H, W)
            assert got == want                                         # bit-identical floats, or None on both sides
            n_none += got is None
    assert eng.project_point(np.array([0.1, 0.1, -1.0], np.float32), g["fxfycxcy"], H, W) is None
    with pytest.raises(ValueError):
        eng.project_point([0.1, 0.2, 1.0 / 3.0], g["fxfycxcy"], H, W)
    # a two-subset dataset on disk in the release layout
    base, meta = tmp_path / "tapvid3d_dataset", tmp_path / "tapvid3d_dataset" / "meta_data"
    ok, jpg = cv2.imencode(".jpg", np.full((36, 64, 3), 90, np.uint8))
    for sub, seeds in (("adt", (31, 32, 33)), ("pstudio", (41, 42))):
        (base / sub).mkdir(parents=True)
        (meta / sub).mkdir(parents=True)
        names = []
        for seed in seeds:
            v = _clip(seed, 16, 36, 2)
            np.savez(str(base / sub / f"{v.video_id}.npz"), images_jpeg_bytes=np.array([jpg.tobytes()] * 16, dtype=object),
                     fx_fy_cx_cy=v.fxfycxcy, tracks_XYZ=v.tracks_xyz, visibility=v.visibility, extrinsics_w2c=v.w2c)
            names.append(v.video_id)
        (meta / sub / "train.txt").write_text("\n".join(names[:-1]) + "\n")
        (meta / sub / "val.txt").write_text(names[-1] + "\n")
    one = str(base / "adt" / "video00031.npz")
    random.seed(3)
    a = eng.generate_qa_training_single_scene(one, 5, 1e8, "", True, 0.5)
    random.seed(3)
    b = tapvid.generate_video_samples(tapvid.load_video(one), "video00031", "tapvid3d_total_distance", 5, 1e8, None, None, True, 0.5)
    assert a == b and len(a) > 0
    random.seed(0)
    eng.generate_qa_training_data(["video00031", "video00032"], str(base / "adt"), str(tmp_path), str(tmp_path / "t.jsonl"),
                                  str(tmp_path / "img"), 3, 5, True, 0.05, max_samples=40)
    rows = [json.loads(l) for l in open(tmp_path / "t.jsonl")]
    assert len(rows) == 40 and os.path.exists(tmp_path / "img" / "video00032" / "00000.jpg")
    tapvid.main(["coord", "--base_npz_file", str(base), "--meta_data_dir", str(meta), "--train_root", str(tmp_path / "tr"),
                 "--val_root", str(tmp_path / "va"), "--img_output_dir", str(tmp_path / "img2"), "--seed", "1"])
    for sub in ("adt", "pstudio"):
        for q in ("tapvid3d_total_distance", "tapvid3d_displacement_vector"):
            tr = tmp_path / "tr" / "v1_0" / sub / f"{sub}_{q}_train_15points_30pairs.jsonl"
            va = tmp_path / "va" / "v1_0" / sub / f"{sub}_{q}_val.jsonl"
            assert tr.exists() and va.exists()
            val_rows = [json.loads(l) for l in open(va)]
            assert val_rows and all("text" in r and r["question_type"] == q for r in val_rows)
    tapvid.main(["dot", "--base_npz_file", str(base), "--meta_data_dir", str(meta), "--train_root", str(tmp_path / "trd"),
                 "--val_root", str(tmp_path / "vad"), "--sub_datasets", "pstudio", "--seed", "1"])
    d = tmp_path / "vad" / "v1_0" / "pstudio"
    assert (d / "pstudio_tapvid3d_total_distance_val.jsonl").exists() and (d / "pstudio_tapvid3d_total_distance_val_orig.jsonl").exists()
    row = json.loads(open(d / "pstudio_tapvid3d_total_distance_val.jsonl").readline())
    assert row["id"].endswith("_ann") and (d / "tapvid3d_total_distance_images" / row["image"][0]).exists()
