"""Host logic of the drop-in QA engines on CPU: the engines run unchanged, the device layer underneath them (DeviceScene's
pair_overlap / pair_annotate, SceneSession's per-frame tables) is replaced by a stand-in that answers the same requests
from the NumPy oracle.  What is checked is everything above the C ABI — row grouping, the speculative replay of Python's
RNG stream, sampling, formatting — against the goldens of the unmodified reference engines
(tests/golden/make_golden_pair_engines.py, make_golden_depth_engines.py).  The kernels' side of the same goldens is the
`-m gpu` suite (tests/test_pair_engines.py, tests/test_gpu_engines.py)."""
import gzip
import importlib.util
import json
import os
import random
from itertools import combinations

import numpy as np
import pandas as pd
import pytest
import torch

from multi_spatialmllm_b200 import synth
from multi_spatialmllm_b200.engine import PAIR_RECORD_DTYPE
from oracle import geometry as og
from oracle import ref_layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleScene:
    """Answers DeviceScene's pair requests from the oracle (float64 NumPy, the reference's own arithmetic)."""

    def __init__(self, s):
        self.s, self.pts = s, s.points.astype(np.float64)
        self.H, self.W = s.image_hw
        self.valid = [bool(np.isfinite(s.c2w[f]).all()) for f in range(s.n_frames)]
        self.E = [og.aligned_pose(s.axis_align, s.c2w[f]) if self.valid[f] else None for f in range(s.n_frames)]
        with np.errstate(all="ignore"):
            self.masks = [og.frame_visibility(self.pts, s.K, self.E[f], s.depth[f], self.H, self.W)[2] if self.valid[f]
                          else np.zeros(len(self.pts), bool) for f in range(s.n_frames)]

    def pair_overlap(self, pairs):
        pairs = np.asarray(pairs).reshape(-1, 2)
        inter = np.array([int((self.masks[a] & self.masks[b]).sum()) for a, b in pairs], np.int32)
        uni = np.array([int((self.masks[a] | self.masks[b]).sum()) for a, b in pairs], np.int32)
        with np.errstate(all="ignore"):
            return torch.from_numpy(inter), torch.from_numpy(uni), torch.from_numpy(inter / uni * 100.0)

    def pair_annotate(self, pairs, sel, sel_mode=0, out=None):
        assert sel_mode == 0
        rec = np.zeros(len(pairs), PAIR_RECORD_DTYPE)
        for r, ((a, b), k) in zip(rec, zip(np.asarray(pairs).tolist(), np.asarray(sel).tolist())):
            common = np.flatnonzero(self.masks[a] & self.masks[b])
            r["inter"], r["uni"] = len(common), int((self.masks[a] | self.masks[b]).sum())
            rel = og.pair_relations(self.E[a], self.E[b], self.masks[a], self.masks[b])
            r["overlap"], r["distance"], r["dyaw"], r["dpitch"] = rel["overlap"], rel["distance"], rel["yaw"], rel["pitch"]
            r["rel_disp"] = og.relative_displacement(self.E[a], self.E[b])[0]
            r["point_id"] = common[k] if k < len(common) else -1
            if r["point_id"] >= 0:
                for slot, f in enumerate((a, b)):
                    an = og.annotate_point(self.pts[r["point_id"]], self.s.K, self.E[f], self.s.depth[f], self.H, self.W)
                    if an is not None:
                        r["flags"] |= 1 << slot
                        r["permille"][2 * slot:2 * slot + 2] = (an["x_permille"], an["y_permille"])
                        r["depth_mm"][slot] = an["depth_mm"]
                        r["pix"][2 * slot:2 * slot + 2] = an["pix"]
        return rec

    @staticmethod
    def records_to_numpy(rec):
        return rec


class OracleCache:
    def __init__(self, scenes):
        self.scenes = {s.scene_id: (OracleScene(s), {i: f for f, i in enumerate(s.image_ids)}) for s in scenes}

    def get(self, scene_id):
        return self.scenes[scene_id]


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def pair_setup(tmp_path_factory):
    from multi_spatialmllm_b200.handler.info_handler import SceneInfoHandler
    gen = _load(os.path.join(ROOT, "tests", "golden", "make_golden_pair_engines.py"), "_golden_pair_engines_cpu")
    scenes = gen.scenes()
    base = tmp_path_factory.mktemp("pair_engines_cpu")
    root = str(base / "data" / "scannet")
    info = ref_layout.write_split(scenes, root, name="scenes_train_val_info_i_D5.pkl")
    h = SceneInfoHandler(info, posed_images_root=os.path.join(root, "posed_images"),
                         instance_data_root=os.path.join(root, "scannet_instance_data"))
    h._test_info, h._test_root = info, root                       # for the live-reference differential test
    cache = OracleCache(scenes)
    rows = []
    for s in scenes:                                               # the relations table, CFR:139-197 order
        sc, _ = cache.get(s.scene_id)
        valid = [f for f in range(s.n_frames) if sc.valid[f]]
        for a, b in combinations(valid, 2):
            with np.errstate(all="ignore"):
                rel = og.pair_relations(sc.E[a], sc.E[b], sc.masks[a], sc.masks[b])
            rows.append({"scene_id": s.scene_id, "image_id1": s.image_ids[a], "image_id2": s.image_ids[b], **rel})
    parquet = str(base / "camera_info_D5.parquet")
    pd.DataFrame(rows).to_parquet(parquet, index=False)
    with gzip.open(os.path.join(ROOT, "tests", "golden", "pair_engines.json.gz"), "rt") as f:
        golden = json.load(f)
    return h, cache, parquet, golden


@pytest.mark.parametrize("run_index", range(7))
def test_pair_engine_host_logic_matches_reference_cpu(run_index, pair_setup, tmp_path):
    from multi_spatialmllm_b200 import pair_engines
    from multi_spatialmllm_b200.annotate import default_atlas
    h, cache, parquet, golden = pair_setup
    run = golden["runs"][run_index]
    cls = {"camera_movement": pair_engines.CameraMovementEngine, "visual_correspondence": pair_engines.VisualCorrespondenceCoorEngine,
           "visual_correspondence_dot": pair_engines.VisualCorrespondenceDotEngine}[run["engine"]]
    eng = cls(h, templates=golden["templates"][run["engine"]], rnd=random, cache=cache)
    logged = []
    if run["engine"] == "visual_correspondence_dot":
        def log_only(image_output_dir, *a):                        # painting / JPEG are device work: keep the draw calls only
            atlas = default_atlas()
            for _, _, _, marks in eng.pending_marks:
                for (x, y, r, glyph, ox, oy, color, _) in marks:
                    logged.append(["circle", x, y, r, list(color), -1])
                    if glyph >= 0:
                        logged.append(["text", atlas.chars[glyph], ox, oy, 0, 1.0, list(color), 2])
            eng.pending_marks = []
        eng.flush_images = log_only
    kwargs = dict(run["kwargs"])
    if run["engine"] != "camera_movement":
        kwargs["warning_file"] = str(tmp_path / "warn.txt")
    od = str(tmp_path / "out")
    os.makedirs(od)
    random.seed(run["seed"])
    np.random.seed(run["seed"])
    getattr(eng, run["fn"])(parquet_path=parquet, output_dir=od, overlap_min=golden["bins"][0], overlap_max=golden["bins"][1],
                            interval=golden["bins"][2], **kwargs)
    after, np_after = random.random(), float(np.random.random())
    got = [json.loads(l) for l in open(os.path.join(od, run["file"]))]
    assert [s["id"] for s in got] == [s["id"] for s in run["samples"]]
    assert got == run["samples"]
    assert after == run["random_after"] and np_after == run["np_random_after"]
    assert logged == run["draw_log"] and eng.replans == 0


# ------------------------------------------------------------------------------------------------
# depth engines: SceneSession replaced by an oracle-backed session
# ------------------------------------------------------------------------------------------------
class OracleSession:
    def __init__(self, scene_info, scene_id, device=None):
        self.scene_id = scene_id
        s = OracleSession.scenes[scene_id]
        self.sc, self.row = OracleScene(s), {i: f for f, i in enumerate(s.image_ids)}

    def visible_points(self, image_id):
        from multi_spatialmllm_b200.engines import _IntSeq
        return _IntSeq(np.flatnonzero(self.sc.masks[self.row[image_id]]))

    def lookup(self, image_id, point):
        f, sc = self.row[image_id], self.sc
        an = og.annotate_point(sc.pts[point], sc.s.K, sc.E[f], sc.s.depth[f], sc.H, sc.W)
        if an is None:
            return (False, 0, 0, 0, (0, 0))
        return (True, an["x_permille"], an["y_permille"], an["depth_mm"], an["pix"])


@pytest.mark.parametrize("run_index", range(6))
def test_depth_engine_host_logic_matches_reference_cpu(run_index, golden_scannet, tmp_path, monkeypatch):
    from multi_spatialmllm_b200 import engines
    from multi_spatialmllm_b200.annotate import default_atlas
    g = golden_scannet
    ids = [str(v) for v in g["image_ids"]]
    a = synth.SynthScene("scene0011_00", g["points"], g["instance_ids"], int(g["num_objects"]), g["axis_align"], g["c2w"],
                         g["K"], tuple(int(v) for v in g["image_hw"]), tuple(int(v) for v in g["depth_hw"]),
                         depth=g["depth"], image_ids=ids)
    keep = [0, 2, 3, 4, 6]
    b = synth.SynthScene("scene0012_00", g["points"], g["instance_ids"], int(g["num_objects"]), g["axis_align"], g["c2w"][keep],
                         g["K"], a.image_hw, a.depth_hw, depth=g["depth"][keep], image_ids=[ids[k] for k in keep])
    root = str(tmp_path / "data" / "scannet")
    info = ref_layout.write_split([a, b], root, name="scenes_val_info_i_D5.pkl")
    with gzip.open(os.path.join(ROOT, "tests", "golden", "depth_engines.json.gz"), "rt") as f:
        golden = json.load(f)
    run = golden["runs"][run_index]
    key, cls_name, n_max, npp, seed, mode = run["run"]
    OracleSession.scenes = {a.scene_id: a, b.scene_id: b}
    monkeypatch.setattr(engines, "SceneSession", OracleSession)
    eng = getattr(engines, cls_name)(info, "v1_0", n_max, str(tmp_path / "images"), None, npp, str(tmp_path / "warn.txt"),
                                     posed_images_root=os.path.join(root, "posed_images"),
                                     instance_data_root=os.path.join(root, "scannet_instance_data"),
                                     templates=golden["templates"][key], rnd=random)
    logged = []

    def log_only(scene_id):
        atlas = default_atlas()
        for _, _, marks in eng.pending_marks:
            for (x, y, r, glyph, ox, oy, color, _) in marks:
                logged.append(["circle", x, y, r, list(color), -1])
                if glyph >= 0:
                    logged.append(["text", atlas.chars[glyph], ox, oy, 0, 1.0, list(color), 2])
        eng.pending_marks = []

    eng._flush_images = log_only
    random.seed(seed)
    with np.errstate(all="ignore"):
        if mode == "train":
            data = eng.generate_qa_training_data(str(tmp_path / "out"), save_file=False)
        else:
            eng.generate_qa_eval_data(str(tmp_path / "out"))
            data = [json.loads(l) for l in open(tmp_path / "out" / f"{eng.task_name}.jsonl")]
    assert json.loads(json.dumps(data)) == run["samples"]
    assert random.random() == run["random_after"] and logged == run["draw_log"]


@pytest.mark.parametrize("name", ["visual_correspondence", "visual_correspondence_dot"])
def test_speculative_replay_equals_row_by_row_cpu(name, pair_setup):
    """Device answers that contradict the dry run's prediction (failed re-check; a distractor on the correct pixel): planning
    every row at once must give what planning one row at a time gives (the reference's sequential order by construction)."""
    from multi_spatialmllm_b200 import pair_engines
    h, cache, parquet, golden = pair_setup
    df = pair_engines.sample_dataframe(pd.read_parquet(parquet), 10 ** 6, 0, 6, 35, 1)
    cls = pair_engines.VisualCorrespondenceCoorEngine if name == "visual_correspondence" else pair_engines.VisualCorrespondenceDotEngine
    outs = []
    for plan_rows in (1 << 18, 1):
        eng = cls(h, templates=golden["templates"][name], cache=cache)
        real = eng._device_answers

        def corrupt(ctx, plan, first, real=real):
            ans = real(ctx, plan, first)
            for r, a in enumerate(ans, start=first):
                for rec in a or ():
                    if name == "visual_correspondence" and r % 11 == 3:
                        rec["flags"] = 1
                    if name == "visual_correspondence_dot" and r % 7 == 2:
                        rec["flags"] = 2 if r % 14 == 2 else 3
                        rec["pix"][2:] = (5, 5)
            return ans

        eng._device_answers = corrupt
        rnd = random.Random(77)
        if name == "visual_correspondence_dot":
            real_randint = rnd.randint
            rnd.randint = lambda a, b, real_randint=real_randint: 5 if (b > 255 and real_randint(0, 3) == 0) else real_randint(a, b)
        eng.rnd = rnd
        samples = eng.build_samples(df, None, plan_rows=plan_rows, max_points_per_pair=1)
        outs.append((json.loads(json.dumps(samples)), rnd.random(), eng.replans))
    assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1]
    assert outs[0][2] > 3 and any(s is None for s in outs[0][0]) and any(s is not None for s in outs[0][0])


# ------------------------------------------------------------------------------------------------
# TAPVid-3D worker: DeviceVideo and the all-pair distance launch replaced by the oracle
# ------------------------------------------------------------------------------------------------
class OracleVideo:
    def __init__(self, tracks_xyz, w2c, fxfycxcy, image_hw, device=None):
        self.tracks = np.asarray(tracks_xyz, np.float32)
        self.T, self.N = self.tracks.shape[:2]
        self.w2c_h = np.asarray(w2c, np.float64).reshape(self.T, 4, 4)
        self.w2c = torch.from_numpy(self.w2c_h.reshape(self.T, 16).copy())
        self.intr, self.image_hw, self.device = np.asarray(fxfycxcy, np.float64), tuple(int(v) for v in image_hw), "cpu"
        self.world = self.uvn = self.permille = self.valid = None

    def project(self):
        cam = self.tracks.astype(np.float64)
        H, W = self.image_hw
        world = og.tracks_to_world(cam, self.w2c_h)
        uvn, pm = np.zeros((self.T, self.N, 2)), np.zeros((self.T, self.N, 2), np.int32)
        valid = np.zeros((self.T, self.N), np.uint8)
        for t in range(self.T):
            for n in range(self.N):
                r = og.track_project_point(cam[t, n], self.intr, H, W)
                if r is not None:
                    uvn[t, n], valid[t, n] = r, 1
                    pm[t, n] = (round(r[0] * 1000), round(r[1] * 1000))
        self.world_h = world
        self.world, self.uvn, self.permille, self.valid = (torch.from_numpy(np.ascontiguousarray(a)) for a in (world, uvn, pm, valid))
        return self

    def pair_displacement(self, triples):
        out = np.zeros((len(triples), 5))
        for k, (f1, f2, p) in enumerate(np.asarray(triples).reshape(-1, 3).tolist()):
            d = self.world_h[f2, p] - self.world_h[f1, p]
            c1, c2 = np.linalg.inv(self.w2c_h[f1]), np.linalg.inv(self.w2c_h[f2])
            out[k, 0] = np.linalg.norm(d)
            out[k, 1:4] = (self.w2c_h[f1] @ np.concatenate([d, [0.0]]))[:3]
            out[k, 4] = np.linalg.norm(c2[:3, 3] - c1[:3, 3])
        return torch.from_numpy(out)

    def rigid_body_groups(self, threshold=0.1, smoothing_factor=0.01, min_size=5):
        return og.rigid_body_groups(self.tracks, threshold, smoothing_factor, min_size)


def _oracle_all_pair_distances(vid, track_ids, visible_frames):
    if vid.world is None:
        vid.project()
    return [og.track_pair_distances(vid.world_h[:, p], np.asarray(vf)) for p, vf in zip(track_ids, visible_frames)]


@pytest.mark.parametrize("case_index", [0, 1])
def test_tapvid_worker_host_logic_matches_reference_cpu(case_index, monkeypatch):
    """sample_video_pairs (static / moving split, 10-bin sampling, augmentation: the reference's RNG stream) and the sample
    formatter over oracle answers reproduce the reference's generate_qa_training_single_scene."""
    from multi_spatialmllm_b200 import streams, tapvid
    with gzip.open(os.path.join(ROOT, "tests", "golden", "tapvid_sampling.json.gz"), "rt") as f:
        g = json.load(f)
    case = g["cases"][case_index]
    vs, T, N, G, qt, npg, ratio, rs = case["case"]
    v = synth.make_video(vs, n_frames=T, n_tracks=N, n_groups=G, image_hw=(36, 64))
    v.tracks_xyz[:, : N // 6] = v.tracks_xyz[:1, : N // 6]
    monkeypatch.setattr(tapvid, "DeviceVideo", OracleVideo)
    monkeypatch.setattr(tapvid, "all_pair_distances", _oracle_all_pair_distances)
    video = {"images_jpeg_bytes": None, "fx_fy_cx_cy": v.fxfycxcy, "tracks_XYZ": v.tracks_xyz, "visibility": v.visibility,
             "extrinsics_w2c": v.w2c}
    random.seed(rs)
    out = tapvid.generate_video_samples(video, v.video_id, qt, npoints_per_group=npg, npairs_per_bin=1e8, image_hw=(36, 64),
                                        augment=True, augment_ratio=ratio,
                                        templates=streams.Templates({f"object_movement:{qt}": g["templates"]}), rnd=random)
    after = random.random()
    got = json.loads(json.dumps(out))
    assert [e["id"] for e in got] == [w["id"] for w in case["samples"]]
    for e, w in zip(got, case["samples"]):
        for key in ("gt_value", "image", "conversations", "p1", "p2", "point_moving", "cam_moving"):
            assert e[key] == w[key], (e["id"], key)
    assert after == case["random_after"]


@pytest.mark.parametrize("engine,seed,count,kw", [("visual_correspondence", 101, 50, {"max_points_per_pair": 2}),
                                                   ("camera_movement", 103, 70, {"qtype": "pitch_movement"})])
def test_pair_engines_differential_against_live_reference_cpu(engine, seed, count, kw, pair_setup, tmp_path):
    """Seeds the goldens do not hold: the unmodified reference engine (when its checkout is mounted) and this repo's engine
    over oracle answers, same tables, same seed -> the same JSONL and the same RNG positions."""
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference checkout not mounted")
    import contextlib
    import io
    import pickle
    import sys
    from multi_spatialmllm_b200 import pair_engines
    h, cache, parquet, golden = pair_setup
    ref = ref_shim.import_reference()
    sys.path.insert(0, os.path.join(ref_shim.REFERENCE_ROOT, "spatial_engine", "camera_movement"))
    quiet = contextlib.redirect_stdout(io.StringIO())
    with quiet:
        if engine == "camera_movement":
            from spatial_engine.camera_movement import camera_movement_engine_train_val as mod
        else:
            from spatial_engine.visual_correspondence import visual_correspondence_qa_engine_coor_2_coor as mod
    t = golden["templates"][engine]
    if engine == "camera_movement":
        mod.TASK_DESCRIPTION, mod.QUESTION_TEMPLATES, mod.ANSWER_TEMPLATES = t["task"], t["questions"], t["answers"]
    else:
        mod.TASK_DESCRIPTION, mod.TEMPLATES = t["task"], {k: v for k, v in t.items() if k != "task"}
    # the reference handler on the same disk scenes (its own geometry), the oracle's visibility lists as its table
    with quiet:
        rh = ref.info_handler.SceneInfoHandler(h._test_info, posed_images_root=os.path.join(h._test_root, "posed_images"),
                                               instance_data_root=os.path.join(h._test_root, "scannet_instance_data"))
    rh.get_image_size = rh.get_image_shape
    vis = {sid: {"image_to_points": {i: np.flatnonzero(sc.masks[f]).tolist() for i, f in row_of.items() if sc.valid[f]}}
           for sid, (sc, row_of) in cache.scenes.items()}
    vis_pkl = str(tmp_path / "vis.pkl")
    with open(vis_pkl, "wb") as f:
        pickle.dump(vis, f)
    args = dict(parquet_path=parquet, overlap_min=6, overlap_max=35, interval=1, desired_count=count)
    ref_dir, our_dir = str(tmp_path / "ref"), str(tmp_path / "ours")
    os.makedirs(ref_dir), os.makedirs(our_dir)
    random.seed(seed), np.random.seed(seed)
    with quiet, np.errstate(all="ignore"):
        if engine == "camera_movement":
            mod.build_val_dataset(output_dir=ref_dir, scene_infos=rh, qtype=kw["qtype"], **args)
        else:
            mod.build_train_dataset(output_dir=ref_dir, scene_infos=rh, visibility_info_path=vis_pkl, warning_file=str(tmp_path / "rw.txt"),
                                    max_points_per_pair=kw["max_points_per_pair"], **args)
    want_after = (random.random(), float(np.random.random()))
    cls = pair_engines.CameraMovementEngine if engine == "camera_movement" else pair_engines.VisualCorrespondenceCoorEngine
    eng = cls(h, templates=t, rnd=random, cache=cache)
    random.seed(seed), np.random.seed(seed)
    with quiet:
        if engine == "camera_movement":
            eng.build_val_dataset(output_dir=our_dir, qtype=kw["qtype"], **args)
        else:
            eng.build_train_dataset(output_dir=our_dir, warning_file=str(tmp_path / "ow.txt"), max_points_per_pair=kw["max_points_per_pair"], **args)
    assert (random.random(), float(np.random.random())) == want_after
    (name,) = os.listdir(ref_dir)
    want = [json.loads(l) for l in open(os.path.join(ref_dir, name))]
    got = [json.loads(l) for l in open(os.path.join(our_dir, name))]
    assert len(want) > 20 and got == want


@pytest.mark.parametrize("clip", [(31, 26, 54, 3, "tapvid3d_total_distance", 4, 0.7, 21), (32, 18, 40, 2, "tapvid3d_displacement_vector", 6, 0.2, 22),
                                  (33, 40, 30, 1, "tapvid3d_total_distance", 2, 1.0, 23)])
def test_tapvid_worker_differential_against_live_reference_cpu(clip, monkeypatch, tmp_path):
    """Clips and seeds outside the goldens: the unmodified TwoFrameVideoQAEngine.generate_qa_training_single_scene (when the
    checkout is mounted) against this repo's worker over oracle answers — same samples, text and RNG position."""
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference checkout not mounted")
    import contextlib
    from multi_spatialmllm_b200 import streams, tapvid
    gen = _load(os.path.join(ROOT, "tests", "golden", "make_golden_tapvid_sampling.py"), "_golden_tapvid_cpu")
    ref_shim.import_reference()
    quiet = contextlib.redirect_stdout(open(os.devnull, "w"))
    with quiet:
        from spatial_engine.object_movement import single_object_movement_engine_coord as om
    vs, T, N, G, qt, npg, ratio, rs = clip
    v = gen.make_clip(vs, T, N, G)
    path = str(tmp_path / f"{v.video_id}.npz")
    gen.write_npz(path, v)
    eng = om.TwoFrameVideoQAEngine(qt, "adt")
    eng.task_description, eng.question_templates, eng.answer_templates = (gen.TEMPLATES["task"], gen.TEMPLATES["questions"],
                                                                         gen.TEMPLATES["answers"])
    random.seed(rs)
    with quiet, np.errstate(all="ignore"):
        want = eng.generate_qa_training_single_scene(path, npoints_per_group=npg, npairs_per_bin=1e8,
                                                     img_output_dir=str(tmp_path / "img"), augment=True, augment_ratio=ratio)
    want_after = random.random()
    monkeypatch.setattr(tapvid, "DeviceVideo", OracleVideo)
    monkeypatch.setattr(tapvid, "all_pair_distances", _oracle_all_pair_distances)
    video = {"images_jpeg_bytes": None, "fx_fy_cx_cy": v.fxfycxcy, "tracks_XYZ": v.tracks_xyz, "visibility": v.visibility,
             "extrinsics_w2c": v.w2c}
    random.seed(rs)
    got = tapvid.generate_video_samples(video, v.video_id, qt, npoints_per_group=npg, npairs_per_bin=1e8, image_hw=(36, 64),
                                        augment=True, augment_ratio=ratio,
                                        templates=streams.Templates({f"object_movement:{qt}": gen.TEMPLATES}), rnd=random)
    assert random.random() == want_after
    got, want = json.loads(json.dumps(got)), json.loads(json.dumps(want))
    assert [e["id"] for e in got] == [e["id"] for e in want] and len(want) > 0
    for e, w in zip(got, want):
        for key in ("gt_value", "image", "conversations", "p1", "p2", "point_moving", "cam_moving", "height_list", "width_list"):
            assert e[key] == w[key], (e["id"], key)


@pytest.mark.parametrize("key,module,cls_name,n_max,npp,seed", [
    ("depth_estimation", "depth_estimation_coor_engine", "DepthEstimationCoorQAEngine", 14, 3, 41),
    ("depth_comparison", "depth_comparison_coor_engine", "DepthComparisonCoorQAEngine", 40, 1, 42)])
def test_depth_engines_differential_against_live_reference_cpu(key, module, cls_name, n_max, npp, seed, tmp_path, monkeypatch):
    """Seeds / sizes outside the goldens: the unmodified depth engine (live, when the checkout is mounted) against this repo's
    engine over oracle answers."""
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference checkout not mounted")
    import contextlib
    import io
    import pickle
    from multi_spatialmllm_b200 import engines
    gen = _load(os.path.join(ROOT, "tests", "golden", "make_golden_depth_engines.py"), "_golden_depth_cpu")
    ref = ref_shim.import_reference()
    quiet = contextlib.redirect_stdout(io.StringIO())
    scenes = gen.scenes_from_fixture()
    root = str(tmp_path / "data" / "scannet")
    info = ref_layout.write_split(scenes, root, name="scenes_val_info_i_D5.pkl")
    cwd = os.getcwd()
    os.chdir(tmp_path)                                       # the reference engines use the relative default roots
    try:
        with quiet, np.errstate(all="ignore"):
            rh = ref.info_handler.SceneInfoHandler(info)
            vis = {s.scene_id: ref.mvi.process_scene(s.scene_id, rh, str(tmp_path / "w.txt"))[1] for s in scenes}
            pkl = os.path.join(root, "scannet_instance_data", "val_visibility_info_D5.pkl")
            with open(pkl, "wb") as f:
                pickle.dump(vis, f)
            ref.mvi.convert_pkl_to_parquet(pkl)
            mod = __import__(f"spatial_engine.depth_perception.{module}", fromlist=[cls_name])
            reng = getattr(mod, cls_name)(scene_info_path=info, version_num="v1_0", all_max_samples=n_max, image_output_dir=None,
                                          visibility_info_path=pkl.replace(".pkl", ".parquet"), max_n_points_per_image=npp,
                                          warning_file=str(tmp_path / "warn.txt"))
            t = gen.TEMPLATES[key]
            reng.task_description, reng.templates = t["task"], {k: v for k, v in t.items() if k != "task"}
            random.seed(seed)
            want = reng.generate_qa_training_data(str(tmp_path / "ref_out"), save_file=False)
        want_after = random.random()
    finally:
        os.chdir(cwd)
    OracleSession.scenes = {s.scene_id: s for s in scenes}
    monkeypatch.setattr(engines, "SceneSession", OracleSession)
    eng = getattr(engines, cls_name)(info, "v1_0", n_max, None, None, npp, str(tmp_path / "warn2.txt"),
                                     posed_images_root=os.path.join(root, "posed_images"),
                                     instance_data_root=os.path.join(root, "scannet_instance_data"), templates=t, rnd=random)
    random.seed(seed)
    with quiet, np.errstate(all="ignore"):
        got = eng.generate_qa_training_data(str(tmp_path / "our_out"), save_file=False)
    assert random.random() == want_after
    assert json.loads(json.dumps(got)) == json.loads(json.dumps(want)) and len(want) >= min(n_max, 10)
