"""Drop-in pair-table engines (camera_movement, visual_correspondence coor / dot) and the object_perception writer against
the reference's own modules run end to end — tests/golden/make_golden_pair_engines.py.

Parity bar: identical sample dicts in identical order, identical warnings, draw calls and image file names, identical
positions of Python's and NumPy's global RNG streams afterwards.  The L2 tables the engines read are produced by this
repo's GPU path (tables.run_relations_split), the goldens from the reference's own tables."""
import gzip
import importlib.util
import json
import os
import pickle
import random

import numpy as np
import pandas as pd
import pytest

from oracle import ref_layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_generator():
    spec = importlib.util.spec_from_file_location("_golden_pair_engines", os.path.join(ROOT, "tests", "golden", "make_golden_pair_engines.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def gen():
    return _load_generator()


@pytest.fixture(scope="module")
def golden():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "pair_engines.json.gz"), "rt") as f:
        return json.load(f)


@pytest.fixture(scope="module")
def disk(gen, tmp_path_factory):
    scenes = gen.scenes()
    base = tmp_path_factory.mktemp("pair_engines")
    root = str(base / "data" / "scannet")
    info = ref_layout.write_split(scenes, root, name="scenes_train_val_info_i_D5.pkl")
    return scenes, str(base), root, info


def _handler(root, info):
    from multi_spatialmllm_b200.handler.info_handler import SceneInfoHandler
    return SceneInfoHandler(info, posed_images_root=os.path.join(root, "posed_images"),
                            instance_data_root=os.path.join(root, "scannet_instance_data"))


# ------------------------------------------------------------------------------------------------
# CPU: sampler, object_perception writer, merge
# ------------------------------------------------------------------------------------------------
def test_sample_dataframe_matches_reference_cpu(gen, golden):
    """Same rows in the same order and the same NumPy RNG position as the reference's sample_dataframe, for pandas 3
    (observed bins only) and for the pinned pandas 2.2.3 behaviour (empty bins share the quota)."""
    from multi_spatialmllm_b200 import pair_engines
    assert len(golden["sampler"]) == 2 * len(gen.sampler_cases())
    sampled_any = False
    for g in golden["sampler"]:
        seed, rows, n_all, n_zero, bins = g["case"]
        df = gen.sampler_table(seed, rows)
        df["row"] = np.arange(rows)
        np.random.seed(seed)
        got = pair_engines.sample_dataframe(df, n_all, n_zero, *bins, observed=g["observed"])
        assert [int(v) for v in got["row"]] == g["rows"], g["case"]
        after = float(np.random.random())
        assert after == g["np_random_after"], g["case"]
        np.random.seed(seed)
        sampled_any |= float(np.random.random()) != after
        assert list(got.index) == list(range(len(got))) and "overlap_group" not in got.columns
    assert sampled_any                                            # at least one case drew from the RNG


def test_sample_dataframe_edge_cases_cpu():
    from multi_spatialmllm_b200 import pair_engines
    empty = pd.DataFrame({"overlap": np.array([], float)})
    assert len(pair_engines.sample_dataframe(empty, 10, 0, 6, 35, 1)) == 0
    only_out = pd.DataFrame({"overlap": [0.0, 0.0, 50.0, np.nan, 5.999]})
    assert len(pair_engines.sample_dataframe(only_out, 10, 0, 6, 35, 1)) == 0
    assert list(pair_engines.sample_dataframe(only_out, 10, 5, 6, 35, 1)["overlap"]) == [0.0, 0.0]
    edge = pd.DataFrame({"overlap": [6.0, 7.0, 35.0, 35.0000001, 5.9999999]})
    assert sorted(pair_engines.sample_dataframe(edge, 10, 0, 6, 35, 1)["overlap"]) == [6.0, 7.0, 35.0]


def test_object_perception_writer_matches_reference_cpu(gen, golden, disk, tmp_path):
    from multi_spatialmllm_b200 import object_perception as op
    scenes, _, root, info = disk
    h = _handler(root, info)
    cov = str(tmp_path / "merged_train_object_coverage.pkl")
    with open(cov, "wb") as f:
        pickle.dump(gen.object_coverage_info(scenes), f)
    for g in golden["object_perception"]:
        od = str(tmp_path / g["dimension"])
        random.seed(g["seed"])
        op.build_lwh_qa_samples(h, cov, g["dimension"], "train", od, max_k=6, max_samples=g["max_samples"],
                                templates=golden["templates"]["object_perception"])
        assert random.random() == g["random_after"]
        files = {f: [json.loads(l) for l in open(os.path.join(od, f))] for f in sorted(os.listdir(od))}
        assert sorted(files) == sorted(g["files"])
        for name, samples in files.items():
            assert samples == g["files"][name], name


def test_object_perception_train_val_and_merge_cpu(gen, disk, tmp_path):
    """merge_object_coverage + build_train_and_val_datasets: shard files -> merged files -> per-k JSONL, val converted."""
    from multi_spatialmllm_b200 import object_perception as op
    scenes, _, root, info = disk
    h = _handler(root, info)
    full = gen.object_coverage_info(scenes)
    for split, base in (("train", tmp_path / "training"), ("val", tmp_path / "evaluation")):
        for k, (sid, v) in enumerate(full.items()):                      # one shard directory per scene, out of order
            d = base / f"{split}_{10 - 5 * k}_{15 - 5 * k}"
            d.mkdir(parents=True)
            for dim in ("height", "length", "width"):
                with open(d / f"{split}_object_coverage_{dim}_{10 - 5 * k}_{15 - 5 * k}.pkl", "wb") as f:
                    pickle.dump({sid: v}, f)
        (base / f"{split}_notashard").mkdir()
        merged = op.merge_split(split, str(base), str(base))
        assert set(merged) == {"height", "length", "width"} and merged["height"] == full
        assert list(merged["height"]) == list(full)[::-1]                # ascending shard start
    random.seed(1)
    op.build_train_and_val_datasets(h, str(tmp_path / "training"), str(tmp_path / "evaluation"), val_max_samples=3)
    tr = sorted(f for f in os.listdir(tmp_path / "training") if f.endswith(".jsonl"))
    va = sorted(f for f in os.listdir(tmp_path / "evaluation") if f.endswith(".jsonl"))
    assert tr == [f"object_perception_{d}_k{k}_train_-1.jsonl" for d in ("height", "length", "width") for k in (1, 2, 3)]
    assert va == [f"object_perception_{d}_k{k}_val_3.jsonl" for d in ("height", "length", "width") for k in (1, 2, 3)]
    assert not (tmp_path / "evaluation" / "temp").exists()
    for f in va:
        rows = [json.loads(l) for l in open(tmp_path / "evaluation" / f)]
        assert len(rows) == 3 and all("text" in r and "conversations" not in r for r in rows)


# ------------------------------------------------------------------------------------------------
# GPU: the three pair-table engines end to end
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def relations(disk):
    """The camera-info parquet, written by this repo's GPU path."""
    from multi_spatialmllm_b200 import tables
    _, base, root, info = disk
    path = os.path.join(base, "camera_info_D5.parquet")
    tables.run_relations_split(_handler(root, info), path)
    return path


def _engine(name, h, golden, **kw):
    from multi_spatialmllm_b200 import pair_engines
    cls = {"camera_movement": pair_engines.CameraMovementEngine, "visual_correspondence": pair_engines.VisualCorrespondenceCoorEngine,
           "visual_correspondence_dot": pair_engines.VisualCorrespondenceDotEngine}[name]
    return cls(h, templates=golden["templates"][name], rnd=random, **kw)


def _marks_as_log(pending, atlas):
    log = []
    for _, _, _, marks in pending:
        for (x, y, r, glyph, ox, oy, color, _) in marks:
            log.append(["circle", x, y, r, list(color), -1])
            if glyph >= 0:
                log.append(["text", atlas.chars[glyph], ox, oy, 0, 1.0, list(color), 2])
    return log


@pytest.mark.gpu
@pytest.mark.parametrize("run_index", range(7))
def test_pair_engine_matches_reference(run_index, disk, relations, golden, tmp_path):
    _, _, root, info = disk
    run = golden["runs"][run_index]
    h = _handler(root, info)
    eng = _engine(run["engine"], h, golden)
    od = str(tmp_path / "out")
    os.makedirs(od)
    kwargs = dict(run["kwargs"])
    if run["engine"] != "camera_movement":
        kwargs["warning_file"] = str(tmp_path / "warn.txt")
    logged = []
    if run["engine"] == "visual_correspondence_dot":
        from multi_spatialmllm_b200.annotate import default_atlas
        flush = eng.flush_images

        def spy(image_output_dir, *a):
            logged.extend(_marks_as_log(eng.pending_marks, default_atlas()))
            return flush(image_output_dir, *a)

        eng.flush_images = spy
    random.seed(run["seed"])
    np.random.seed(run["seed"])
    getattr(eng, run["fn"])(parquet_path=relations, output_dir=od, overlap_min=golden["bins"][0], overlap_max=golden["bins"][1],
                            interval=golden["bins"][2], **kwargs)
    after, np_after = random.random(), float(np.random.random())
    files = sorted(f for f in os.listdir(od) if f.endswith(".jsonl"))
    assert files == [run["file"]]
    got = [json.loads(l) for l in open(os.path.join(od, files[0]))]
    assert [s["id"] for s in got] == [s["id"] for s in run["samples"]]
    assert got == run["samples"]
    assert after == run["random_after"] and np_after == run["np_random_after"]
    assert eng.replans == 0
    warn = open(tmp_path / "warn.txt").read() if os.path.exists(tmp_path / "warn.txt") else ""
    assert warn == run["warnings"]
    if run["engine"] == "visual_correspondence_dot":
        assert logged == run["draw_log"]
        img_dir = os.path.join(od, "images")
        images = sorted(os.path.join(d, f)[len(img_dir) + 1:] for d, _, fs in os.walk(img_dir) for f in fs)
        assert images == run["images"]
        import cv2
        img = cv2.imread(os.path.join(img_dir, images[0]))
        assert img is not None and img.shape[:2] == (968, 1296) and img.max() > 0          # something was painted


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["visual_correspondence", "visual_correspondence_dot"])
def test_speculative_replay_equals_row_by_row(name, disk, relations, golden, monkeypatch):
    """When device answers contradict the dry run's prediction (a failed visibility re-check; a distractor landing on the
    correct pixel), planning all rows at once must give what planning one row at a time gives — the latter is the
    reference's sequential order by construction."""
    from multi_spatialmllm_b200 import pair_engines
    _, _, root, info = disk
    h = _handler(root, info)
    df = pair_engines.sample_dataframe(pd.read_parquet(relations), 10 ** 6, 0, 6, 35, 1)
    outs = []
    for plan_rows in (None, 1):
        eng = _engine(name, h, golden)
        real = eng._device_answers
        seen = {"n": 0}

        def corrupt(ctx, plan, first, real=real, seen=seen):
            ans = real(ctx, plan, first)
            for r, a in enumerate(ans, start=first):
                if a is None:
                    continue
                for rec in a:
                    if name == "visual_correspondence" and r % 11 == 3:
                        rec["flags"] = 1                                  # not visible in the second frame
                    if name == "visual_correspondence_dot" and r % 7 == 2:
                        rec["flags"] = 2 if r % 14 == 2 else 3
                        rec["pix"][2:] = (5, 5)                           # a pixel the patched randint below hits
            return ans

        eng._device_answers = corrupt
        rnd = random.Random(77)
        if name == "visual_correspondence_dot":
            real_randint = rnd.randint
            rnd.randint = lambda a, b: 5 if (b > 255 and real_randint(0, 3) == 0) else real_randint(a, b)
        eng.rnd = rnd
        samples = eng.build_samples(df, None, max_points_per_pair=1, **({"plan_rows": plan_rows} if plan_rows else {}))
        outs.append((json.loads(json.dumps(samples)), rnd.random(), eng.replans))
    assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1]
    assert outs[0][2] > 3                                              # the speculation was in fact broken several times
    assert any(s is None for s in outs[0][0]) and any(s is not None for s in outs[0][0])


@pytest.mark.gpu
def test_pair_engine_cli(disk, relations, tmp_path, monkeypatch):
    """main(): val then train, <root>/<engine>/<version>/ layout, eval `text` field, reference file names."""
    from multi_spatialmllm_b200 import pair_engines
    _, base, root, info = disk
    monkeypatch.chdir(base)                                            # default roots are relative: data/scannet/...
    common = ["--info_path", info, "--train_parquet_path", relations, "--val_parquet_path", relations,
              "--train_output_root", str(tmp_path / "training_data"), "--val_output_root", str(tmp_path / "evaluation_data"),
              "--train_max_samples", "12", "--val_max_samples", "5"]
    pair_engines.main(["visual_correspondence_coor_2_coor", *common])
    tr = [json.loads(l) for l in open(tmp_path / "training_data/visual_correspondence_coor_2_coor/v1_0/train_visual_correspondence_coor_2_coor.jsonl")]
    va = [json.loads(l) for l in open(tmp_path / "evaluation_data/visual_correspondence_coor_2_coor/v1_0/val_visual_correspondence_coor_2_coor.jsonl")]
    assert len(tr) == 12 and len(va) == 5 and all("text" in s for s in va) and all("conversations" in s for s in tr)
    pair_engines.main(["camera_movement", *common])
    d = tmp_path / "training_data/camera_movement/v1_0"
    assert sorted(os.listdir(d)) == sorted(f"{q}_train.jsonl" for q in pair_engines.CAMERA_QUESTION_TYPES)
    assert len(open(d / "total_distance_train.jsonl").readlines()) == 12


@pytest.mark.gpu
def test_l2_tables_cli_single_rank(disk, relations, tmp_path):
    """`python -m multi_spatialmllm_b200.tables <table>`: the main() of MVI / CFR / COV on one GPU; the relations file is
    the one the split driver writes, the visibility file is readable by VisibilityInfoHandler."""
    from multi_spatialmllm_b200 import tables
    from multi_spatialmllm_b200.handler.info_handler import VisibilityInfoHandler
    scenes, _, root, info = disk
    common = ["--splits", "val", "--val_info_path", info, "--posed_images_root", os.path.join(root, "posed_images"),
              "--instance_data_root", os.path.join(root, "scannet_instance_data")]
    res = tables.main(["relations", *common, "--val_output", str(tmp_path / "cam.parquet")])
    assert res["val"]["rows"] == len(pd.read_parquet(relations))
    pd.testing.assert_frame_equal(pd.read_parquet(tmp_path / "cam.parquet"), pd.read_parquet(relations))
    tables.main(["visibility", *common, "--val_output", str(tmp_path / "vis.parquet")])
    v = VisibilityInfoHandler(str(tmp_path / "vis.parquet"))
    ids = v.get_image_to_points_info(scenes[1].scene_id, scenes[1].image_ids[2])
    assert ids == sorted(ids) and len(ids) > 0
    tables.main(["object_visibility", *common, "--val_output", str(tmp_path / "ov.pkl")])
    with open(tmp_path / "ov.pkl", "rb") as f:
        ov = pickle.load(f)
    assert list(ov) == [s.scene_id for s in scenes] and all("object_to_images" in x for x in ov.values())


def test_coverage_shard_drivers_layout_cpu(gen, disk, tmp_path, monkeypatch):
    """object_perception.coverage_main / process_split_objects: the shard directories and file names of
    single_object_coverage_finder.py (SOCF:380-420, 296-340) feed merge_split and the writer unchanged.  The per-scene search is
    stubbed here (it needs the GPU; tests/test_handler_tables.py covers it) — this checks the host-side plumbing."""
    from multi_spatialmllm_b200 import object_perception as op, tables
    scenes, _, root, info = disk
    h = _handler(root, info)
    full = gen.object_coverage_info(scenes)
    ov = {s.scene_id: {"object_to_images": {o: [] for o in full[s.scene_id]}} for s in scenes}
    ov_file = tmp_path / "object_visibility.pkl"
    with open(ov_file, "wb") as f:
        pickle.dump(ov, f)
    seen = []
    monkeypatch.setattr(tables, "device_scene_from_handler", lambda hh, sid, dev="cuda:0": (None, []))

    def fake_scene(scene_infos, scene_id, sc, image_ids, object_visibility, max_images=5):
        seen.append(scene_id)
        assert object_visibility == ov[scene_id]
        return {o: {"height": d, "length": {1: d.get(1, [])}, "width": {}} for o, d in full[scene_id].items()}

    monkeypatch.setattr(tables, "object_coverage_scene", fake_scene)
    splits = {"val": {"scene_info_path": h, "visibility_parquet_file": None, "object_visibility_file": str(ov_file),
                      "output_dir": str(tmp_path / "evaluation_data" / "object_perception")}}
    ids = h.get_all_scene_ids()
    op.coverage_main(["--split", "val", "--start", "1", "--end", "2"], splits=splits)
    op.coverage_main(["--split", "val", "--start", "0", "--end", "1"], splits=splits)
    assert seen == [ids[1], ids[0]]
    base = tmp_path / "evaluation_data" / "object_perception"
    assert sorted(os.listdir(base)) == ["val_0_1", "val_1_2"]
    assert sorted(os.listdir(base / "val_0_1")) == [f"val_object_coverage_{d}_0_1.pkl" for d in ("height", "length", "width")]
    merged = op.merge_split("val", str(base), str(base))
    assert merged["height"] == {sid: full[sid] for sid in ids} and list(merged["height"]) == ids
    assert merged["width"] == {sid: {o: {} for o in full[sid]} for sid in ids}
    with pytest.raises(ValueError):
        op.coverage_main(["--split", "test", "--start", "0"], splits=splits)
    res = op.process_split_objects("val", h, None, str(ov_file), str(tmp_path / "whole"))
    assert sorted(os.listdir(tmp_path / "whole")) == ["coverage_finding_warning_objects.txt"] + \
        [f"val_object_coverage_{d}.pkl" for d in ("height", "length", "width")]
    assert res["height"] == merged["height"]
    random.seed(1)
    out = op.build_lwh_qa_samples(h, str(base / "merged_val_object_coverage_height.pkl"), "height", "val", str(tmp_path / "qa"))
    assert sum(len(v) for v in out.values()) > 0 and all(f.startswith("object_perception_height_k") for f in os.listdir(tmp_path / "qa"))


def test_reference_templates_are_read_without_importing_cpu(tmp_path):
    """load_reference_templates parses the phrasings out of a checkout through the AST: module-level side effects of the
    reference files (RNG seeding, heavy imports, a failing import) never run.  A miniature checkout stands in here; when the
    real one is mounted (build container) it is read too."""
    from multi_spatialmllm_b200 import object_perception as op, pair_engines
    root = tmp_path / "ref"
    (root / "spatial_engine/camera_movement").mkdir(parents=True)
    (root / "spatial_engine/visual_correspondence").mkdir(parents=True)
    (root / "spatial_engine/object_perception").mkdir(parents=True)
    (root / "spatial_engine/camera_movement/TEMPLATES.py").write_text(
        'QUESTION_TEMPLATES = {"x_movement": ["q {x_movement}"]}\nTASK_DESCRIPTION = ["t0", "t1"]\nANSWER_TEMPLATES = {"x_movement": ["a"]}\n')
    for name in ("visual_correspondence_qa_engine_coor_2_coor.py", "visual_correspondence_qa_engine_dot_2_multichoice.py"):
        (root / "spatial_engine/visual_correspondence" / name).write_text(
            'import this_module_does_not_exist\nraise RuntimeError("must not run")\nTASK_DESCRIPTION = ["vt"]\n'
            'TEMPLATES = {"questions": ["vq {x1}"], "answers": ["va {x2}"]}\n')
    (root / "spatial_engine/object_perception/single_object_perception_engine.py").write_text(
        'import random\nrandom.seed(1)\nASK_DESCRIPTION = ["ot"]\nQUESTION_TEMPLATES = ["oq {dimension}"]\nANSWER_TEMPLATES = ["oa {value_mm}"]\n')
    t = pair_engines.load_reference_templates(str(root), "camera_movement")
    assert t == {"task": ["t0", "t1"], "questions": {"x_movement": ["q {x_movement}"]}, "answers": {"x_movement": ["a"]}}
    for eng in ("visual_correspondence", "visual_correspondence_dot"):
        assert pair_engines.load_reference_templates(str(root), eng) == {"task": ["vt"], "questions": ["vq {x1}"], "answers": ["va {x2}"]}
    assert op.load_reference_templates(str(root)) == {"task": ["ot"], "questions": ["oq {dimension}"], "answers": ["oa {value_mm}"]}
    real = os.environ.get("MSP_REFERENCE_ROOT", "/root/reference")
    if os.path.isdir(os.path.join(real, "spatial_engine")):
        state = random.getstate()
        cm = pair_engines.load_reference_templates(real, "camera_movement")
        assert set(cm["questions"]) == set(pair_engines.CAMERA_QUESTION_TYPES) and len(cm["task"]) >= 5
        vc = pair_engines.load_reference_templates(real, "visual_correspondence")
        assert all("{x1}" in q for q in vc["questions"]) and len(vc["answers"]) >= 5
        vd = pair_engines.load_reference_templates(real, "visual_correspondence_dot")
        assert all("{correct_label}" in a for a in vd["answers"])
        so = op.load_reference_templates(real)
        assert len(so["task"]) >= 5 and all("{value_mm}" in a for a in so["answers"])
        assert random.getstate() == state                      # nothing was executed: the global RNG is untouched


def test_sample_dataframe_differential_against_live_reference_cpu():
    """When the reference checkout is mounted (build container): 60 random overlap tables through the reference's own
    sample_dataframe and through this repo's, same rows in the same order and the same NumPy RNG position every time."""
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference checkout not mounted")
    import contextlib
    import io
    import sys
    from multi_spatialmllm_b200 import pair_engines
    ref_shim.import_reference()
    sys.path.insert(0, os.path.join(ref_shim.REFERENCE_ROOT, "spatial_engine", "camera_movement"))
    with contextlib.redirect_stdout(io.StringIO()):
        from spatial_engine.camera_movement import camera_movement_engine_train_val as cme
    rng = np.random.default_rng(7)
    for case in range(60):
        rows = int(rng.integers(0, 300))
        ov = rng.uniform(-5, 110, rows)
        ov[rng.random(rows) < 0.15] = 0.0
        ov[rng.random(rows) < 0.05] = np.nan
        snap = rng.random(rows) < 0.2
        ov[snap] = np.round(ov[snap])
        lo = int(rng.integers(0, 40))
        step = int(rng.choice([1, 2, 5, 10]))
        hi = lo + step * int(rng.integers(1, 30))
        n_all, n_zero = int(rng.integers(0, 2 * rows + 2)), int(rng.integers(0, 20))
        df = pd.DataFrame({"overlap": ov, "row": np.arange(rows)})
        np.random.seed(case)
        with contextlib.redirect_stdout(io.StringIO()):
            want = cme.sample_dataframe(df, n_all, n_zero, lo, hi, step)
        want_after = float(np.random.random())
        np.random.seed(case)
        with contextlib.redirect_stdout(io.StringIO()):
            got = pair_engines.sample_dataframe(df, n_all, n_zero, lo, hi, step)
        assert [int(v) for v in got["row"]] == [int(v) for v in want["row"]], (case, rows, lo, hi, step, n_all, n_zero)
        assert float(np.random.random()) == want_after, case


def test_scene_cache_evicts_least_recently_used_cpu(monkeypatch):
    """SceneCache keeps scenes resident up to its byte budget and drops the least recently used first."""
    import torch
    from multi_spatialmllm_b200 import pair_engines, tables
    loads = []

    class FakeScene:
        def __init__(self, nbytes):
            self.pts, self.depth, self.mask = torch.zeros(nbytes // 2, dtype=torch.uint8), torch.zeros(nbytes // 2, dtype=torch.uint8), None

        def project_visibility(self):
            self.mask = torch.zeros(16, dtype=torch.int32)

    def fake_load(scene_infos, scene_id, device="cuda:0"):
        loads.append(scene_id)
        return FakeScene(1000), ["00000", "00005"]

    monkeypatch.setattr(tables, "device_scene_from_handler", fake_load)
    cache = pair_engines.SceneCache(None, device="cpu", budget_bytes=2300)          # two scenes of ~1064 B fit, three do not
    a, rows = cache.get("a")
    assert rows == {"00000": 0, "00005": 1} and a.mask is not None
    cache.get("b")
    assert cache.get("a")[0] is a and loads == ["a", "b"]                            # a hit, and "a" is now the most recent
    cache.get("c")                                                                   # evicts "b", the least recently used
    assert cache.get("a")[0] is a and loads == ["a", "b", "c"]
    cache.get("b")
    assert loads == ["a", "b", "c", "b"]
