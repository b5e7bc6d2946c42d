"""Drop-in depth engines (multi_spatialmllm_b200.engines) against the reference's own engine classes run end to end
(constructor -> generate_qa_training_data / generate_qa_eval_data) on two synthetic disk scenes —
tests/golden/make_golden_depth_engines.py.  Parity bar: identical sample dicts in identical order, identical annotation
draw calls and file names, identical position of Python's RNG stream afterwards."""
import gzip
import json
import os
import random

import numpy as np
import pytest

from multi_spatialmllm_b200 import synth
from oracle import ref_layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _golden():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "depth_engines.json.gz"), "rt") as f:
        return json.load(f)


@pytest.fixture(scope="module")
def disk(golden_scannet, tmp_path_factory):
    g = golden_scannet
    ids = [str(v) for v in g["image_ids"]]
    a = synth.SynthScene("scene0011_00", g["points"], g["instance_ids"], int(g["num_objects"]), g["axis_align"], g["c2w"],
                         g["K"], tuple(int(v) for v in g["image_hw"]), tuple(int(v) for v in g["depth_hw"]),
                         depth=g["depth"], image_ids=ids)
    keep = [0, 2, 3, 4, 6]
    b = synth.SynthScene("scene0012_00", g["points"], g["instance_ids"], int(g["num_objects"]), g["axis_align"], g["c2w"][keep],
                         g["K"], a.image_hw, a.depth_hw, depth=g["depth"][keep], image_ids=[ids[k] for k in keep])
    root = str(tmp_path_factory.mktemp("engines") / "data" / "scannet")
    info = ref_layout.write_split([a, b], root, name="scenes_val_info_i_D5.pkl")
    return root, info


def _marks_as_log(pending, atlas):
    """The queued marks in the order cv2 would have been called: circle, then putText when the mark carries a glyph."""
    log = []
    for _, _, marks in pending:
        for (x, y, r, glyph, ox, oy, color, _) in marks:
            log.append(["circle", x, y, r, list(color), -1])
            if glyph >= 0:
                log.append(["text", atlas.chars[glyph], ox, oy, 0, 1.0, list(color), 2])       # FONT_HERSHEY_SIMPLEX == 0
    return log


@pytest.mark.parametrize("run_index", range(6))
def test_depth_engine_matches_reference(run_index, disk, tmp_path):
    from multi_spatialmllm_b200 import engines
    from multi_spatialmllm_b200.annotate import default_atlas
    root, info = disk
    g = _golden()
    run = g["runs"][run_index]
    key, cls_name, n_max, npp, seed, mode = run["run"]
    cls = getattr(engines, cls_name)
    img_dir = str(tmp_path / "images")
    eng = cls(info, "v1_0", n_max, img_dir, None, npp, str(tmp_path / "warn.txt"),
              posed_images_root=os.path.join(root, "posed_images"), instance_data_root=os.path.join(root, "scannet_instance_data"),
              templates=g["templates"][key], rnd=random)
    logged = []
    flush = eng._flush_images

    def spy_flush(scene_id):
        logged.extend(_marks_as_log(eng.pending_marks, default_atlas()))
        flush(scene_id)

    eng._flush_images = spy_flush
    random.seed(seed)
    if mode == "train":
        data = eng.generate_qa_training_data(str(tmp_path / "out"), save_file=False)
    else:
        eng.generate_qa_eval_data(str(tmp_path / "out"))
        data = [json.loads(l) for l in open(tmp_path / "out" / f"{eng.task_name}.jsonl")]
    after = random.random()
    got = json.loads(json.dumps(data))                       # tuples -> lists, like the golden
    assert [s["id"] for s in got] == [s["id"] for s in run["samples"]]
    assert got == run["samples"]
    assert after == run["random_after"]
    assert logged == run["draw_log"]
    files = sorted(os.path.join(d, f)[len(img_dir) + 1:] for d, _, fs in os.walk(img_dir) for f in fs) if os.path.isdir(img_dir) else []
    assert files == run["files"]
    if files:
        import cv2
        img = cv2.imread(os.path.join(img_dir, files[0]))
        assert img is not None and img.shape[:2] == (968, 1296)


def test_engine_cli_writes_both_splits(disk, tmp_path, monkeypatch):
    """The __main__ of the reference (argparse flags, output layout <dir>/<version>/<task>.jsonl, eval `text` field)."""
    from multi_spatialmllm_b200 import engines
    root, info = disk
    monkeypatch.chdir(os.path.dirname(os.path.dirname(root)))       # default roots are relative: data/scannet/...
    engines.main(["depth_estimation_coor", "--train_scene_info_path", info, "--val_scene_info_path", info,
                  "--train_all_max_samples", "7", "--val_all_max_samples", "3",
                  "--output_dir_train", str(tmp_path / "train"), "--output_dir_val", str(tmp_path / "val")])
    tr = [json.loads(l) for l in open(tmp_path / "train" / "v1_0" / "depth_estimation_coor.jsonl")]
    va = [json.loads(l) for l in open(tmp_path / "val" / "v1_0" / "depth_estimation_coor.jsonl")]
    assert len(tr) == 7 and len(va) == 3 and all("text" in s for s in va) and all("text" not in s for s in tr)
    assert {"id", "image", "conversations", "height_list", "width_list", "question_type", "gt_value", "ori_coordinates"} <= set(tr[0])
