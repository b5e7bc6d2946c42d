"""The end-to-end regeneration driver on synthetic scenes (GPU) and its host-side sampler (CPU)."""
import json
import os

import numpy as np
import pytest


def test_overlap_bin_sampler_cpu():
    """regen samples pairs with the exact replica of the reference's sample_dataframe (bins [6,35], step 1)."""
    from multi_spatialmllm_b200.pair_engines import sample_rows_by_overlap
    rng = np.random.default_rng(0)
    ov = np.concatenate([rng.uniform(0, 60, 5000), [np.nan] * 10, [6.0, 35.0, 35.0001]])
    idx = sample_rows_by_overlap(ov, 290, 0, overlap_min=6, overlap_max=35, interval=1, np_random=np.random.default_rng(1),
                                 log=lambda *a, **k: None)
    assert len(idx) == 290 and len(set(idx.tolist())) == len(idx)
    assert ((ov[idx] >= 6.0) & (ov[idx] <= 35.0)).all()                  # [6, 35] (include_lowest), NaN never selected
    counts = np.histogram(ov[idx], bins=np.arange(6, 36))[0]
    assert counts.max() - counts.min() <= 2                               # equal share per 1 %-wide bin
    assert len(sample_rows_by_overlap(np.array([1.0, 50.0]), 10, 0, overlap_min=6, overlap_max=35, interval=1,
                                      log=lambda *a, **k: None)) == 0


@pytest.mark.gpu
def test_regen_synthetic_end_to_end(tmp_path):
    from multi_spatialmllm_b200 import regen
    out = str(tmp_path / "multispa")
    assert regen.main(["--synthetic", "3", "--out", out, "--points", "30000", "--frames", "24", "--pairs-per-scene", "60",
                       "--depth-frames-per-scene", "12"]) == 0
    man = json.load(open(os.path.join(out, "manifest.json")))
    assert man["total_samples"] > 300 and len(man["parts"]) == 1 and len(man["parts"][0]["scenes"]) == 3
    vc = [json.loads(l) for l in open(os.path.join(out, "visual_correspondence_coor_2_coor.jsonl"))]
    cm = [json.loads(l) for l in open(os.path.join(out, "camera_movement.jsonl"))]
    de = [json.loads(l) for l in open(os.path.join(out, "depth_estimation_coor.jsonl"))]
    dc = [json.loads(l) for l in open(os.path.join(out, "depth_comparison_coor.jsonl"))]
    assert len(vc) > 50 and len(cm) == 9 * len({(s["image"][0], s["image"][1]) for s in cm}) and len(de) > 20 and len(dc) > 10
    for s in vc:
        assert set(s) >= {"id", "image", "conversations", "height_list", "width_list", "question_type", "p1_list", "p2_list", "gt_value"}
        assert all(0 <= v <= 1000 for p in s["p1_list"] + s["p2_list"] for v in p) and s["gt_value"] == s["p2_list"][0]
    for s in de:
        assert s["gt_value"] > 0 and 0 <= s["ori_coordinates"][0] < 1296 and 0 <= s["ori_coordinates"][1] < 968
    for s in dc:
        d = {tuple([q["x"], q["y"]]): q["depth"] for q in s["points_info"]}
        other = [v for k, v in d.items() if list(k) != s["gt_value"]]
        if other:
            assert (d[tuple(s["gt_value"])] <= other[0]) == s["is_closer_question"]
