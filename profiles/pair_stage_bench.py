"""All-pairs stage of one config-4-sized scene (shared-frame model of SURVEY.md §8d: N = 200 k points, F = 375 frames,
P = F(F-1)/2 = 70 125 pairs): kernel times by CUDA events for msp_pair_overlap and msp_pair_annotate over every i<j pair,
pairs/s and the algorithmic rate B_pairstage = 2*N/8 + record bytes per pair against the measured HBM peak.

Usage (GPU box):  python profiles/pair_stage_bench.py > gpurun_out/pair_stage.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from multi_spatialmllm_b200 import engine, synth  # noqa: E402

N, F = int(os.environ.get("PB_POINTS", 200_000)), int(os.environ.get("PB_FRAMES", 375))
s = synth.make_scene(seed=0, n_points=N, n_frames=F, n_objects=12)
sc = engine.DeviceScene(s.points, s.K, s.c2w, axis_align=s.axis_align, instance_ids=s.instance_ids,
                        num_objects=s.num_objects, image_hw=s.image_hw, depth_hw=s.depth_hw)
sc.render_depth(1)
sc.project_visibility()
ii, jj = np.triu_indices(F, k=1)
pairs = torch.from_numpy(np.stack([ii, jj], 1).astype(np.int32)).cuda()
P = int(pairs.shape[0])
sel = torch.zeros(P, dtype=torch.int32, device="cuda")
rec = torch.empty((P, 128), dtype=torch.uint8, device="cuda")
all_pairs = getattr(sc, "all_pairs_overlap", None)


def events(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except OSError:
    pass
peak = float(peaks.get("hbm_gbs", 7700.0))
out = {"points": N, "frames": F, "pairs": P, "mask_bytes_per_frame": sc.stride * 4, "hbm_peak_gbs": peak,
       "peak_source": "MEASURED_PEAKS.json" if peaks else "B200_PROFILING.md fallback"}
for name, fn, rec_bytes in (("pair_overlap", lambda: sc.pair_overlap(pairs), 16),
                            ("pair_annotate", lambda: sc.pair_annotate(pairs, sel, 0, out=rec), 128)):
    ms = events(fn)
    alg = P * (2 * N / 8 + rec_bytes)
    out[name] = {"ms": ms, "pairs_per_s": P / (ms * 1e-3), "algorithmic_gbs": alg / (ms * 1e-3) / 1e9,
                 "frac_of_hbm_peak": alg / (ms * 1e-3) / 1e9 / peak}
if all_pairs is not None:
    ms = events(lambda: all_pairs())
    alg = P * (2 * N / 8 + 8)
    out["all_pairs_overlap_tiled"] = {"ms": ms, "pairs_per_s": P / (ms * 1e-3), "algorithmic_gbs": alg / (ms * 1e-3) / 1e9,
                                      "frac_of_hbm_peak": alg / (ms * 1e-3) / 1e9 / peak}
    inter, uni = all_pairs()
    i0, u0, _ = sc.pair_overlap(pairs)
    out["all_pairs_overlap_tiled"]["identical_counts"] = bool(torch.equal(inter[ii, jj].cpu() if inter.dim() == 2 else inter.cpu(), i0.cpu())
                                                              and torch.equal(uni[ii, jj].cpu() if uni.dim() == 2 else uni.cpu(), u0.cpu()))
print(json.dumps(out))
