"""End-to-end cost of the two L2 tables for one config-4-sized scene (200 k points x 375 frames):
make_visibility_info.py -> parquet and calculate_frames_relations.py -> parquet, three ways:

  device   tables.visibility_arrow_table / relations_arrow_table + the row-group streaming writer (this repo's path)
  dict     the reference-shaped dict path (engine.process_scene_* + pandas writer): GPU geometry, Python lists / json.dumps
  cpu      the reference's algorithm (oracle port, NumPy) on a sample of frames / pairs, extrapolated to the scene

Usage (GPU box):  python profiles/table_bench.py > gpurun_out/tables.json
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from multi_spatialmllm_b200 import engine, synth, tables  # noqa: E402
from multi_spatialmllm_b200._lib import check, lib  # noqa: E402
from oracle import geometry as og  # noqa: E402

N, F = int(os.environ.get("TB_POINTS", 200_000)), int(os.environ.get("TB_FRAMES", 375))
s = synth.make_scene(seed=0, n_points=N, n_frames=F, n_objects=12)
ids = ["%05d" % (5 * f) for f in range(F)]
sc = engine.DeviceScene(s.points, s.K, s.c2w, axis_align=s.axis_align, instance_ids=s.instance_ids,
                        num_objects=s.num_objects, image_hw=s.image_hw, depth_hw=s.depth_hw)
sc.render_depth(1)
sc.project_visibility()
torch.cuda.synchronize()
out = {"points": N, "frames": F, "pairs": F * (F - 1) // 2}
tmp = tempfile.mkdtemp()


def wall(fn, reps=1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, r


def events(fn, reps=5):
    fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# ---- kernels alone (CUDA events)
out["k1_ms"] = events(sc.project_visibility)
rows = torch.arange(F, dtype=torch.int32, device="cuda")
sizes = torch.empty(F + N, dtype=torch.int64, device="cuda")
idlen = torch.full((F,), 5, dtype=torch.uint8, device="cuda")
idtab = torch.from_numpy(np.frombuffer("".join(ids).encode(), np.uint8).reshape(F, 5).copy()).cuda()
st = torch.cuda.current_stream().cuda_stream


def size_kernels():
    check(lib.msp_visibility_json_points_size(sc.mask.data_ptr(), N, sc.stride, rows.data_ptr(), F, sizes.data_ptr(), st))
    check(lib.msp_visibility_json_images_size(sc.mask.data_ptr(), N, sc.stride, rows.data_ptr(), F, idlen.data_ptr(),
                                              sizes[F:].data_ptr(), st))


out["json_size_kernels_ms"] = events(size_kernels)
offs = torch.zeros(F + N + 1, dtype=torch.int64, device="cuda")
torch.cumsum(sizes, 0, out=offs[1:])
total = int(offs[-1].item())
text = torch.empty(total, dtype=torch.uint8, device="cuda")


def write_kernels():
    check(lib.msp_visibility_json_points_write(sc.mask.data_ptr(), N, sc.stride, rows.data_ptr(), F, offs.data_ptr(),
                                               text.data_ptr(), st))
    check(lib.msp_visibility_json_images_write(sc.mask.data_ptr(), N, sc.stride, rows.data_ptr(), F, idtab.data_ptr(), 5,
                                               idlen.data_ptr(), offs[F:].data_ptr(), text.data_ptr(), st))


out["json_write_kernels_ms"] = events(write_kernels)
out["json_text_bytes"] = total
out["json_write_gbs"] = total / (out["json_write_kernels_ms"] * 1e-3) / 1e9
del text

# ---- device path end to end (one scene -> parquet file)
def device_visibility():
    w = tables._ParquetStream(os.path.join(tmp, "vis_device.parquet"))
    w.write(tables.visibility_arrow_table(sc, s.scene_id, ids))
    return w.close()


device_visibility()
out["visibility_device_s"], nrows = wall(device_visibility)
out["visibility_device_arrow_only_s"], _ = wall(lambda: tables.visibility_arrow_table(sc, s.scene_id, ids))
out["visibility_rows"] = nrows


def device_relations():
    w = tables._ParquetStream(os.path.join(tmp, "cam_device.parquet"))
    w.write(tables.relations_arrow_table(sc, s.scene_id, ids))
    return w.close()


device_relations()
out["relations_device_s"], _ = wall(device_relations)


# ---- a stream of scenes through the visibility table writer (what tables.run_visibility_split does per scene): the single
#      reference-shaped file (one encode thread, overlapped with the next scene's kernels) vs the dataset layout with 8 encode threads
def stream_scenes(writers, n_scenes=8):
    import shutil
    path = os.path.join(tmp, f"vis_stream_{writers}.parquet")
    w = tables._ParquetStream(path, writers=writers)
    t0 = time.perf_counter()
    for k in range(n_scenes):
        sc.project_visibility()
        w.write(tables.visibility_arrow_table(sc, f"scene{k:04d}_00", ids))
    rows = w.close()
    dt = time.perf_counter() - t0
    if os.path.isdir(path):
        shutil.rmtree(path)
    else:
        os.remove(path)
    return n_scenes * F / dt, rows


stream_scenes(1, 2)
out["visibility_stream_frames_per_s"] = {"single_file_1_writer": stream_scenes(1)[0], "dataset_8_writers": stream_scenes(8)[0],
                                         "dataset_8_writers_24_scenes": stream_scenes(8, 24)[0],
                                         "dataset_16_writers_32_scenes": stream_scenes(16, 32)[0]}
if os.environ.get("TB_STREAM_ONLY"):
    print(json.dumps(out))
    sys.exit(0)

# ---- dict path (GPU geometry, Python formatting) — what this repo did before the device formatter
def dict_visibility():
    tab = engine.process_scene_visibility(sc, ids)
    return tables.save_visibility_parquet({s.scene_id: tab}, os.path.join(tmp, "vis_dict.parquet"))


out["visibility_dict_s"], _ = wall(dict_visibility)


def dict_relations():
    rel = engine.process_scene_relations(sc, ids)
    return tables.save_overlap_parquet({s.scene_id: rel}, os.path.join(tmp, "cam_dict.parquet"))


out["relations_dict_s"], _ = wall(dict_relations)

import pandas as pd  # noqa: E402
a, b = pd.read_parquet(os.path.join(tmp, "vis_dict.parquet")), pd.read_parquet(os.path.join(tmp, "vis_device.parquet"))
out["visibility_files_identical_rows"] = bool(a.equals(b))
a, b = pd.read_parquet(os.path.join(tmp, "cam_dict.parquet")), pd.read_parquet(os.path.join(tmp, "cam_device.parquet"))
out["relations_files_identical_rows"] = bool(a.equals(b))

# ---- the reference's algorithm on the host (oracle port), sampled
pts = s.points.astype(np.float64)
H, W = s.image_hw
depth = sc.depth.cpu().numpy()
E = [og.aligned_pose(s.axis_align, s.c2w[f]) for f in range(F)]
nf = 8
t0 = time.perf_counter()
sets = [set() for _ in range(N)]
masks = []
for f in range(nf):                                   # make_visibility_info.process_scene, MVI:93-108
    _, _, m = og.frame_visibility(pts, s.K, E[f], depth[f], H, W)
    vis = np.where(m)[0]
    lst = vis.tolist()
    for idx in vis:
        sets[idx].add(ids[f])
    json.dumps(lst)
    masks.append(m)
per_frame = (time.perf_counter() - t0) / nf
out["visibility_cpu_port_s_per_frame"] = per_frame
out["visibility_cpu_port_s_extrapolated"] = per_frame * F
npairs = 200
t0 = time.perf_counter()
for k in range(npairs):                               # calculate_frames_relations.process_scene pair loop, CFR:176-189
    og.pair_relations(E[k % nf], E[(k + 1) % nf], masks[k % nf], masks[(k + 1) % nf])
per_pair = (time.perf_counter() - t0) / npairs
out["relations_cpu_port_s_per_pair"] = per_pair
out["relations_cpu_port_s_extrapolated"] = per_pair * out["pairs"] + per_frame * F
out["frames_per_s"] = {"device": F / out["visibility_device_s"], "dict": F / out["visibility_dict_s"],
                       "cpu_port_1proc": 1.0 / per_frame, "reference_comment_25procs": 42.0}
print(json.dumps(out))
