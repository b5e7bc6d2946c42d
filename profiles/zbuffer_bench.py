"""K2 (msp_zbuffer_render) time per frame on config-2- and stress-shaped scenes, CUDA events.
Usage (GPU box):  python profiles/zbuffer_bench.py   (build with `make EXTRA=-DMSP_ZBUF_WARP_MIN=0` for the plain-atomic variant)"""
import hashlib
import json
import os
import sys


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from multi_spatialmllm_b200 import engine, synth  # noqa: E402

out = {}
for name, n, shuffle, room in (("mesh_ordered_200k", 200_000, False, (8.0, 8.0, 3.0)), ("shuffled_500k", 500_000, True, (8.0, 8.0, 3.0)),
                               ("large_room_200k", 200_000, False, (30.0, 30.0, 4.0))):
    s = synth.make_scene(seed=3, n_points=n, n_frames=256, n_objects=12, shuffle_points=shuffle, room=room)
    sc = engine.DeviceScene(s.points, s.K, s.c2w, axis_align=s.axis_align, image_hw=s.image_hw, depth_hw=s.depth_hw)
    sc.render_zbuffer()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        zb = sc.render_zbuffer()
    e1.record()
    torch.cuda.synchronize()
    out[name] = {"ms_per_256_frames": e0.elapsed_time(e1) / 5,
                 "sha256": hashlib.sha256(zb.cpu().numpy().tobytes()).hexdigest()[:16]}
print(json.dumps(out))
