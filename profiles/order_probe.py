"""K1 time for the same 500 k points in spatially coherent vs random order (config-5 shape)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multi_spatialmllm_b200 import synth
from multi_spatialmllm_b200.engine import DeviceScene

def morton_order(p, bits=10):
    q = ((p - p.min(0)) / (p.max(0) - p.min(0) + 1e-9) * (2**bits - 1)).astype(np.uint64)
    key = np.zeros(len(p), np.uint64)
    for b in range(bits):
        for a in range(3):
            key |= ((q[:, a] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + a)
    return np.argsort(key, kind="stable")

rng = np.random.default_rng(1234)
pts = rng.uniform([0, 0, 0], [8, 8, 3], (500_000, 3)).astype(np.float32)
s = synth.make_scene(seed=1234, n_points=1000, n_frames=1000)
for name, P in (("random order", pts), ("morton order", pts[morton_order(pts)])):
    sc = DeviceScene(P, s.K, s.c2w, axis_align=s.axis_align)
    sc.render_depth(1)
    for _ in range(2): sc.project_visibility()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): sc.project_visibility()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name}: {ms:.3f} ms per 1000 frames x 500k points = {1000*500_000/ms/1e6:.0f} G point-frames/s, visible {sc.mask_bool()[:8].mean():.3f}")
