"""Differential fuzz of the float32-decision kernel (k_frame_cull) against the all-float64 reference-order kernel
(k_frame_loop, MSP_DISABLE_CULL=1): masks, fused object counts and z-buffers must be bit-identical over random
intrinsics, image / depth sizes, depth scales, pose distributions, coordinate magnitudes and point layouts."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _case(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.choice([777, 5000, 40_000, 120_000]))
    F = int(rng.choice([3, 40, 70, 130]))
    scale = float(rng.choice([1.0, 1.0, 30.0, 1000.0]))              # world size: room .. kilometre-scale coordinates
    offset = rng.choice([0.0, 0.0, 500.0, 20_000.0]) * rng.normal(size=3)
    kind = rng.integers(0, 3)
    if kind == 0:                                                    # surfaces: mesh-like order
        u, v = rng.random(n), rng.random(n)
        face = rng.integers(0, 3, n)
        pts = np.where(face[:, None] == 0, np.stack([u * 8, v * 8, 0 * u], 1),
                       np.where(face[:, None] == 1, np.stack([u * 8, 0 * u, v * 3], 1), np.stack([0 * u, u * 8, v * 3], 1)))
        pts = pts[np.lexsort((pts[:, 0], pts[:, 1], face))]
    elif kind == 1:                                                  # volume, random order
        pts = rng.uniform([0, 0, 0], [8, 8, 3], (n, 3))
    else:                                                            # clustered blobs
        c = rng.uniform([0, 0, 0], [8, 8, 3], (12, 3))
        pts = c[rng.integers(0, 12, n)] + rng.normal(0, 0.15, (n, 3))
    pts = pts * scale + offset
    W = int(rng.choice([1296, 640, 1920, 333]))
    H = int(rng.choice([968, 480, 1080, 251]))
    dw, dh = (int(v) for v in rng.choice([[640, 480], [W, H], [320, 200], [512, 424]]))
    f = float(rng.uniform(0.3, 2.5)) * W
    K = np.eye(4)
    K[0, 0], K[1, 1] = f, f * rng.uniform(0.9, 1.1)
    K[0, 2], K[1, 2] = W * rng.uniform(0.2, 0.8), H * rng.uniform(0.2, 0.8)
    lo, hi = pts.min(0), pts.max(0)
    c2w = np.tile(np.eye(4), (F, 1, 1))
    for i in range(F):
        mode = rng.integers(0, 4)
        if mode == 0:                                                # inside, random orientation
            pos = rng.uniform(lo, hi)
        elif mode == 1:                                              # far outside
            pos = (lo + hi) / 2 + rng.normal(size=3) * (hi - lo).max() * 3
        elif mode == 2:                                              # hugging a point (tiny depths)
            pos = pts[rng.integers(0, n)] + rng.normal(0, 1e-3 * scale, 3)
        else:
            pos = rng.uniform(lo - 0.2 * (hi - lo), hi + 0.2 * (hi - lo))
        c2w[i, :3, :3] = _random_rotation(rng)
        c2w[i, :3, 3] = pos
    if rng.random() < 0.5:
        c2w = np.round(c2w, 6)
    A = np.eye(4)
    if rng.random() < 0.5:
        A[:3, :3] = _random_rotation(rng)
        A[:3, 3] = rng.normal(size=3) * scale
    layout = rng.choice(["f32x3", "f32x4", "f64x3"])
    p32 = pts.astype(np.float32)
    dev_pts = {"f32x3": p32, "f32x4": np.concatenate([p32, np.ones((n, 1), np.float32)], 1), "f64x3": pts}[layout]
    depth_scale = float(rng.choice([0.001, 0.001, 0.00025, 0.01])) * (scale if scale > 1 and rng.random() < 0.7 else 1.0)
    inst = (rng.integers(0, 9, n) * (rng.random(n) < 0.7)).astype(np.uint16)
    if kind == 0:
        inst = np.sort(inst)
    return dict(pts=dev_pts, K=K, c2w=c2w, A=A, hw=(H, W), dhw=(dh, dw), depth_scale=depth_scale, inst=inst, rng=rng)


@pytest.mark.parametrize("seed", range(64))
def test_fast_path_equals_reference_order_kernel(seed, monkeypatch):
    import torch
    from multi_spatialmllm_b200.engine import DeviceScene
    c = _case(seed)
    sc = DeviceScene(c["pts"], c["K"], c["c2w"], axis_align=c["A"], instance_ids=c["inst"], num_objects=8,
                     image_hw=c["hw"], depth_hw=c["dhw"], depth_scale=c["depth_scale"])
    monkeypatch.setenv("MSP_DISABLE_CULL", "1")
    zb_ref = sc.render_zbuffer().cpu().numpy()
    monkeypatch.setenv("MSP_DISABLE_CULL", "0")
    zb_fast = sc.render_zbuffer().cpu().numpy()
    np.testing.assert_array_equal(zb_fast, zb_ref)
    rng = c["rng"]
    depth = sc.render_depth(margin_mm=int(rng.integers(0, 3))).cpu().numpy().astype(np.int64)
    style = rng.integers(0, 4)
    if style == 1:                                                   # sensor-like noise and holes
        depth = depth + rng.integers(-3, 4, depth.shape) * (depth > 0)
        depth[rng.random(depth.shape) < 0.1] = 0
    elif style == 2:                                                 # unrelated random depth
        depth = rng.integers(0, 65536, depth.shape)
    elif style == 3:                                                 # everything passes the depth test
        depth[:] = 65535
    sc.set_depth(np.clip(depth, 0, 65535).astype(np.uint16))
    monkeypatch.setenv("MSP_DISABLE_CULL", "1")
    sc.project_visibility()
    mask_ref = sc.mask.cpu().numpy().copy()
    cnt_ref = sc.object_counts().cpu().numpy()
    monkeypatch.setenv("MSP_DISABLE_CULL", "0")
    sc.mask.fill_(-1)
    _, cnt = sc.project_visibility_objects()
    np.testing.assert_array_equal(sc.mask.cpu().numpy(), mask_ref)
    np.testing.assert_array_equal(cnt.cpu().numpy(), cnt_ref)
    torch.cuda.synchronize()
