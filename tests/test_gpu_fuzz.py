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


@pytest.mark.parametrize("n", [1, 31, 32, 33, 127, 128, 129, 4097, 200_001])
def test_pair_stage_on_random_masks(n):
    """K3 on arbitrary bit patterns written straight into the mask rows: counts, k-th common bit (word / quad / warp-segment
    boundaries), both selection modes, empty and full rows."""
    import torch
    from multi_spatialmllm_b200 import synth
    from multi_spatialmllm_b200.engine import DeviceScene
    rng = np.random.default_rng(n)
    F = 9
    s = synth.make_scene(seed=1, n_points=max(n, 64), n_frames=F)
    sc = DeviceScene(s.points[:n], s.K, s.c2w, axis_align=s.axis_align)
    sc.set_depth(np.full((F, 480, 640), 65535, np.uint16))
    dens = [0.0, 1.0, 0.5, 0.03, 0.9, 0.5, 0.001, 0.2, 0.7]
    bits = np.stack([rng.random(n) < d for d in dens])
    words = np.zeros((F, sc.stride * 32), np.uint8)
    words[:, :n] = bits
    packed = np.packbits(words.reshape(F, sc.stride, 32), axis=2, bitorder="little").view("<u4").reshape(F, sc.stride)
    sc.mask = torch.from_numpy(packed.view(np.int32).copy()).to(sc.device)
    pairs = np.array([(a, b) for a in range(F) for b in range(F)], np.int32)
    inter, uni, ov = (t.cpu().numpy() for t in sc.pair_overlap(pairs))
    for (a, b), i_, u_, o_ in zip(pairs, inter, uni, ov):
        ci, cu = int((bits[a] & bits[b]).sum()), int((bits[a] | bits[b]).sum())
        assert (i_, u_) == (ci, cu)
        assert (np.isnan(o_) and cu == 0) or o_ == ci / cu * 100
    ks, expect, sel_pairs = [], [], []
    for a, b in pairs:
        common = np.flatnonzero(bits[a] & bits[b])
        cand = {0, 1, len(common) // 2, len(common) - 1, len(common), 31, 32, 127, 128} | set(rng.integers(0, max(len(common), 1), 3).tolist())
        for k in sorted(c for c in cand if c >= 0):
            sel_pairs.append((a, b)); ks.append(k); expect.append(int(common[k]) if k < len(common) else -1)
    rec = sc.records_to_numpy(sc.pair_annotate(np.array(sel_pairs, np.int32), np.array(ks, np.uint32), 0))
    np.testing.assert_array_equal(rec["point_id"], expect)
    r32 = rng.integers(0, 2**32, len(sel_pairs), dtype=np.uint64)
    rec1 = sc.records_to_numpy(sc.pair_annotate(np.array(sel_pairs, np.int32), r32.astype(np.uint32), 1))
    for (a, b), rv, r in zip(sel_pairs, r32, rec1):
        common = np.flatnonzero(bits[a] & bits[b])
        assert r["point_id"] == (int(common[(int(rv) * len(common)) >> 32]) if len(common) else -1)


def test_annotation_integers_on_rounding_boundaries():
    """Points constructed so that per-mille / mm values sit within a few float64 ulps of .5 boundaries: the kernel's
    round-half-even / truncation must follow the float64 reference arithmetic exactly."""
    from multi_spatialmllm_b200 import synth
    from multi_spatialmllm_b200.engine import DeviceScene
    from oracle import geometry as og
    s = synth.make_scene(seed=3, n_points=64, n_frames=4)
    H, W = s.image_hw
    rng = np.random.default_rng(0)
    pts = []
    E = [og.aligned_pose(s.axis_align, s.c2w[f]) for f in range(4)]
    fx, fy, cx, cy = s.K[0, 0], s.K[1, 1], s.K[0, 2], s.K[1, 2]
    for f in range(4):
        for _ in range(300):
            xm = rng.integers(1, 999) + 0.5                        # per-mille boundary
            u = np.nextafter(xm * W / 1000.0, rng.choice([-np.inf, np.inf])) if rng.random() < 0.7 else xm * W / 1000.0
            v = (rng.integers(1, 999) + 0.5) * H / 1000.0
            z = (rng.integers(400, 6000) + 0.5) / 1000.0             # mm boundary
            cam = np.array([(u - cx) / fx * z, (v - cy) / fy * z, z, 1.0])
            pts.append((E[f] @ cam)[:3])
    pts = np.array(pts)                                              # float64 points
    sc = DeviceScene(pts, s.K, s.c2w, axis_align=s.axis_align)
    sc.set_depth(np.full((4, 480, 640), 65535, np.uint16))
    frames = np.repeat(np.arange(4), 300)
    a = sc.annotate_points(np.arange(len(pts)), frames)
    pm, mm, pix, vis = (a[k].cpu().numpy() for k in ("permille", "depth_mm", "pix", "visible"))
    # does this machine's NumPy use the same inverse order as the committed fixture (and hence as the device)?
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "inverse_small.npz"))
    numpy_matches_fixture = np.array_equal(np.linalg.inv(g["mats"]), g["inv"])
    n_ok = n_off = 0
    for i, f in enumerate(frames):
        r = og.annotate_point(pts[i], s.K, E[f], np.full((480, 640), 65535, np.uint16), H, W)
        if r is None:
            assert not vis[i]
            continue
        for got, want, val in ((pm[i][0], r["x_permille"], r["u"] / W * 1000), (pm[i][1], r["y_permille"], r["v"] / H * 1000),
                               (mm[i], r["depth_mm"], r["z"] * 1000)):
            if got != want:
                # only legitimate when the host BLAS differs from the device order in the last ulp of inv(E) AND the value
                # sits on the boundary to within that ulp
                assert abs(got - want) == 1 and abs(abs(val - np.floor(val)) - 0.5) < 1e-9
                n_off += 1
            else:
                n_ok += 1
    assert n_ok > 2000
    if numpy_matches_fixture:
        assert n_off == 0, f"{n_off} boundary values differ although the host inverse order equals the device order"


@pytest.mark.gpu
def test_depth_to_u16_guard_band_equals_float64_definition():
    """msp_depth_to_u16's float32 fast path + float64 guard band against the all-float64 definition (oracle.depth_to_u16):
    z values sitting exactly on, and one ulp either side of, millimetre boundaries, far depths that saturate, +inf pixels,
    and an element count that is not a multiple of 8."""
    import torch
    from multi_spatialmllm_b200._lib import check, lib
    from oracle import geometry as og
    rng = np.random.default_rng(11)
    k = rng.integers(1, 70000, 60000)
    on = (k * 0.001).astype(np.float32)
    z = np.concatenate([on, np.nextafter(on, np.float32(np.inf)), np.nextafter(on, np.float32(0)),
                        (k.astype(np.float64) * 0.001).astype(np.float32) + np.float32(1e-7),
                        rng.uniform(0.0, 70.0, 100003).astype(np.float32),
                        np.array([np.inf, 0.0, 1e-30, 65.534, 65.535, 65.5351, 1e6, 3e38], np.float32)])
    z[rng.integers(0, len(z), 5000)] = np.inf
    zt = torch.from_numpy(z).cuda()
    for margin in (0, 1, 2, 7):
        out = torch.empty(len(z), dtype=torch.uint16, device="cuda")
        check(lib.msp_depth_to_u16(zt.data_ptr(), len(z), margin, 0.001, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
        np.testing.assert_array_equal(out.cpu().numpy(), og.depth_to_u16(z, margin))


@pytest.mark.gpu
def test_cull9_nonstandard_frames_and_unaligned_depth(monkeypatch):
    """Corners of the round-2 K1 (msp_cull9.cuh): a pose whose inverse has a non-[0,0,0,1] bottom row (every lane goes through the
    deferred exact path), depth maps that start on an odd 2-byte boundary and depth maps with an odd pixel count (both take the
    round-1 kernel: the new one gathers aligned 32-bit words) — all against the all-float64 reference-order kernel."""
    import torch
    from multi_spatialmllm_b200 import synth
    from multi_spatialmllm_b200.engine import DeviceScene
    s = synth.make_scene(seed=31, n_points=37_777, n_frames=7, n_objects=5)
    c2w = s.c2w.copy()
    c2w[2, 3, 3] = 2.0                                   # homogeneous scale: w2c bottom row becomes [0,0,0,0.5]
    c2w[5, 3, :] = [0.0, 1e-3, 0.0, 1.0]                 # a perspective row
    for dhw in ((480, 640), (37, 51)):                   # 37 * 51 is odd: depth_stride is odd
        sc = DeviceScene(s.points, s.K, c2w, axis_align=s.axis_align, instance_ids=s.instance_ids, num_objects=s.num_objects,
                         depth_hw=dhw)
        depth = sc.render_depth(1).clone()
        monkeypatch.setenv("MSP_DISABLE_CULL", "1")
        sc.project_visibility()
        ref_mask = sc.mask.cpu().numpy().copy()
        ref_cnt = sc.object_counts().cpu().numpy()
        monkeypatch.setenv("MSP_DISABLE_CULL", "0")
        for shift in (0, 1):                             # shift 1: the maps start 2 bytes off a 4-byte boundary
            big = torch.zeros(depth.numel() + 8, dtype=torch.uint16, device="cuda")
            view = big[shift:shift + depth.numel()].view(depth.shape)
            view.copy_(depth)
            assert (view.data_ptr() % 4 == 0) == (shift == 0)
            sc.depth = view
            sc.mask.fill_(-1)
            _, cnt = sc.project_visibility_objects()
            np.testing.assert_array_equal(sc.mask.cpu().numpy(), ref_mask)
            np.testing.assert_array_equal(cnt.cpu().numpy(), ref_cnt)
        assert ref_mask.any()
