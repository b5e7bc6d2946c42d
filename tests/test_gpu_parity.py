"""CUDA path (through the C ABI) against the reference goldens and the NumPy oracle.

Bars: bit-exact masks / indices / integer annotations, and bit-exact float64 arrays against the committed reference
fixtures (the device mirrors NumPy's operation orders).  Comparisons against the oracle evaluated on the
machine running the tests keep a 1e-9 relative tolerance because that machine's BLAS build may order the 4-term products
differently in the last ulp (the judged tolerance is 1e-5 relative, BASELINE.json).
"""
import numpy as np
import pytest

from oracle import geometry as og

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def _scene_from_golden(g, **kw):
    from multi_spatialmllm_b200.engine import DeviceScene
    pts = kw.pop("points", g["points"])
    return DeviceScene(pts, g["K"], g["c2w"], axis_align=g["axis_align"], depth=g["depth"],
                       instance_ids=g["instance_ids"], num_objects=int(g["num_objects"]),
                       image_hw=tuple(g["image_hw"]), depth_hw=tuple(g["depth_hw"]), **kw)


def _valid_rows(g):
    ids = list(g["image_ids"])
    return [ids.index(v) for v in g["valid_ids"]]


@pytest.mark.parametrize("layout", ["f32x3", "f32x4", "f64x3"])
def test_golden_projection_and_masks(golden_scannet, layout):
    g = golden_scannet
    pts = g["points"]
    if layout == "f32x4":
        pts = np.concatenate([pts, np.ones((len(pts), 1), np.float32)], 1)
    elif layout == "f64x3":
        pts = pts.astype(np.float64)
    sc = _scene_from_golden(g, points=pts)
    mask, xy, z = sc.project_visibility(materialize=True)
    rows = _valid_rows(g)
    mb = sc.mask_bool()
    np.testing.assert_array_equal(sc.E.cpu().numpy().reshape(-1, 4, 4)[rows], g["ref_E_aligned"])
    np.testing.assert_array_equal(mb[rows], g["ref_mask"])
    xy, z = xy.cpu().numpy(), z.cpu().numpy()
    # bit-identical floats: the device reproduces NumPy's operation order for E = A @ c2w, inv(E) (LAPACK gesv order),
    # both 4-term matrix products (k-sequential FMA chains) and the IEEE division — no tolerance needed on the fixture
    np.testing.assert_array_equal(xy[rows], g["ref_xy"])
    np.testing.assert_array_equal(z[rows], g["ref_z"])
    invalid = sorted(set(range(len(g["image_ids"]))) - set(rows))
    assert invalid and not mb[invalid].any()                       # -inf poses see nothing
    np.testing.assert_array_equal(sc.valid.cpu().numpy()[invalid], 0)
    # pad bits beyond n are zero
    w = sc.mask.cpu().numpy().view(np.uint32)
    assert int(np.unpackbits(w.view(np.uint8), axis=1, bitorder="little")[:, sc.n:].sum()) == 0


def test_golden_pair_relations_and_tables(golden_scannet):
    from multi_spatialmllm_b200 import engine
    g = golden_scannet
    sc = _scene_from_golden(g)
    ids = [str(s) for s in g["image_ids"]]
    rel = engine.process_scene_relations(sc, ids)
    assert len(rel) == len(g["ref_pair_ids"])
    for (a, b), vals in zip(g["ref_pair_ids"], g["ref_pair_vals"]):
        r = rel[(str(a), str(b))]
        assert r["overlap"] == vals[0] or (np.isnan(r["overlap"]) and np.isnan(vals[0]))   # same integers -> same double
        np.testing.assert_allclose([r["distance"], r["yaw"], r["pitch"]], vals[1:], rtol=1e-12, atol=1e-12)
    tab = engine.process_scene_visibility(sc, ids)
    off = g["ref_i2p_off"]
    assert list(tab["image_to_points"].keys()) == [str(v) for v in g["valid_ids"]]
    for k, v in enumerate(g["valid_ids"]):
        assert tab["image_to_points"][str(v)] == g["ref_i2p_flat"][off[k]:off[k + 1]].tolist()
    assert [len(tab["point_to_images"][p]) for p in range(sc.n)] == g["ref_p2i_cnt"].tolist()


def test_golden_point_annotation(golden_scannet):
    g = golden_scannet
    sc = _scene_from_golden(g)
    rows = _valid_rows(g)
    spot = g["ref_spot"]
    fr = [rows[int(k)] for k in spot[:, 0]]
    a = sc.annotate_points(spot[:, 1].astype(np.int32), fr)
    # per-point calls: bit-identical floats (single-column matmul order + LAPACK-order inverse)
    np.testing.assert_array_equal(a["uvz"].cpu().numpy(), spot[:, 2:5])
    np.testing.assert_array_equal(a["permille"].cpu().numpy(), spot[:, 5:7].astype(np.int32))
    np.testing.assert_array_equal(a["depth_mm"].cpu().numpy(), spot[:, 7].astype(np.int32))
    np.testing.assert_array_equal(a["pix"].cpu().numpy(), spot[:, 8:10].astype(np.int32))
    assert a["visible"].cpu().numpy().all()


def test_kth_common_point_selection(golden_scannet):
    g = golden_scannet
    sc = _scene_from_golden(g)
    sc.project_visibility()
    rows = _valid_rows(g)
    ref = g["ref_mask"]
    pairs, sel, expect = [], [], []
    for i in range(len(rows)):
        for j in range(len(rows)):
            if i == j:
                continue
            common = og.common_visible_points(np.flatnonzero(ref[i]), np.flatnonzero(ref[j]))
            ks = sorted({0, len(common) // 2, max(len(common) - 1, 0), len(common), len(common) + 7})
            for k in ks:
                pairs.append((rows[i], rows[j])); sel.append(k)
                expect.append(int(common[k]) if k < len(common) else -1)
    rec = sc.records_to_numpy(sc.pair_annotate(np.array(pairs, np.int32), np.array(sel, np.uint32), sel_mode=0))
    np.testing.assert_array_equal(rec["point_id"], expect)
    got = rec[rec["point_id"] >= 0]
    assert (got["flags"] == 3).all()            # table <-> re-check invariant (DE-dot:203-213): zero warnings
    # sel_mode 1: k = floor(r * inter / 2^32)
    r = np.random.default_rng(0).integers(0, 2**32, len(pairs), dtype=np.uint64).astype(np.uint32)
    rec1 = sc.records_to_numpy(sc.pair_annotate(np.array(pairs, np.int32), r, sel_mode=1))
    for (a, b), rv, rr in zip(pairs, r, rec1):
        common = np.flatnonzero(ref[rows.index(a)] & ref[rows.index(b)])
        k = (int(rv) * len(common)) >> 32
        assert rr["point_id"] == (int(common[k]) if len(common) else -1)
        assert rr["inter"] == len(common)


def test_pair_record_integers_match_oracle(golden_scannet):
    g = golden_scannet
    H, W = g["image_hw"]
    sc = _scene_from_golden(g)
    sc.project_visibility()
    rows = _valid_rows(g)
    pairs = np.array([(rows[i], rows[j]) for i in range(len(rows)) for j in range(i + 1, len(rows))], np.int32)
    rec = sc.records_to_numpy(sc.pair_annotate(pairs, np.full(len(pairs), 3, np.uint32), sel_mode=0))
    E = g["ref_E_aligned"]
    pts = g["points"].astype(np.float64)
    for (a, b), r in zip(pairs, rec):
        i, j = rows.index(a), rows.index(b)
        d, _ = og.relative_displacement(E[i], E[j])
        np.testing.assert_allclose(r["rel_disp"], d, rtol=1e-12, atol=1e-13)
        assert r["inter"] == int((g["ref_mask"][i] & g["ref_mask"][j]).sum())
        assert r["uni"] == int((g["ref_mask"][i] | g["ref_mask"][j]).sum())
        if r["point_id"] < 0:
            continue
        for side, (k, f) in enumerate(((i, a), (j, b))):
            an = og.annotate_point(pts[r["point_id"]], g["K"], E[k], g["depth"][f], H, W)
            assert an is not None
            assert list(r["permille"][2 * side:2 * side + 2]) == [an["x_permille"], an["y_permille"]]
            assert r["depth_mm"][side] == an["depth_mm"]
            assert tuple(r["pix"][2 * side:2 * side + 2]) == an["pix"]


def test_object_counts(golden_scannet):
    g = golden_scannet
    sc = _scene_from_golden(g)
    sc.project_visibility()
    counts = sc.object_counts().cpu().numpy()
    mb = sc.mask_bool()
    for f in range(sc.F):
        np.testing.assert_array_equal(counts[f], og.object_visible_counts(mb[f], g["instance_ids"], int(g["num_objects"])))
    mask_before = sc.mask.cpu().numpy().copy()
    _, fused = sc.project_visibility_objects()            # K1 with the counting fused in
    np.testing.assert_array_equal(fused.cpu().numpy(), counts)
    np.testing.assert_array_equal(sc.mask.cpu().numpy(), mask_before)


def test_fused_object_counts_mixed_groups_and_fallback():
    """Instance ids that change inside 32-point groups, ids above n_objects, more than 63 / 255 objects (fused for any count)."""
    from multi_spatialmllm_b200 import synth
    from multi_spatialmllm_b200.engine import DeviceScene
    s = synth.make_scene(seed=21, n_points=70_003, n_frames=67, n_objects=9, shuffle_points=False)
    rng = np.random.default_rng(5)
    # ScanNet-like: 255 / 2000 instances as runs of consecutive vertices (uniform 64-point super-groups, tiles with fewer and
    # with more than 15 distinct ids: the per-tile id table of msp_cull9.cuh and its global-atomic overflow path)
    runs255 = np.repeat(rng.integers(0, 256, 2000), rng.integers(5, 130, 2000))[:s.n_points].astype(np.uint16)
    runs2000 = np.repeat(rng.integers(0, 2001, 9000), rng.integers(1, 20, 9000))
    runs2000 = np.resize(runs2000, s.n_points).astype(np.uint16)
    # few objects in short runs: every 64-point super-group carries 3-5 ids that all fit the tile's 15-entry table (the
    # per-lane `match_any` path of msp_cull9.cuh without its global-atomic overflow)
    runs20 = np.resize(np.repeat(np.arange(9), 20), s.n_points).astype(np.uint16)
    runs27 = np.resize(np.repeat(rng.integers(0, 9, 5000), 27), s.n_points).astype(np.uint16)
    for n_obj, ids in ((9, s.instance_ids.copy()), (8, runs20), (8, runs27), (40, rng.integers(0, 45, s.n_points).astype(np.uint16)),
                       (100, rng.integers(0, 101, s.n_points).astype(np.uint16)), (255, runs255), (2000, runs2000)):
        ids[::977] = 60000                                  # out-of-range ids are ignored (treated as "no object")
        sc = DeviceScene(s.points, s.K, s.c2w, axis_align=s.axis_align, instance_ids=ids, num_objects=n_obj)
        sc.render_depth(1)
        _, fused = sc.project_visibility_objects()
        sep = sc.object_counts().cpu().numpy()
        np.testing.assert_array_equal(fused.cpu().numpy(), sep)
        mb = sc.mask_bool()
        good = np.where(ids <= n_obj, ids, 0)
        for f in (0, 33, 66):
            np.testing.assert_array_equal(sep[f], og.object_visible_counts(mb[f], good, n_obj))
        assert sep.sum() > 0


def test_zbuffer_matches_oracle(golden_scannet):
    g = golden_scannet
    H, W = g["image_hw"]
    dh, dw = g["depth_hw"]
    sc = _scene_from_golden(g)
    zb = sc.render_zbuffer().cpu().numpy()
    d16 = sc.render_depth(margin_mm=2).cpu().numpy()
    pts = g["points"].astype(np.float64)
    for k, f in enumerate(_valid_rows(g)):
        ref = og.zbuffer_render(pts, g["K"], g["ref_E_aligned"][k], H, W, dh, dw)
        np.testing.assert_array_equal(zb[f], ref)
        np.testing.assert_array_equal(d16[f], og.depth_to_u16(ref, 2))
    # a surface is visible against its own render: every point that won a pixel passes the strict test
    sc.project_visibility()
    assert sc.mask_bool().sum() > 0


@pytest.mark.parametrize("n,frames,shuffle", [(1, 3, False), (31, 2, False), (1025, 5, True), (50_000, 6, True)])
def test_random_scenes_vs_oracle(n, frames, shuffle):
    from multi_spatialmllm_b200 import synth
    from multi_spatialmllm_b200.engine import DeviceScene
    s = synth.make_scene(seed=100 + n, n_points=max(n, 64), n_frames=frames, n_objects=4, shuffle_points=shuffle)
    pts = s.points[:n]
    sc = DeviceScene(pts, s.K, s.c2w, axis_align=s.axis_align, image_hw=s.image_hw, depth_hw=s.depth_hw)
    depth = sc.render_depth(margin_mm=1).cpu().numpy()
    # perturb: holes + noise so that the depth test is not trivially true
    rng = np.random.default_rng(n)
    noisy = depth.astype(np.int64) + rng.integers(-2, 3, depth.shape) * (depth > 0)
    noisy[rng.random(depth.shape) < 0.05] = 0
    noisy = np.clip(noisy, 0, 65535).astype(np.uint16)
    sc.set_depth(noisy)
    _, xy, z = sc.project_visibility(materialize=True)
    mb = sc.mask_bool()
    H, W = s.image_hw
    for f in range(frames):
        E = og.aligned_pose(s.axis_align, s.c2w[f])
        oxy, oz, om = og.frame_visibility(pts.astype(np.float64), s.K, E, noisy[f], H, W)
        np.testing.assert_array_equal(mb[f], om)
        np.testing.assert_allclose(xy[f].cpu().numpy(), oxy, rtol=RTOL, atol=1e-6)
        np.testing.assert_allclose(z[f].cpu().numpy(), oz, rtol=RTOL, atol=1e-9)


def test_generic_intrinsics_path():
    from multi_spatialmllm_b200 import synth
    from multi_spatialmllm_b200.engine import DeviceScene
    s = synth.make_scene(seed=5, n_points=3000, n_frames=3, n_objects=2)
    K = s.K.copy()
    K[0, 1] = 0.7          # skew: not the [fx 0 cx 0] form -> generic 4-term path
    sc = DeviceScene(s.points, K, s.c2w, axis_align=s.axis_align)
    sc.render_depth(margin_mm=1)
    _, xy, z = sc.project_visibility(materialize=True)
    mb, depth = sc.mask_bool(), sc.depth.cpu().numpy()
    for f in range(3):
        E = og.aligned_pose(s.axis_align, s.c2w[f])
        oxy, oz, om = og.frame_visibility(s.points.astype(np.float64), K, E, depth[f], *s.image_hw)
        np.testing.assert_array_equal(mb[f], om)
        np.testing.assert_allclose(xy[f].cpu().numpy(), oxy, rtol=RTOL, atol=1e-6)


def test_full_size_properties():
    """BASELINE config-2 scene size: size-independent properties instead of a CPU oracle pass."""
    import torch
    from multi_spatialmllm_b200 import synth
    from multi_spatialmllm_b200.engine import DeviceScene
    s = synth.make_scene(seed=42, n_points=200_000, n_frames=48, n_objects=12)
    sc = DeviceScene(s.points, s.K, s.c2w, axis_align=s.axis_align, instance_ids=s.instance_ids, num_objects=12)
    sc.render_depth(margin_mm=1)
    sc.project_visibility()
    mb = sc.mask_bool()
    cnt = mb.sum(1)
    assert (cnt > 1000).all()
    F = sc.F
    ij = np.array([(i, j) for i in range(F) for j in range(F)], np.int32)
    inter, uni, ov = [t.cpu().numpy() for t in sc.pair_overlap(ij)]
    inter, uni, ov = inter.reshape(F, F), uni.reshape(F, F), ov.reshape(F, F)
    np.testing.assert_array_equal(inter, inter.T)                      # symmetry
    np.testing.assert_array_equal(uni, uni.T)
    np.testing.assert_array_equal(np.diag(inter), cnt)                 # |A&A| = |A|
    np.testing.assert_array_equal(inter + uni, cnt[:, None] + cnt[None, :])   # inclusion-exclusion
    np.testing.assert_array_equal(np.diag(ov), 100.0)
    # three spot frames against the oracle at full size
    depth = sc.depth.cpu().numpy()
    for f in (0, 17, 47):
        E = og.aligned_pose(s.axis_align, s.c2w[f])
        _, _, om = og.frame_visibility(s.points.astype(np.float64), s.K, E, depth[f], *s.image_hw)
        np.testing.assert_array_equal(mb[f], om)
    # object counts sum to the number of visible object points
    oc = sc.object_counts().cpu().numpy()
    np.testing.assert_array_equal(oc.sum(1), (mb & (s.instance_ids > 0)[None]).sum(1))
    np.testing.assert_array_equal(sc.project_visibility_objects()[1].cpu().numpy(), oc)
    # every selected common point passes the re-check in both frames
    pairs = ij[(ij[:, 0] < ij[:, 1])]
    rec = sc.records_to_numpy(sc.pair_annotate(pairs, torch.randint(0, 2**31, (len(pairs),), dtype=torch.int32), sel_mode=1))
    has = rec["point_id"] >= 0
    assert has.sum() > 0 and (rec["flags"][has] == 3).all()
    pid = rec["point_id"][has]
    assert mb[pairs[has, 0], pid].all() and mb[pairs[has, 1], pid].all()


def test_tapvid_tracks(golden_tapvid):
    import ctypes as C
    import torch
    from multi_spatialmllm_b200._lib import check, lib
    g = golden_tapvid
    T, N, _ = g["tracks_xyz"].shape
    H, W = [int(v) for v in g["image_hw"]]
    dev = "cuda:0"
    tr = torch.from_numpy(g["tracks_xyz"]).to(dev)
    w2c = torch.from_numpy(g["w2c"].reshape(T, 16)).to(dev)
    intr = torch.from_numpy(g["fxfycxcy"]).to(dev)
    world = torch.empty((T, N, 3), dtype=torch.float64, device=dev)
    uvn = torch.empty((T, N, 2), dtype=torch.float64, device=dev)
    pm = torch.empty((T, N, 2), dtype=torch.int32, device=dev)
    ok = torch.empty((T, N), dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    check(lib.msp_track_project(tr.data_ptr(), w2c.data_ptr(), intr.data_ptr(), T, N, H, W, world.data_ptr(),
                                uvn.data_ptr(), pm.data_ptr(), ok.data_ptr(), s))
    np.testing.assert_array_equal(world.cpu().numpy(), g["ref_world"])      # einsum order + LAPACK-order inverse: exact
    okh = ok.cpu().numpy().astype(bool)
    np.testing.assert_array_equal(okh, np.isfinite(g["ref_proj"][..., 0]))
    np.testing.assert_allclose(uvn.cpu().numpy()[okh], g["ref_proj"][okh], rtol=1e-13)
    np.testing.assert_array_equal(pm.cpu().numpy()[okh], np.round(g["ref_proj"][okh] * 1000).astype(np.int32))
    # pair displacement vs oracle
    rng = np.random.default_rng(1)
    tri = np.stack([rng.integers(0, T, 64), rng.integers(0, T, 64), rng.integers(0, N, 64)], 1).astype(np.int32)
    trid = torch.from_numpy(tri).to(dev)
    out = torch.empty((64, 5), dtype=torch.float64, device=dev)
    check(lib.msp_track_pair_displacement(world.data_ptr(), w2c.data_ptr(), T, N, trid.data_ptr(), 64, out.data_ptr(), s))
    out = out.cpu().numpy()
    wr = g["ref_world"]
    for (f1, f2, p), o in zip(tri, out):
        d = wr[f2, p] - wr[f1, p]
        # same operation orders as NumPy (1-D norm = sqrt of an FMA chain, (4,4)@(4,) pairwise): exact on SkylakeX OpenBLAS
        c1, c2 = np.linalg.inv(g["w2c"][f1]), np.linalg.inv(g["w2c"][f2])
        want = [np.linalg.norm(d), *(g["w2c"][f1] @ np.concatenate([d, [0]]))[:3], np.linalg.norm(c2[:3, 3] - c1[:3, 3])]
        np.testing.assert_allclose(o, want, rtol=1e-12, atol=1e-14)
        exact_hits = int(np.sum(o == np.array(want)))
        assert exact_hits >= 4


def test_bad_arguments_fail_loudly(golden_scannet):
    import torch
    from multi_spatialmllm_b200._lib import MspError
    sc = _scene_from_golden(golden_scannet)
    sc.project_visibility()
    sc.pts = torch.cat([torch.zeros(1, device=sc.device), sc.pts.flatten()])[1:].view(-1, 3)   # 4-byte offset: misaligned
    with pytest.raises(MspError):
        sc.project_visibility()


def _boundary_points(s, frames, per_frame, rng):
    """Points back-projected from the image border / the z=0 plane of given frames (float32-rounded):
    they sit inside or just outside the float32 cull's guard band."""
    out = []
    H, W = s.image_hw
    fx, fy, cx, cy = s.K[0, 0], s.K[1, 1], s.K[0, 2], s.K[1, 2]
    for f in frames:
        E = og.aligned_pose(s.axis_align, s.c2w[f])
        n = per_frame
        edge = rng.integers(0, 5, n)
        u = rng.uniform(0, W, n)
        v = rng.uniform(0, H, n)
        z = rng.uniform(0.3, 6.0, n)
        eps = rng.choice([0.0, 1e-9, -1e-9, 1e-6, -1e-6, 1e-4, -1e-4, 1e-2, -1e-2], n)
        u = np.where(edge == 0, 0.0 + eps, np.where(edge == 1, W + eps, u))
        v = np.where(edge == 2, 0.0 + eps, np.where(edge == 3, H + eps, v))
        z = np.where(edge == 4, eps, z)                     # on the camera plane
        cam = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z, np.ones(n)], 1)
        out.append((E @ cam.T).T[:, :3])
    return np.concatenate(out).astype(np.float32)


@pytest.mark.parametrize("layout", ["f32x3", "f32x4", "f64x3"])
def test_cull_path_equals_reference_order_path(layout, monkeypatch):
    """The float32-cull kernel must give the same bits as the all-float64 kernel, also on points that sit on
    the frustum boundary, with shuffled point order, ragged N and a frame count that crosses CTA chunks."""
    from multi_spatialmllm_b200 import synth
    from multi_spatialmllm_b200.engine import DeviceScene
    rng = np.random.default_rng(77)
    s = synth.make_scene(seed=9, n_points=120_001, n_frames=70, n_objects=6, shuffle_points=True, invalid_pose_every=33)
    valid = [f for f in range(70) if og.pose_is_valid(s.c2w[f])]
    extra = _boundary_points(s, valid[:20], 400, rng)
    pts = np.concatenate([s.points, extra, np.array([[np.nan, 0, 0], [np.inf, 1, 1], [1e30, 1e30, 1e30]], np.float32)])
    pts = pts[rng.permutation(len(pts))]
    if layout == "f32x4":
        dev_pts = np.concatenate([pts, np.ones((len(pts), 1), np.float32)], 1)
    elif layout == "f64x3":
        dev_pts = pts.astype(np.float64) + rng.normal(0, 1e-9, pts.shape)      # not float32-representable
    else:
        dev_pts = pts
    sc = DeviceScene(dev_pts, s.K, s.c2w, axis_align=s.axis_align)
    monkeypatch.setenv("MSP_DISABLE_CULL", "1")
    depth_ref = sc.render_depth(margin_mm=1).cpu().numpy()
    sc.project_visibility()
    mask_ref = sc.mask.cpu().numpy().copy()
    monkeypatch.setenv("MSP_DISABLE_CULL", "0")
    sc.mask.zero_()
    depth_cull = sc.render_depth(margin_mm=1).cpu().numpy()
    sc.project_visibility()
    np.testing.assert_array_equal(depth_cull, depth_ref)
    np.testing.assert_array_equal(sc.mask.cpu().numpy(), mask_ref)
    assert np.unpackbits(mask_ref.view(np.uint8)).sum() > 100_000
    # and both equal the oracle on a few frames
    mb = sc.mask_bool()
    p64 = np.asarray(dev_pts[:, :3], dtype=np.float64)
    for f in (valid[0], valid[7], valid[-1]):
        E = og.aligned_pose(s.axis_align, s.c2w[f])
        _, _, om = og.frame_visibility(p64, s.K, E, depth_ref[f], *s.image_hw)
        np.testing.assert_array_equal(mb[f], om)


def test_pose_prologue_is_bit_identical_to_numpy_fixture():
    """k_pose_prepare: aligned pose = axis_align @ c2w and its inverse, against np.linalg.inv values committed from the
    build container (tests/golden/inverse_small.npz) — exact equality, no tolerance."""
    import os
    import torch
    from multi_spatialmllm_b200._lib import check, lib
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "inverse_small.npz"))
    dev = "cuda:0"

    def prep(mats, A):
        F = len(mats)
        c2w = torch.from_numpy(np.ascontiguousarray(mats.reshape(F, 16))).to(dev)
        Ad = None if A is None else torch.from_numpy(np.ascontiguousarray(A.reshape(16))).to(dev)
        E = torch.empty((F, 16), dtype=torch.float64, device=dev)
        w = torch.empty((F, 16), dtype=torch.float64, device=dev)
        yp = torch.empty((F, 2), dtype=torch.float64, device=dev)
        v = torch.empty((F,), dtype=torch.int32, device=dev)
        check(lib.msp_pose_prepare(c2w.data_ptr(), None if Ad is None else Ad.data_ptr(), F, E.data_ptr(), w.data_ptr(),
                                   yp.data_ptr(), v.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return E.cpu().numpy().reshape(F, 4, 4), w.cpu().numpy().reshape(F, 4, 4), v.cpu().numpy()

    _, w, v = prep(g["mats"], None)
    assert v.all()
    np.testing.assert_array_equal(w, g["inv"])
    n = len(g["aligned"])
    E, w, _ = prep(g["mats"][:n], g["axis_align"])
    np.testing.assert_array_equal(E, g["aligned"])
    np.testing.assert_array_equal(w, g["aligned_inv"])


def test_morton_reorder_is_transparent():
    """reorder="morton": same masks (after un-permutation), counts, extents, annotations; pair records refer to the
    caller's point numbering."""
    import torch
    from multi_spatialmllm_b200 import synth
    from multi_spatialmllm_b200.engine import DeviceScene
    s = synth.make_scene(seed=31, n_points=60_001, n_frames=20, n_objects=7, shuffle_points=True)
    kw = dict(axis_align=s.axis_align, instance_ids=s.instance_ids, num_objects=7)
    a = DeviceScene(s.points, s.K, s.c2w, **kw)
    b = DeviceScene(s.points, s.K, s.c2w, reorder="morton", **kw)
    depth = a.render_depth(1)
    b.set_depth(depth)
    np.testing.assert_array_equal(b.render_zbuffer().cpu().numpy(), a.render_zbuffer().cpu().numpy())
    _, ca = a.project_visibility_objects()
    _, cb = b.project_visibility_objects()
    ma, mb = a.mask_bool(), b.mask_bool()
    np.testing.assert_array_equal(mb, ma)
    np.testing.assert_array_equal(cb.cpu().numpy(), ca.cpu().numpy())
    np.testing.assert_array_equal(b.object_extents().cpu().numpy(), a.object_extents().cpu().numpy())
    vis = np.argwhere(ma)[::997]
    ra = a.annotate_points(vis[:, 1], vis[:, 0])
    rb = b.annotate_points(vis[:, 1], vis[:, 0])
    for k in ra:
        np.testing.assert_array_equal(rb[k].cpu().numpy(), ra[k].cpu().numpy())
    pairs = np.array([(i, j) for i in range(20) for j in range(i + 1, 20)], np.int32)
    sel = torch.randint(-2**31, 2**31 - 1, (len(pairs),), dtype=torch.int32)
    rec = b.records_to_numpy(b.pair_annotate(pairs, sel, 1))
    ref = a.records_to_numpy(a.pair_annotate(pairs, sel, 1))
    np.testing.assert_array_equal(rec["inter"], ref["inter"])
    np.testing.assert_array_equal(rec["uni"], ref["uni"])
    has = rec["point_id"] >= 0
    np.testing.assert_array_equal(has, ref["point_id"] >= 0)
    assert ma[pairs[has, 0], rec["point_id"][has]].all() and ma[pairs[has, 1], rec["point_id"][has]].all()
    assert (rec["flags"][has] == 3).all()
    with pytest.raises(ValueError):
        b.pair_annotate(pairs, np.zeros(len(pairs), np.uint32), 0)
