"""BASELINE.json configs as parity cases (configs[1] is the bench workload; the others are tested here).

  configs[0]  one scene0000_00-shaped scene (N = 81 369, 65 orbit frames -> 64 adjacent pairs), depth_perception
  configs[2]  TAPVid-3D object_movement: 3D track -> 2D projection + visibility across frames (scaled-down video count)
  configs[4]  stress shape: 500 k points/scene, random poses, z-buffer depth + correspondence (one scene, properties)
plus empty / ragged inputs through the C ABI.
"""
import numpy as np
import pytest

from oracle import geometry as og

pytestmark = pytest.mark.gpu


def test_config0_depth_perception_scene0000_shape():
    from multi_spatialmllm_b200 import streams, synth
    from multi_spatialmllm_b200.engine import DeviceScene
    s = synth.make_scene(seed=0, n_points=81_369, n_frames=65, n_objects=20, orbit=True, scene_id="scene0000_00")
    sc = DeviceScene(s.points, s.K, s.c2w, axis_align=s.axis_align, instance_ids=s.instance_ids, num_objects=20)
    depth = sc.render_depth(margin_mm=1).cpu().numpy()
    sc.project_visibility()
    mb = sc.mask_bool()
    rng = np.random.default_rng(5)                      # DE-dot seeds with 5 (depth_estimation_dot_engine.py:15-16)
    frames, points = [], []
    for f in range(65):
        vis = np.flatnonzero(mb[f])
        assert len(vis) > 100
        pick = rng.choice(vis, 3, replace=False)
        frames += [f] * 3
        points += pick.tolist()
    out = streams.depth_estimation_samples(sc, s.scene_id, s.image_ids, frames, points, variant="dot")
    assert len(out) == len(frames)                      # table <-> re-check invariant: nothing is skipped
    pts = s.points.astype(np.float64)
    for smp, f, p in zip(out, frames, points):
        E = og.aligned_pose(s.axis_align, s.c2w[f])
        a = og.annotate_point(pts[p], s.K, E, depth[f], *s.image_hw)
        assert smp["gt_value"] == a["depth_mm"] and tuple(smp["ori_coordinates"]) == a["pix"]
        assert smp["image"] == [f"scene0000_00/{s.image_ids[f]}_p{p}_annotated.jpg"]
    # the 64 adjacent pairs: overlap table against the oracle
    pairs = np.array([(i, i + 1) for i in range(64)], np.int32)
    rec = sc.records_to_numpy(sc.pair_annotate(pairs, np.zeros(64, np.uint32), 0))
    for (a, b), r in zip(pairs, rec):
        assert r["overlap"] == og.camera_overlap(mb[a], mb[b])


def test_config2_tapvid_videos():
    from multi_spatialmllm_b200 import streams, synth
    from multi_spatialmllm_b200.engine import DeviceVideo
    total = 0
    for seed, (T, N) in enumerate([(300, 1024), (150, 512), (150, 64)]):         # ADT-like, PStudio-like, tiny
        v = synth.make_video(seed=seed, n_frames=T, n_tracks=N, n_groups=5)
        vid = DeviceVideo(v.tracks_xyz, v.w2c, v.fxfycxcy, v.image_hw).project()
        world = og.tracks_to_world(v.tracks_xyz, v.w2c)
        np.testing.assert_allclose(vid.world.cpu().numpy(), world, rtol=1e-12, atol=1e-12)
        ok = vid.valid.cpu().numpy().astype(bool)
        rng = np.random.default_rng(seed)
        tt, nn = rng.integers(0, T, 400), rng.integers(0, N, 400)
        for t, n in zip(tt, nn):
            r = og.track_project_point(v.tracks_xyz[t, n], v.fxfycxcy, *v.image_hw)
            assert ok[t, n] == (r is not None)
        # all pairs of one track's visible frames (OM-coord:484-498)
        p = int(np.argmax(v.visibility.sum(0)))
        vf = np.flatnonzero(v.visibility[:, p])[:40]
        d, f1, f2 = og.track_pair_distances(world[:, p], vf)
        tri = np.stack([f1, f2, np.full_like(f1, p)], 1)
        got = vid.pair_displacement(tri).cpu().numpy()
        np.testing.assert_allclose(got[:, 0], d, rtol=1e-10, atol=1e-13)
        out = streams.object_movement_samples(vid, v.video_id, tri)
        total += len(out)
        for smp in out[:20]:
            f1_, f2_, p_ = (int(x) for x in smp["id"].split("_")[1:])
            ref = og.track_pair_sample_values(world, v.tracks_xyz, v.w2c, v.fxfycxcy, f1_, f2_, p_, *v.image_hw)
            assert tuple(smp["p1"]) == ref["p1"] and tuple(smp["p2"]) == ref["p2"]
            assert (smp["point_moving"], smp["cam_moving"]) == (ref["point_moving"], ref["cam_moving"])
    assert total > 100


def test_config4_stress_shape_properties():
    """500 k uniformly random points (no spatial order at all), random inward-looking poses, rendered depth."""
    import torch
    from multi_spatialmllm_b200 import synth
    from multi_spatialmllm_b200.engine import DeviceScene
    rng = np.random.default_rng(1234)
    pts = rng.uniform([0, 0, 0], [8, 8, 3], (500_000, 3)).astype(np.float32)
    s = synth.make_scene(seed=1234, n_points=1000, n_frames=32)
    sc = DeviceScene(pts, s.K, s.c2w, axis_align=s.axis_align)
    depth = sc.render_depth(margin_mm=1).cpu().numpy()
    sc.project_visibility()
    mb = sc.mask_bool()
    for f in (0, 13, 31):
        E = og.aligned_pose(s.axis_align, s.c2w[f])
        zb = og.zbuffer_render(pts.astype(np.float64), s.K, E, *s.image_hw, *s.depth_hw)
        np.testing.assert_array_equal(depth[f], og.depth_to_u16(zb, 1))
        _, _, om = og.frame_visibility(pts.astype(np.float64), s.K, E, depth[f], *s.image_hw)
        np.testing.assert_array_equal(mb[f], om)
    # the nearest point of every filled depth pixel passes the strict test against its own render (+ margin), so the
    # number of visible points is at least the number of filled pixels
    filled = (depth > 0).reshape(32, -1).sum(1)
    assert (mb.sum(1) >= filled).all() and (filled > 1000).all()
    pairs = torch.randint(0, 32, (256, 2), dtype=torch.int32)
    rec = sc.records_to_numpy(sc.pair_annotate(pairs, torch.randint(0, 2**31, (256,), dtype=torch.int32), 1))
    for (a, b), r in zip(pairs.tolist(), rec):
        assert r["inter"] == int((mb[a] & mb[b]).sum()) and r["uni"] == int((mb[a] | mb[b]).sum())
        if r["point_id"] >= 0:
            assert mb[a, r["point_id"]] and mb[b, r["point_id"]] and r["flags"] == 3


def test_empty_and_ragged_inputs():
    import ctypes as C
    import torch
    from multi_spatialmllm_b200 import _lib, synth
    from multi_spatialmllm_b200.engine import DeviceScene
    lib = _lib.lib
    s = synth.make_scene(seed=2, n_points=200, n_frames=2)
    sc = DeviceScene(s.points[:33], s.K, s.c2w, axis_align=s.axis_align)       # N not a multiple of 32
    sc.render_depth(1)
    sc.project_visibility()
    assert sc.mask.shape == (2, 4)
    cam = C.byref(sc.cam)
    st = torch.cuda.current_stream().cuda_stream
    # n == 0 / F == 0 / P == 0 / M == 0 are accepted and do nothing
    assert lib.msp_project_visibility(None, 0, 0, sc.K_host.ctypes.data, sc.w2c.data_ptr(), 2, sc.depth.data_ptr(), cam, sc.mask.data_ptr(), 4, None, None, st) == 0
    assert lib.msp_project_visibility(sc.pts.data_ptr(), 33, 0, sc.K_host.ctypes.data, sc.w2c.data_ptr(), 0, None, cam, None, 4, None, None, st) == 0
    assert lib.msp_pair_annotate(None, 33, 4, None, 0, None, 0, None, 0, None, None, None, None, None, cam, None, st) == 0
    assert lib.msp_annotate_points(None, 0, 0, None, None, 0, None, None, None, cam, None, None, None, None, None, st) == 0
    assert lib.msp_zbuffer_render(None, 0, 0, sc.K_host.ctypes.data, sc.w2c.data_ptr(), 0, cam, None, st) == 0
    # a frame that sees nothing: all-zero mask row, NaN overlap for an empty union (CFR:136)
    far = np.tile(np.eye(4), (2, 1, 1))
    far[:, :3, 3] = [1e4, 1e4, 1e4]
    sc2 = DeviceScene(s.points, s.K, far)
    sc2.set_depth(np.full((2, 480, 640), 5000, np.uint16))
    sc2.project_visibility()
    assert not sc2.mask_bool().any()
    rec = sc2.records_to_numpy(sc2.pair_annotate(np.array([[0, 1]], np.int32), np.zeros(1, np.uint32), 0))
    assert rec["uni"][0] == 0 and np.isnan(rec["overlap"][0]) and rec["point_id"][0] == -1


@pytest.mark.gpu
def test_scene_on_second_device_while_first_is_current():
    """A DeviceScene on cuda:1 driven while cuda:0 is the current device (libmsp launches on the current device;
    the per-device shared-memory attribute).  Needs two GPUs; skipped on a one-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from multi_spatialmllm_b200 import synth
    from multi_spatialmllm_b200.engine import DeviceScene
    s = synth.make_scene(seed=77, n_points=30_000, n_frames=9, n_objects=5)
    out = []
    torch.cuda.set_device(0)
    for dev in ("cuda:0", "cuda:1"):
        sc = DeviceScene(s.points.astype(np.float64), s.K, s.c2w, axis_align=s.axis_align, instance_ids=s.instance_ids,
                         num_objects=s.num_objects, device=dev)
        sc.render_depth(1)
        mask, counts = sc.project_visibility_objects()
        assert mask.device == torch.device(dev)
        out.append((sc.mask_bool(), counts.cpu().numpy()))
    assert torch.cuda.current_device() == 0
    np.testing.assert_array_equal(out[0][0], out[1][0])
    np.testing.assert_array_equal(out[0][1], out[1][1])
    assert out[0][0].sum() > 0


@pytest.mark.gpu
def test_pair_annotator_public_api_matches_resident_path():
    """pipeline.PairAnnotator (host buffers in, records out; persistent slots reused across scenes of different sizes)
    returns exactly what a resident DeviceScene computes."""
    import torch
    from multi_spatialmllm_b200 import synth
    from multi_spatialmllm_b200.engine import DeviceScene
    from multi_spatialmllm_b200.pipeline import HostScene, PairAnnotator
    ann = PairAnnotator("cuda:0", n_streams=2)
    jobs = []
    for seed, n, frames in ((3, 40_000, 12), (4, 9_000, 30), (5, 65_000, 8), (6, 20_000, 16), (7, 50_001, 10)):
        s = synth.make_scene(seed=seed, n_points=n, n_frames=frames, n_objects=7)
        sc = DeviceScene(s.points, s.K, s.c2w, axis_align=s.axis_align, instance_ids=s.instance_ids, num_objects=s.num_objects)
        depth = sc.render_depth(1)
        pairs = torch.arange(frames, dtype=torch.int32).view(-1, 2)
        sel = torch.arange(frames // 2, dtype=torch.int32) * 977 + 5
        _, counts = sc.project_visibility_objects()
        rec = sc.records_to_numpy(sc.pair_annotate(pairs, sel, sel_mode=1))
        hs = HostScene(torch.from_numpy(s.points), s.K, s.axis_align, torch.from_numpy(s.c2w), depth.cpu(),
                       torch.from_numpy(s.instance_ids), s.num_objects, s.image_hw, s.depth_hw)
        if seed % 2:
            hs = hs.pin()
        jobs.append((ann.submit(hs, pairs, sel, sel_mode=1), rec, counts.cpu().numpy()))
    done = ann.drain()
    assert len(done) == 5                       # (three of them were already finished when their slot was reused)
    for h, rec, counts in jobs:
        assert "slot" not in h
        for f in ("inter", "uni", "point_id", "permille", "depth_mm", "pix"):
            np.testing.assert_array_equal(h["records"][f], rec[f])
        np.testing.assert_array_equal(h["object_counts"], counts)
