"""Drop-in handler + table writers.  GPU tests compare against the reference goldens; CPU tests cover the readers."""
import os

import numpy as np
import pandas as pd
import pytest

from multi_spatialmllm_b200 import synth
from oracle import geometry as og
from oracle import ref_layout


def _golden_scene(g):
    s = synth.SynthScene("scene0011_00", g["points"], g["instance_ids"], int(g["num_objects"]), g["axis_align"],
                         g["c2w"], g["K"], tuple(int(v) for v in g["image_hw"]), tuple(int(v) for v in g["depth_hw"]),
                         depth=g["depth"], image_ids=[str(v) for v in g["image_ids"]])
    return s


@pytest.fixture(scope="module")
def disk_scene(golden_scannet, tmp_path_factory):
    root = str(tmp_path_factory.mktemp("scannet"))
    s = _golden_scene(golden_scannet)
    info = ref_layout.write_split([s], root)
    return s, root, info


def _handler(root, info):
    from multi_spatialmllm_b200.handler.info_handler import SceneInfoHandler
    return SceneInfoHandler(info, posed_images_root=os.path.join(root, "posed_images"),
                            instance_data_root=os.path.join(root, "scannet_instance_data"))


def test_visibility_table_roundtrip_cpu(golden_scannet, tmp_path):
    from multi_spatialmllm_b200 import tables
    from multi_spatialmllm_b200.handler.info_handler import VisibilityInfoHandler
    g = golden_scannet
    valid = [str(v) for v in g["valid_ids"]]
    off = g["ref_i2p_off"]
    tab = {"image_to_points": {v: g["ref_i2p_flat"][off[k]:off[k + 1]].tolist() for k, v in enumerate(valid)},
           "point_to_images": {p: [valid[k] for k in np.flatnonzero(g["ref_mask"][:, p])] for p in range(50)}}
    path = str(tmp_path / "vis.parquet")
    n = tables.save_visibility_parquet({"scene0011_00": tab}, path)
    assert n == len(valid) + 50
    df = pd.read_parquet(path)
    assert list(df.columns) == ["key", "values"] and df["key"][0] == f"scene0011_00:image_to_points:{valid[0]}"
    h = VisibilityInfoHandler(path)
    assert h.get_image_to_points_info("scene0011_00", valid[2]) == tab["image_to_points"][valid[2]]
    assert h.get_point_to_images_info("scene0011_00", 7) == tab["point_to_images"][7]
    with pytest.raises(ValueError):
        h.get_image_to_points_info("scene0011_00", "99999")


def test_overlap_table_schema_cpu(tmp_path):
    from multi_spatialmllm_b200 import tables
    info = {"s": {("00000", "00005"): {"overlap": 12.5, "distance": 1.0, "yaw": 3.0, "pitch": -1.0},
                  ("00000", "00010"): {"overlap": 0.0, "distance": 2.0, "yaw": 30.0, "pitch": 1.0},
                  ("00005", "00010"): {"overlap": float("nan"), "distance": 2.5, "yaw": 9.0, "pitch": 0.0}}}
    path = str(tmp_path / "cam.parquet")
    assert tables.save_overlap_parquet(info, path) == (3, 2)            # NaN != 0.0 -> kept in _nonzero (CFR:67)
    df = pd.read_parquet(path)
    assert list(df.columns) == ["scene_id", "image_id1", "image_id2", "overlap", "distance", "yaw", "pitch"]
    assert len(pd.read_parquet(path.replace(".parquet", "_nonzero.parquet"))) == 2


def test_handler_getters_cpu(disk_scene, golden_scannet):
    s, root, info = disk_scene
    h = _handler(root, info)
    sid = s.scene_id
    assert h.get_all_extrinsic_valid_image_ids(sid) == [str(v) for v in golden_scannet["valid_ids"]]
    assert h.get_image_shape(sid) == (968, 1296) == h.get_image_size(sid)
    assert (h.image_height, h.image_width) == (968, 1296)
    assert h.convert_image_id_to_key(35) == "00035" and h.convert_image_id_to_key(-1) is None
    np.testing.assert_array_equal(h.get_extrinsic_matrix_align(sid, "00005"), golden_scannet["ref_E_aligned"][1])
    assert h.get_scene_points_align(sid).shape == (s.n_points, 6)
    np.testing.assert_array_equal(h.get_object_point_index(sid, 2), np.flatnonzero(s.instance_ids == 3))
    assert h.get_depth_image(sid, 0).dtype == np.uint16


@pytest.mark.gpu
def test_handler_geometry_matches_reference(disk_scene, golden_scannet):
    g = golden_scannet
    s, root, info = disk_scene
    h = _handler(root, info)
    sid = s.scene_id
    pts = h.get_scene_points_align(sid)[:, :3]
    for k, image_id in enumerate(h.get_all_extrinsic_valid_image_ids(sid)):
        xy, z = h.project_3d_point_to_image(sid, image_id, pts)
        assert xy.dtype == np.float64 and xy.shape == (len(pts), 2) and z.shape == (len(pts),)
        np.testing.assert_array_equal(xy, g["ref_xy"][k])          # bit-identical to the reference's floats
        np.testing.assert_array_equal(z, g["ref_z"][k])
        # the two-step API on the REFERENCE's own xy/z must reproduce its masks bit for bit
        m = h.check_point_visibility(sid, image_id, g["ref_xy"][k], g["ref_z"][k])
        assert m.dtype == bool
        np.testing.assert_array_equal(m, g["ref_mask"][k])
        f = list(g["image_ids"]).index(image_id)
        np.testing.assert_array_equal(h.check_point_in_image_boundary(sid, g["ref_xy"][k]),
                                      og.in_image_boundary(g["ref_xy"][k], 968, 1296))
        np.testing.assert_array_equal(h.check_point_visibility_by_depth(sid, image_id, g["ref_xy"][k], g["ref_z"][k]),
                                      og.visible_by_depth(g["ref_xy"][k], g["ref_z"][k], g["depth"][f], 968, 1296))
    valid = [str(v) for v in g["valid_ids"]]
    for row in g["ref_spot"][::5]:
        xy, z = h.get_point_2d_coordinates_in_image(sid, valid[int(row[0])], int(row[1]), align=True,
                                                    check_visible=True, return_depth=True)
        assert len(xy) == 1
        np.testing.assert_array_equal([xy[0][0], xy[0][1], z[0]], row[2:5])      # single-point call: exact as well
        assert [round((xy[0][0] / 1296) * 1000), round((xy[0][1] / 968) * 1000), round(z[0] * 1000)] == [int(v) for v in row[5:8]]
    # module-level project_points
    from multi_spatialmllm_b200.handler.info_handler import project_points
    xy, z = project_points(og.homogenize(pts[:100]), g["K"], g["ref_E_aligned"][0])
    np.testing.assert_array_equal(xy, g["ref_xy"][0][:100])
    with pytest.raises(ValueError):
        project_points(np.ones((3, 4)) * 2.0, g["K"], g["ref_E_aligned"][0])


@pytest.mark.gpu
def test_split_drivers_write_reference_tables(disk_scene, golden_scannet, tmp_path):
    from multi_spatialmllm_b200 import tables
    from multi_spatialmllm_b200.handler.info_handler import VisibilityInfoHandler
    g = golden_scannet
    s, root, info = disk_scene
    h = _handler(root, info)
    vis_file, cam_file = str(tmp_path / "vis_D5.parquet"), str(tmp_path / "cam_D5.parquet")
    tables.run_visibility_split(h, vis_file)
    tables.run_relations_split(h, cam_file)
    v = VisibilityInfoHandler(vis_file)
    off = g["ref_i2p_off"]
    for k, image_id in enumerate(g["valid_ids"]):
        assert v.get_image_to_points_info(s.scene_id, str(image_id)) == g["ref_i2p_flat"][off[k]:off[k + 1]].tolist()
    for p in (0, 17, s.n_points - 1):
        assert len(v.get_point_to_images_info(s.scene_id, p)) == g["ref_p2i_cnt"][p]
    df = pd.read_parquet(cam_file)
    assert len(df) == len(g["ref_pair_ids"])
    for (a, b), vals, (_, row) in zip(g["ref_pair_ids"], g["ref_pair_vals"], df.iterrows()):
        assert (row["image_id1"], row["image_id2"]) == (str(a), str(b))
        np.testing.assert_allclose([row["overlap"], row["distance"], row["yaw"], row["pitch"]], vals, rtol=1e-12, atol=1e-12)
    ov = tables.run_object_visibility_split(h, str(tmp_path / "objvis.pkl"))[s.scene_id]
    valid_rows = [list(g["image_ids"]).index(vv) for vv in g["valid_ids"]]
    totals = np.bincount(g["instance_ids"].astype(np.int64), minlength=int(g["num_objects"]) + 1)[1:]
    for k, image_id in enumerate(g["valid_ids"]):
        counts = og.object_visible_counts(g["ref_mask"][k], g["instance_ids"], int(g["num_objects"]))
        want = og.object_visibility_rows(counts, totals)
        got = [(e["object_id"], e["intersection_count"], e["visibility"]) for e in ov["image_to_objects"].get(str(image_id), [])]
        assert got == want


@pytest.mark.gpu
def test_object_coverage_and_perception_stream(disk_scene, golden_scannet):
    """Scene-level coverage driver (SOCF:222-284) + object_perception writer (SOPE:129-227) on the golden scene;
    the BFS itself is pinned to the reference in tests/test_streams.py."""
    import json
    import random
    from multi_spatialmllm_b200 import coverage, streams, tables
    s, root, info = disk_scene
    h = _handler(root, info)
    sc, image_ids = tables.device_scene_from_handler(h, s.scene_id)
    ov = tables.object_visibility_scene(h, s.scene_id, sc, image_ids)
    random.seed(0)
    cov = tables.object_coverage_scene(h, s.scene_id, sc, image_ids, ov)
    assert set(cov) == set(ov["object_to_images"])
    mb = sc.mask_bool()
    pts = golden_scannet["points"].astype(np.float64)
    inst = golden_scannet["instance_ids"]
    n_checked = 0
    for obj, res in cov.items():
        for name, axis, target in (("height", 2, h.get_object_height(s.scene_id, obj)),):
            for k, combos in res[name].items():
                for combo in combos:
                    union = np.zeros(len(pts), bool)
                    for img in combo:
                        union |= mb[image_ids.index(img)]
                    sel = union & (inst == obj + 1)
                    covv = pts[sel][:, axis].max() - pts[sel][:, axis].min()       # SOCF:57-65 on the union mask
                    assert coverage.covers_dimension(covv, target, 0.1) and len(combo) == k
                    n_checked += 1
    assert n_checked > 3
    dims = {o: h.get_object_height(s.scene_id, o) for o in cov}
    cats = {o: h.get_object_raw_category(s.scene_id, o) for o in cov}
    out = streams.object_perception_samples(s.scene_id, h.get_image_shape(s.scene_id), "height", dims, cats,
                                            {o: r["height"] for o, r in cov.items()})
    flat = [x for v in out.values() for x in v]
    assert len(flat) == n_checked
    for x in flat:
        o = int(x["id"].split("_")[2])
        assert x["gt_value"] == int(round(dims[o] * 1000)) and x["question_type"] == "object_perception_height_estimation"
        assert len(x["image"]) == len(x["height_list"]) == int(x["id"].split("_")[3]) and json.dumps(x)


def test_parquet_stream_cpu(tmp_path):
    """Row-group-per-scene writer: concatenation in write order, no file when nothing was written (CFR:52-54)."""
    import pyarrow as pa
    from multi_spatialmllm_b200 import tables
    path = str(tmp_path / "sub" / "t.parquet")
    w = tables._ParquetStream(path)
    w.write(pa.table({"key": ["a", "b"], "values": ["[]", "[1, 2]"]}))
    w.write(pa.table({"key": pa.array([], pa.string()), "values": pa.array([], pa.string())}))
    w.write(pa.table({"key": ["c"], "values": ['["00000"]']}))
    assert w.close() == 3
    df = pd.read_parquet(path)
    assert df["key"].tolist() == ["a", "b", "c"] and df["values"].tolist() == ["[]", "[1, 2]", '["00000"]']
    empty = tables._ParquetStream(str(tmp_path / "none.parquet"))
    assert empty.close() == 0 and not os.path.exists(str(tmp_path / "none.parquet"))
    with pytest.raises(ValueError):
        tables._check_json_safe(['quo"te'])
    tables._check_json_safe(["00005", "frame_12", "a b"])


def test_parquet_stream_dataset_layout_reads_like_the_single_file(tmp_path):
    """_ParquetStream(writers > 1): a *.parquet directory of per-scene files whose rows, read through the path the way
    VisibilityInfoHandler does (pd.read_parquet, IH:451-460), equal the single file's rows in the same order."""
    import pandas as pd
    import pyarrow as pa
    from multi_spatialmllm_b200 import tables
    tabs = [pa.table({"key": [f"scene{k:04d}_00:point_to_images:{i}" for i in range(50 + k)],
                      "values": [str(list(range(i % 7))) for i in range(50 + k)]}) for k in range(13)]
    one, many = str(tmp_path / "one.parquet"), str(tmp_path / "many.parquet")
    w1, wn = tables._ParquetStream(one), tables._ParquetStream(many, writers=4)
    for t in tabs:
        w1.write(t)
        wn.write(t)
    assert w1.close() == wn.close() == sum(t.num_rows for t in tabs)
    assert os.path.isfile(one) and os.path.isdir(many) and len(os.listdir(many)) == 13
    a, b = pd.read_parquet(one), pd.read_parquet(many)
    assert a["key"].tolist() == b["key"].tolist() and a["values"].tolist() == b["values"].tolist()
