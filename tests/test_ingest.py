"""ScanNet ingestion (multi_spatialmllm_b200.ingest) against goldens produced by the reference's own scripts
(batch_load_scannet_data.py, extract_posed_images.py, update_info_file_with_images.py — tests/golden/make_golden_ingest.py)
on seeded synthetic raw scans.  Parity bar: exact arrays, character-identical text files."""
import os
import pickle

import cv2
import numpy as np
import pytest

from multi_spatialmllm_b200 import ingest
from oracle import ref_layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCANS = [("scene0900_00", 3, True), ("scene0901_00", 4, False)]


@pytest.fixture(scope="module")
def golden():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "ingest_small.npz"), allow_pickle=False))


@pytest.fixture(scope="module")
def raw(tmp_path_factory):
    tmp = str(tmp_path_factory.mktemp("raw"))
    scans = os.path.join(tmp, "scans")
    tsv = ref_layout.write_label_map(os.path.join(tmp, "meta", "labels.tsv"))
    written = {scan: ref_layout.write_raw_scan(scans, scan, seed=seed, with_alignment=al) for scan, seed, al in SCANS}
    names = os.path.join(tmp, "meta", "scans.txt")
    with open(names, "w") as f:
        f.write("\n".join(s for s, _, _ in reversed(SCANS)) + "\n")           # unsorted on purpose (BLS:283 sorts)
    return tmp, scans, tsv, names, written


def test_ply_reader_binary_and_ascii(raw, tmp_path):
    tmp, scans, _, _, written = raw
    scan = SCANS[0][0]
    v = ingest.read_mesh_vertices_rgb(os.path.join(scans, scan, scan + "_vh_clean_2.ply"))
    assert v.dtype == np.float32 and v.shape == (600, 6)
    np.testing.assert_array_equal(v[:, :3], written[scan]["xyz"])
    np.testing.assert_array_equal(v[:, 3:], written[scan]["rgba"][:, :3].astype(np.float32))
    p = str(tmp_path / "a.ply")
    with open(p, "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex 2\nproperty float x\nproperty float y\nproperty float z\n"
                "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n0.5 1 2 10 20 30\n-1 0 3.25 1 2 3\n")
    np.testing.assert_array_equal(ingest.read_mesh_vertices_rgb(p), [[0.5, 1, 2, 10, 20, 30], [-1, 0, 3.25, 1, 2, 3]])
    with open(p, "w") as f:
        f.write("plx\n")
    with pytest.raises(ValueError):
        ingest.read_ply_vertices(p)


def test_batch_export_matches_reference(raw, golden):
    tmp, scans, tsv, names, _ = raw
    out_dir = os.path.join(tmp, "data", "scannet", "scannet_instance_data")
    pkl = ingest.batch_export(None, out_dir, names, tsv, scans)
    assert os.path.basename(pkl) == "scenes_train_val_info.pkl"
    infos = pickle.load(open(pkl, "rb"))
    assert list(infos) == sorted(s for s, _, _ in SCANS)
    for scan, _, _ in SCANS:
        g = {k.split("/", 1)[1]: v for k, v in golden.items() if k.startswith(scan + "/")}
        inst = os.path.join(out_dir, scan)
        for name in ("unaligned_points", "aligned_points", "semantic_mask", "instance_mask"):
            a = np.load(os.path.join(inst, name + ".npy"))
            assert a.dtype == g[name].dtype, name
            np.testing.assert_array_equal(a, g[name], err_msg=name)
        cats = np.load(os.path.join(inst, "raw_category_mask.npy"), allow_pickle=True)
        assert cats.dtype == object and (cats.astype(str) == g["raw_category_mask"]).all()    # None -> "None" on both sides
        info = infos[scan]
        n_obj = info["num_objects"]
        assert n_obj == len(g["aligned_bbox"])
        np.testing.assert_array_equal(np.stack([info[o]["aligned_bbox"] for o in range(n_obj)]), g["aligned_bbox"])
        np.testing.assert_array_equal(np.stack([info[o]["unaligned_bbox"] for o in range(n_obj)]), g["unaligned_bbox"])
        assert [info[o]["raw_category"] for o in range(n_obj)] == g["raw_category"].tolist()
        np.testing.assert_array_equal(info["axis_align_matrix"], g["axis_align_matrix"])
        rows = [len(np.load(os.path.join(inst, f"object_{o}_aligned_points.npy"), allow_pickle=True)) for o in range(n_obj)]
        assert rows == g["object_rows"].tolist()
        np.testing.assert_array_equal(np.load(os.path.join(inst, "object_0_aligned_points.npy")), g["object_0_aligned_points"])
        assert np.load(os.path.join(inst, "object_0_unaligned_points.npy")).dtype == np.float32
    assert (golden["scene0901_00/axis_align_matrix"] == np.eye(4)).all()          # scan without an axisAlignment line


def test_sens_extraction_matches_reference(raw, golden):
    tmp, scans, _, _, written = raw
    for scan, _, _ in SCANS:
        g = {k.split("/", 1)[1]: v for k, v in golden.items() if k.startswith(scan + "/")}
        posed = os.path.join(tmp, "data", "scannet", "posed_images", scan)
        sens = os.path.join(scans, scan, scan + ".sens")
        rd = ingest.SensReader(sens, 1)
        assert [rd.color_width, rd.color_height, rd.depth_width, rd.depth_height, rd.depth_shift, len(rd.frames)] == g["header"].tolist()
        assert rd.sensor_name == b"StructureSensor" and rd.color_compression_type == "jpeg" and rd.depth_compression_type == "zlib_ushort"
        np.testing.assert_array_equal(rd.depth_stack(), g["depth"])
        assert bytes(rd.frames[2].color_data) == written[scan]["jpgs"][2]
        assert rd.frames[5].timestamp_color == 1005 and rd.frames[5].timestamp_depth == 2005
        rd.close()
        assert ingest.extract_posed_images(sens, posed, 1) == 7
        assert ingest.extract_posed_images(sens, posed, 1) == 0                      # exists -> skipped (EPI:168-171)
        files = sorted(os.listdir(posed))
        assert files == g["posed_files"].tolist()
        texts = [open(os.path.join(posed, f)).read() for f in files if f.endswith(".txt")]
        assert texts == g["pose_text"].tolist()                                       # incl. "-inf" rows and intrinsic.txt
        depth = np.stack([cv2.imread(os.path.join(posed, f), -1) for f in files if f.endswith(".png")])
        assert depth.dtype == np.uint16
        np.testing.assert_array_equal(depth, g["depth"])
        assert open(os.path.join(posed, "00004.jpg"), "rb").read() == written[scan]["jpgs"][4]
        # frame_skip keeps frames 0, 3, 6 and renumbers them (EPI:101-106, index_to_str)
        rd3 = ingest.SensReader(sens, 3)
        assert len(rd3.frames) == 3
        np.testing.assert_array_equal(rd3.depth_image(1), g["depth"][3])
        p = rd3.poses()
        assert p.shape == (3, 4, 4) and p.dtype == np.float64 and np.isinf(p[1]).all()
        rd3.close()


def test_update_info_matches_reference(raw, golden):
    tmp, scans, tsv, names, _ = raw
    out_dir = os.path.join(tmp, "data", "scannet", "scannet_instance_data")
    posed_root = os.path.join(tmp, "data", "scannet", "posed_images")
    pkl = os.path.join(out_dir, "scenes_train_val_info.pkl")
    if not os.path.exists(pkl):
        ingest.batch_export(None, out_dir, names, tsv, scans)
    for scan, _, _ in SCANS:
        ingest.extract_posed_images(os.path.join(scans, scan, scan + ".sens"), os.path.join(posed_root, scan), 1)
    infos = ingest.update_info_with_images(pickle.load(open(pkl, "rb")), posed_root, frame_skip=5)
    for scan, _, _ in SCANS:
        g = {k.split("/", 1)[1]: v for k, v in golden.items() if k.startswith(scan + "/")}
        ii = infos[scan]["images_info"]
        assert list(ii) == g["info_image_ids"].tolist() == ["00000", "00005"]
        np.testing.assert_array_equal(np.stack([ii[k]["extrinsic_matrix"] for k in ii]), g["info_extrinsics"])
        assert [[ii[k]["image_path"], ii[k]["depth_image_path"]] for k in ii] == g["info_paths"].tolist()
        np.testing.assert_array_equal(infos[scan]["intrinsic_matrix"], g["info_intrinsic"])
        assert infos[scan]["num_posed_images"] == int(g["info_num_posed"])


def test_ingested_scene_feeds_the_handler_schema(raw):
    """The info dict produced by the ingestion has every key the geometry path reads (info_handler.py:7-30)."""
    tmp, scans, tsv, names, _ = raw
    out_dir = os.path.join(tmp, "data2", "scannet_instance_data")
    posed_root = os.path.join(tmp, "data2", "posed_images")
    pkl = ingest.batch_export(None, out_dir, names, tsv, scans)
    for scan, _, _ in SCANS:
        ingest.extract_posed_images(os.path.join(scans, scan, scan + ".sens"), os.path.join(posed_root, scan), 1)
    infos = ingest.update_info_with_images(pickle.load(open(pkl, "rb")), posed_root, frame_skip=2)
    info = infos[SCANS[0][0]]
    assert {"axis_align_matrix", "num_objects", "images_info", "intrinsic_matrix", "num_posed_images"} <= set(info)
    assert info["intrinsic_matrix"].shape == (4, 4) and info["num_posed_images"] == 4
    assert not np.isfinite(info["images_info"]["00003"]["extrinsic_matrix"]).any() if "00003" in info["images_info"] else True
    assert set(info[0]) == {"aligned_bbox", "unaligned_bbox", "raw_category"}
