"""Golden vectors for the ingestion path, produced by the UNMODIFIED reference scripts (build container only):

  batch_load_scannet_data.export_one_scan   (BLS:138-249)  on a synthetic raw scan written by oracle/ref_layout.write_raw_scan
  extract_posed_images.SensorData + exports (EPI:59-157)
  update_info_file_with_images.py           (UIF, a script: run with runpy from a scratch cwd)

The reference imports ``plyfile`` and ``imageio`` (absent here): ``plyfile.PlyData`` is stubbed by a 20-line numpy
reader of the vertex block written right here (independent of multi_spatialmllm_b200.ingest), ``imageio`` by cv2.

    python tests/golden/make_golden_ingest.py        ->  tests/golden/ingest_small.npz
"""
import os
import pickle
import runpy
import sys
import tempfile
import types

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_layout, ref_shim  # noqa: E402

SCANS = [("scene0900_00", 3, True), ("scene0901_00", 4, False)]


def _stub_plyfile():
    class _El:
        def __init__(self, data):
            self.data, self.count = data, len(data)

    class PlyData(dict):
        @staticmethod
        def read(f):
            assert f.readline().strip() == b"ply"
            props, count, in_vertex = [], 0, False
            while True:
                t = f.readline().decode().split()
                if t[0] == "element":
                    in_vertex = t[1] == "vertex"
                    if in_vertex:
                        count = int(t[2])
                elif t[0] == "property" and in_vertex:
                    props.append((t[2], {"float": "<f4", "uchar": "u1"}[t[1]]))
                elif t[0] == "end_header":
                    break
            dt = np.dtype(props)
            out = PlyData()
            out["vertex"] = _El(np.frombuffer(f.read(count * dt.itemsize), dt))
            return out

    m = types.ModuleType("plyfile")
    m.PlyData = PlyData
    sys.modules["plyfile"] = m


def _stub_imageio():
    v2 = types.ModuleType("imageio.v2")

    def imread(data):
        return cv2.cvtColor(cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)

    def imwrite(path, arr):
        cv2.imwrite(path, arr if arr.ndim == 2 else cv2.cvtColor(arr, cv2.COLOR_RGB2BGR))

    v2.imread, v2.imwrite = imread, imwrite
    pkg = types.ModuleType("imageio")
    pkg.v2 = v2
    sys.modules["imageio"], sys.modules["imageio.v2"] = pkg, v2


def main():
    assert ref_shim.reference_available()
    ref_shim._install_mmengine_stub()
    _stub_plyfile()
    _stub_imageio()
    sys.dont_write_bytecode = True
    # EPI:118 calls np.fromstring in binary mode (fine under the pinned numpy 1.26.4, removed in numpy 2.x): same bytes -> array
    np.fromstring = lambda data, dtype=float, **k: np.frombuffer(data, dtype=dtype).copy()
    su_dir = os.path.join(ref_shim.REFERENCE_ROOT, "spatial_engine", "utils", "scannet_utils")
    sys.path.insert(0, su_dir)
    import batch_load_scannet_data as bls          # the reference module, unmodified
    import extract_posed_images as epi
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        scans = os.path.join(tmp, "scans")
        tsv = ref_layout.write_label_map(os.path.join(tmp, "meta", "labels.tsv"))
        infos = {}
        for scan, seed, align in SCANS:
            ref_layout.write_raw_scan(scans, scan, seed=seed, with_alignment=align)
            inst = os.path.join(tmp, "data", "scannet", "scannet_instance_data", scan)
            infos.update(bls.export_one_scan(scan, inst, None, tsv, scans))
            for name in ("unaligned_points", "aligned_points", "semantic_mask", "instance_mask"):
                out[f"{scan}/{name}"] = np.load(os.path.join(inst, name + ".npy"))
            out[f"{scan}/raw_category_mask"] = np.load(os.path.join(inst, "raw_category_mask.npy"), allow_pickle=True).astype(str)
            n_obj = infos[scan]["num_objects"]
            out[f"{scan}/aligned_bbox"] = np.stack([infos[scan][o]["aligned_bbox"] for o in range(n_obj)])
            out[f"{scan}/unaligned_bbox"] = np.stack([infos[scan][o]["unaligned_bbox"] for o in range(n_obj)])
            out[f"{scan}/raw_category"] = np.array([infos[scan][o]["raw_category"] for o in range(n_obj)])
            out[f"{scan}/axis_align_matrix"] = infos[scan]["axis_align_matrix"]
            out[f"{scan}/object_rows"] = np.array([len(np.load(os.path.join(inst, f"object_{o}_aligned_points.npy"), allow_pickle=True))
                                                   for o in range(n_obj)])
            out[f"{scan}/object_0_aligned_points"] = np.load(os.path.join(inst, "object_0_aligned_points.npy"), allow_pickle=True)
            # .sens -> posed images with the reference's SensorData
            posed = os.path.join(tmp, "data", "scannet", "posed_images", scan)
            sd = epi.SensorData(os.path.join(scans, scan, scan + ".sens"), 1)
            sd.export_color_images(posed), sd.export_intrinsics(posed), sd.export_poses(posed), sd.export_depth_images(posed)
            files = sorted(os.listdir(posed))
            out[f"{scan}/posed_files"] = np.array(files)
            out[f"{scan}/pose_text"] = np.array([open(os.path.join(posed, f)).read() for f in files if f.endswith(".txt")])
            out[f"{scan}/depth"] = np.stack([cv2.imread(os.path.join(posed, f), -1) for f in files if f.endswith(".png")])
            out[f"{scan}/header"] = np.array([sd.color_width, sd.color_height, sd.depth_width, sd.depth_height, sd.depth_shift,
                                              len(sd.frames)], np.float64)
        # update_info_file_with_images.py is a script with relative paths: run it from the scratch tree
        pkl = os.path.join(tmp, "data", "scannet", "scannet_instance_data", "scenes_train_val_info.pkl")
        with open(pkl, "wb") as f:
            pickle.dump(infos, f)
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            runpy.run_path(os.path.join(su_dir, "update_info_file_with_images.py"), run_name="__main__")
        finally:
            os.chdir(cwd)
        upd = pickle.load(open(pkl.replace(".pkl", "_i_D5.pkl"), "rb"))
        for scan, _, _ in SCANS:
            ii = upd[scan]["images_info"]
            out[f"{scan}/info_image_ids"] = np.array(list(ii))
            out[f"{scan}/info_extrinsics"] = np.stack([ii[k]["extrinsic_matrix"] for k in ii])
            out[f"{scan}/info_paths"] = np.array([[ii[k]["image_path"], ii[k]["depth_image_path"]] for k in ii])
            out[f"{scan}/info_intrinsic"] = upd[scan]["intrinsic_matrix"]
            out[f"{scan}/info_num_posed"] = np.array(upd[scan]["num_posed_images"])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ingest_small.npz"), **out)
    print("wrote ingest_small.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
