"""Drop-in for the reference's ``spatial_engine/utils/scannet_utils/handler/info_handler.py``.

Same public names, argument order and return types, so the reference's engines run with

    from multi_spatialmllm_b200.handler.info_handler import SceneInfoHandler, VisibilityInfoHandler

instead of the reference import.  The geometry methods run on the GPU through libmsp.so:

    project_points                    <- IH:46-72     msp_pose_prepare + msp_project_points (float64, NumPy's op order for the shape)
    project_3d_point_to_image         <- IH:313-335
    check_point_in_image_boundary     <- IH:337-344   msp_check_visibility(mode=1)
    check_point_visibility_by_depth   <- IH:346-373   msp_check_visibility(mode=2)
    check_point_visibility            <- IH:375-386   msp_check_visibility(mode=3)
    get_point_2d_coordinates_in_image <- IH:291-305

and return NumPy float64 / bool arrays like the reference.  Three members the reference's engines call but
its handler lacks are provided: ``get_image_size``, ``image_height``, ``image_width``.
Differences by design: files are decoded once and cached (the reference re-reads the .npy / .jpg / .png on
every call, IH:133-139,153-157,272-289); homogeneous inputs to ``project_points`` must have w == 1 (every
reference call site appends ones, IH:331).  There is no CPU fallback: without libmsp.so or a CUDA device the
geometry methods raise.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import pickle

import cv2
import numpy as np
import pandas as pd
import torch

from .. import _lib
from .._lib import MspCamera, check, lib


def _cuda_device():
    if not torch.cuda.is_available():
        raise RuntimeError("multi_spatialmllm_b200.handler needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _project_on_device(points_xyz: np.ndarray, K: np.ndarray, E_c2w: np.ndarray):
    """(N,3) float64 world points -> xy (N,2), z (N,) float64 on the host, reference operation order."""
    dev = _cuda_device()
    n = int(points_xyz.shape[0])
    pts = torch.from_numpy(np.ascontiguousarray(points_xyz, dtype=np.float64)).to(dev)
    c2w = torch.from_numpy(np.ascontiguousarray(E_c2w, dtype=np.float64).reshape(1, 16)).to(dev)
    E = torch.empty((1, 16), dtype=torch.float64, device=dev)
    w2c = torch.empty((1, 16), dtype=torch.float64, device=dev)
    yp = torch.empty((1, 2), dtype=torch.float64, device=dev)
    valid = torch.empty((1,), dtype=torch.int32, device=dev)
    check(lib.msp_pose_prepare(c2w.data_ptr(), None, 1, E.data_ptr(), w2c.data_ptr(), yp.data_ptr(), valid.data_ptr(),
                               _stream()), "msp_pose_prepare")
    if int(valid.item()) == 0:
        # np.linalg.inv raises LinAlgError on a singular pose and propagates NaN for non-finite ones (IH:57)
        if np.all(np.isfinite(E_c2w)):
            raise np.linalg.LinAlgError("Singular matrix")
    uvz = torch.empty((n, 3), dtype=torch.float64, device=dev)
    Kh = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(16))
    check(lib.msp_project_points(pts.data_ptr(), n, _lib.MSP_PTS_F64X3, Kh.ctypes.data, w2c.data_ptr(), uvz.data_ptr(),
                                 _stream()), "msp_project_points")
    out = uvz.cpu().numpy()
    return out[:, :2], out[:, 2].copy()


def project_points(points, K, E):
    """IH:46-72: (N,4) homogeneous world points, 4x4 K, 4x4 camera->world E -> (xy (N,2), depth (N,))."""
    points = np.asarray(points, dtype=np.float64)
    if points.ndim != 2 or points.shape[1] != 4:
        raise ValueError("points must be (N,4) homogeneous coordinates")
    if not np.all(points[:, 3] == 1.0):
        raise ValueError("project_points on the GPU path requires w == 1 (the reference always appends ones, IH:331)")
    return _project_on_device(points[:, :3], K, E)


class SceneInfoHandler:
    """Scene-info dict access + projection / visibility (reference: IH:74-443)."""

    def __init__(self, info_path, posed_images_root="data/scannet/posed_images",
                 instance_data_root="data/scannet/scannet_instance_data", mask_image_root="data/scannet/scans",
                 depth_value_scale=0.001):
        try:
            with open(info_path, "rb") as f:
                self.infos = pickle.load(f)
            print(f"Data from {info_path} loaded successfully.")
        except Exception as e:  # the reference prints and exits (IH:77-82)
            print(f"Failed to load data from {info_path}: {e}")
            raise SystemExit(1)
        self.posed_images_root = posed_images_root
        self.instance_data_root = instance_data_root
        self.mask_image_root = mask_image_root
        self.depth_value_scale = depth_value_scale
        self._shape_cache, self._npy_cache, self._depth_cache = {}, {}, {}

    # ---- plain dictionary access --------------------------------------------------------------
    def __len__(self):
        return len(self.infos)

    def get_sorted_keys(self):
        return sorted(self.infos.keys())

    def get_all_scene_ids(self):
        return list(self.infos.keys())

    def _scene(self, scene_id):
        return self.infos[scene_id]

    def get_intrinsic_matrix(self, scene_id, image_id=None):
        return self._scene(scene_id)["intrinsic_matrix"]

    def get_world_to_axis_align_matrix(self, scene_id, image_id=None):
        return self._scene(scene_id)["axis_align_matrix"]

    def get_num_posed_images(self, scene_id):
        return self._scene(scene_id)["num_posed_images"]

    def get_num_objects(self, scene_id):
        return self._scene(scene_id)["num_objects"]

    def get_all_image_ids(self, scene_id):
        return list(self._scene(scene_id)["images_info"].keys())

    def convert_image_id_to_key(self, image_id):
        """IH:159-173: ints / numeric strings -> five-digit key, None for negatives or junk."""
        try:
            v = int(image_id)
        except Exception as e:
            print(f"Failed to convert image_id: {image_id}: {e}")
            return None
        return None if v < 0 else f"{v:05d}"

    def get_extrinsic_matrix(self, scene_id, image_id, warning=True):
        """Camera->world pose as stored (IH:100-111)."""
        key = self.convert_image_id_to_key(image_id)
        E = self._scene(scene_id)["images_info"][key]["extrinsic_matrix"]
        if warning and not np.all(np.isfinite(E)):
            print(f"[SceneInfoHanlder] Warning: extrinsics matrix of {scene_id}: {key} contains inf or nan.")
        return E

    def get_extrinsic_matrix_align(self, scene_id, image_id):
        """IH:113-124: axis_align @ c2w (host NumPy: a 4x4 product, same BLAS call as the reference)."""
        return self.get_world_to_axis_align_matrix(scene_id) @ self.get_extrinsic_matrix(scene_id, image_id)

    def is_posed_image_valid(self, scene_id, image_id):
        key = self.convert_image_id_to_key(image_id)
        if key is None:
            return False
        return bool(np.all(np.isfinite(self.get_extrinsic_matrix(scene_id, key, warning=False))))

    def get_all_extrinsic_valid_image_ids(self, scene_id):
        return [i for i in self.get_all_image_ids(scene_id) if self.is_posed_image_valid(scene_id, i)]

    def _posed_path(self, scene_id, image_id, ext):
        key = self.convert_image_id_to_key(image_id)
        return None if key is None else os.path.join(self.posed_images_root, scene_id, f"{key}.{ext}")

    def get_image_path(self, scene_id, image_id):
        return self._posed_path(scene_id, image_id, "jpg")

    def get_depth_image_path(self, scene_id, image_id):
        return self._posed_path(scene_id, image_id, "png")

    def get_image_shape(self, scene_id, image_id=None):
        """(H, W) of the colour images; decoded once per scene (the reference decodes a JPEG per call, IH:133-139)."""
        if scene_id not in self._shape_cache:
            if image_id is None:
                image_id = self.get_all_image_ids(scene_id)[0]
            img = cv2.imread(self.get_image_path(scene_id, image_id))
            self._shape_cache[scene_id] = tuple(img.shape[:2])
        return self._shape_cache[scene_id]

    # members the reference's engines use although its handler does not define them
    def get_image_size(self, scene_id):
        return self.get_image_shape(scene_id)

    @property
    def image_height(self):
        return self.get_image_shape(self.get_all_scene_ids()[0])[0]

    @property
    def image_width(self):
        return self.get_image_shape(self.get_all_scene_ids()[0])[1]

    def get_depth_image_shape(self, scene_id, image_id=0):
        return tuple(self.get_depth_image(scene_id, image_id).shape[:2])

    def get_depth_image(self, scene_id, image_id):
        key = (scene_id, self.convert_image_id_to_key(image_id))
        d = self._depth_cache.get(key)
        if d is None:                                        # called from decoder threads too: never index after a clear()
            d = cv2.imread(self.get_depth_image_path(scene_id, image_id), -1)
            if len(self._depth_cache) > 256:
                self._depth_cache.clear()
            self._depth_cache[key] = d
        return d

    # ---- objects --------------------------------------------------------------------------------
    def get_object_gt_bbox(self, scene_id, object_id, axis_aligned=True, with_class_id=False):
        box = self._scene(scene_id)[object_id]["aligned_bbox" if axis_aligned else "unaligned_bbox"]
        return box if with_class_id else box[0:-1]

    def get_object_raw_category(self, scene_id, object_id):
        return self._scene(scene_id)[object_id]["raw_category"]

    def get_scene_raw_categories(self, scene_id):
        return [self.get_object_raw_category(scene_id, o) for o in range(self.get_num_objects(scene_id))]

    def _extent(self, scene_id, object_id):
        b = self.get_object_gt_bbox(scene_id, object_id, axis_aligned=True)
        return b[3], b[4], b[5]

    def get_object_height(self, scene_id, object_id):
        return self._extent(scene_id, object_id)[2]

    def get_object_length(self, scene_id, object_id):
        dx, dy, _ = self._extent(scene_id, object_id)
        return max(dx, dy)

    def get_object_width(self, scene_id, object_id):
        dx, dy, _ = self._extent(scene_id, object_id)
        return min(dx, dy)

    def get_object_volume(self, scene_id, object_id):
        dx, dy, dz = self._extent(scene_id, object_id)
        return dx * dy * dz

    def get_object_length_axis_aligned(self, scene_id, object_id):
        dx, dy, _ = self._extent(scene_id, object_id)
        return 0 if dx > dy else 1

    def get_object_width_axis_aligned(self, scene_id, object_id):
        dx, dy, _ = self._extent(scene_id, object_id)
        return 0 if dx < dy else 1

    # ---- per-scene arrays (cached) ----------------------------------------------------------------
    def _load_npy(self, scene_id, name, **kw):
        key = (scene_id, name)
        if key not in self._npy_cache:
            if len(self._npy_cache) > 64:
                self._npy_cache.clear()
            self._npy_cache[key] = np.load(os.path.join(self.instance_data_root, scene_id, name), **kw)
        return self._npy_cache[key]

    def get_scene_points_align(self, scene_id):
        return self._load_npy(scene_id, "aligned_points.npy")

    def get_scene_points(self, scene_id):
        return self._load_npy(scene_id, "unaligned_points.npy")

    def get_scene_instance_mask(self, scene_id):
        return self._load_npy(scene_id, "instance_mask.npy")

    def get_object_points_aligned(self, scene_id, object_id):
        return self._load_npy(scene_id, f"object_{object_id}_aligned_points.npy", allow_pickle=True)

    def get_object_point_index(self, scene_id, object_id):
        """IH:250-261: indices of the scene points that belong to the (0-indexed) object; the mask is 1-indexed."""
        idx = np.where(self.get_scene_instance_mask(scene_id) == object_id + 1)[0]
        if len(idx) == 0:
            print(f"[SceneInfoHanlder] Warning: {scene_id} does not have object {object_id}.")
        return idx

    def get_point_3d_coordinates(self, scene_id, point_id, align=True):
        pts = self.get_scene_points_align(scene_id) if align else self.get_scene_points(scene_id)
        return pts[point_id]

    # ---- geometry (GPU) ---------------------------------------------------------------------------
    def project_3d_point_to_image(self, scene_id, image_id, points_3d, align=True):
        """IH:313-335: (N,3) or (3,) -> (xy (N,2) float64, depth (N,) float64)."""
        K = self.get_intrinsic_matrix(scene_id, image_id)
        E = self.get_extrinsic_matrix_align(scene_id, image_id) if align else self.get_extrinsic_matrix(scene_id, image_id)
        p = np.asarray(points_3d, dtype=np.float64)
        if p.ndim == 1:
            p = p[None, :]
        return _project_on_device(p[:, :3], K, E)

    def _camera(self, scene_id, depth_shape=(1, 1)):
        H, W = self.get_image_shape(scene_id)
        return MspCamera(int(H), int(W), int(depth_shape[0]), int(depth_shape[1]), float(self.depth_value_scale))

    def _check(self, scene_id, image_id, points_2d, points_depth, mode):
        dev = _cuda_device()
        xy = torch.from_numpy(np.ascontiguousarray(points_2d, dtype=np.float64)).to(dev)
        n = int(xy.shape[0])
        z = depth = None
        cam = self._camera(scene_id)
        if mode & 2:
            dimg = self.get_depth_image(scene_id, image_id)
            cam = self._camera(scene_id, dimg.shape[:2])
            depth = torch.from_numpy(np.ascontiguousarray(dimg)).to(dev)
            z = torch.from_numpy(np.ascontiguousarray(points_depth, dtype=np.float64)).to(dev)
        out = torch.empty((n,), dtype=torch.uint8, device=dev)
        check(lib.msp_check_visibility(xy.data_ptr(), None if z is None else z.data_ptr(), n,
                                       None if depth is None else depth.data_ptr(), C.byref(cam), mode, out.data_ptr(),
                                       _stream()), "msp_check_visibility")
        return out.cpu().numpy().astype(bool)

    def check_point_in_image_boundary(self, scene_id, points_2d):
        return self._check(scene_id, None, points_2d, None, 1)

    def check_point_visibility_by_depth(self, scene_id, image_id, points_2d, points_depth):
        return self._check(scene_id, image_id, points_2d, points_depth, 2)

    def check_point_visibility(self, scene_id, image_id, points_2d, points_depth):
        return self._check(scene_id, image_id, points_2d, points_depth, 3)

    def get_point_2d_coordinates_in_image(self, scene_id, image_id, point_id, align=True, check_visible=False,
                                          return_depth=False):
        """IH:291-305: one scene point -> xy (k,2) [, depth (k,)] with k = 0 when the visibility re-check fails."""
        p3 = self.get_point_3d_coordinates(scene_id, point_id, align)[:3]
        xy, z = self.project_3d_point_to_image(scene_id, image_id, p3, align)
        if check_visible:
            vis = self.check_point_visibility(scene_id, image_id, xy, z)
            xy, z = xy[vis], z[vis]
        return (xy, z) if return_depth else xy

    # ---- not on the geometry path -------------------------------------------------------------------
    def get_instance_mask(self, scene_id, image_id, target_id) -> np.ndarray:
        """IH:420-443: binary (H,W) mask of instance ``target_id`` from the instance-filt PNG."""
        path = os.path.join(self.mask_image_root, scene_id, "instance-filt", f"{int(image_id)}.png")
        img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if img is None:
            raise FileNotFoundError(f"Mask image not found at path: {path}")
        return np.where(img == target_id + 1, 1, 0)

    def project_image_to_3d_with_mask(self, scene_id, image_id, mask=None, with_color=False):
        raise NotImplementedError("depth->3D unprojection (ops.project_mask_to_3d, needs open3d) is outside the "
                                  "geometry hot path this package covers; no reference engine calls it")


class VisibilityInfoHandler:
    """Reader of the frame<->points visibility table (reference: IH:445-538).

    Accepts the reference's two on-disk formats: the parquet written by convert_pkl_to_parquet
    (key "{scene}:image_to_points:{image_id}" / "{scene}:point_to_images:{point}", JSON string values,
    MVI:59-65) and the nested-dict pkl (MVI:16-29).
    """

    def __init__(self, visibility_info_path):
        self.visibility_info_path = visibility_info_path
        print(f"[VisibilityInfoHandler] Reading visibility info from {self.visibility_info_path}.")
        if visibility_info_path.endswith(".parquet"):
            self.info_format = "parquet"
            df = pd.read_parquet(visibility_info_path)
            print("[VisibilityInfoHandler] Converting parquet file to dict.")
            self.visibility_info = self.convert_parquet_to_dict(df)
        elif visibility_info_path.endswith(".pkl"):
            self.info_format = "pkl"
            with open(visibility_info_path, "rb") as f:
                self.visibility_info = pickle.load(f)
        else:
            raise ValueError(f"Unsupported file format: {self.visibility_info_path}")

    def convert_parquet_to_dict(self, parquet_df):
        return dict(zip(parquet_df["key"].tolist(), parquet_df["values"].tolist()))

    def _lookup(self, scene_id, kind, item, what):
        if self.info_format == "parquet":
            key = f"{scene_id}:{kind}:{item}"
            if key not in self.visibility_info:
                raise ValueError(f"Key {key} not found in visibility info.")
            return json.loads(self.visibility_info[key])
        if scene_id not in self.visibility_info:
            raise ValueError(f"Scene {scene_id} not found in visibility info.")
        table = self.visibility_info[scene_id][kind]
        if item not in table:
            raise ValueError(f"{what} {item} not found in visibility info for scene {scene_id}.")
        return table[item]

    def get_image_to_points_info(self, scene_id, image_id):
        return self._lookup(scene_id, "image_to_points", image_id, "Image")

    def get_point_to_images_info(self, scene_id, point_index):
        return self._lookup(scene_id, "point_to_images", point_index, "Point")
