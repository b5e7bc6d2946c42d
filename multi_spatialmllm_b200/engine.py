"""Device-resident scene + the per-scene workers of the MultiSPA geometry path.

This is the host side above the C ABI (include/msp.h).  PyTorch is used only for
device memory, streams and host<->device copies; all geometry runs in libmsp.so.

Reference surface mirrored here (file:line relative to the reference checkout):
  DeviceScene.project_visibility  <- MVI:93-100 / CFR:152-157 frame loop (IH:313-386)
  DeviceScene.pair_annotate       <- CFR:176-189 + VC-coor:303-344 + CME:177-191
  DeviceScene.annotate_points     <- IH:291-305 + DE-dot:216-218,261
  DeviceScene.object_counts       <- COV:110-136
  process_scene_visibility        <- make_visibility_info.process_scene (MVI:75-125)
  process_scene_relations         <- calculate_frames_relations.process_scene (CFR:139-197)
"""
from __future__ import annotations

import ctypes as C
from itertools import combinations

import numpy as np
import torch

from . import _lib
from ._lib import MspCamera, check, lib

PAIR_RECORD_DTYPE = np.dtype([
    ("inter", "<u4"), ("uni", "<u4"), ("overlap", "<f8"),
    ("distance", "<f8"), ("dyaw", "<f8"), ("dpitch", "<f8"), ("rel_disp", "<f8", (3,)),
    ("point_id", "<i4"), ("flags", "<i4"), ("permille", "<i4", (4,)),
    ("depth_mm", "<i4", (2,)), ("pix", "<i4", (4,)), ("reserved", "<i4", (4,)),
])
assert PAIR_RECORD_DTYPE.itemsize == 128


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("multi_spatialmllm_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _on_device(fn):
    """Run a method with ``self.device`` as the CURRENT CUDA device: libmsp launches on the current device and on the
    stream handed to it, so a scene that lives on cuda:1 must not be driven while cuda:0 is current (multi-GPU processes,
    torchrun ranks that did not call ``torch.cuda.set_device``)."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *a, **kw):
        with torch.cuda.device(self.device):
            return fn(self, *a, **kw)
    return wrapper


def _ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def mask_stride_words(n: int) -> int:
    return ((n + 31) // 32 + 3) // 4 * 4


def _to_device(a, dtype, device, non_blocking=True):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=dtype, non_blocking=non_blocking).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dtype, non_blocking=non_blocking)


class DeviceScene:
    """One scene resident in HBM: points, poses (prepared on the device), depth maps, masks."""

    def __init__(self, points, K, c2w, axis_align=None, depth=None, instance_ids=None, num_objects=0,
                 image_hw=(968, 1296), depth_hw=(480, 640), depth_scale=0.001, device="cuda:0", reorder=None):
        """``reorder="morton"`` stores the points in Z-order (30-bit Morton keys, sorted once at upload) so that the
        32-point groups K1 culls are spatially compact even when the input order is arbitrary (2.7x faster on uniformly
        random clouds, profiles/order_probe.py).  Mask bits are then in the INTERNAL order: ``mask_bool`` un-permutes them,
        ``annotate_points`` / ``pair_annotate`` translate point ids both ways, counts and extents are order-invariant;
        selecting the k-th common point by its original index (sel_mode 0) is not available on a reordered scene."""
        _require_cuda()
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        pts = points
        if not isinstance(pts, torch.Tensor):
            pts = torch.from_numpy(np.ascontiguousarray(pts))
        if pts.dtype == torch.float64:
            assert pts.ndim == 2 and pts.shape[1] == 3, "float64 points must be (N,3)"
            self.pts_dtype = _lib.MSP_PTS_F64X3
        elif pts.dtype == torch.float32 and pts.shape[1] == 4:
            self.pts_dtype = _lib.MSP_PTS_F32X4
        elif pts.dtype == torch.float32 and pts.shape[1] == 3:
            self.pts_dtype = _lib.MSP_PTS_F32X3
        else:
            raise ValueError(f"unsupported point layout {tuple(pts.shape)} {pts.dtype}")
        self.pts = pts.to(self.device, non_blocking=True).contiguous()
        self.n = int(self.pts.shape[0])
        self.perm = self.inv_perm = self._perm_i32 = None
        self._zb = None
        if reorder == "morton" and self.n > 1:
            xyz = self.pts[:, :3].to(torch.float64)
            finite = torch.isfinite(xyz).all(1, keepdim=True)
            safe = torch.where(finite, xyz, torch.zeros_like(xyz))
            lo, hi = safe.min(0).values, safe.max(0).values
            q = ((safe - lo) / (hi - lo).clamp_min(1e-300) * 1023.0).to(torch.int64).clamp_(0, 1023)
            key = torch.zeros(self.n, dtype=torch.int64, device=self.device)
            for b in range(10):
                for a in range(3):
                    key |= ((q[:, a] >> b) & 1) << (3 * b + a)
            self.perm = torch.sort(key, stable=True).indices            # internal position -> original index
            self.inv_perm = torch.empty_like(self.perm)
            self.inv_perm[self.perm] = torch.arange(self.n, device=self.device)
            self.pts = self.pts[self.perm].contiguous()
        elif reorder not in (None, "morton"):
            raise ValueError(f"unknown reorder mode {reorder!r}")
        self.K_host = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(16))
        self.cam = MspCamera(int(image_hw[0]), int(image_hw[1]), int(depth_hw[0]), int(depth_hw[1]), float(depth_scale))
        self.image_hw, self.depth_hw = tuple(image_hw), tuple(depth_hw)
        self.stride = mask_stride_words(self.n)
        if instance_ids is None:
            self.instance_ids = None
        elif isinstance(instance_ids, torch.Tensor):
            self.instance_ids = (instance_ids if instance_ids.dtype == torch.uint16 else instance_ids.to(torch.int32).to(torch.uint16)
                                 ).to(self.device, non_blocking=True).contiguous()
        else:
            self.instance_ids = _to_device(np.asarray(instance_ids).astype(np.uint16), torch.uint16, self.device)
        if self.instance_ids is not None and self.perm is not None:
            self.instance_ids = self.instance_ids.view(torch.int16)[self.perm].view(torch.uint16).contiguous()
        self.num_objects = int(num_objects)
        self.depth = None
        self.mask = None
        self.set_poses(c2w, axis_align)
        if depth is not None:
            self.set_depth(depth)

    # ---- poses ---------------------------------------------------------------------------------
    @_on_device
    def set_poses(self, c2w, axis_align=None):
        c2w_d = _to_device(np.asarray(c2w, dtype=np.float64).reshape(-1, 16) if not isinstance(c2w, torch.Tensor) else c2w.reshape(-1, 16),
                           torch.float64, self.device)
        self.F = int(c2w_d.shape[0])
        A = None if axis_align is None else _to_device(np.asarray(axis_align, dtype=np.float64).reshape(16), torch.float64, self.device)
        self.c2w = c2w_d
        self.E = torch.empty((self.F, 16), dtype=torch.float64, device=self.device)
        self.w2c = torch.empty((self.F, 16), dtype=torch.float64, device=self.device)
        self.yawpitch = torch.empty((self.F, 2), dtype=torch.float64, device=self.device)
        self.valid = torch.empty((self.F,), dtype=torch.int32, device=self.device)
        check(lib.msp_pose_prepare(_ptr(c2w_d), _ptr(A), self.F, _ptr(self.E), _ptr(self.w2c), _ptr(self.yawpitch),
                                   _ptr(self.valid), _stream()), "msp_pose_prepare")
        self.mask = None

    @_on_device
    def set_depth(self, depth):
        d = depth if isinstance(depth, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(depth))
        assert d.dtype == torch.uint16 and tuple(d.shape) == (self.F, *self.depth_hw), (d.dtype, d.shape)
        self.depth = d.to(self.device, non_blocking=True).contiguous()

    # ---- z-buffer (not in the reference) -------------------------------------------------------
    @_on_device
    def render_zbuffer(self) -> torch.Tensor:
        zb = torch.empty((self.F, *self.depth_hw), dtype=torch.float32, device=self.device)
        check(lib.msp_zbuffer_render(_ptr(self.pts), self.n, self.pts_dtype, self.K_host.ctypes.data, _ptr(self.w2c),
                                     self.F, C.byref(self.cam), _ptr(zb), _stream()), "msp_zbuffer_render")
        return zb

    @_on_device
    def render_depth(self, margin_mm: int = 1, chunk: int = 256, keep_buffers: bool = False) -> torch.Tensor:
        """Synthetic sensor depth for the scene's own frames (chunked so the fp32 buffer stays small)."""
        out = self.depth if (self.depth is not None and tuple(self.depth.shape) == (self.F, *self.depth_hw)) else \
            torch.empty((self.F, *self.depth_hw), dtype=torch.uint16, device=self.device)
        px = self.depth_hw[0] * self.depth_hw[1]
        nz = min(chunk, self.F)
        zbuf = self._zb if (self._zb is not None and self._zb.shape[0] >= nz) else \
            torch.empty((nz, *self.depth_hw), dtype=torch.float32, device=self.device)
        self._zb = zbuf if keep_buffers else None           # the float32 splat buffer (1.2 MB per frame) is only kept on request
        for f0 in range(0, self.F, chunk):
            f1 = min(self.F, f0 + chunk)
            zb = zbuf[:f1 - f0]
            check(lib.msp_zbuffer_render(_ptr(self.pts), self.n, self.pts_dtype, self.K_host.ctypes.data,
                                         self.w2c[f0:].data_ptr(), f1 - f0, C.byref(self.cam), _ptr(zb), _stream()),
                  "msp_zbuffer_render")
            check(lib.msp_depth_to_u16(_ptr(zb), (f1 - f0) * px, int(margin_mm), self.cam.depth_scale,
                                       out[f0:].data_ptr(), _stream()), "msp_depth_to_u16")
        self.depth = out
        return out

    # ---- K1 ------------------------------------------------------------------------------------
    @_on_device
    def project_visibility(self, materialize: bool = False):
        assert self.depth is not None, "no depth maps: call set_depth() or render_depth() first"
        if self.mask is None:
            self.mask = torch.zeros((self.F, self.stride), dtype=torch.int32, device=self.device)
        xy = z = None
        if materialize:
            xy = torch.empty((self.F, self.n, 2), dtype=torch.float64, device=self.device)
            z = torch.empty((self.F, self.n), dtype=torch.float64, device=self.device)
        check(lib.msp_project_visibility(_ptr(self.pts), self.n, self.pts_dtype, self.K_host.ctypes.data, _ptr(self.w2c),
                                         self.F, _ptr(self.depth), C.byref(self.cam), _ptr(self.mask), self.stride,
                                         _ptr(xy), _ptr(z), _stream()), "msp_project_visibility")
        return (self.mask, xy, z) if materialize else self.mask

    @_on_device
    def project_visibility_objects(self, out_counts: torch.Tensor | None = None):
        """Masks and per-(frame, object) visible counts in one pass (msp_project_visibility_objects)."""
        assert self.depth is not None and self.instance_ids is not None and self.num_objects > 0
        if self.mask is None:
            self.mask = torch.zeros((self.F, self.stride), dtype=torch.int32, device=self.device)
        counts = out_counts if out_counts is not None else torch.empty((self.F, self.num_objects), dtype=torch.int32, device=self.device)
        assert counts.dtype == torch.int32 and counts.is_contiguous() and tuple(counts.shape) == (self.F, self.num_objects)
        check(lib.msp_project_visibility_objects(_ptr(self.pts), self.n, self.pts_dtype, self.K_host.ctypes.data,
                                                 _ptr(self.w2c), self.F, _ptr(self.depth), C.byref(self.cam),
                                                 _ptr(self.mask), self.stride, _ptr(self.instance_ids), self.num_objects,
                                                 _ptr(counts), _stream()), "msp_project_visibility_objects")
        return self.mask, counts

    def mask_bool(self) -> np.ndarray:
        """(F,N) bool on the host — the reference's in_bounds mask per frame (IH:386)."""
        w = self.mask.cpu().numpy().view(np.uint32)
        bits = np.unpackbits(w.view(np.uint8), axis=1, bitorder="little")[:, :self.n].astype(bool)
        if self.inv_perm is not None:                     # internal (Morton) order -> the caller's point order
            bits = bits[:, self.inv_perm.cpu().numpy()]
        return bits

    # ---- pair stage ----------------------------------------------------------------------------
    @_on_device
    def pair_overlap(self, pairs):
        pairs_d = _to_device(pairs if isinstance(pairs, torch.Tensor) else np.asarray(pairs, dtype=np.int32).reshape(-1, 2),
                             torch.int32, self.device)
        P = int(pairs_d.shape[0])
        inter = torch.empty((P,), dtype=torch.int32, device=self.device)
        uni = torch.empty((P,), dtype=torch.int32, device=self.device)
        ov = torch.empty((P,), dtype=torch.float64, device=self.device)
        check(lib.msp_pair_overlap(_ptr(self.mask), self.n, self.stride, _ptr(pairs_d), P, _ptr(inter), _ptr(uni),
                                   _ptr(ov), _stream()), "msp_pair_overlap")
        return inter, uni, ov

    @_on_device
    def pair_annotate(self, pairs, sel, sel_mode: int = 0, out: torch.Tensor | None = None) -> torch.Tensor:
        pairs_d = _to_device(pairs if isinstance(pairs, torch.Tensor) else np.asarray(pairs, dtype=np.int32).reshape(-1, 2),
                             torch.int32, self.device)
        P = int(pairs_d.shape[0])
        sel_d = sel if isinstance(sel, torch.Tensor) else torch.from_numpy(np.asarray(sel).astype(np.uint32).view(np.int32))
        sel_d = sel_d.to(self.device, non_blocking=True).contiguous()
        if self.perm is not None and int(sel_mode) == 0:
            raise ValueError("sel_mode 0 (k-th common point by original index) is not available on a reordered scene")
        rec = out if out is not None else torch.empty((P, 128), dtype=torch.uint8, device=self.device)
        check(lib.msp_pair_annotate(_ptr(self.mask), self.n, self.stride, _ptr(pairs_d), P, _ptr(sel_d), int(sel_mode),
                                    _ptr(self.pts), self.pts_dtype, self.K_host.ctypes.data, _ptr(self.E), _ptr(self.w2c),
                                    _ptr(self.yawpitch), _ptr(self.depth), C.byref(self.cam), _ptr(rec), _stream()),
              "msp_pair_annotate")
        if self.perm is not None:                          # point_id (int32 at byte 64) back to the caller's numbering
            pid = rec.view(torch.int32).view(P, 32)[:, 16]
            if self._perm_i32 is None:
                self._perm_i32 = self.perm.to(torch.int32)
            # no boolean indexing: that would synchronise the stream (nonzero) in the middle of the step
            pid.copy_(torch.where(pid >= 0, self._perm_i32[pid.clamp_min(0).long()], pid))
        return rec

    @staticmethod
    def records_to_numpy(rec: torch.Tensor) -> np.ndarray:
        return rec.cpu().numpy().view(PAIR_RECORD_DTYPE).reshape(-1)

    @_on_device
    def annotate_points(self, point_ids, frame_ids, use_depth: bool = True):
        pid = _to_device(np.asarray(point_ids, dtype=np.int32), torch.int32, self.device)
        if self.inv_perm is not None:
            pid = self.inv_perm[pid.long()].to(torch.int32)
        fid = _to_device(np.asarray(frame_ids, dtype=np.int32), torch.int32, self.device)
        M = int(pid.shape[0])
        uvz = torch.empty((M, 3), dtype=torch.float64, device=self.device)
        pm = torch.empty((M, 2), dtype=torch.int32, device=self.device)
        dmm = torch.empty((M,), dtype=torch.int32, device=self.device)
        pix = torch.empty((M, 2), dtype=torch.int32, device=self.device)
        vis = torch.empty((M,), dtype=torch.uint8, device=self.device)
        check(lib.msp_annotate_points(_ptr(self.pts), self.n, self.pts_dtype, _ptr(pid), _ptr(fid), M,
                                      self.K_host.ctypes.data, _ptr(self.w2c), _ptr(self.depth) if use_depth else None,
                                      C.byref(self.cam), _ptr(uvz), _ptr(pm), _ptr(dmm), _ptr(pix), _ptr(vis), _stream()),
              "msp_annotate_points")
        return {"uvz": uvz, "permille": pm, "depth_mm": dmm, "pix": pix, "visible": vis}

    @_on_device
    def object_counts(self) -> torch.Tensor:
        assert self.instance_ids is not None and self.num_objects > 0
        counts = torch.empty((self.F, self.num_objects), dtype=torch.int32, device=self.device)
        check(lib.msp_object_counts(_ptr(self.mask), self.n, self.stride, self.F, _ptr(self.instance_ids),
                                    self.num_objects, _ptr(counts), _stream()), "msp_object_counts")
        return counts


    @_on_device
    def object_extents(self) -> torch.Tensor:
        """(F, n_objects, 6) float64: min xyz / max xyz of each object's visible points per frame (SOCF:57-65)."""
        assert self.instance_ids is not None and self.num_objects > 0 and self.mask is not None
        ext = torch.empty((self.F, self.num_objects, 6), dtype=torch.float64, device=self.device)
        check(lib.msp_object_extents(_ptr(self.mask), self.n, self.stride, self.F, _ptr(self.pts), self.pts_dtype,
                                     _ptr(self.instance_ids), self.num_objects, _ptr(ext), _stream()), "msp_object_extents")
        return ext


# ------------------------------------------------------------------------------------------------
# per-scene workers with the reference's return shapes
# ------------------------------------------------------------------------------------------------

def process_scene_visibility(scene: DeviceScene, image_ids):
    """make_visibility_info.process_scene (MVI:75-125): image_to_points / point_to_images.

    ``image_ids`` are the ids of the scene's frames in row order; frames with an
    invalid pose are skipped exactly like get_all_extrinsic_valid_image_ids (IH:184-189).
    """
    scene.project_visibility()
    valid = scene.valid.cpu().numpy().astype(bool)
    m = scene.mask_bool()
    image_to_points = {}
    for f, image_id in enumerate(image_ids):
        if valid[f]:
            image_to_points[image_id] = np.flatnonzero(m[f]).tolist()
    ids = np.asarray(image_ids)
    mv, idv = m[valid], ids[valid]
    order = np.argsort(idv, kind="stable")          # sorted(list(set)) of string ids (MVI:116-118)
    mv, idv = mv[order], idv[order]
    # all (point, frame) incidences at once, grouped by point; points never seen keep an empty list (MVI:116-118)
    p_idx, f_idx = np.nonzero(mv.T)
    cuts = np.cumsum(np.bincount(p_idx, minlength=scene.n))[:-1]
    point_to_images = dict(enumerate(a.tolist() for a in np.split(idv[f_idx], cuts)))
    return {"image_to_points": image_to_points, "point_to_images": point_to_images}


def process_scene_relations(scene: DeviceScene, image_ids):
    """calculate_frames_relations.process_scene (CFR:139-197): all i<j pairs of valid frames."""
    scene.project_visibility()
    valid = np.flatnonzero(scene.valid.cpu().numpy())
    pairs = np.array(list(combinations(valid.tolist(), 2)), dtype=np.int32).reshape(-1, 2)
    if len(pairs) == 0:
        return {}
    rec = scene.records_to_numpy(scene.pair_annotate(pairs, np.zeros(len(pairs), np.uint32), sel_mode=0))
    out = {}
    for (i, j), r in zip(pairs.tolist(), rec):
        out[(image_ids[i], image_ids[j])] = {
            "overlap": float(r["overlap"]), "distance": float(r["distance"]),
            "yaw": float(r["dyaw"]), "pitch": float(r["dpitch"]),
        }
    return out


class DeviceVideo:
    """One TAPVid-3D clip resident in HBM (single_object_movement_engine_coord.py:441-451)."""

    def __init__(self, tracks_xyz, w2c, fxfycxcy, image_hw, device="cuda:0"):
        _require_cuda()
        self.device = torch.device(device)
        self.tracks = _to_device(np.asarray(tracks_xyz, dtype=np.float32), torch.float32, self.device)
        self.T, self.N = int(self.tracks.shape[0]), int(self.tracks.shape[1])
        self.w2c = _to_device(np.asarray(w2c, dtype=np.float64).reshape(self.T, 16), torch.float64, self.device)
        self.intr = _to_device(np.asarray(fxfycxcy, dtype=np.float64).reshape(4), torch.float64, self.device)
        self.image_hw = (int(image_hw[0]), int(image_hw[1]))
        self.world = self.uvn = self.permille = self.valid = None

    @_on_device
    def project(self):
        """World transform + pinhole projection of every (frame, track) (OM-coord:446-451, 293-315)."""
        T, N = self.T, self.N
        self.world = torch.empty((T, N, 3), dtype=torch.float64, device=self.device)
        self.uvn = torch.empty((T, N, 2), dtype=torch.float64, device=self.device)
        self.permille = torch.empty((T, N, 2), dtype=torch.int32, device=self.device)
        self.valid = torch.empty((T, N), dtype=torch.uint8, device=self.device)
        check(lib.msp_track_project(_ptr(self.tracks), _ptr(self.w2c), _ptr(self.intr), T, N, self.image_hw[0],
                                    self.image_hw[1], _ptr(self.world), _ptr(self.uvn), _ptr(self.permille),
                                    _ptr(self.valid), _stream()), "msp_track_project")
        return self

    @_on_device
    def pair_displacement(self, triples):
        """(M,3) int32 (frame1, frame2, track) -> (M,5) float64: dist, d_cam1 xyz, camera baseline (OM-coord:329-356)."""
        if self.world is None:
            self.project()
        tri = _to_device(np.asarray(triples, dtype=np.int32).reshape(-1, 3), torch.int32, self.device)
        out = torch.empty((int(tri.shape[0]), 5), dtype=torch.float64, device=self.device)
        check(lib.msp_track_pair_displacement(_ptr(self.world), _ptr(self.w2c), self.T, self.N, _ptr(tri),
                                              int(tri.shape[0]), _ptr(out), _stream()), "msp_track_pair_displacement")
        return out

    @_on_device
    def rigid_body_groups(self, threshold: float = 0.1, smoothing_factor: float = 0.01, min_size: int = 5):
        """rigid_body_segmentation + filter_large_groups (OM-coord:49-105): the O(T N^2) accumulation on the GPU,
        average-linkage clustering of the N x N result on the host (scipy, like the reference)."""
        from scipy.cluster.hierarchy import fcluster, linkage
        n = self.N
        loss = torch.empty((n * (n - 1) // 2,), dtype=torch.float64, device=self.device)
        check(lib.msp_track_cumulative_loss(_ptr(self.tracks), self.T, n, float(smoothing_factor), _ptr(loss), _stream()),
              "msp_track_cumulative_loss")
        labels = fcluster(linkage(loss.cpu().numpy(), method="average"), threshold, criterion="distance")
        groups = [np.where(labels == i)[0].tolist() for i in range(1, int(labels.max()) + 1)]
        return [g for g in groups if len(g) > min_size], loss
