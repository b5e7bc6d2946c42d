"""Public batch API: fully annotate frame pairs of scenes that live in HOST memory.

One call per scene does what the reference spreads over three scripts
(make_visibility_info.py -> calculate_frames_relations.py ->
visual_correspondence_qa_engine_*.py / compute_object_visibility.py): upload the
scene, build both frames' visibility masks, the pair relations, one sampled
common point annotated in both frames and the per-object visible counts, and
bring the records back.  Scenes are independent, so the
annotator round-robins them over a few CUDA streams to overlap the PCIe
upload of one scene with the kernels of another; there is no inter-GPU step.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from .engine import PAIR_RECORD_DTYPE, DeviceScene


@dataclass
class HostScene:
    """Inputs of one scene in (preferably pinned) host memory."""
    points: torch.Tensor          # (N,3) float32 | (N,4) float32 | (N,3) float64
    K: np.ndarray                 # (4,4) float64
    axis_align: np.ndarray | None  # (4,4) float64
    c2w: torch.Tensor             # (F,4,4) float64
    depth: torch.Tensor           # (F,dh,dw) uint16
    instance_ids: torch.Tensor | None = None   # (N,) uint16
    num_objects: int = 0
    image_hw: tuple = (968, 1296)
    depth_hw: tuple = (480, 640)
    depth_scale: float = 0.001

    def h2d_bytes(self) -> int:
        b = self.points.numel() * self.points.element_size() + self.c2w.numel() * 8 + self.depth.numel() * 2
        if self.instance_ids is not None:
            b += self.instance_ids.numel() * 2
        return int(b)

    def pin(self) -> "HostScene":
        self.points, self.c2w, self.depth = self.points.pin_memory(), self.c2w.pin_memory(), self.depth.pin_memory()
        if self.instance_ids is not None:
            self.instance_ids = self.instance_ids.pin_memory()
        return self


class PairAnnotator:
    def __init__(self, device="cuda:0", n_streams: int = 2):
        self.device = torch.device(device)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(n_streams)]
        self._slot = 0
        self._pending = []

    def submit(self, scene: HostScene, pairs: torch.Tensor, sel: torch.Tensor, sel_mode: int = 1):
        """Enqueue one scene; returns a handle whose tensors are valid after ``drain()``."""
        st = self.streams[self._slot]
        self._slot = (self._slot + 1) % len(self.streams)
        P = int(pairs.shape[0])
        with torch.cuda.stream(st):
            sc = DeviceScene(scene.points, scene.K, scene.c2w, axis_align=scene.axis_align, depth=scene.depth,
                             instance_ids=scene.instance_ids, num_objects=scene.num_objects,
                             image_hw=scene.image_hw, depth_hw=scene.depth_hw, depth_scale=scene.depth_scale,
                             device=self.device)
            cnt = None
            if scene.instance_ids is not None and scene.num_objects > 0:
                _, cnt = sc.project_visibility_objects()
            else:
                sc.project_visibility()
            rec = sc.pair_annotate(pairs, sel, sel_mode)
            rec_h = torch.empty((P, 128), dtype=torch.uint8, pin_memory=True)
            rec_h.copy_(rec, non_blocking=True)
            cnt_h = None
            if cnt is not None:
                cnt_h = torch.empty(tuple(cnt.shape), dtype=torch.int32, pin_memory=True)
                cnt_h.copy_(cnt, non_blocking=True)
        h = {"records": rec_h, "object_counts": cnt_h, "scene": sc, "h2d_bytes": scene.h2d_bytes() + P * 12,
             "d2h_bytes": P * 128 + (0 if cnt_h is None else cnt_h.numel() * 4)}
        self._pending.append(h)
        return h

    def drain(self):
        for st in self.streams:
            st.synchronize()
        done, self._pending = self._pending, []
        for h in done:
            h["scene"] = None                      # release the device buffers
            h["records"] = h["records"].numpy().view(PAIR_RECORD_DTYPE).reshape(-1)
        return done
