"""Public batch API: fully annotate frame pairs of scenes that live in HOST memory.

One call per scene does what the reference spreads over three scripts
(make_visibility_info.py -> calculate_frames_relations.py ->
visual_correspondence_qa_engine_*.py / compute_object_visibility.py): upload the
scene, build both frames' visibility masks, the pair relations, one sampled
common point annotated in both frames and the per-object visible counts, and
bring the records back.  Scenes are independent, so the
annotator round-robins them over a few CUDA streams to overlap the PCIe
upload of one scene with the kernels of another; there is no inter-GPU step.

Every stream owns a SLOT of persistent buffers (device inputs, masks, records,
counts, and the pinned host buffers the results land in) that only ever grow:
after the first scenes a ``submit`` performs no ``cudaMalloc`` / ``cudaHostAlloc``
and no memset — the depth maps (1.2 MB per pair) cross PCIe exactly once and
everything else is stream-ordered behind them.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from .engine import PAIR_RECORD_DTYPE, DeviceScene, mask_stride_words


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


class _Slot:
    """Grow-only buffers of one stream."""

    def __init__(self, device, stream):
        self.device, self.stream = device, stream
        self.dev = {}      # name -> flat device tensor
        self.host = {}     # name -> flat pinned host tensor
        self.busy = None   # event recorded after the slot's last D2H copy
        self.handle = None # the handle whose results still live in this slot's pinned buffers

    def dbuf(self, name, numel, dtype):
        t = self.dev.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            t = torch.empty((max(numel, 1),), dtype=dtype, device=self.device)
            self.dev[name] = t
        return t[:numel]

    def hbuf(self, name, numel, dtype):
        t = self.host.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            t = torch.empty((max(numel, 1),), dtype=dtype, pin_memory=True)
            self.host[name] = t
        return t[:numel]


class PairAnnotator:
    def __init__(self, device="cuda:0", n_streams: int = 2):
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(self.device):
            self.slots = [_Slot(self.device, torch.cuda.Stream(device=self.device)) for _ in range(n_streams)]
        self._next = 0
        self._pending = []

    def submit(self, scene: HostScene, pairs: torch.Tensor, sel: torch.Tensor, sel_mode: int = 1):
        """Enqueue one scene; returns a handle whose ``records`` / ``object_counts`` are filled after ``drain()``."""
        slot = self.slots[self._next]
        self._next = (self._next + 1) % len(self.slots)
        if slot.busy is not None:
            # the slot's buffers (and pinned result buffers) are about to be overwritten: copy its previous scene's results out
            self._finish(slot.handle)
        P = int(pairs.shape[0])
        F = int(scene.c2w.shape[0])
        n = int(scene.points.shape[0])
        with torch.cuda.device(self.device), torch.cuda.stream(slot.stream):
            pts = slot.dbuf("pts", scene.points.numel(), scene.points.dtype).view(scene.points.shape)
            pts.copy_(scene.points, non_blocking=True)
            c2w = slot.dbuf("c2w", F * 16, torch.float64).view(F, 16)
            c2w.copy_(scene.c2w.reshape(F, 16), non_blocking=True)
            depth = slot.dbuf("depth", scene.depth.numel(), torch.uint16).view(scene.depth.shape)
            depth.copy_(scene.depth, non_blocking=True)
            inst = None
            if scene.instance_ids is not None and scene.num_objects > 0:
                inst = slot.dbuf("inst", n, torch.uint16)
                inst.copy_(scene.instance_ids, non_blocking=True)
            pairs_d = slot.dbuf("pairs", P * 2, torch.int32).view(P, 2)
            pairs_d.copy_(pairs, non_blocking=True)
            sel_d = slot.dbuf("sel", P, torch.int32)
            sel_d.copy_(sel.view(torch.int32) if sel.dtype != torch.int32 else sel, non_blocking=True)
            sc = DeviceScene(pts, scene.K, c2w, axis_align=scene.axis_align, depth=depth, instance_ids=inst,
                             num_objects=scene.num_objects, image_hw=scene.image_hw, depth_hw=scene.depth_hw,
                             depth_scale=scene.depth_scale, device=self.device)
            stride = mask_stride_words(n)
            sc.mask = slot.dbuf("mask", F * stride, torch.int32).view(F, stride)   # every word is written by K1: no memset
            cnt = None
            if inst is not None:
                cnt = slot.dbuf("counts", F * scene.num_objects, torch.int32).view(F, scene.num_objects)
                sc.project_visibility_objects(out_counts=cnt)
            else:
                sc.project_visibility()
            rec = slot.dbuf("rec", P * 128, torch.uint8).view(P, 128)
            sc.pair_annotate(pairs_d, sel_d, sel_mode, out=rec)
            rec_h = slot.hbuf("rec", P * 128, torch.uint8).view(P, 128)
            rec_h.copy_(rec, non_blocking=True)
            cnt_h = None
            if cnt is not None:
                cnt_h = slot.hbuf("counts", cnt.numel(), torch.int32).view(tuple(cnt.shape))
                cnt_h.copy_(cnt, non_blocking=True)
            slot.busy = torch.cuda.Event()
            slot.busy.record(slot.stream)
        h = {"records": rec_h, "object_counts": cnt_h, "slot": slot, "h2d_bytes": scene.h2d_bytes() + P * 12,
             "d2h_bytes": P * 128 + (0 if cnt_h is None else cnt_h.numel() * 4)}
        slot.handle = h
        self._pending.append(h)
        return h

    @staticmethod
    def _finish(h):
        slot = h.pop("slot", None)
        if slot is None:
            return                                   # already finished
        slot.busy.synchronize()
        slot.busy, slot.handle = None, None
        h["records"] = h["records"].numpy().view(PAIR_RECORD_DTYPE).reshape(-1).copy()
        if h["object_counts"] is not None:
            h["object_counts"] = h["object_counts"].numpy().copy()

    def drain(self):
        """Wait for every submitted scene and hand the results out as NumPy arrays (copied out of the pinned slots)."""
        done, self._pending = self._pending, []
        for h in done:
            self._finish(h)
        return done
