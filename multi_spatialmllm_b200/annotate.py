"""Dot / letter annotation of colour frames on the device (the `*_dot` streams).

The reference decodes the JPEG, calls ``cv2.circle`` / ``cv2.putText`` and re-encodes per sample (DE-dot:216-229,
DC-dot:333-346, VC-dot:357-397, OM-dot:330-340).  Here a batch of BGR uint8 frames lives in HBM and one
``msp_draw_marks`` launch paints every mark of the batch; the raster is OpenCV's (pixel-exact circles, glyph masks
rendered once by cv2 and shipped in ``data/glyphs_hershey_simplex_s1_t2.npz``).  JPEG decode/encode stay outside
(I/O; nvJPEG or cv2 at the caller's choice).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from ._lib import check, lib
from .engine import _require_cuda, _stream

MARK_DTYPE = np.dtype([("image", "<i4"), ("x", "<i4"), ("y", "<i4"), ("radius", "<i4"), ("glyph", "<i4"),
                       ("org_x", "<i4"), ("org_y", "<i4"), ("bgr", "u1", (3,)), ("reserved", "u1")])
assert MARK_DTYPE.itemsize == 32

_ATLAS_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "glyphs_hershey_simplex_s1_t2.npz")


class GlyphAtlas:
    """FONT_HERSHEY_SIMPLEX, scale 1.0, thickness 2 masks for A-Z0-9 (what DC-dot:338-341 / VC-dot:391 draw)."""

    def __init__(self, path: str = _ATLAS_PATH):
        d = np.load(path)
        self.chars = bytes(d["chars"]).decode()
        self.rows = np.ascontiguousarray(d["rows"], dtype=np.uint32)            # (G, 32)
        self.meta = np.ascontiguousarray(d["meta"], dtype=np.int32)             # (G, 4): off_x, off_y, w, h
        self._dev = {}

    def index(self, ch: str) -> int:
        i = self.chars.find(ch)
        if i < 0:
            raise KeyError(f"glyph {ch!r} is not in the atlas ({self.chars})")
        return i

    def mask(self, ch: str):
        """(bool mask (h, w), (off_x, off_y)) — the host view the oracle and the tests use."""
        i = self.index(ch)
        ox, oy, w, h = (int(v) for v in self.meta[i])
        bits = (self.rows[i, :h, None] >> np.arange(w, dtype=np.uint32)[None, :]) & 1
        return bits.astype(bool), (ox, oy)

    def device(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = (torch.from_numpy(self.rows.view(np.int32)).to(device),
                              torch.from_numpy(self.meta).to(device))
        return self._dev[key]


_default_atlas = None


def default_atlas() -> GlyphAtlas:
    global _default_atlas
    if _default_atlas is None:
        _default_atlas = GlyphAtlas()
    return _default_atlas


def make_marks(n: int) -> np.ndarray:
    m = np.zeros(n, MARK_DTYPE)
    m["radius"] = -1
    m["glyph"] = -1
    return m


def draw_marks(images: torch.Tensor, marks: np.ndarray, atlas: GlyphAtlas | None = None) -> torch.Tensor:
    """Paint ``marks`` (MARK_DTYPE, runs of equal ``image`` contiguous, array order = paint order) into the
    (B, H, W, 3) uint8 BGR device batch, in place."""
    _require_cuda()
    if not (isinstance(images, torch.Tensor) and images.is_cuda and images.dtype == torch.uint8 and images.dim() == 4
            and images.shape[3] == 3 and images.is_contiguous()):
        raise ValueError("images must be a contiguous (B, H, W, 3) uint8 CUDA tensor")
    marks = np.ascontiguousarray(marks, dtype=MARK_DTYPE)
    if len(marks) == 0:
        return images
    img = marks["image"]
    change = np.flatnonzero(img[1:] != img[:-1])
    if len(np.unique(img[np.r_[0, change + 1]])) != len(change) + 1:
        raise ValueError("marks of one image must be contiguous")
    atlas = atlas or default_atlas()
    with torch.cuda.device(images.device):
        rows, meta = atlas.device(images.device)
        dev_marks = torch.from_numpy(marks.view(np.uint8).reshape(-1, 32)).to(images.device)
        B, H, W, _ = images.shape
        check(lib.msp_draw_marks(images.data_ptr(), B, H, W, dev_marks.data_ptr(), len(marks), rows.data_ptr(),
                                 meta.data_ptr(), len(atlas.chars), _stream()), "msp_draw_marks")
    return images


def distinct_colors(n: int, rng, max_retries: int = 10):
    """`generate_distinct_colors` (DE-dot:22-33): random colours pairwise more than 300 apart in L1, padded from a fixed list."""
    colors, retries = [], 0
    while len(colors) < n and retries < max_retries:
        c = tuple(int(v) for v in rng.integers(0, 256, 3))
        if all(sum(abs(a - b) for a, b in zip(c, e)) > 300 for e in colors):
            colors.append(c)
        retries += 1
    if len(colors) < n:
        fixed = [(255, 0, 0), (0, 255, 0), (0, 0, 255), (0, 0, 0), (255, 255, 255)]
        colors += [fixed[i] for i in rng.permutation(len(fixed))[:n - len(colors)]]
    return colors


# ---- JPEG glue (nvJPEG through the C ABI; cv2.imread / cv2.imwrite of the reference's dot engines) ------------------
def jpeg_decode(data: bytes, device="cuda:0") -> torch.Tensor:
    """JPEG bytes -> (H, W, 3) BGR uint8 device tensor (what ``cv2.imread`` returns, decoded on the GPU)."""
    import ctypes as C
    _require_cuda()
    buf = np.frombuffer(data, np.uint8)
    h, w = C.c_int32(), C.c_int32()
    check(lib.msp_jpeg_info(buf.ctypes.data, len(buf), C.byref(h), C.byref(w)), "msp_jpeg_info")
    with torch.cuda.device(device):
        out = torch.empty((h.value, w.value, 3), dtype=torch.uint8, device=device)
        check(lib.msp_jpeg_decode_bgr(buf.ctypes.data, len(buf), out.data_ptr(), h.value, w.value, _stream()),
              "msp_jpeg_decode_bgr")
        torch.cuda.current_stream().synchronize()          # `buf` must outlive the decode
    return out


def jpeg_encode(image: torch.Tensor, quality: int = 95) -> bytes:
    """(H, W, 3) BGR uint8 device tensor -> JPEG bytes (``cv2.imwrite``'s defaults: quality 95, 4:2:0)."""
    import ctypes as C
    _require_cuda()
    if not (image.is_cuda and image.dtype == torch.uint8 and image.dim() == 3 and image.shape[2] == 3 and image.is_contiguous()):
        raise ValueError("image must be a contiguous (H, W, 3) uint8 CUDA tensor")
    H, W, _ = image.shape
    n = C.c_size_t(0)
    with torch.cuda.device(image.device):
        out = np.empty(H * W * 3 + 65536, np.uint8)
        check(lib.msp_jpeg_encode_bgr(image.data_ptr(), H, W, int(quality), out.ctypes.data, out.size, C.byref(n), _stream()),
              "msp_jpeg_encode_bgr")
    return out[:n.value].tobytes()


_pool, _pool_size, _tls = None, 0, None


def _worker_pool(workers: int):
    """One long-lived pool: nvJPEG keeps a handle per host thread (msp_jpeg.inl), created on the thread's first call and
    worth ~0.1 s — short-lived threads would pay it again and again."""
    global _pool, _pool_size, _tls
    import threading
    from concurrent.futures import ThreadPoolExecutor
    if _pool is None or _pool_size < workers:
        if _pool is not None:
            _pool.shutdown(wait=True)
        _pool, _pool_size, _tls = ThreadPoolExecutor(workers, thread_name_prefix="msp-jpeg"), workers, threading.local()
    return _pool


def write_annotated(jobs, read_source, output_dir: str, device="cuda:0", workers: int = 4, batch: int = 16) -> int:
    """The ``cv2.imread -> cv2.circle / cv2.putText -> cv2.imwrite`` tail of the dot engines for a list of jobs
    ``(source key, output path relative to output_dir, [mark tuples without the image index])``.

    Every distinct source frame is decoded ONCE (nvJPEG), replicated on the device for the jobs that annotate it, painted
    by one ``msp_draw_marks`` launch per batch and encoded per output (nvJPEG; the reference decodes and encodes per
    sample).  ``workers`` host threads, each with its own CUDA stream and nvJPEG state, share the sources between them:
    the Huffman stages of nvJPEG's decode and encode run on the host and release the GIL."""
    _require_cuda()
    if not jobs:
        return 0
    by_source = {}
    for key, path, marks in jobs:
        by_source.setdefault(key, []).append((path, marks))
    sources = list(by_source.items())
    for path in {os.path.dirname(os.path.join(output_dir, p)) for _, p, _ in jobs}:
        os.makedirs(path, exist_ok=True)

    def work(share, own_stream=True):
        with torch.cuda.device(device):
            if own_stream:
                if getattr(_tls, "stream", None) is None or _tls.stream.device != torch.device(device):
                    _tls.stream = torch.cuda.Stream(device)
                stream = _tls.stream
            else:
                stream = torch.cuda.current_stream()
            with torch.cuda.stream(stream):
                for key, outs in share:
                    src = jpeg_decode(read_source(key), device)
                    for k0 in range(0, len(outs), batch):
                        chunk = outs[k0:k0 + batch]
                        stack = src.unsqueeze(0).repeat(len(chunk), 1, 1, 1)
                        rows = [(n, *m) for n, (_, marks) in enumerate(chunk) for m in marks]
                        if rows:
                            draw_marks(stack, np.array(rows, dtype=MARK_DTYPE))
                        for n, (path, _) in enumerate(chunk):
                            with open(os.path.join(output_dir, path), "wb") as f:
                                f.write(jpeg_encode(stack[n]))
        return len(share)

    workers = max(1, min(workers, len(sources)))
    if workers == 1:
        work(sources, own_stream=False)
    else:
        list(_worker_pool(workers).map(work, [sources[w::workers] for w in range(workers)]))
    return len(jobs)
