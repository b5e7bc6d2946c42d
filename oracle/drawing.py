"""CPU restatement of the raster operations the reference's ``*_dot`` engines use (TEST INFRASTRUCTURE — see oracle/__init__).

The reference marks points with ``cv2.circle(img, (x, y), r, color, -1)`` and labels them with
``cv2.putText(img, letter, org, cv2.FONT_HERSHEY_SIMPLEX, 1.0, color, 2)``
(DE-dot:171, DC-dot:337-341, VC-dot:363,390-391, OM-dot:337).  OpenCV (4.x, `imgproc/src/drawing.cpp`, not vendored in
/root/reference) rasterises the filled LINE_8 circle with an integer midpoint walk; this file restates that walk and is
pinned against cv2 itself (tests/test_oracle_golden.py::test_drawing_matches_cv2, tests/golden/drawing_small.npz).
Glyphs are not re-derived: the product ships 1-bit masks rendered once by cv2 (multi_spatialmllm_b200/data/make_glyph_atlas.py),
and stamping a mask is the restated operation.
"""
from __future__ import annotations

import numpy as np


def circle_halfwidths(radius: int) -> np.ndarray:
    """hw[k] = half-width of the filled span on rows centre.y +- k (k = 0..radius) of OpenCV's filled circle.

    Midpoint walk (dx, dy) from (radius, 0) while dx >= dy: rows +-dy get the span +-dx, rows +-dx get the span +-dy;
    a later visit to the same row is never narrower for the first kind and the second kind only touches rows > dy."""
    hw = np.full(radius + 1, -1, np.int64)
    err, dx, dy, plus, minus = 0, radius, 0, 1, (radius << 1) - 1
    while dx >= dy:
        hw[dy] = max(hw[dy], dx)
        hw[dx] = max(hw[dx], dy)
        dy += 1
        err += plus
        plus += 2
        if err > 0:
            err -= minus
            dx -= 1
            minus -= 2
    return hw


def draw_filled_circle(img: np.ndarray, x: int, y: int, radius: int, color) -> None:
    """In-place equivalent of ``cv2.circle(img, (x, y), radius, color, -1)`` for uint8 (H, W, 3) images."""
    h, w = img.shape[:2]
    hw = circle_halfwidths(radius)
    for k in range(-radius, radius + 1):
        yy = y + k
        if 0 <= yy < h:
            x0, x1 = max(x - hw[abs(k)], 0), min(x + hw[abs(k)], w - 1)
            if x0 <= x1:
                img[yy, x0:x1 + 1] = color


def stamp_glyph(img: np.ndarray, mask: np.ndarray, offset, org, color) -> None:
    """Paint ``color`` where the glyph mask is set; mask[0, 0] sits at pixel (org.x + offset.x, org.y + offset.y)."""
    h, w = img.shape[:2]
    gx, gy = org[0] + offset[0], org[1] + offset[1]
    ys, xs = np.nonzero(mask)
    py, px = ys + gy, xs + gx
    ok = (py >= 0) & (py < h) & (px >= 0) & (px < w)
    img[py[ok], px[ok]] = color


def draw_marks(img: np.ndarray, marks, atlas=None) -> None:
    """marks: sequence of dicts {x, y, radius, color, [glyph, org]} drawn in order (later marks overwrite earlier)."""
    for m in marks:
        draw_filled_circle(img, m["x"], m["y"], m["radius"], m["color"])
        if m.get("glyph") is not None:
            mask, off = atlas[m["glyph"]]
            stamp_glyph(img, mask, off, m["org"], m["color"])
