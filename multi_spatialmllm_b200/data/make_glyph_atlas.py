"""Render the label glyph atlas the dot-annotation kernel stamps (run once, output committed).

The reference labels marks with ``cv2.putText(img, letter, org, cv2.FONT_HERSHEY_SIMPLEX, 1.0, color, 2)``
(DC-dot:338-341, VC-dot:391).  The rendering is translation invariant for integer origins, so each glyph is stored as a
1-bit mask plus the offset of its top-left corner from ``org``; `msp_draw_marks` paints the mask.  Characters: A-Z, 0-9.

    python multi_spatialmllm_b200/data/make_glyph_atlas.py      # needs cv2; writes glyphs_hershey_simplex_s1_t2.npz
"""
import os

import cv2
import numpy as np

CHARS = "ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789"
CELL_W, CELL_H = 32, 32                       # every glyph fits; rows are packed into one uint32 each


def render(ch, scale=1.0, thickness=2):
    canvas = np.zeros((128, 128), np.uint8)
    org = (48, 80)
    cv2.putText(canvas, ch, org, cv2.FONT_HERSHEY_SIMPLEX, scale, 255, thickness)
    ys, xs = np.nonzero(canvas)
    x0, y0 = xs.min(), ys.min()
    mask = canvas[y0:ys.max() + 1, x0:xs.max() + 1] > 0
    return mask, (int(x0 - org[0]), int(y0 - org[1]))


def main():
    rows = np.zeros((len(CHARS), CELL_H), np.uint32)
    meta = np.zeros((len(CHARS), 4), np.int32)           # off_x, off_y, width, height
    for i, ch in enumerate(CHARS):
        mask, off = render(ch)
        h, w = mask.shape
        assert h <= CELL_H and w <= CELL_W, (ch, mask.shape)
        for r in range(h):
            rows[i, r] = sum(1 << c for c in range(w) if mask[r, c])
        meta[i] = (off[0], off[1], w, h)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "glyphs_hershey_simplex_s1_t2.npz")
    np.savez_compressed(out, chars=np.frombuffer(CHARS.encode(), np.uint8), rows=rows, meta=meta,
                        cv2_version=np.bytes_(cv2.__version__))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
