"""Golden dot annotations: images painted by cv2 itself (circle fill + putText), the marks that produced them.

    python tests/golden/make_golden_draw.py       # needs cv2 (4.13 here); writes tests/golden/drawing_small.npz

Circles may overlap each other and the borders; label glyphs are kept inside the image (the pixel-exact domain of
msp_draw_marks, include/msp.h) except in the last image, whose expected pixels therefore come from cv2 only for the
circles and from the documented per-pixel clipping for the glyphs (checked separately against the oracle).
"""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from multi_spatialmllm_b200.annotate import MARK_DTYPE, GlyphAtlas  # noqa: E402  (host-only pieces)


def main():
    rng = np.random.default_rng(2024)
    atlas = GlyphAtlas()
    B, H, W = 6, 96, 128
    yy, xx, bb, cc = np.mgrid[0:H, 0:W, 0:B, 0:3]
    base = np.ascontiguousarray(((xx * 3 + yy * 5 + cc * 40 + bb * 17) % 256).astype(np.uint8).transpose(2, 0, 1, 3))
    out = base.copy()
    marks = []
    for b in range(B):
        for k in range(int(rng.integers(3, 9))):
            r = int(rng.choice([0, 1, 2, 5, 10, 12, 13, 25]))
            x, y = int(rng.integers(-r, W + r)), int(rng.integers(-r, H + r))
            col = tuple(int(c) for c in rng.integers(0, 256, 3))
            cv2.circle(out[b], (x, y), r, col, -1)
            glyph, org = -1, (0, 0)
            if rng.random() < 0.7:
                ch = atlas.chars[int(rng.integers(len(atlas.chars)))]
                ox, oy, gw, gh = (int(v) for v in atlas.meta[atlas.index(ch)])
                org = (int(rng.integers(1 - ox, W - 1 - ox - gw)), int(rng.integers(1 - oy, H - 1 - oy - gh)))
                cv2.putText(out[b], ch, org, cv2.FONT_HERSHEY_SIMPLEX, 1.0, col, 2)
                glyph = atlas.index(ch)
            marks.append((b, x, y, r, glyph, org[0], org[1], col, 0))
    marks = np.array(marks, dtype=MARK_DTYPE)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "drawing_small.npz")
    np.savez_compressed(path, base=base, expected=out, marks=marks.view(np.uint8).reshape(-1, 32),
                        cv2_version=np.bytes_(cv2.__version__))
    print("wrote", path, os.path.getsize(path), "bytes,", len(marks), "marks")


if __name__ == "__main__":
    main()
