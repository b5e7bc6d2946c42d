"""Object-coverage search (single_object_coverage_finder.py:76-220) on per-frame extents instead of union masks.

The reference finds, per object and axis, all minimal sets of <= 5 images whose visible object points together span
the object's bounding-box dimension within 10 %; every BFS node ORs two N-point boolean masks and takes
max - min of the masked coordinates (SOCF:57-65).  The extent of a union is max(per-image max) - min(per-image min),
so with the (F, n_objects, 6) table of msp_object_extents (one kernel pass per scene) a node costs two scalar
min/max updates.  The search order, pruning rules and the two ``random.sample`` calls (25-image cap, 5000-node cap)
follow the reference so that, given the same ``random`` state, the result is the same dictionary.
"""
from __future__ import annotations

import random


def coverage_of(lo: float, hi: float):
    """SOCF:57-65: None for an empty union, else max - min."""
    return None if hi < lo else hi - lo


def covers_dimension(coverage, target, tolerance) -> bool:
    """SOCF:68-74."""
    return coverage is not None and abs(coverage - target) <= tolerance * target


def find_minimal_combinations(visible_images, ext_min, ext_max, target_dim, tolerance, max_images=5, rng=random):
    """visible_images: image ids in table order; ext_min/ext_max: per-image min / max of the visible object points
    along the searched axis (+inf / -inf when the image sees none).  Returns {k: [tuples of image ids]}."""
    images = list(visible_images)
    lo_all = {img: float(a) for img, a in zip(images, ext_min)}
    hi_all = {img: float(b) for img, b in zip(images, ext_max)}
    if len(images) > 25:                                   # SOCF:118-119
        images = rng.sample(images, 25)
    n = len(images)
    lo = [lo_all[i] for i in images]
    hi = [hi_all[i] for i in images]
    # suffix unions: what the images from index i to the end can still add (SOCF:121-126)
    suf_lo, suf_hi = [0.0] * n, [0.0] * n
    for i in range(n - 1, -1, -1):
        suf_lo[i] = lo[i] if i == n - 1 else min(lo[i], suf_lo[i + 1])
        suf_hi[i] = hi[i] if i == n - 1 else max(hi[i], suf_hi[i + 1])

    def ok(a, b):
        return covers_dimension(coverage_of(a, b), target_dim, tolerance)

    minimal_bits = []                                       # bit sets of the minimal solutions found so far
    solutions = {}
    level = [((images[i],), lo[i], hi[i], i, 1 << i) for i in range(n)]
    singles = []                                            # uncovered single images: the only expansion material
    k = 1
    while k <= max_images and level:
        expand, new_bits = [], []
        for comb, a, b, last, bits in level:
            if any((m & bits) == m for m in minimal_bits):  # superset of a known minimal set (SOCF:131-141)
                continue
            if ok(a, b):
                new_bits.append(bits)
                solutions.setdefault(k, []).append(tuple(comb))
                continue
            if last < n - 1 and not ok(min(a, suf_lo[last]), max(b, suf_hi[last])):   # SOCF:181-185
                continue
            expand.append((comb, a, b, last, bits))
            if k == 1:
                singles.append((comb, a, b, last, bits))
        minimal_bits.extend(new_bits)
        nxt = []
        if k < max_images:
            for comb, a, b, last, bits in expand:
                for c1, a1, b1, last1, bits1 in singles:
                    if last1 > last:
                        nxt.append((comb + c1, min(a, a1), max(b, b1), last1, bits | bits1))
        if len(nxt) > 5000:                                 # SOCF:212-214
            nxt = rng.sample(nxt, 5000)
        level = nxt
        k += 1
    return solutions
