"""Host-side helpers of the bench / harness that need no GPU: the screen footprint used by the footprint read-back."""
import numpy as np

import bench
from ascent_b200 import _lib, datasets
from oracle import oracle as O


def test_footprint_rect_contains_every_domains_subset_and_is_group_aligned():
    doms = datasets.braid_uniform_blocks(9, 2, dtype=np.float32)
    bl = [datasets.domain_bounds(d) for d in doms]
    gb = datasets.union_bounds(bl)
    W, H = 640, 360
    for az in (0.0, 33.0, 170.0):
        cam = O.camera_reset_to_bounds(gb)
        O.camera_azimuth(cam, az)
        x0, y0, x1, y1 = bench.footprint_rect(cam, W, H, bl)
        assert x0 % 4 == 0 and (x1 % 4 == 0 or x1 == W) and 0 <= x0 < x1 <= W and 0 <= y0 < y1 <= H
        for b in bl:
            sx, sy, sw, sh = _lib.find_subset(cam, W, H, b)
            assert x0 <= sx and sx + sw <= x1 and y0 <= sy and sy + sh <= y1
        # ... and of the global bounds' subset (what the C++ mirror uses)
        sx, sy, sw, sh = _lib.find_subset(cam, W, H, gb)
        assert (x1 - x0) * (y1 - y0) <= ((sw + 8) * (sh + 2)) * 1.05


def test_union_rect():
    assert bench.union_rect(None, (1, 2, 3, 4)) == (1, 2, 3, 4)
    assert bench.union_rect((0, 0, 0, 0), (1, 2, 3, 4)) == (1, 2, 3, 4)
    assert bench.union_rect((4, 1, 8, 3), (1, 2, 6, 9)) == (1, 1, 8, 9)
