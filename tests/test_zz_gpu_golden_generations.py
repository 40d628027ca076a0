"""The product side of tests/test_oracle_golden.py: with the first-sample offset of the VTK-m generation that rendered
them switched in through the ABI (vr_set_first_sample_offset), the CUDA path reproduces the reference's older
pure-volume goldens uint8 for uint8, bit-identical to the oracle under the same switch; switched back, it is
bit-identical to the oracle's default (meshEpsilon) again.  (Runs last: it changes a context-wide setting.)"""
import os

import numpy as np
import pytest

from ascent_b200 import _lib
from oracle import oracle as O
import scenes
from test_gpu_parity import gpu_path_a_canvas
from test_oracle_golden import OLDER_GENERATION, crop_diffs, first_sample

pytestmark = pytest.mark.gpu


@pytest.fixture()
def ctx():
    c = _lib.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("which,name,exact", [(0, "render_0100", 0.998), (1, "render_1100", 1.0)])
def test_product_reproduces_the_older_goldens_when_switched(ctx, golden_dir, which, name, exact):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    sc = scenes.multi_render_scene(which)
    W, H = sc["W"], sc["H"]
    ctx.set_first_sample_offset(*OLDER_GENERATION)
    gpu_path_a_canvas(ctx, sc["doms"][0], sc)
    ctx.image_from_canvas()
    u8, d = ctx.image_download(W, H)
    with first_sample(*OLDER_GENERATION):
        o_u8, _, _ = scenes.oracle_path_a(sc)
    assert np.array_equal(u8, o_u8)
    can, _ = O.image_to_canvas(u8, d)
    diff = crop_diffs(scenes.png_bytes(can, W, H), g["rgb"], g["rects"])
    assert (diff == 0).mean() >= exact and diff.max() <= 1
    # back to the default: the oracle's default again, and a different image
    ctx.set_first_sample_offset()
    gpu_path_a_canvas(ctx, sc["doms"][0], sc)
    ctx.image_from_canvas()
    u8_default, _ = ctx.image_download(W, H)
    o_default, _, _ = scenes.oracle_path_a(sc)
    assert np.array_equal(u8_default, o_default)
    assert not np.array_equal(u8_default, u8)
