"""The CPU model of the device PNG encoder (tests/png_model.py) produces valid files: zlib inflates the stream to the
filtered scanlines, PIL decodes the file to the input pixels -- for random content (9-bit literals), flat content
(long distance-1 matches across segment boundaries), widths that do not divide into the 256 segments, 1x1."""
import io
import zlib

import numpy as np
import pytest

import png_model


@pytest.mark.parametrize("W,H", [(64, 48), (203, 77), (1, 1), (3, 2), (1920, 6), (4096, 2)])
def test_model_files_decode_to_the_input(W, H):
    from PIL import Image
    g = np.random.default_rng(W * 1000 + H)
    img = np.zeros((H, W, 4), np.uint8)
    img[H // 4:H // 2 + 1, W // 4:W // 2 + 1] = g.integers(0, 256, (H // 2 + 1 - H // 4, W // 2 + 1 - W // 4, 4))
    img[..., 3] = 255
    if H > 4:
        img[H - 3:] = g.integers(0, 256, (3, W, 4))
    png, raw, z = png_model.encode(img)
    assert zlib.decompress(z) == raw
    back = np.array(Image.open(io.BytesIO(png)).convert("RGBA"))
    assert np.array_equal(back, img)
    # never more than 9/8 of the raw scanlines (every byte a 9-bit literal) + per-row and file overhead
    assert len(png) <= (W * 4 + 1) * H * 9 // 8 + 6 * H + 80
