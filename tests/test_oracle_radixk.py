"""C4 across ranks: the radix-k surface compositor (RadixKCompositor.cpp:35-180 + the DIY it vendors).

Three layers, each pinned to the one above:
  1. the REFERENCE's own reduce_images / CollectImages / DIY partners, n blocks in one process
     (oracle/_ref/libradixk_ref.so, built by oracle/Makefile from the sources under /root/reference);
  2. oracle.radixk_simulate (the literal round-by-round restatement) and oracle.radixk_zbuffer (its closed form:
     per-piece visiting order + "last of the nearest wins");
  3. the product's host schedule vr_radixk_schedule (what the z-select kernel consumes), through the C ABI.
The GPU kernel itself is checked against (2) in tests/test_gpu_single_process.py and the multi-rank worker."""
import numpy as np
import pytest

from ascent_b200 import _lib
from oracle import oracle as O

SIZES = [(37, 23), (64, 48), (16, 9), (101, 7), (9, 5), (320, 200)]
TIED = np.array([0.1, 0.25, 0.25, 0.5, 0.75, 1.0, 1.001, 1.5], np.float32)  # few values: many exact ties; two > 1


def _images(n, W, H, seed):
    g = np.random.default_rng(seed)
    return g.integers(0, 256, (n, H * W, 4), dtype=np.uint8), g.choice(TIED, (n, H * W))


needs_ref = pytest.mark.skipif(O.radixk_ref is None, reason="oracle/_ref/libradixk_ref.so not built (no /root/reference)")


@needs_ref
@pytest.mark.parametrize("n", list(range(1, 17)))
def test_restatements_match_the_reference_radixk(n):
    for k, (W, H) in enumerate(SIZES):
        c, d = _images(n, W, H, 100 * n + k)
        try:
            rc, rd, info = O.ref_radixk_zbuffer(c, d, W, H)
        except RuntimeError:
            # the reference throws "Unable to decompose domain": so must the restatement
            with pytest.raises(RuntimeError):
                O.radixk_divisions(n, W, H)
            continue
        div = O.radixk_divisions(n, W, H)
        assert div == info["divisions"]
        assert [(a, b) for a, b, _ in O.radixk_rounds(div)] == info["rounds"]
        sc, sd = O.radixk_simulate(c, d, W, H)
        assert np.array_equal(sc, rc) and np.array_equal(sd, rd), (n, W, H, "literal restatement")
        zc, zd = O.radixk_zbuffer(c, d, W, H)
        assert np.array_equal(zc, rc) and np.array_equal(zd, rd), (n, W, H, "closed form")


@needs_ref
def test_reference_tie_break_is_not_rank_order():
    """What the round-1 kernel got wrong: with every fragment at the same depth the reference's winner depends on
    the piece of the frame, and is neither the lowest nor (everywhere) the highest rank."""
    n, W, H = 8, 64, 48
    c = np.zeros((n, H * W, 4), np.uint8)
    for r in range(n):
        c[r, :, 0] = r + 1
    d = np.full((n, H * W), 0.5, np.float32)
    rc, _, info = O.ref_radixk_zbuffer(c, d, W, H)
    assert info["divisions"] == [4, 2] and info["rounds"] == [(0, 4), (1, 2)]
    winners = set(np.unique(rc[:, 0]).tolist())
    assert winners == {3, 4, 7, 8}
    zc, _ = O.radixk_zbuffer(c, d, W, H)
    assert np.array_equal(zc, rc)


@needs_ref
def test_full_hd_frame_eight_ranks():
    n, W, H = 8, 1920, 1080
    c, d = _images(n, W, H, 7)
    rc, rd, _ = O.ref_radixk_zbuffer(c, d, W, H)
    zc, zd = O.radixk_zbuffer(c, d, W, H)
    assert np.array_equal(zc, rc) and np.array_equal(zd, rd)


@pytest.mark.parametrize("n", list(range(1, 17)))
def test_product_schedule_equals_the_oracle(n):
    """vr_radixk_schedule (ascent_b200/csrc/vr_radixk.hpp, host only) against oracle.radixk_schedule."""
    for W, H in SIZES + [(1920, 1080), (3840, 2160), (4096, 4096), (1, 1), (2, 1), (5, 3)]:
        got = _lib.radixk_schedule(n, W, H)
        try:
            want = O.radixk_schedule(n, W, H)
        except RuntimeError:
            assert got is None, (n, W, H)
            continue
        assert got is not None, (n, W, H)
        assert got["divisions"] == want["divisions"], (n, W, H)
        assert got["lo"] == want["lo"], (n, W, H)
        assert got["seq"] == want["seq"], (n, W, H)


def test_schedule_rejects_bad_arguments():
    assert _lib.radixk_schedule(17, 64, 64) is None  # more ranks than the exchange supports
    assert _lib.radixk_schedule(11, 9, 5) is None    # the reference throws here too
