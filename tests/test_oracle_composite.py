"""Pin the compositing oracle (oracle/composite_oracle.c):
 (a) against the reference's golden PNGs for the apcomp scenes (t_apcomp_c_order.cpp,
     t_apcomp_volume_partials.cpp, t_apcomp_zbuffer.cpp) incl. the known-answer pixels of
     SURVEY 8(c), and
 (b) bit-exactly against the reference's own apcomp sources (oracle/_ref/libapcomp_ref.so)."""
import os

import numpy as np
import pytest

import scenes
from oracle import oracle as O

W = H = 1024
needs_ref = pytest.mark.skipif(O.ref is None, reason="oracle/_ref not built (no /root/reference)")


def _c_order_layers():
    imgs = [scenes.apcomp_image(i) for i in range(4)]
    return np.stack([i[0] for i in imgs]), np.stack([i[1] for i in imgs])


def _oracle_c_order():
    rgba, depth = _c_order_layers()
    q = [O.image_init(rgba[i], depth[i], 2) for i in range(4)]
    out, od = O.ordered_composite(np.stack([x[0] for x in q]), np.stack([x[1] for x in q]),
                                  np.arange(4))
    return out.reshape(H, W, 4), od


@pytest.mark.parametrize("name", ["apcomp_c_order", "apcomp_c_order_mpi"])
def test_c_order_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "apcomp_goldens.npz"))[name]
    out, _ = _oracle_c_order()
    assert np.array_equal(out, g)
    # known-answer pixels
    assert tuple(out[600, 450]) == (255, 128, 65, 222)
    assert tuple(out[700, 350]) == (255, 128, 0, 190)
    assert tuple(out[700, 250]) == (255, 0, 0, 127)
    assert tuple(out[10, 10]) == (0, 0, 0, 0)


@pytest.mark.parametrize("name", ["apcomp_volume_partial", "apcomp_volume_partial_mpi"])
def test_volume_partial_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "apcomp_goldens.npz"))[name]
    res = O.composite_partials([scenes.apcomp_partials(i) for i in range(4)])
    img = scenes.partials_to_image(res, W, H)
    assert np.array_equal(img, g)
    assert tuple(img[600, 450]) == (255, 127, 63, 223)
    assert tuple(img[700, 350]) == (255, 127, 0, 191)


@pytest.mark.parametrize("name", ["apcomp_zbuffer", "apcomp_zbuffer_mpi"])
def test_zbuffer_golden(golden_dir, name):
    """t_apcomp_zbuffer.cpp:26-68: squares at y=400, x=200+100*i, colour 0.1+0.1*i opaque,
    depth 0.05*i, Z_BUFFER_SURFACE_GL (images composited as they are added)."""
    g = np.load(os.path.join(golden_dir, "apcomp_goldens.npz"))[name]
    front = None
    for i in range(4):
        c = np.float32(0.1) + np.float32(i) * np.float32(0.1)
        colors = np.tile(np.array([c, c, c, 1.0], np.float32), (4, 1))
        px = np.zeros((H, W, 4), np.float32)
        dp = np.full((H, W), 1.01, np.float32)
        px[400:700, 200 + 100 * i:500 + 100 * i] = colors[i]
        dp[400:700, 200 + 100 * i:500 + 100 * i] = np.float32(i) * np.float32(0.05)
        q, d = O.image_init(px.reshape(-1, 4), dp.reshape(-1), 1)
        if front is None:
            front, fd = q.copy(), d.copy()
        else:
            O.zbuffer_composite(front, fd, q, d, gl_depth=True)
    img = front.reshape(H, W, 4)
    assert tuple(img[600, 450]) == (25, 25, 25, 255)
    assert np.array_equal(img, g)


@needs_ref
def test_restatement_matches_reference_zbuffer():
    rgba, depth = [], []
    for i in range(4):
        c = np.float32(0.1) + np.float32(i) * np.float32(0.1)
        px = np.zeros((H, W, 4), np.float32)
        dp = np.full((H, W), 1.01, np.float32)
        px[400:700, 200 + 100 * i:500 + 100 * i] = [c, c, c, 1.0]
        dp[400:700, 200 + 100 * i:500 + 100 * i] = np.float32(i) * np.float32(0.05)
        rgba.append(px.reshape(-1, 4)); depth.append(dp.reshape(-1))
    ref_out, ref_d = O.ref_composite_zbuffer(np.stack(rgba), np.stack(depth), 4, W, H)
    front, fd = O.image_init(rgba[0], depth[0], 1)
    for i in range(1, 4):
        q, d = O.image_init(rgba[i], depth[i], 1)
        O.zbuffer_composite(front, fd, q, d, gl_depth=True)
    assert np.array_equal(front, ref_out) and np.array_equal(fd, ref_d)


@needs_ref
def test_restatement_matches_reference_c_order():
    rgba, depth = _c_order_layers()
    ref_out, ref_d = O.ref_composite_vis_order(rgba, depth, np.arange(4), W, H)
    out, od = _oracle_c_order()
    assert np.array_equal(out.reshape(-1, 4), ref_out)
    assert np.array_equal(od, ref_d)


@needs_ref
@pytest.mark.parametrize("n_img,seed", [(2, 0), (3, 1), (8, 2)])
def test_restatement_matches_reference_random_images(n_img, seed):
    rng = np.random.default_rng(seed)
    w, h = 97, 61
    alpha = rng.random((n_img, h * w, 1), dtype=np.float32)
    rgba = np.concatenate([rng.random((n_img, h * w, 3), dtype=np.float32) * alpha, alpha], axis=2)
    rgba[rng.random((n_img, h * w)) < 0.3] = 0
    depth = rng.random((n_img, h * w), dtype=np.float32) * 1.2
    order = rng.permutation(n_img).astype(np.int32)
    ref_out, ref_d = O.ref_composite_vis_order(rgba, depth, order, w, h)
    q = [O.image_init(rgba[i], depth[i], 2) for i in range(n_img)]
    out, od = O.ordered_composite(np.stack([x[0] for x in q]), np.stack([x[1] for x in q]), order)
    assert np.array_equal(out, ref_out)
    assert np.array_equal(od, ref_d)


@needs_ref
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_restatement_matches_reference_random_partials(seed):
    """distinct depths per pixel -> the unstable std::sort has a unique answer."""
    rng = np.random.default_rng(seed)
    n_pix, n_lists = 500, 5
    lists = []
    for l in range(n_lists):
        ids = rng.choice(n_pix, size=rng.integers(1, n_pix), replace=False)
        p = np.zeros(ids.size, O.PARTIAL_DTYPE)
        p["pixel_id"] = ids + 1000
        p["depth"] = rng.random(ids.size, dtype=np.float32) + l
        a = rng.random(ids.size, dtype=np.float32)
        a[rng.random(ids.size) < 0.1] = 1.0
        p["alpha"] = a
        p["rgb"] = rng.random((ids.size, 3), dtype=np.float32) * a[:, None]
        lists.append(p)
    ref_out = O.ref_composite_partials(lists)
    out = O.composite_partials(lists)
    assert out.size == ref_out.size
    assert out.tobytes() == ref_out.tobytes()


def test_partial_owner_matches_diy_rule():
    """1-D RegularDecomposer: width=(max-min+1)/n, owner=min((p-min)/width, n-1)."""
    assert O.partial_owner(0, 0, 99, 4) == 0
    assert O.partial_owner(24, 0, 99, 4) == 0
    assert O.partial_owner(25, 0, 99, 4) == 1
    assert O.partial_owner(99, 0, 99, 4) == 3
    assert O.partial_owner(102, 3, 102, 3) == 2   # 100 px / 3 -> width 33, last rank takes the tail
    assert O.partial_owner(3 + 99, 3, 102, 3) == 2
