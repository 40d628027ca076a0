"""Pin the ray-cast oracle (oracle/raycast_oracle.c) against the reference's own golden images.

The goldens are whole Ascent renders with annotations burned in, so the comparison runs over
hand-picked annotation-free crops (tests/golden/make_golden.py) with the reference's own metric:
a pixel differs if any channel is off by more than 4/255 (ascent_png_compare.cpp:35,138-141),
and the test fails if the differing fraction exceeds the tolerance the reference test used
(check_test_image default 0.001; 0.01 where the test passes 0.01f)."""
import os

import numpy as np
import pytest

import scenes


def crop_diffs(mine, gold, rects):
    return np.concatenate([np.abs(mine[y0:y1, x0:x1, :3].astype(int) - gold[y0:y1, x0:x1].astype(int)).max(axis=2)
                           .reshape(-1) for (y0, y1, x0, x1) in rects])


def crop_stats(mine, gold, rects):
    bad = tot = 0
    for (y0, y1, x0, x1) in rects:
        d = np.abs(mine[y0:y1, x0:x1, :3].astype(int) - gold[y0:y1, x0:x1].astype(int)).max(axis=2)
        bad += int((d > 4).sum())
        tot += d.size
    return bad / tot


@pytest.mark.parametrize("which,name,tol", [(0, "render_0100", 0.001), (1, "render_1100", 0.01)])
def test_multi_render_golden(golden_dir, which, name, tol):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    sc = scenes.multi_render_scene(which)
    _, _, canvas = scenes.oracle_path_a(sc)
    mine = scenes.png_bytes(canvas, sc["W"], sc["H"])
    assert crop_stats(mine, g["rgb"], g["rects"]) <= tol


def test_mpi_volume_golden(golden_dir):
    """2-rank path A: rectilinear sampler, visibility order, uint8 ordered fold."""
    g = np.load(os.path.join(golden_dir, "tout_render_mpi_3d_diy_volume100.npz"))
    sc = scenes.mpi_volume_scene()
    _, _, canvas = scenes.oracle_path_a(sc)
    mine = scenes.png_bytes(canvas, sc["W"], sc["H"])
    assert crop_stats(mine, g["rgb"], g["rects"]) <= 0.01


def test_default_camera_geometry():
    """SURVEY K0 corroboration: in render_0100.png the cube's front face spans
    10/(24.64*tan30) = 0.703 of the image -> subset rect of the 20^3 braid at 512^2."""
    from oracle import oracle as O
    sc = scenes.multi_render_scene(0)
    sx, sy, sw, sh = O.find_subset(sc["cam"], 512, 512, sc["bounds"])
    assert abs(sw / 512.0 - 0.703) < 0.01 and abs(sh / 512.0 - 0.703) < 0.01
    assert abs((sx + sw / 2) - 256) <= 1 and abs((sy + sh / 2) - 256) <= 1


@pytest.mark.parametrize("which,name,within1,within2", [(0, "render_0100", 0.99, 0.997), (1, "render_1100", 0.98, 0.995),
                                                        (2, "tout_render_mpi_3d_diy_volume100", 0.995, 0.998)])
def test_goldens_pin_the_oracle_far_tighter_than_the_reference_tolerance(golden_dir, which, name, within1, within2):
    """The reference's PNGCompare only asks for <= 0.1 % (1 %) of pixels off by more than 4/255.  The
    restated K0-K8 chain actually reproduces the reference's pixels much more closely: >= 98-99.5 % of
    the annotation-free pixels within 1/255 and >= 99.5 % within 2/255 (the rest sit on silhouette
    edges).  Asserting that keeps any drift of the restatement from hiding inside the loose tolerance."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    sc = scenes.multi_render_scene(which) if which < 2 else scenes.mpi_volume_scene()
    _, _, canvas = scenes.oracle_path_a(sc)
    d = crop_diffs(scenes.png_bytes(canvas, sc["W"], sc["H"]), g["rgb"], g["rects"])
    assert (d <= 1).mean() >= within1, (d <= 1).mean()
    assert (d <= 2).mean() >= within2, (d <= 2).mean()
    assert (d <= 4).mean() >= 0.9995


@pytest.mark.parametrize("which,name", [(0, "render_0100"), (1, "render_1100"), (2, "tout_render_mpi_3d_diy_volume100")])
def test_residual_against_the_goldens_is_unbiased_noise(golden_dir, which, name):
    """What is left between the restated chain and the reference's PNGs is not a geometric disagreement: a
    least-squares fit of the per-pixel difference against the golden's image gradients (difference = shift .
    gradient + bias, the signature of a camera / pixel-centre / FindSubset offset) finds a sub-pixel shift below
    0.05 px and a bias below 0.15 grey levels, explaining < 5 % of the variance; the RMS difference is below
    0.7 levels of 255.  The > 1-level differences sit where the IMAGE changes fast (they are 40x more frequent on
    pixels whose golden gradient exceeds 8 levels/px than on flat ones) in the interior of the volume as much as
    on its silhouette -- per-ray rounding of sample positions by a different build of VTK-m, not a convention of
    K0-K8 that the restatement misses."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    sc = scenes.multi_render_scene(which) if which < 2 else scenes.mpi_volume_scene()
    _, _, canvas = scenes.oracle_path_a(sc)
    mine = scenes.png_bytes(canvas, sc["W"], sc["H"])
    gold = g["rgb"]
    A, Y = [], []
    for (y0, y1, x0, x1) in g["rects"]:
        for ch in range(3):
            m = mine[y0:y1, x0:x1, ch].astype(float)
            gd = gold[y0:y1, x0:x1, ch].astype(float)
            gy, gx = np.gradient(gd)
            ok = (gd > 2) & (gd < 253) & (m > 2) & (m < 253)  # (saturated channels carry no signal)
            A.append(np.stack([gx[ok], gy[ok], np.ones(int(ok.sum()))], 1))
            Y.append((m - gd)[ok])
    A, Y = np.concatenate(A), np.concatenate(Y)
    sol = np.linalg.lstsq(A, Y, rcond=None)[0]
    explained = 1.0 - ((Y - A @ sol) ** 2).sum() / ((Y - Y.mean()) ** 2).sum()
    assert abs(sol[0]) < 0.05 and abs(sol[1]) < 0.05, sol
    assert abs(sol[2]) < 0.15, sol
    assert explained < 0.05, explained
    assert np.sqrt((Y ** 2).mean()) < 0.7
