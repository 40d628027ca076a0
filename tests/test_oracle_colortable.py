"""K8 (transfer-function table), V3 (opacity correction) and K0 (camera): the oracle's restatement
(oracle/colortable_oracle.c, raycast_oracle.c) is pinned against the reference's own output -- the
colour bars VTK-m draws into the golden images are ColorTable::Sample(bar height) of the same tables
-- and the product's host code (csrc/vr_color_table.hpp, vr_host_math.hpp, reached through the C ABI
from Python and from the C++ mirror) is held bit-equal to the oracle."""
import math
import os

import numpy as np
import pytest

from ascent_b200 import _lib, camera, color_table
from oracle import oracle as O

BARS = [("cool to warm", "cool_to_warm_179"), ("cool to warm", "cool_to_warm_359"),
        ("rainbow desaturated", "rainbow_desaturated_359")]


@pytest.mark.parametrize("preset,key", BARS)
def test_oracle_table_reproduces_the_reference_colour_bars_exactly(golden_dir, preset, key):
    """897 golden samples, every channel equal.  (Sampling positions as start + i * delta, or the
    colour-space math in double with i * delta positions, miss two of them by one level: the bars
    pin the Float32-accumulated positions.)"""
    bar = np.load(os.path.join(golden_dir, "colorbars.npz"))[key]
    mine = O.ColorTable(preset).sample_u8(bar.shape[0])[:, :3]
    assert np.array_equal(mine, bar)


@pytest.mark.parametrize("preset,key", BARS)
def test_product_table_reproduces_the_reference_colour_bars_exactly(golden_dir, preset, key):
    bar = np.load(os.path.join(golden_dir, "colorbars.npz"))[key]
    assert np.array_equal(color_table.ColorTable(preset).sample_u8(bar.shape[0])[:, :3], bar)


TABLES = [
    {"name": "cool to warm", "control_points": [{"type": "alpha", "position": 0., "alpha": 0.},
                                                {"type": "alpha", "position": 1., "alpha": 1.}]},
    {"name": "rainbow desaturated", "control_points": [{"type": "alpha", "position": 0., "alpha": 0.2},
                                                       {"type": "alpha", "position": 1., "alpha": 0.5}]},
    {"control_points": [{"type": "rgb", "position": 0., "color": [1, 0, 0]},
                        {"type": "rgb", "position": .5, "color": [0, 1, 0]},
                        {"type": "rgb", "position": 1., "color": [0, 0, 1]},
                        {"type": "alpha", "position": 0., "alpha": .1},
                        {"type": "alpha", "position": .3, "alpha": .9},
                        {"type": "alpha", "position": 1., "alpha": .4}]},
    {"name": "black-body radiation"},
    {},
]


@pytest.mark.parametrize("node", TABLES)
def test_product_lut_is_bit_equal_to_the_oracle(node):
    for samples in (100, 7, 887):
        for n in (1024, 359, 2):
            a = color_table.parse_color_table(node).corrected_opacity(samples)
            b = O.parse_color_table(node).correct_opacity(samples)
            assert np.array_equal(a.sample_u8(n), b.sample_u8(n)), (node.get("name"), samples, n)
            assert a.lut(n).tobytes() == b.lut(n).tobytes()


def test_default_volume_table_quirk():
    """VolumeRenderer.cpp:404-405 adds both alpha points at x = 0: the second one replaces
    the first, so the table starts at the corrected 0.5, never at 0.02."""
    a = color_table.default_volume_table().corrected_opacity(100).lut()
    b = O.default_volume_table().correct_opacity(100).lut()
    assert a.tobytes() == b.tobytes()
    k = int(np.float32(np.float32(1.0 - math.pow(1.0 - 0.5, float(np.float32(10.0) / np.float32(100.0)))) * 255 + 0.5))
    assert abs(a[0, 3] * 255 - k) < 1e-3


def test_correct_opacity_matches_the_oracle():
    """V3 directly (VolumeRenderer.cpp:448-466): f32 ratio, f64 pow, stored back as Float32."""
    lib = _lib.load()
    import ctypes as C
    for samples in (1, 7, 100, 887, 1000):
        for alpha in (0.0, 0.02, 0.5, 0.999, 1.0, 0.123456789):
            mine = lib.vr_correct_opacity(C.c_float(alpha), C.c_float(samples))
            want = np.float32(O.correct_opacity(float(np.float32(alpha)), samples))
            assert np.float32(mine) == want, (samples, alpha)


def test_python_and_abi_camera_are_bit_equal_to_the_oracle():
    """K0: ResetToBounds, Azimuth, Elevation, Zoom and the cinema orbit, on random bounds/angles."""
    rng = np.random.default_rng(7)
    for t in range(100):
        b = np.sort(rng.uniform(-50, 50, (3, 2)), axis=1).reshape(-1)
        c = camera.Camera().reset_to_bounds(b)
        o = O.camera_reset_to_bounds(b)
        for _ in range(3):
            az, el = rng.uniform(-180, 180, 2)
            c.azimuth(az).elevation(el)
            O.camera_azimuth(o, az)
            O.camera_elevation(o, el)
        z = rng.uniform(-1, 1)
        c.zoom_by(z)
        O.camera_zoom(o, z)
        assert bytes(c.to_struct()) == bytes(o), t
        phi, theta = rng.uniform(-180, 180), rng.uniform(0, 180)
        cc = camera.cinema_cameras(b, [phi], [theta])[0]
        assert bytes(cc.to_struct()) == bytes(O.camera_cinema(b, phi, theta)), t
    assert camera.cinema_angles(8, 8) == O.cinema_angles(8, 8)


def test_braid_generators_agree():
    """the harness's braid (ascent_b200.datasets) against the oracle's point-by-point restatement."""
    from ascent_b200 import datasets
    a = datasets.braid_values(9, 8, 7, 1, 2, 3, 20, 21, 22)
    b = O.braid_values(9, 8, 7, 1, 2, 3, 20, 21, 22)
    assert np.abs(a - b).max() < 1e-12
