"""The C++ host mirror (ascent_b200/csrc/host/vtkh_b200.hpp: DataSet, Camera, ColorTable, Render,
VolumeRenderer, Scene, Compositor with vtk-h's method names) driven by the C++ test program
t_vtkh_b200_volume_renderer, which reads like the reference's t_vtk-h_volume_renderer.cpp.

CPU: the host-only classes against the Python mirrors / oracle, and the loud failure without a GPU.
GPU: the reference's vtkh_parallel_render body (2 blocks -> partial path) and its 1-block variant
(image path) pixel-checked against the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

import scenes
from ascent_b200 import _lib, camera, color_table, datasets
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "ascent_b200", "t_vtkh_b200_volume_renderer")


@pytest.fixture(scope="module")
def exe():
    if not os.path.exists(EXE):
        from ascent_b200 import build
        build.build_host()
    return EXE


def get_block(block, num_blocks, total):
    divs = [[list(total[0]), list(total[1])]]
    dim = 0
    while len(divs) < num_blocks:
        for i in range(len(divs)):
            if len(divs) >= num_blocks:
                break
            mins, maxs = divs[i]
            size = maxs[dim] - mins[dim] + 1
            if size <= 1:
                continue
            right = [list(mins), list(maxs)]
            maxs[dim] = mins[dim] + size // 2 - 1
            right[0][dim] = maxs[dim] + 1
            divs.append(right)
        dim = (dim + 1) % 3
    return divs[block]


def create_test_data(block, num_blocks, base_size):
    """CreateTestData (src/tests/vtkh/t_vtkm_test_utils.hpp:196-252) as a harness domain."""
    mins, maxs = get_block(block, num_blocks, ([0, 0, 0], [num_blocks * base_size - 1] * 3))
    dims = [maxs[k] - mins[k] + 2 for k in range(3)]
    ax = [np.float32(mins[k]) + np.arange(dims[k], dtype=np.float32) for k in range(3)]
    Z, Y, X = np.meshgrid(ax[2].astype(np.float64), ax[1].astype(np.float64), ax[0].astype(np.float64),
                          indexing="ij")
    field = np.sqrt(X * X + Y * Y + Z * Z) + 1.0
    return dict(kind="uniform", dims=tuple(dims), origin=[float(m) for m in mins], spacing=[1., 1., 1.],
                field=field.reshape(-1), assoc="point")


def test_host_classes_match_python_mirrors(exe, tmp_path):
    out = str(tmp_path / "h.bin")
    subprocess.check_call([exe, "host", out])
    raw = open(out, "rb").read()
    cam = np.frombuffer(raw, np.float32, 15, 0)
    # same operations on the Python mirror and on the oracle's K0
    b = [-10, 10.5, -3, 7, 0, 31]
    c = camera.Camera().reset_to_bounds(b)
    o = O.camera_reset_to_bounds(b)
    for az, el in [(45.0, -10.0), (10.0, 33.0)]:
        c.azimuth(az).elevation(el)
        O.camera_azimuth(o, az)
        O.camera_elevation(o, el)
    c.zoom_by(0.5)
    O.camera_zoom(o, 0.5)
    ref = np.frombuffer(bytes(o), np.float32)
    assert np.allclose(cam, ref, rtol=1e-5, atol=1e-4)
    assert np.allclose(cam, np.frombuffer(bytes(c.to_struct()), np.float32), rtol=1e-5, atol=1e-4)
    # colour tables: ColorTable::Sample(1024) uint8, byte for byte
    off = 60
    t1 = color_table.ColorTable("cool to warm")
    t1.add_point_alpha(0.0, 0.01)
    t1.add_point_alpha(1.0, 0.6)
    t2 = color_table.default_volume_table()
    t3 = color_table.parse_color_table(scenes.MULTI_RENDER_TF | {"name": "cool to warm"})
    for t in (t1, t2, t3):
        got = np.frombuffer(raw, np.uint8, 4096, off).reshape(1024, 4)
        exp = t.sample_u8(1024)
        assert np.abs(got.astype(int) - exp.astype(int)).max() <= 1
        assert (got == exp).mean() > 0.999
        off += 4096
    gb = np.frombuffer(raw, np.float64, 6, off); off += 48
    rr = np.frombuffer(raw, np.float64, 2, off); off += 16
    dims = np.frombuffer(raw, np.int32, 6, off)
    doms = [create_test_data(i, 2, 32) for i in range(2)]
    assert list(dims) == list(doms[0]["dims"]) + list(doms[1]["dims"]) == [33, 65, 65, 33, 65, 65]
    assert np.array_equal(gb, datasets.union_bounds([datasets.domain_bounds(d) for d in doms]))
    assert rr[0] == min(d["field"].min() for d in doms) and rr[1] == max(d["field"].max() for d in doms)


def test_no_gpu_fails_loudly(exe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([exe, "errors"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU path" in r.stderr


def read_render(path):
    raw = open(path, "rb").read()
    W, H, nb, path_a = struct.unpack_from("4i", raw, 0)
    cam = O.Camera.from_buffer_copy(raw[16:76])
    rmin, rmax = struct.unpack_from("2d", raw, 76)
    launches, = struct.unpack_from("Q", raw, 92)
    off = 100
    rgba = np.frombuffer(raw, np.float32, W * H * 4, off).reshape(-1, 4)
    depth = np.frombuffer(raw, np.float32, W * H, off + W * H * 16)
    return W, H, nb, bool(path_a), cam, rmin, rmax, launches, rgba, depth


@pytest.mark.gpu
@pytest.mark.parametrize("num_blocks", [2, 1, 4])
def test_reference_volume_test_body_against_oracle(exe, tmp_path, num_blocks):
    out = str(tmp_path / "r.bin")
    r = subprocess.run([exe, "render", str(num_blocks), "512", "512", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    W, H, nb, path_a, cam, rmin, rmax, launches, rgba, depth = read_render(out)
    assert path_a == (num_blocks == 1) and launches > 0
    doms = [create_test_data(i, num_blocks, 32) for i in range(num_blocks)]
    gb = datasets.union_bounds([datasets.domain_bounds(d) for d in doms])
    t = color_table.ColorTable("cool to warm")
    t.add_point_alpha(0.0, 0.01)
    t.add_point_alpha(1.0, 0.6)
    sc = dict(doms=doms, W=W, H=H, cam=cam, lut=t.corrected_opacity(100).lut(),
              sample_dist=O.sample_distance(gb, 100), rmin=np.float32(rmin), rmax=np.float32(rmax))
    assert (rmin, rmax) == scenes.field_range(doms)
    if path_a:
        _, _, ref = scenes.oracle_path_a(sc)
    else:
        _, ref, _ = scenes.oracle_path_b(sc)
    d = np.abs(rgba - ref).max(axis=1)
    assert (d <= 1 / 255).mean() >= 0.999 and d.max() <= 3 / 255
    assert (ref[:, 3] > 0).sum() > 10000 and np.array_equal(rgba[:, 3] > 0, ref[:, 3] > 0)
    # Render::Save: the PNG the renderer encoded on the device = RenderBackground + PNGEncoder's conversion of the
    # very canvas that came back (oracle's epilogue), rows flipped
    from PIL import Image
    png = np.array(Image.open(out + ".png").convert("RGBA"))
    c = np.ascontiguousarray(rgba, np.float32).copy()
    O.blend_background(c, (0.2, 0.3, 0.4, 1.0))
    assert np.array_equal(png, O.encode_rgba8(c, W, H, flip=True))


@pytest.mark.gpu
def test_unstructured_domain_forces_path_b(exe, tmp_path):
    """VolumeRenderer::SetInput's classification (VolumeRenderer.cpp:874-903): one domain per rank, but unstructured
    -> m_has_unstructured -> RenderMultipleDomainsPerRank with the unstructured wrapper; against the N4 oracle"""
    out = str(tmp_path / "u.bin")
    r = subprocess.run([exe, "render_unstructured", "384", "320", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    W, H, nb, path_a, cam, rmin, rmax, launches, rgba, depth = read_render(out)
    assert not path_a and launches > 0
    dom = create_test_data(0, 1, 16)
    pts, conn = datasets.structured_to_hexes(dom["dims"], dom["origin"], dom["spacing"])
    gb = datasets.domain_bounds(dom)
    t = color_table.ColorTable("cool to warm")
    t.add_point_alpha(0.0, 0.01)
    t.add_point_alpha(1.0, 0.6)
    sc = dict(points=pts, conn=conn, field=dom["field"].reshape(-1), W=W, H=H, cam=cam,
              lut=t.corrected_opacity(100).lut(), sample_dist=O.sample_distance(gb, 100), rmin=np.float32(rmin),
              rmax=np.float32(rmax))
    _, ref, _ = scenes.oracle_unstructured_path_b(sc)
    d = np.abs(rgba - ref).max(axis=1)
    assert (d <= 1 / 255).mean() >= 0.999 and d.max() <= 3 / 255
    assert (ref[:, 3] > 0).sum() > 10000 and np.array_equal(rgba[:, 3] > 0, ref[:, 3] > 0)


@pytest.mark.gpu
def test_error_behaviour(exe):
    r = subprocess.run([exe, "errors"], capture_output=True, text=True)
    assert r.returncode == 0 and "errors ok" in r.stdout, r.stdout + r.stderr
