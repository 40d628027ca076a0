"""The per-ray geometry of the unstructured kernel (csrc/vr_umesh_geom.hpp: external-face mask, bin traversal,
ray/face crossings), compiled for the HOST (tests/umesh_geom_host.cpp) and checked against the oracle's brute
force over every external face (oracle/raycast_oracle.c, um_boundary_crossings): the signed, sorted crossing
distances must be bit-identical for every ray -- which is what makes the GPU's sample positions the oracle's."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from ascent_b200 import datasets
from oracle import oracle as O
import scenes

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("umesh") / "libumesh_geom_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", out,
                           os.path.join(HERE, "umesh_geom_host.cpp")])
    lib = C.CDLL(out)
    lib.umesh_crossings_host.restype = C.c_int
    return lib


def _rays(um, cam, W, H):
    """origin, direction and the span inside the point bounds of every ray that meets the bounds"""
    r, t = O.trace_umesh(um, cam, W, H, np.zeros((1024, 4), np.float32), 1.0, 0.0, 1.0, structured_conventions=True)
    O.rays_free(r)
    hit = t.min_dist >= 0
    n = int(hit.sum())
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = np.asarray(cam.position, np.float32)
    rays[:, 3:6] = t.dir[hit]
    rays[:, 6] = t.min_dist[hit]
    rays[:, 7] = t.max_dist[hit]
    return rays


def _both(host_lib, um, rays, bins=0):
    n = rays.shape[0]
    want = np.zeros((n, 32), np.float32)
    want_n = np.zeros(n, np.int32)
    O.lib.orc_umesh_crossings_batch(C.byref(um.m), rays.ctypes.data_as(C.c_void_p), n, want.ctypes.data_as(C.c_void_p),
                                    want_n.ctypes.data_as(C.c_void_p))
    got = np.zeros((n, 32), np.float32)
    got_n = np.zeros(n, np.int32)
    mask = np.zeros(um.conn.shape[0], np.uint8)
    rc = host_lib.umesh_crossings_host(um.points.ctypes.data_as(C.c_void_p), um.points.shape[0],
                                       um.conn.ctypes.data_as(C.c_void_p), um.conn.shape[0], um.conn.shape[1], bins,
                                       rays.ctypes.data_as(C.c_void_p), n, got.ctypes.data_as(C.c_void_p),
                                       got_n.ctypes.data_as(C.c_void_p), mask.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return want, want_n, got, got_n, mask


def _check(host_lib, um, cams, W, H, bins=(0,)):
    total = 0
    for cam in cams:
        rays = _rays(um, cam, W, H)
        for b in bins:
            want, want_n, got, got_n, mask = _both(host_lib, um, rays, b)
            assert int(np.unpackbits(mask).sum()) == um.m.n_ext
            assert np.array_equal(want_n, got_n)
            assert np.array_equal(want.view(np.uint32), got.view(np.uint32))
        total += int((want_n > 0).sum())
        # crossings alternate for a watertight boundary: enter, leave, enter, leave ... for nearly every ray
        k = want_n.max()
        sign = np.sign(want[:, :k])
        clean = np.all((sign[:, 0::2] >= 0) & (np.pad(sign[:, 1::2], ((0, 0), (0, sign[:, 0::2].shape[1] - sign[:, 1::2].shape[1]))) <= 0), axis=1)
        assert clean.mean() > 0.98
    return total


def _cams(bounds):
    out = []
    for az, el in ((0.0, 0.0), (25.0, 20.0), (170.0, 11.0), (-50.0, -35.0), (90.0, 0.0), (45.0, 89.0)):
        cam = O.camera_reset_to_bounds(bounds)
        O.camera_azimuth(cam, az)
        O.camera_elevation(cam, el)
        out.append(cam)
    return out


def test_notched_box_of_the_ghost_golden(host_lib):
    sc = scenes.ghost_volume_scene()
    um = O.OracleUMesh(sc["points"], sc["conn"], sc["field"])
    assert um.m.n_ext == 6 * 81  # the box's faces; the two missing corner cells close 5 of them and open 5 others
    assert _check(host_lib, um, _cams(sc["bounds"]), 160, 120, bins=(0, 1, 3, 20)) > 20000


@pytest.mark.parametrize("shape", ["hex", "tet"])
def test_warped_cells_and_a_hollow_mesh(host_lib, shape):
    """interior points pushed off the lattice (non-planar internal faces), a block of cells removed from the middle
    (a cavity: rays enter, leave, enter, leave) and one from a corner"""
    n = 9
    dom = datasets.braid_uniform(n, dtype=np.float32)
    drop = [((k * (n - 1) + j) * (n - 1) + i) for k in range(3, 5) for j in range(3, 6) for i in range(2, 5)] + [0, 1, 8]
    pts, conn = datasets.structured_to_hexes(dom["dims"], dom["origin"], dom["spacing"], drop_cells=drop)
    g = np.random.default_rng(11)
    idx = np.arange(pts.shape[0])
    i, j, k = idx % n, (idx // n) % n, idx // (n * n)
    inner = (i > 0) & (i < n - 1) & (j > 0) & (j < n - 1) & (k > 0) & (k < n - 1)
    pts = pts.copy()
    pts[inner] += (g.random((int(inner.sum()), 3), dtype=np.float32) - 0.5) * np.float32(0.3 * float(dom["spacing"][0]))
    if shape == "tet":
        conn = datasets.hexes_to_tets(conn)
    um = O.OracleUMesh(pts, conn, dom["field"].reshape(-1))
    b = [pts[:, 0].min(), pts[:, 0].max(), pts[:, 1].min(), pts[:, 1].max(), pts[:, 2].min(), pts[:, 2].max()]
    assert _check(host_lib, um, _cams(b), 128, 96, bins=(0, 2, 17)) > 10000


def test_more_crossings_than_the_list_holds(host_lib):
    """20 slabs of cells separated by gaps: a ray along z crosses the boundary 40 times, the list keeps the nearest
    32 on both sides (oracle: brute force; product: bin traversal) -- and the stretches they describe"""
    dims, origin, spacing = (4, 4, 41), (-1.5, -1.5, -20.0), (1.0, 1.0, 1.0)
    ncx, ncy, ncz = 3, 3, 40
    drop = [((k * ncy + j) * ncx + i) for k in range(1, ncz, 2) for j in range(ncy) for i in range(ncx)]
    pts, conn = datasets.structured_to_hexes(dims, origin, spacing, drop_cells=drop)
    field = np.zeros(pts.shape[0], np.float32)
    um = O.OracleUMesh(pts, conn, field)
    b = [-1.5, 1.5, -1.5, 1.5, -20.0, 20.0]
    cam = O.camera_reset_to_bounds(b)          # looks down -z: rays run along the slabs' normal
    rays = _rays(um, cam, 96, 96)
    want, want_n, got, got_n, _ = _both(host_lib, um, rays, 0)
    assert want_n.max() == 32 and (want_n == 32).sum() > 10
    assert np.array_equal(want_n, got_n) and np.array_equal(want.view(np.uint32), got.view(np.uint32))
    full = want[want_n == 32]
    assert (full[:, 0::2] > 0).all() and (full[:, 1::2] < 0).all()        # enter, leave, enter, leave ...
    assert (np.diff(np.abs(full), axis=1) > 0).all()                      # ... in order of distance
    cam2 = O.camera_reset_to_bounds(b)
    O.camera_azimuth(cam2, 35.0)
    O.camera_elevation(cam2, 15.0)
    rays = _rays(um, cam2, 96, 96)
    for bins in (0, 5):
        want, want_n, got, got_n, _ = _both(host_lib, um, rays, bins)
        assert np.array_equal(want_n, got_n) and np.array_equal(want.view(np.uint32), got.view(np.uint32))
