"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/vr_b200.h declares; host-only helpers (no GPU needed) agree with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import scenes
from ascent_b200 import _lib
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "vr_b200.h")).read()
    return sorted(set(re.findall(r"VR_API\s+[\w\s\*]+?\b(vr_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = header_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), "missing export " + name
    assert sorted(_lib.SYMBOLS) == declared


def test_camera_struct_layout_matches_oracle():
    assert C.sizeof(_lib.CameraStruct) == C.sizeof(O.Camera) == 15 * 4
    assert _lib.PARTIAL_DTYPE.itemsize == 24 == O.PARTIAL_DTYPE.itemsize


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.VRError):
        _lib.Context(0)


def test_host_helpers_match_oracle():
    sc = scenes.mpi_volume_scene()
    assert _lib.sample_distance(sc["bounds"], 100) == sc["sample_dist"]
    o_order, _ = O.visibility_order(np.array(sc["dom_bounds"]), sc["cam"])
    assert list(_lib.visibility_order(sc["dom_bounds"], sc["cam"])) == list(o_order)
    for b in sc["dom_bounds"]:
        assert _lib.find_subset(sc["cam"], 512, 512, b) == O.find_subset(sc["cam"], 512, 512, b)
    rng = np.random.default_rng(0)
    for _ in range(50):
        cam = O.camera_reset_to_bounds([-10, 10, -10, 10, -10, 10])
        O.camera_azimuth(cam, float(rng.uniform(-180, 180)))
        O.camera_elevation(cam, float(rng.uniform(-80, 80)))
        b = np.sort(rng.uniform(-10, 10, (3, 2)), axis=1).reshape(-1)
        W, H = int(rng.integers(16, 2000)), int(rng.integers(16, 2000))
        assert _lib.find_subset(cam, W, H, b) == O.find_subset(cam, W, H, b)
    # visibility order of a 4x4x4 block grid from several cameras: integer order is bit-exact
    doms = []
    for k in range(4):
        for j in range(4):
            for i in range(4):
                doms.append([i, i + 1, j, j + 1, k, k + 1])
    doms = np.array(doms, np.float64)
    for az in (0, 33, 91, 200):
        cam = O.camera_reset_to_bounds([0, 4, 0, 4, 0, 4])
        O.camera_azimuth(cam, az)
        O.camera_elevation(cam, 17)
        assert list(_lib.visibility_order(doms, cam)) == list(O.visibility_order(doms, cam)[0])


def test_python_camera_mirror_equals_oracle_camera():
    """ascent_b200.camera (host mirror of vtkm::rendering::Camera) vs the oracle's K0: same bits."""
    from ascent_b200 import camera
    b = [-10, 10.5, -3, 7, 0, 31]
    c = camera.Camera().reset_to_bounds(b)
    o = O.camera_reset_to_bounds(b)
    for deg_a, deg_e in [(45.0, -10.0), (10.0, 33.0)]:
        c.azimuth(deg_a).elevation(deg_e)
        O.camera_azimuth(o, deg_a)
        O.camera_elevation(o, deg_e)
    c.zoom_by(0.5)
    O.camera_zoom(o, 0.5)
    assert bytes(c.to_struct()) == bytes(o)
    cams = camera.cinema_cameras(b, *camera.cinema_angles(8, 8))
    assert len(cams) == 64
    assert bytes(cams[1 * 8 + 1].to_struct()) == bytes(O.camera_cinema(b, -135.0, 22.5))


def test_header_is_plain_c_and_library_exports_only_the_abi(tmp_path):
    """The boundary is a C ABI: include/vr_b200.h must compile as C99 (no C++ or torch types in the
    signatures) and libvr_b200.so must export nothing but the vr_* entry points it declares."""
    import subprocess
    src = tmp_path / "abi_check.c"
    src.write_text('#include "vr_b200.h"\n'
                   "int main(void) { vr_camera c; vr_partial p; (void)c; (void)p;\n"
                   "  return (int)sizeof(vr_partial) == 24 && VR_IPC_HANDLE_BYTES == 64 ? 0 : 1; }\n")
    exe = tmp_path / "abi_check"
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(exe)]).returncode == 0
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert exported == header_symbols(), sorted(set(exported) ^ set(header_symbols()))
