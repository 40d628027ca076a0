"""Double-entry check of the C oracle's sampler (oracle/raycast_oracle.c, K3 + K4 + K6).

The ray caster's arithmetic lives in VTK-m, which is not in the reference tree, so the C oracle is a
restatement from the spec in SURVEY.md section 8(a).  This file restates the SAME spec a second
time, independently of the C source, as scalar numpy float32 code (every operation rounded on its
own, like the x86 build) and demands bit-identical per-ray results on small cases: a slip in
either transcription (a corner index, an operand order, a comparison direction) shows up as a
mismatch.  Small on purpose: pure-Python loops."""
import numpy as np
import pytest

import scenes
from ascent_b200 import color_table, datasets
from oracle import oracle as O

f32 = np.float32


def slab(o, d, lo, hi):
    """K3 CalcRayStart: returns (min_distance or -1, max_distance)."""
    tnear, tfar = [], []
    for k in range(3):
        dk = d[k]
        inv = f32(1.0) / (f32(1e-8) if abs(dk) < f32(1e-8) else dk)
        od = o[k] * inv
        a = lo[k] * inv - od
        b = hi[k] * inv - od
        tnear.append(min(a, b))
        tfar.append(max(a, b))
    mn = max(f32(0.0), max(tnear))
    mx = min(f32(np.inf), min(tfar))
    return (f32(-1.0) if mx < mn else mn), mx


def march(o, d, t_min, t_max, dims, origin, spacing, field, cell_assoc, lut, sd, rmin, rmax, eps, axes=None):
    """K4 / K5 / K6: returns the ray's RGBA after the epilogue clamp.  axes != None: rectilinear block
    (K5): the cell is the one with x[c] <= p < x[c+1] per axis, its own spacing, the last cell for a
    point exactly on the upper boundary (which also keeps the previous cell's inverse spacing)."""
    nx, ny, nz = dims
    if axes is None:
        minp = [f32(origin[k]) for k in range(3)]
        maxp = [f32(origin[k]) + f32(spacing[k]) * f32(dims[k] - 1) for k in range(3)]
        inv_sp = [f32(1.0) / f32(spacing[k]) for k in range(3)]
    else:
        ax = [np.asarray(a, np.float64).astype(f32) for a in axes]
        minp = [ax[k][0] for k in range(3)]
        maxp = [ax[k][-1] for k in range(3)]
        inv_sp = [f32(0.0)] * 3

    def inside(p):
        return all(not (p[k] < minp[k] or p[k] > maxp[k]) for k in range(3))

    c = [f32(0.0)] * 4
    if t_min == f32(-1.0):
        return c
    t = t_min + eps
    p = [o[k] + t * d[k] for k in range(3)]
    while not inside(p) and t < t_max:
        t = t + sd
        p = [o[k] + t * d[k] for k in range(3)]
    step = [sd * d[k] for k in range(3)]
    inv_delta = f32(1.0) / (rmax - rmin) if (rmax - rmin) != 0 else rmin
    size = f32(lut.shape[0] - 1)
    cell, bl, s, tx = None, None, None, [f32(2.0)] * 3
    new_cell = True
    while inside(p) and t < t_max:
        if not new_cell:
            tx = [(p[k] - bl[k]) * inv_sp[k] for k in range(3)]
            new_cell = max(tx) > f32(1.0) or min(tx) < f32(0.0)
        if new_cell and axes is not None:
            cell = []
            for k in range(3):
                if p[k] == maxp[k]:
                    cell.append(dims[k] - 2)
                else:
                    ck = int(np.searchsorted(ax[k], p[k], side="right")) - 1
                    cell.append(ck)
                    inv_sp[k] = f32(1.0) / (ax[k][ck + 1] - ax[k][ck])
            bl = [ax[k][cell[k]] for k in range(3)]
            tx = [(p[k] - bl[k]) * inv_sp[k] for k in range(3)]
        elif new_cell:
            cell = []
            for k in range(3):
                tk = (p[k] - minp[k]) * inv_sp[k]
                if tk == f32(dims[k] - 1):
                    tk = f32(dims[k] - 2)
                cell.append(int(tk))
            bl = [f32(np.float64(f32(origin[k])) + np.float64(f32(spacing[k])) * cell[k]) for k in range(3)]
            tx = [(p[k] - bl[k]) * inv_sp[k] for k in range(3)]
        if new_cell:
            if cell_assoc:
                s = f32(field[(cell[2] * (ny - 1) + cell[1]) * (nx - 1) + cell[0]])
            else:
                i0 = (cell[2] * ny + cell[1]) * nx + cell[0]
                i1 = i0 + 1; i2 = i1 + nx; i3 = i2 - 1
                i4 = i0 + nx * ny; i5 = i4 + 1; i6 = i5 + nx; i7 = i6 - 1
                s = [f32(field[i]) for i in (i0, i1, i2, i3, i4, i5, i6, i7)]
            new_cell = False
        if cell_assoc:
            v = s
        else:
            l76 = s[7] + tx[0] * (s[6] - s[7])
            l45 = s[4] + tx[0] * (s[5] - s[4])
            top = l45 + tx[1] * (l76 - l45)
            l01 = s[0] + tx[0] * (s[1] - s[0])
            l32 = s[3] + tx[0] * (s[2] - s[3])
            bot = l01 + tx[1] * (l32 - l01)
            v = bot + tx[2] * (top - bot)
        v = (v - rmin) * inv_delta
        idx = int(min(max(v * size, f32(0.0)), size))
        sc = lut[idx]
        a = sc[3] * (f32(1.0) - c[3])
        c = [c[0] + sc[0] * a, c[1] + sc[1] * a, c[2] + sc[2] * a, a + c[3]]
        if c[3] >= f32(1.0):
            break
        t = t + sd
        p = [p[k] + step[k] for k in range(3)]
    return [min(x, f32(1.0)) for x in c]


@pytest.mark.parametrize("kind,cell_assoc,az,samples", [("uniform", False, 0.0, 100), ("uniform", False, 33.0, 17),
                                                        ("uniform", True, -58.0, 40), ("uniform", False, 90.0, 400),
                                                        ("rectilinear", False, 0.0, 100), ("rectilinear", False, 41.0, 30),
                                                        ("rectilinear", True, -77.0, 60)])
def test_numpy_restatement_matches_the_c_oracle_bit_for_bit(kind, cell_assoc, az, samples):
    rng = np.random.default_rng(11)
    dims = (9, 8, 7)
    origin, spacing = [-1.0, 0.5, 2.0], [0.25, 0.375, 0.5]
    n = int(np.prod([d - 1 for d in dims])) if cell_assoc else int(np.prod(dims))
    field = rng.random(n, dtype=np.float32) * f32(3.0) - f32(1.0)
    axes = None
    if kind == "rectilinear":
        axes = [origin[k] + spacing[k] * (dims[k] - 1) * (np.arange(dims[k]) / (dims[k] - 1.0)) ** 1.4 for k in range(3)]
        blk = O.OracleBlock(dims, field, axes=axes, cell_assoc=cell_assoc)
    else:
        blk = O.OracleBlock(dims, field, origin=origin, spacing=spacing, cell_assoc=cell_assoc)
    b = blk.bounds()
    cam = O.camera_reset_to_bounds(b)
    O.camera_azimuth(cam, az)
    O.camera_elevation(cam, az / 3.0)
    W, H = 26, 19
    lut = color_table.parse_color_table(scenes.RAMP_TF).corrected_opacity(samples).lut()
    sd = f32(O.sample_distance(b, samples))
    rmin, rmax = f32(field.min()), f32(field.max())
    rays, res = O.trace_block(blk, cam, W, H, lut, sd, rmin, rmax)
    O.rays_free(rays)
    assert res.n > 100
    o = [f32(v) for v in cam.position]
    lo = [f32(b[0]), f32(b[2]), f32(b[4])]
    hi = [f32(b[1]), f32(b[3]), f32(b[5])]
    ext = [f32(b[1] - b[0]), f32(b[3] - b[2]), f32(b[5] - b[4])]
    eps = f32(np.sqrt(ext[0] * ext[0] + ext[1] * ext[1] + ext[2] * ext[2])) * f32(0.0001)  # meshEpsilon (the default)
    hit = 0
    with np.errstate(over="ignore", invalid="ignore"):
        for r in range(res.n):
            d = [f32(v) for v in res.dir[r]]
            assert abs(np.sqrt(sum(np.float64(x) ** 2 for x in d)) - 1.0) < 1e-6     # K1: unit directions
            t_min, t_max = slab(o, d, lo, hi)
            assert t_min == res.min_dist[r] and (t_min == f32(-1.0) or t_max == res.max_dist[r]), r
            c = march(o, d, t_min, t_max, dims, origin, spacing, field, cell_assoc, lut, sd, rmin, rmax, eps, axes)
            assert np.array_equal(np.array(c, f32).view(np.uint32), res.rgba[r].view(np.uint32)), (r, c, res.rgba[r])
            hit += c[3] > 0
    assert hit > 50
