"""Host-side colour table -> 1024 x float4 lookup table.

The drop-in boundary takes the *sampled* table: in a real Ascent build the caller keeps
``vtkm::cont::ColorTable`` and hands ``vr_set_tf`` the array that
``vtkh::detail::convert_table`` (src/libs/vtkh/rendering/VolumeRenderer.cpp:64-91) /
``Mapper::SetActiveColorTable`` produce.  This module is the stand-alone harness's own
generator of that array so that tests and bench can run without VTK-m: control points,
RGB / Lab / Diverging (Msh) interpolation, piecewise-linear alpha, ``Sample(1024)`` rounded
to uint8, then ``* (1/255.f)``.

Mirrors (argument meaning, not code): ``vtkm::cont::ColorTable`` as used by
src/libs/ascent/runtimes/flow_filters/ascent_runtime_conduit_to_vtkm_parsing.cpp:199-305.
"""
import math

import numpy as np

# presets used by the reference's volume tests (name -> (space, [(x, r, g, b)]))
_PRESETS = {
    # [VTK-m, recalled] SURVEY B22
    "cool to warm": ("diverging", [(0.0, 0.23137254902, 0.298039215686, 0.752941176471),
                                   (0.5, 0.865, 0.865, 0.865),
                                   (1.0, 0.705882352941, 0.0156862745098, 0.149019607843)]),
    "black-body radiation": ("rgb", [(0.0, 0.0, 0.0, 0.0), (0.4, 0.9, 0.0, 0.0),
                                     (0.8, 0.9, 0.9, 0.0), (1.0, 1.0, 1.0, 1.0)]),
    "grayscale": ("rgb", [(0.0, 0.0, 0.0, 0.0), (1.0, 1.0, 1.0, 1.0)]),
}

_REF_X, _REF_Y, _REF_Z = 0.9505, 1.000, 1.089


def _rgb_to_lab(rgb):
    def lin(c):
        return ((c + 0.055) / 1.055) ** 2.4 if c > 0.04045 else c / 12.92
    r, g, b = (lin(c) for c in rgb)
    x = r * 0.4124 + g * 0.3576 + b * 0.1805
    y = r * 0.2126 + g * 0.7152 + b * 0.0722
    z = r * 0.0193 + g * 0.1192 + b * 0.9505

    def f(t):
        return t ** (1.0 / 3.0) if t > 0.008856 else 7.787 * t + 16.0 / 116.0
    fx, fy, fz = f(x / _REF_X), f(y / _REF_Y), f(z / _REF_Z)
    return (116.0 * fy - 16.0, 500.0 * (fx - fy), 200.0 * (fy - fz))


def _lab_to_rgb(lab):
    L, a, b = lab
    vy = (L + 16.0) / 116.0
    vx = a / 500.0 + vy
    vz = vy - b / 200.0

    def finv(v):
        return v ** 3 if v ** 3 > 0.008856 else (v - 16.0 / 116.0) / 7.787
    x, y, z = _REF_X * finv(vx), _REF_Y * finv(vy), _REF_Z * finv(vz)
    r = x * 3.2406 + y * -1.5372 + z * -0.4986
    g = x * -0.9689 + y * 1.8758 + z * 0.0415
    bb = x * 0.0557 + y * -0.2040 + z * 1.0570

    def gam(c):
        return 1.055 * (c ** (1.0 / 2.4)) - 0.055 if c > 0.0031308 else 12.92 * c
    r, g, bb = gam(r), gam(g), gam(bb)
    m = max(r, g, bb)
    if m > 1.0:
        r, g, bb = r / m, g / m, bb / m
    return (max(r, 0.0), max(g, 0.0), max(bb, 0.0))


def _lab_to_msh(lab):
    L, a, b = lab
    M = math.sqrt(L * L + a * a + b * b)
    s = math.acos(L / M) if M > 0.001 else 0.0
    h = math.atan2(b, a) if s > 0.001 else 0.0
    return [M, s, h]


def _msh_to_lab(msh):
    M, s, h = msh
    return (M * math.cos(s), M * math.sin(s) * math.cos(h), M * math.sin(s) * math.sin(h))


def _angle_diff(a1, a2):
    d = abs(a1 - a2)
    while d >= 2.0 * math.pi:
        d -= 2.0 * math.pi
    if d > math.pi:
        d = 2.0 * math.pi - d
    return d


def _adjust_hue(msh, unsat_m):
    if msh[0] >= unsat_m - 0.1:
        return msh[2]
    spin = msh[1] * math.sqrt(unsat_m * unsat_m - msh[0] * msh[0]) / (msh[0] * math.sin(msh[1]))
    return msh[2] + spin if msh[2] > -0.3 * math.pi else msh[2] - spin


def _interp_diverging(rgb1, rgb2, w):
    msh1 = _lab_to_msh(_rgb_to_lab(rgb1))
    msh2 = _lab_to_msh(_rgb_to_lab(rgb2))
    if msh1[1] > 0.05 and msh2[1] > 0.05 and _angle_diff(msh1[2], msh2[2]) > 0.33 * math.pi:
        mmid = max(88.0, max(msh1[0], msh2[0]))
        if w < 0.5:
            msh2 = [mmid, 0.0, 0.0]
            w = 2.0 * w
        else:
            msh1 = [mmid, 0.0, 0.0]
            w = 2.0 * w - 1.0
    if msh1[1] < 0.05 and msh2[1] > 0.05:
        msh1[2] = _adjust_hue(msh2, msh1[0])
    elif msh2[1] < 0.05 and msh1[1] > 0.05:
        msh2[2] = _adjust_hue(msh1, msh2[0])
    tmp = [(1.0 - w) * msh1[k] + w * msh2[k] for k in range(3)]
    return _lab_to_rgb(_msh_to_lab(tmp))


def _interp_lab(rgb1, rgb2, w):
    l1, l2 = _rgb_to_lab(rgb1), _rgb_to_lab(rgb2)
    return _lab_to_rgb(tuple((1.0 - w) * l1[k] + w * l2[k] for k in range(3)))


def _interp_rgb(rgb1, rgb2, w):
    return tuple((1.0 - w) * rgb1[k] + w * rgb2[k] for k in range(3))


_INTERP = {"rgb": _interp_rgb, "lab": _interp_lab, "diverging": _interp_diverging}


class ColorTable:
    """Control-point colour table over the normalised scalar range [0, 1].

    ``ColorTable(name)`` loads a preset with opaque alpha end points (0,1),(1,1)
    [VTK-m, recalled]; ``add_point`` / ``add_point_alpha`` at an existing position
    overwrite that node."""

    def __init__(self, name="cool to warm", space=None):
        key = name.lower()
        if key not in _PRESETS:
            raise ValueError("unknown color table preset '%s'" % name)
        sp, pts = _PRESETS[key]
        self.space = space or sp
        self.rgb_points = [(p[0], (p[1], p[2], p[3])) for p in pts]
        self.alpha_points = [(0.0, 1.0), (1.0, 1.0)]

    def clear_colors(self):
        self.rgb_points = []

    def clear_alpha(self):
        self.alpha_points = []

    @staticmethod
    def _insert(points, x, v):
        for i, (px, _) in enumerate(points):
            if px == x:
                points[i] = (x, v)
                return
        points.append((x, v))
        points.sort(key=lambda p: p[0])

    def add_point(self, x, rgb):
        self._insert(self.rgb_points, float(x), tuple(min(1.0, max(0.0, float(c))) for c in rgb))

    def add_point_alpha(self, x, alpha):
        self._insert(self.alpha_points, float(x), min(1.0, max(0.0, float(alpha))))

    def reverse_colors(self):
        self.rgb_points = sorted([(1.0 - x, c) for x, c in self.rgb_points], key=lambda p: p[0])

    def copy(self):
        c = ColorTable.__new__(ColorTable)
        c.space = self.space
        c.rgb_points = list(self.rgb_points)
        c.alpha_points = list(self.alpha_points)
        return c

    def corrected_opacity(self, samples):
        """``VolumeRenderer::CorrectOpacity`` (VolumeRenderer.cpp:448-466):
        alpha' = 1 - (1 - alpha)^(10 / samples) on every alpha control point, in f64, with
        the ratio formed in f32."""
        ratio = float(np.float32(10.0) / np.float32(samples))
        c = self.copy()
        c.alpha_points = [(x, 1.0 - math.pow(1.0 - a, ratio)) for x, a in self.alpha_points]
        return c

    def _color_at(self, x):
        pts = self.rgb_points
        if not pts:
            return (0.0, 0.0, 0.0)
        if x <= pts[0][0]:
            return pts[0][1]
        if x >= pts[-1][0]:
            return pts[-1][1]
        for (x0, c0), (x1, c1) in zip(pts[:-1], pts[1:]):
            if x0 <= x <= x1:
                if x == x0:
                    return c0
                if x == x1:
                    return c1
                w = float(np.float32((x - x0) / (x1 - x0)))
                return _INTERP[self.space](c0, c1, w)
        return pts[-1][1]

    def _alpha_at(self, x):
        pts = self.alpha_points
        if not pts:
            return 1.0
        if x <= pts[0][0]:
            return pts[0][1]
        if x >= pts[-1][0]:
            return pts[-1][1]
        for (x0, a0), (x1, a1) in zip(pts[:-1], pts[1:]):
            if x0 <= x <= x1:
                w = (x - x0) / (x1 - x0)
                return (1.0 - w) * a0 + w * a1
        return pts[-1][1]

    def sample_u8(self, n=1024):
        """``ColorTable::Sample(n, Vec4ui_8)``: n positions from range min to max
        (last one exact), each channel ``(uint8)(c * 255.0f + 0.5f)``."""
        out = np.zeros((n, 4), np.uint8)
        delta = np.float32(1.0) / np.float32(n - 1)
        for i in range(n):
            x = 1.0 if i == n - 1 else float(np.float32(0.0) + delta * np.float32(i))
            r, g, b = self._color_at(x)
            a = self._alpha_at(x)
            for k, c in enumerate((r, g, b, a)):
                out[i, k] = int(np.float32(np.float32(c) * np.float32(255.0) + np.float32(0.5)))
        return out

    def lut(self, n=1024):
        """``convert_table`` (VolumeRenderer.cpp:64-91): uint8 samples * (1/255.f) -> float4."""
        return (self.sample_u8(n).astype(np.float32) * np.float32(1.0 / 255.0)).astype(np.float32)


def default_volume_table():
    """``VolumeRenderer::VolumeRenderer`` default (VolumeRenderer.cpp:395-408): "Cool to Warm"
    plus AddPointAlpha(0, .02) then AddPointAlpha(0, .5) -- both at x = 0."""
    t = ColorTable("cool to warm")
    t.add_point_alpha(0.0, 0.02)
    t.add_point_alpha(0.0, 0.5)
    return t


def parse_color_table(node):
    """``parse_color_table`` (parsing.cpp:199-305) for a dict shaped like the actions YAML:
    {"name": ..., "control_points": [{"type": "rgb"|"alpha", "position": p, "color"|"alpha": v}],
     "reverse": "true"}."""
    name = "cool to warm"
    name_provided = "name" in node
    if name_provided and str(node["name"]).lower() in _PRESETS:
        name = str(node["name"]).lower()
    table = ColorTable(name)
    cps = node.get("control_points", [])
    if any(p["type"] == "rgb" for p in cps) and not name_provided:
        table.clear_colors()
    for p in cps:
        if p["type"] == "rgb":
            table.add_point(p["position"], p["color"])
        elif p["type"] == "alpha":
            table.add_point_alpha(p["position"], p["alpha"])
    if str(node.get("reverse", "false")) == "true":
        table.reverse_colors()
    return table
