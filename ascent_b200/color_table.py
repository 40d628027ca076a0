"""Host-side colour table -> 1024 x float4 lookup table.

The drop-in boundary takes the *sampled* table: in a real Ascent build the caller keeps
``vtkm::cont::ColorTable`` and hands ``vr_set_tf`` the array that
``vtkh::detail::convert_table`` (src/libs/vtkh/rendering/VolumeRenderer.cpp:64-91) /
``Mapper::SetActiveColorTable`` produce.  This module is the stand-alone harness's own
generator of that array so that tests and bench can run without VTK-m: it keeps the control
points (presets, ``AddPoint`` / ``AddPointAlpha`` semantics, ``parse_color_table``) and hands them
to the library's host helper ``vr_color_table_sample`` (csrc/vr_color_table.hpp: RGB / Lab /
diverging-Msh interpolation in Float32, Float32-accumulated sample positions, uint8 rounding,
``* (1/255.f)``) -- one implementation for Python and C++ callers.

Mirrors (argument meaning, not code): ``vtkm::cont::ColorTable`` as used by
src/libs/ascent/runtimes/flow_filters/ascent_runtime_conduit_to_vtkm_parsing.cpp:199-305.
"""
import ctypes as C

import numpy as np

from . import _lib

_SPACES = {"rgb": 0, "lab": 1, "diverging": 2}

# presets used by the reference's volume tests (name -> (space, [(x, r, g, b)]))
_PRESETS = {
    # [VTK-m, recalled] SURVEY B22; both confirmed by the colour bars of the reference's goldens
    "cool to warm": ("diverging", [(0.0, 0.23137254902, 0.298039215686, 0.752941176471),
                                   (0.5, 0.865, 0.865, 0.865),
                                   (1.0, 0.705882352941, 0.0156862745098, 0.149019607843)]),
    "rainbow desaturated": ("rgb", [(0.0, 0.278431372549, 0.278431372549, 0.858823529412),
                                    (0.143, 0.0, 0.0, 0.360784313725), (0.285, 0.0, 1.0, 1.0),
                                    (0.429, 0.0, 0.501960784314, 0.0), (0.571, 1.0, 1.0, 0.0),
                                    (0.714, 1.0, 0.380392156863, 0.0), (0.857, 0.419607843137, 0.0, 0.0),
                                    (1.0, 0.878431372549, 0.301960784314, 0.301960784314)]),
    "black-body radiation": ("rgb", [(0.0, 0.0, 0.0, 0.0), (0.4, 0.9, 0.0, 0.0),
                                     (0.8, 0.9, 0.9, 0.0), (1.0, 1.0, 1.0, 1.0)]),
    "grayscale": ("rgb", [(0.0, 0.0, 0.0, 0.0), (1.0, 1.0, 1.0, 1.0)]),
}


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
        # node colours are stored as Float32
        self._insert(self.rgb_points, float(x), tuple(float(np.float32(min(1.0, max(0.0, float(c))))) for c in rgb))

    def add_point_alpha(self, x, alpha):
        # alpha nodes are stored as Float32
        self._insert(self.alpha_points, float(x), float(np.float32(min(1.0, max(0.0, float(alpha))))))

    def reverse_colors(self):
        self.rgb_points = sorted([(1.0 - x, c) for x, c in self.rgb_points], key=lambda p: p[0])

    def copy(self):
        c = ColorTable.__new__(ColorTable)
        c.space = self.space
        c.rgb_points = list(self.rgb_points)
        c.alpha_points = list(self.alpha_points)
        return c

    def corrected_opacity(self, samples):
        """``VolumeRenderer::CorrectOpacity`` (VolumeRenderer.cpp:448-466) on a deep copy:
        alpha' = 1 - (1 - alpha)^(10 / samples) on every alpha control point (vr_correct_opacity)."""
        lib = _lib.load()
        c = self.copy()
        c.alpha_points = [(x, float(lib.vr_correct_opacity(C.c_float(a), C.c_float(samples))))
                          for x, a in self.alpha_points]
        return c

    def sample_u8(self, n=1024, want_float=False):
        """``ColorTable::Sample(n, Vec4ui_8)`` (and, with want_float, ``convert_table``'s float4 form)."""
        cx = np.array([p[0] for p in self.rgb_points], np.float64)
        cc = np.array([p[1] for p in self.rgb_points], np.float32).reshape(-1, 3)
        ax = np.array([p[0] for p in self.alpha_points], np.float64)
        av = np.array([p[1] for p in self.alpha_points], np.float32)
        u8 = np.zeros((n, 4), np.uint8)
        f4 = np.zeros((n, 4), np.float32)
        dp, fp = C.POINTER(C.c_double), C.POINTER(C.c_float)
        st = _lib.load().vr_color_table_sample(
            _SPACES[self.space], len(cx), cx.ctypes.data_as(dp), cc.ctypes.data_as(fp), len(ax),
            ax.ctypes.data_as(dp), av.ctypes.data_as(fp), int(n), u8.ctypes.data_as(C.c_void_p),
            f4.ctypes.data_as(fp))
        if st != _lib.VR_OK:
            raise _lib.VRError("vr_color_table_sample: invalid colour table")
        return f4 if want_float else u8

    def lut(self, n=1024):
        """``convert_table`` (VolumeRenderer.cpp:64-91): uint8 samples * (1/255.f) -> float4."""
        return self.sample_u8(n, want_float=True)


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
