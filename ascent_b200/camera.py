"""Host-side camera: the subset of ``vtkm::rendering::Camera`` that Ascent's render parsing drives
(src/libs/ascent/runtimes/flow_filters/ascent_runtime_conduit_to_vtkm_parsing.cpp:97-173,
src/libs/vtkh/rendering/Render.cpp:314-347, cinema orbit
ascent_runtime_rendering_filters.cpp:906-960).  All state is float32 like VTK-m's.  The tracer
consumes the plain ``vr_camera`` struct (``to_struct``)."""
import ctypes as C
import math

import numpy as np

from . import _lib

F = np.float32


def _v(x):
    return np.array(x, dtype=np.float32)


def _normalize(v):
    r = F(1.0) / F(np.sqrt(F(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])))
    return (v * r).astype(np.float32)


class Camera:
    """State = one ``vr_camera`` struct; every operation is the library's host helper
    (csrc/vr_host_math.hpp), so Python and C++ callers get the same bits."""

    def __init__(self):
        self.c = _lib.CameraStruct()
        _lib.load().vr_camera_default(C.byref(self.c))

    # -- vector-valued fields as float32 numpy views of the struct
    def _get(self, name):
        return np.array(list(getattr(self.c, name)), np.float32)

    def _set(self, name, v):
        getattr(self.c, name)[:] = [float(F(x)) for x in v]

    look_at = property(lambda self: self._get("look_at"), lambda self, v: self._set("look_at", v))
    position = property(lambda self: self._get("position"), lambda self, v: self._set("position", v))
    up = property(lambda self: self._get("up"), lambda self, v: self._set("up", v))

    def _scalar(name):
        return property(lambda self: F(getattr(self.c, name)), lambda self, v: setattr(self.c, name, float(F(v))))

    fov = _scalar("fov")
    zoom = _scalar("zoom")
    xpan = _scalar("xpan")
    ypan = _scalar("ypan")
    near_plane = _scalar("near_plane")
    far_plane = _scalar("far_plane")
    del _scalar

    # -- setters named after the vtkm methods parse_camera calls
    def set_look_at(self, v): self.look_at = v
    def set_position(self, v): self.position = v
    def set_view_up(self, v): self.up = v
    def set_field_of_view(self, deg): self.fov = deg

    def set_clipping_range(self, near, far):
        self.near_plane, self.far_plane = near, far

    def reset_to_bounds(self, bounds):
        """``Camera::ResetToBounds(bounds)``: look at the centre from |extent| away along the
        current view direction, fov 60, clip [0.1, 10] x diagonal, pan 0, zoom 1."""
        _lib.load().vr_camera_reset_to_bounds(C.byref(self.c), _lib._d6(bounds))
        return self

    def azimuth(self, deg):
        _lib.load().vr_camera_azimuth(C.byref(self.c), C.c_float(deg))
        return self

    def elevation(self, deg):
        _lib.load().vr_camera_elevation(C.byref(self.c), C.c_float(deg))
        return self

    def zoom_by(self, z):
        """``Camera::Zoom(z)``: zoom *= 4^z."""
        _lib.load().vr_camera_zoom(C.byref(self.c), C.c_float(z))
        return self

    def apply_ascent_zoom(self, user_zoom):
        """``zoom_to_vtkm_zoom`` (parsing.cpp:59-69): Ascent's zoom factor -> log4 -> Zoom()."""
        return self.zoom_by(math.log(user_zoom) / math.log(4.0))

    def pan(self, dx, dy):
        self.xpan, self.ypan = F(self.xpan + F(dx)), F(self.ypan + F(dy))
        return self

    def to_struct(self):
        out = _lib.CameraStruct()
        C.memmove(C.byref(out), C.byref(self.c), C.sizeof(_lib.CameraStruct))
        return out


def parse_camera(node, camera):
    """``parse_camera`` (parsing.cpp:97-173) for a dict shaped like the actions YAML."""
    if "look_at" in node: camera.set_look_at(node["look_at"])
    if "position" in node: camera.set_position(node["position"])
    if "up" in node: camera.set_view_up(_normalize(_v(node["up"])))
    if "fov" in node: camera.set_field_of_view(node["fov"])
    if "xpan" in node or "ypan" in node:
        xpan = 0.0
        if "xpan" in node: xpan = node["xpan"]
        if "ypan" in node: xpan = node["ypan"]  # sic: parsing.cpp:134-136
        camera.pan(xpan, 0.0)
    if "zoom" in node: camera.apply_ascent_zoom(node["zoom"])
    if "near_plane" in node: camera.near_plane = F(node["near_plane"])
    if "far_plane" in node: camera.far_plane = F(node["far_plane"])
    if "azimuth" in node: camera.azimuth(node["azimuth"])
    if "elevation" in node: camera.elevation(node["elevation"])
    return camera


def cinema_cameras(bounds, phi_values, theta_values):
    """``CinemaManager::create_cinema_cameras`` (rendering_filters.cpp:906-960) for the
    phi x theta grid of ``create_cinema_angles`` (:882-893)."""
    cams = []
    for phi in phi_values:
        for theta in theta_values:
            cam = Camera()
            _lib.load().vr_camera_cinema(C.byref(cam.c), _lib._d6(bounds), C.c_float(phi), C.c_float(theta))
            cams.append(cam)
    return cams


def cinema_angles(phi, theta):
    """phi=N, theta=M -> angle lists of CinemaManager (rendering_filters.cpp:568-574,613-619)."""
    ph = [float(F(-180.0 + (360.0 / float(phi)) * a)) for a in range(phi)]
    th = [float(F(0.0 + (180.0 / float(theta)) * a)) for a in range(theta)]
    return ph, th
