"""Host-side camera: the subset of ``vtkm::rendering::Camera`` that Ascent's render parsing drives
(src/libs/ascent/runtimes/flow_filters/ascent_runtime_conduit_to_vtkm_parsing.cpp:97-173,
src/libs/vtkh/rendering/Render.cpp:314-347, cinema orbit
ascent_runtime_rendering_filters.cpp:906-960).  All state is float32 like VTK-m's.  The tracer
consumes the plain ``vr_camera`` struct (``to_struct``)."""
import math

import numpy as np

from . import _lib

F = np.float32
_PI_180 = F(0.01745329251994329547437168059786927)


def _v(x):
    return np.array(x, dtype=np.float32)


def _normalize(v):
    r = F(1.0) / F(np.sqrt(F(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])))
    return (v * r).astype(np.float32)


def _cross(a, b):
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    return np.cross(a64, b64).astype(np.float32)


def _rotate(deg, axis):
    a = _PI_180 * F(deg)
    n = _normalize(_v(axis))
    s, c = F(math.sin(a)), F(math.cos(a))
    m = np.eye(4, dtype=np.float32)
    for i in range(3):
        for j in range(3):
            m[i, j] = n[i] * n[j] * (F(1) - c)
    m[0, 0] += c; m[1, 1] += c; m[2, 2] += c
    m[0, 1] -= n[2] * s; m[0, 2] += n[1] * s
    m[1, 0] += n[2] * s; m[1, 2] -= n[0] * s
    m[2, 0] -= n[1] * s; m[2, 1] += n[0] * s
    return m


def _translate(t):
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = t
    return m


class Camera:
    def __init__(self):
        self.look_at = _v([0, 0, 0])
        self.position = _v([0, 0, 1])
        self.up = _v([0, 1, 0])
        self.fov = F(60)
        self.zoom = F(1)
        self.xpan = F(0)
        self.ypan = F(0)
        self.near_plane = F(0.01)
        self.far_plane = F(1000)

    # -- setters named after the vtkm methods parse_camera calls
    def set_look_at(self, v): self.look_at = _v(v)
    def set_position(self, v): self.position = _v(v)
    def set_view_up(self, v): self.up = _v(v)
    def set_field_of_view(self, deg): self.fov = F(deg)

    def set_clipping_range(self, near, far):
        self.near_plane, self.far_plane = F(near), F(far)

    def reset_to_bounds(self, bounds):
        """``Camera::ResetToBounds(bounds)``: look at the centre from |extent| away along the
        current view direction, fov 60, clip [0.1, 10] x diagonal, pan 0, zoom 1."""
        b = np.asarray(bounds, np.float64)
        d = _normalize(self.position - self.look_at)
        center = _v([(b[0] + b[1]) / 2, (b[2] + b[3]) / 2, (b[4] + b[5]) / 2])
        ext = _v([b[1] - b[0], b[3] - b[2], b[5] - b[4]])
        diag = F(np.sqrt(F(ext[0] * ext[0] + ext[1] * ext[1] + ext[2] * ext[2])))
        self.look_at = center
        self.position = (center + d * diag * F(1.0)).astype(np.float32)
        self.fov = F(60)
        self.near_plane, self.far_plane = F(0.1) * diag, diag * F(10)
        self.xpan = self.ypan = F(0)
        self.zoom = F(1)
        return self

    def _rotate_about_look_at(self, deg, axis):
        m = _translate(self.look_at) @ _rotate(deg, axis) @ _translate(-self.look_at)
        p = m @ np.append(self.position, F(1)).astype(np.float32)
        self.position = p[:3].astype(np.float32)

    def azimuth(self, deg):
        self._rotate_about_look_at(deg, self.up)
        return self

    def elevation(self, deg):
        self._rotate_about_look_at(deg, _cross(self.position - self.look_at, self.up))
        return self

    def zoom_by(self, z):
        """``Camera::Zoom(z)``: zoom *= 4^z."""
        self.zoom = F(self.zoom * F(math.pow(4.0, z)))
        return self

    def apply_ascent_zoom(self, user_zoom):
        """``zoom_to_vtkm_zoom`` (parsing.cpp:59-69): Ascent's zoom factor -> log4 -> Zoom()."""
        return self.zoom_by(math.log(user_zoom) / math.log(4.0))

    def pan(self, dx, dy):
        self.xpan, self.ypan = F(self.xpan + F(dx)), F(self.ypan + F(dy))
        return self

    def to_struct(self):
        c = _lib.CameraStruct()
        c.position[:] = [float(x) for x in self.position]
        c.look_at[:] = [float(x) for x in self.look_at]
        c.up[:] = [float(x) for x in self.up]
        c.fov, c.zoom, c.xpan, c.ypan = float(self.fov), float(self.zoom), float(self.xpan), float(self.ypan)
        c.near_plane, c.far_plane = float(self.near_plane), float(self.far_plane)
        return c


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
    b = np.asarray(bounds, np.float64)
    center = _v([(b[0] + b[1]) / 2, (b[2] + b[3]) / 2, (b[4] + b[5]) / 2])
    ext = _v([b[1] - b[0], b[3] - b[2], b[5] - b[4]])
    radius = F(F(np.sqrt(F(ext[0] * ext[0] + ext[1] * ext[1] + ext[2] * ext[2]))) * 2.5 / 2.0)
    cams = []
    for phi in phi_values:
        for theta in theta_values:
            cam = Camera().reset_to_bounds(b)
            rot = _rotate(F(phi), [0, 0, 1]) @ _rotate(F(theta), [1, 0, 0])
            up = _normalize((rot[:3, :3] @ _v([0, 1, 0])).astype(np.float32))
            pos = (rot @ _v([0, 0, 1, 1]))[:3]
            cam.up = up
            cam.look_at = center
            cam.position = (pos * radius + center).astype(np.float32)
            cams.append(cam)
    return cams


def cinema_angles(phi, theta):
    """phi=N, theta=M -> angle lists of CinemaManager (rendering_filters.cpp:568-574,613-619)."""
    ph = [float(F(-180.0 + (360.0 / float(phi)) * a)) for a in range(phi)]
    th = [float(F(0.0 + (180.0 / float(theta)) * a)) for a in range(theta)]
    return ph, th
