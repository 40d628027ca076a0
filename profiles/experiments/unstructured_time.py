"""First timing of the unstructured producer (N4): a braid field on n^3 points written as (n-1)^3 warped hexahedra,
1920x1080, default camera, 100 samples -- publish (locator build) and partial trace, CUDA-event timed."""
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ascent_b200 import _lib, color_table, datasets  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 129
dom = datasets.braid_uniform(n, dtype=np.float32)
pts, conn = datasets.structured_to_hexes(dom["dims"], dom["origin"], dom["spacing"])
g = np.random.default_rng(0)
pts = pts + (g.random(pts.shape, dtype=np.float32) - 0.5) * np.float32(0.2 * dom["spacing"][0])
field = dom["field"].reshape(-1)
b = [float(pts[:, 0].min()), float(pts[:, 0].max()), float(pts[:, 1].min()), float(pts[:, 1].max()),
     float(pts[:, 2].min()), float(pts[:, 2].max())]
ctx = _lib.Context(0)
W, H = 1920, 1080
from ascent_b200 import camera as camera_mod  # noqa: E402
cam = camera_mod.Camera()
cam.reset_to_bounds(b)
cam.azimuth(30.0)
lut = color_table.parse_color_table({"name": "cool to warm", "control_points": [
    {"type": "alpha", "position": 0., "alpha": 0.}, {"type": "alpha", "position": 1., "alpha": 1.}]}).corrected_opacity(100).lut()
ctx.set_tf(lut)
sd = _lib.sample_distance(b, 100)
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
import time
h0 = time.perf_counter()
ctx.block_unstructured(0, pts, conn, field)
publish_ms = (time.perf_counter() - h0) * 1e3
h0 = time.perf_counter()
ctx.block_unstructured(0, pts, conn, field)  # same topology again: the external-face mask is reused
republish_ms = (time.perf_counter() - h0) * 1e3
rmin, rmax = float(field.min()), float(field.max())
times = []
for it in range(12):
    ctx.canvas_clear(W, H)
    ctx.partials_begin(W, H)
    ctx.synchronize()
    h0 = time.perf_counter()
    ctx.trace_to_partials(0, cam, sd, rmin, rmax, False)
    ctx.synchronize()
    times.append((time.perf_counter() - h0) * 1e3)
npart = ctx.partials_count()
print(json.dumps({"cells": int(conn.shape[0]), "points": int(pts.shape[0]), "image": [W, H], "publish_ms_incl_h2d_and_locator": publish_ms, "republish_ms_same_topology": republish_ms,
                  "trace_ms": float(np.median(times[1:])), "trace_ms_min": float(np.min(times[1:])), "partials": int(npart),
                  "lib": os.environ.get("VR_LIB_NAME", "libvr_b200.so"),
                  "mrays_per_s": W * H / (float(np.median(times[1:])) * 1e-3) / 1e6}))
