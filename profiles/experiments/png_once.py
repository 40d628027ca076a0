"""One c2-like frame (braid 256^3, 1920x1080) encoded as PNG on the device a few times: the target of the ncu capture
of png_rows_kernel / png_finish_kernel / png_compact_kernel."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ascent_b200 import _lib, camera, color_table, datasets  # noqa: E402

dom = datasets.braid_uniform(256, dtype=np.float32)
b = datasets.domain_bounds(dom)
ctx = _lib.Context(0)
W, H = 1920, 1080
cam = camera.Camera()
cam.reset_to_bounds(b)
lut = color_table.parse_color_table({"name": "cool to warm", "control_points": [
    {"type": "alpha", "position": 0., "alpha": 0.}, {"type": "alpha", "position": 1., "alpha": 1.}]}).corrected_opacity(100).lut()
ctx.set_tf(lut)
ctx.block_from_domain(0, dom)
ctx.trace_to_image(0, cam, W, H, _lib.sample_distance(b, 100), float(dom["field"].min()), float(dom["field"].max()),
                   write_canvas=True)
for _ in range(4):
    png = ctx.canvas_encode_png(W, H, np.array([1., 1., 1., 1.], np.float32))
print("png bytes", len(png))
