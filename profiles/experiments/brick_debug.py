import numpy as np, sys, os
sys.path.insert(0, os.getcwd())
from ascent_b200 import _lib, color_table, datasets
from oracle import oracle as O
ctx = _lib.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 100
dom = datasets.braid_uniform(n, dtype=np.float32)
b = datasets.domain_bounds(dom)
cam = O.camera_reset_to_bounds(b)
O.camera_azimuth(cam, 35.0); O.camera_elevation(cam, 20.0)
lut = color_table.parse_color_table({"name": "cool to warm", "control_points": [{"type": "alpha", "position": 0., "alpha": 0.}, {"type": "alpha", "position": 1., "alpha": 1.}]}).corrected_opacity(S).lut()
sd = O.sample_distance(b, S)
W, H = 160, 120
ctx.set_tf(lut)
ctx.block_from_domain(0, dom)
ctx.canvas_clear(W, H)
ctx.trace_to_canvas(0, cam, sd, float(dom["field"].min()), float(dom["field"].max()), False)
rgba, depth = ctx.canvas_download(W, H)
o_rgba, o_depth = O.new_canvas(W, H)
from tests import scenes
O.render_to_canvas(scenes.oracle_block(dom), cam, W, H, lut, sd, float(dom["field"].min()), float(dom["field"].max()), o_rgba, o_depth, use_depth=False)
d = np.abs(rgba - o_rgba).max(axis=1)
print("covered", int((o_rgba[:, 3] > 0).sum()), "differing px", int((d > 0).sum()), "max", float(d.max()))
