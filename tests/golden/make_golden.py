"""Extract golden vectors from the reference's own baseline PNGs into tests/golden/*.npz.

Run HERE (the container that has /root/reference):  python tests/golden/make_golden.py
The GPU box has no /root/reference, so tests read only the committed .npz files.

Sources (all under /root/reference/src/tests/_baseline_images/):
  apcomp/apcomp_{c_order,volume_partial,zbuffer}{,_mpi}.png  <- t_apcomp_{c_order,volume_partials,
                                                              zbuffer}{,_mpi}.cpp
  render_0100.png, render_1100.png   <- t_ascent_render_3d.cpp:1643-1780 (test_render_3d_multi_render)
  tout_render_mpi_3d_diy_volume100.png <- t_ascent_mpi_render_3d.cpp:284-388
  tout_render_3d_multi_default_runtime100.png <- t_ascent_render_3d.cpp:1017-1122 (colour bars, and the pixels
      whose rays meet no contour surface: the scene's opaque part needs the contour filter and the surface ray
      tracer, which are outside the volume path)

Colour bars: VTK-m's ColorBarAnnotation draws ColorTable::Sample(bar height) of the plot's table into
the image (one table sample per pixel row), so one pixel column of a bar is a golden vector for K8:
  colorbars.npz: cool_to_warm_179 (mpi volume golden), cool_to_warm_359 and rainbow_desaturated_359
  (multi_default_runtime golden: the pseudocolor plot's default table and the volume plot's
  "rainbow desaturated"), each n x 3 uint8, index 0 = table position 0.

PNG rows are flipped vertically on save (ascent_png_encoder.cpp:95-96): stored arrays are
un-flipped, i.e. row j of the array is image row y=j in canvas/test coordinates.
The ascent goldens carry annotations (axes, colour bar); `rects` lists hand-picked
annotation-free crops (y0,y1,x0,x1 in canvas coordinates) over which the oracle is compared.
"""
import os

import numpy as np
from PIL import Image

BASE = "/root/reference/src/tests/_baseline_images"
HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    return np.array(Image.open(os.path.join(BASE, name)).convert("RGBA"))[::-1].copy()


def main():
    ap = {}
    for n in ["apcomp_c_order", "apcomp_c_order_mpi", "apcomp_volume_partial",
              "apcomp_volume_partial_mpi", "apcomp_zbuffer", "apcomp_zbuffer_mpi"]:
        ap[n] = load("apcomp/%s.png" % n)
    np.savez_compressed(os.path.join(HERE, "apcomp_goldens.npz"), **ap)

    def flip_rects(rects, H):
        # rects were picked on the PNG (top-down) -> canvas coordinates (bottom-up)
        return np.array([[H - y1, H - y0, x0, x1] for (y0, y1, x0, x1) in rects], np.int32)

    np.savez_compressed(os.path.join(HERE, "render_0100.npz"), rgb=load("render_0100.png")[..., :3],
                        rects=flip_rects([(100, 420, 95, 419)], 512))  # (column 419 carries an axis tick)
    np.savez_compressed(os.path.join(HERE, "render_1100.npz"), rgb=load("render_1100.png")[..., :3],
                        rects=flip_rects([(130, 380, 170, 330)], 400))
    np.savez_compressed(os.path.join(HERE, "tout_render_mpi_3d_diy_volume100.npz"),
                        rgb=load("tout_render_mpi_3d_diy_volume100.png")[..., :3],
                        rects=flip_rects([(200, 350, 100, 250), (170, 350, 262, 420)], 512))
    vol = load("tout_render_mpi_3d_diy_volume100.png")
    multi = load("tout_render_3d_multi_default_runtime100.png")
    # the whole frame of the pseudocolor + volume golden: where no ray meets the contour surface the pixels are the
    # volume plot's alone (tests/test_oracle_golden.py finds those pixels itself)
    np.savez_compressed(os.path.join(HERE, "tout_render_3d_multi_default_runtime100.npz"), rgb=multi[..., :3])
    # (arrays are un-flipped: row index grows upwards, like the table position along a vertical bar)
    np.savez_compressed(os.path.join(HERE, "colorbars.npz"),
                        cool_to_warm_179=vol[512 - 230:512 - 51, 481, :3],
                        cool_to_warm_359=multi[1024 - 461:1024 - 102, 975, :3],
                        rainbow_desaturated_359=multi[1024 - 922:1024 - 563, 975, :3])
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
