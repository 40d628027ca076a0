"""tout_multi_topo_single_ghost_vol_render100.png (t_ascent_multi_topo.cpp:181-252) -> tests/golden/*.npz.
Run HERE (the container that has /root/reference).  No annotations in this golden: the whole frame is used."""
import os

import numpy as np
from PIL import Image

BASE = "/root/reference/src/tests/_baseline_images"
HERE = os.path.dirname(os.path.abspath(__file__))
img = np.array(Image.open(os.path.join(BASE, "tout_multi_topo_single_ghost_vol_render100.png")).convert("RGB"))[::-1].copy()
np.savez_compressed(os.path.join(HERE, "tout_multi_topo_single_ghost_vol_render100.npz"), rgb=img)
print(img.shape, os.path.getsize(os.path.join(HERE, "tout_multi_topo_single_ghost_vol_render100.npz")))
