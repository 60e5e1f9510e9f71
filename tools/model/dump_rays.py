#!/usr/bin/env python3
"""Dump the rays of every `stride`-th tile of the benchmark frame (and the scene) for tools/model/pair_model.cpp."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np  # noqa: E402

import _oracle as O  # noqa: E402

out = Path(sys.argv[1])
stride = int(sys.argv[2]) if len(sys.argv) > 2 else 16
out.mkdir(parents=True, exist_ok=True)
pt = O.NumpyPt.load_scene("Sponza")
w, h = 1920, 1080
rays, kinds = O.frame_rays(pt, w, h, O.fly_camera_array(w, h), O.default_sky_state(), 1, 8, tile_stride=stride)
pt.bvh_nodes.tofile(out / "nodes.bin")
pt.bvh_position_attributes.astype("<f4").tofile(out / "tris9.bin")
rays.tofile(out / "rays.bin")
kinds.tofile(out / "kinds.bin")
print(len(kinds), "rays")
