#!/usr/bin/env python3
"""Small persistent-kernel debug run: Duck, a few sizes, counters vs the oracle."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import _oracle as O  # noqa: E402
import rayfinder_b200 as rf  # noqa: E402

pt = rf.PtFormat.loads(O.duck_pt_bytes())
for (w, h, bounces) in ((96, 64, 4), (200, 120, 8)):
    cam = rf.bvh_visualizer_camera(pt.bvh_nodes, w, h)
    params = rf.RenderParameters((w, h), cam, rf.SamplingParams(1, bounces), rf.Sky(), 1.0)
    ren = rf.ReferencePathTracer(params, (w, h), rf.SceneArrays.from_pt(pt))
    ren.render()
    img, _ = ren.read_hdr()
    orc = O.OracleRenderer(pt, w, h, rf.camera_to_array(cam), rf.sky_state(rf.Sky()), 1, bounces)
    orc.render()
    s, o = ren.stats(), orc.stats()
    print(w, h, bounces, {k: (s[k], o[k]) for k in O.COUNTER_NAMES}, "rmse", O.rmse(img, orc.image), flush=True)
    ren.close()
