#!/usr/bin/env python3
"""Lane occupancy of the traversal warps over the time of one frame (GPU box, instrumented debug build).

    python -m rayfinder_b200._build --timeline
    python tools/occupancy_timeline.py [WxH] [sub_frames] [evict_max] [persistent_kernel] [name=value ...]

Per 16.4 us bucket: loop rounds of all traversal warps, and the mean number of lanes (of 32) that held a ray.
"""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("RAYFINDER_B200_LIB", str(ROOT / "rayfinder_b200" / "librayfinder_b200_timeline.so"))

import rayfinder_b200 as rf  # noqa: E402
from rayfinder_b200 import assets as rfa  # noqa: E402
from rayfinder_b200 import capi  # noqa: E402


def main():
    size = sys.argv[1] if len(sys.argv) > 1 else "672x384"
    sub = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    evict = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    mega = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    options = [a.split("=") for a in sys.argv[5:]]
    w, h = (int(x) for x in size.split("x"))
    capi.lib()
    raw = C.CDLL(str(capi.LIB_PATH))
    raw.rf_debug_timeline_arm.argtypes = [C.c_uint32]
    raw.rf_debug_occupancy_read.argtypes = [C.c_void_p]
    pt = rfa.load_scene("Sponza")
    params = rf.RenderParameters((w, h), rf.fly_camera(w, h), rf.SamplingParams(1, 8), rf.Sky(), 0.25)
    ren = rf.ReferencePathTracer(params, (w, h), rf.SceneArrays.from_pt(pt))
    ren.set_pipeline(sub, mega, 3, 256)
    ren.set_tail_policy(evict)
    for name, value in options:
        ren.set_option(name, int(value))
    for k in range(3):
        params.exposure = 0.25 + 0.01 * k
        ren.set_render_parameters(params)
        ren.render()
    ren.synchronize()
    raw.rf_debug_timeline_read.argtypes = [C.c_void_p, C.c_uint32]
    raw.rf_debug_timeline_read.restype = C.c_uint32
    assert raw.rf_debug_timeline_arm(1 << 16) == 0
    ren.reset_stats()
    params.exposure = 0.3
    ren.set_render_parameters(params)
    ren.render()
    ren.synchronize()
    stats = ren.stats()
    buf = np.zeros(512, dtype=np.uint64)
    raw.rf_debug_occupancy_read(buf.ctypes.data)
    busy, rounds = buf[:256].astype(np.float64), buf[256:].astype(np.float64)
    used = np.nonzero(rounds)[0]
    # the buckets are absolute (mod 4.2 ms): rotate so that the longest empty stretch comes first
    gaps = np.diff(np.concatenate([used, [used[0] + 256]]))
    first = (used[np.argmax(gaps)] + int(gaps.max())) % 256
    print(f"{w}x{h} sub_frames={sub} evict_max={evict} persistent={mega} options={options}: frame {stats['device_ms_total']:.3f} ms")
    print("  t [us]   rounds   lanes/32   share of all lane-rounds")
    total = busy.sum()
    for k in range(256):
        b = (first + k) % 256
        if rounds[b] == 0:
            continue
        print(f"  {k * 16.384:7.0f} {int(rounds[b]):8d}   {busy[b] / rounds[b]:6.2f}   {100 * busy[b] / total:5.1f} %")
    REC = np.dtype([("tag", "<u8"), ("start", "<u8"), ("dry", "<u8"), ("exit", "<u8"), ("rays", "<u4"), ("rounds", "<u4"), ("sm", "<u4"), ("pad", "<u4")])
    rec = np.zeros(1 << 16, dtype=REC)
    rec = rec[:raw.rf_debug_timeline_read(rec.ctypes.data, 1 << 16)]
    shade = rec[rec["tag"] == 2]
    if len(shade):
        span = (shade["exit"] - shade["start"]).astype(np.float64)
        busy = shade["dry"].astype(np.float64)
        t0 = shade["start"].min()
        end = (shade["exit"] - t0) / 1e3
        print(f"shading warps: {len(shade)}; busy (generating / shading) {100 * busy.sum() / span.sum():.1f} % of their time (max {100 * (busy / span).max():.1f} %); "
              f"{shade['rays'].sum() / max(1, shade['rounds'].sum()):.1f} entries per batch, {busy.sum() / max(1, shade['rounds'].sum() + shade['pad'].sum() / 32) / 1e3:.2f} us per batch; "
              f"paths per block min {shade['pad'].min()} max {shade['pad'].max()}; block end [us] p10 {np.percentile(end, 10):.0f} p50 {np.percentile(end, 50):.0f} "
              f"p90 {np.percentile(end, 90):.0f} p99 {np.percentile(end, 99):.0f} max {end.max():.0f}")
    ren.close()


if __name__ == "__main__":
    main()
