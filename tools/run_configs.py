#!/usr/bin/env python3
"""Run BASELINE.json configs[1], [2] and [4] on one B200 and record the results (profiles/r01_configs.json):

  [1] Sponza 1920x1080, 1 spp, 8 bounces           — time + HDR RMSE and work counters vs the oracle at FULL size
  [2] Sponza 1920x1080, 64 spp accumulated          — time, RMSE of the accumulated buffer vs the oracle's 64 frames,
                                                      convergence of the running mean
  [4] throughput sweep: 1/2/4/8/16 bounces x 256^2..4096^2 (1 GPU part; the 8-GPU part is tools/check_multigpu.py +
      bench.py under torchrun)

The oracle legs are the checker (tests-side code); nothing here is a bench value — bench.py is.
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import _oracle as O  # noqa: E402
import rayfinder_b200 as rf  # noqa: E402
from rayfinder_b200 import assets as rfa  # noqa: E402


def run_sweep(ren, quick):
    """configs[4], single-GPU part: square frames x bounce counts, automatic schedule."""
    sweep = []
    for size in ((256, 1024) if quick else (256, 512, 1024, 2048, 4096)):
        for b in (1, 2, 4, 8, 16):
            params = rf.RenderParameters((size, size), rf.fly_camera(size, size), rf.SamplingParams(1, b), rf.Sky(), 0.25)
            frames = 5
            for k in range(frames + 1):
                if k == 1:
                    ren.reset_stats()
                params.exposure = 0.25 + 0.01 * k
                ren.set_render_parameters(params)
                ren.render()
            s = ren.stats()
            rays = (s["closest_rays"] + s["shadow_rays"]) / frames
            sweep.append({"size": size, "bounces": b, "ms_per_frame": s["device_ms_total"] / frames, "rays_per_frame": rays,
                          "mrays_s": rays / (s["device_ms_total"] / frames) / 1e3, "sub_frames": s["sub_frames"], "evict_max": s["evict_max"]})
            print("sweep", sweep[-1], flush=True)
    return sweep


def main():
    quick = "--quick" in sys.argv
    sweep_only = "--sweep-only" in sys.argv  # just configs[4]'s single-GPU sweep -> gpurun_out/sweep_1gpu.json
    pt = rfa.load_scene("Sponza")
    scene = rf.SceneArrays.from_pt(pt)
    sky = rf.sky_state(rf.Sky())
    out = {"scene": "Sponza.pt", "gpu": "B200", "host_threads": O.num_threads()}
    if sweep_only:
        params = rf.RenderParameters((1920, 1080), rf.fly_camera(1920, 1080), rf.SamplingParams(1, 8), rf.Sky(), 0.25)
        ren = rf.ReferencePathTracer(params, (4096, 4096), scene)
        out["config4_sweep_1gpu"] = run_sweep(ren, False)
        (ROOT / "gpurun_out").mkdir(exist_ok=True)
        (ROOT / "gpurun_out" / "sweep_1gpu.json").write_text(json.dumps(out, indent=1))
        return

    # ---- configs[1] ------------------------------------------------------------------------------------
    w, h, bounces = 1920, 1080, 8
    cam = rf.fly_camera(w, h)
    params = rf.RenderParameters((w, h), cam, rf.SamplingParams(1, bounces), rf.Sky(), 0.25)
    ren = rf.ReferencePathTracer(params, (4096, 4096), scene)
    ren.render()
    img, _ = ren.read_hdr()
    s = ren.stats()
    orc = O.OracleRenderer(pt, w, h, rf.camera_to_array(cam), sky, 1, bounces)
    t0 = time.perf_counter()
    orc.render()
    oracle_s = time.perf_counter() - t0
    o = orc.stats()
    out["config1_1080p_1spp_8b"] = {
        "rmse_vs_oracle": O.rmse(img, orc.image), "max_abs_err": float(np.nanmax(np.abs(img[..., :3] - orc.image[..., :3]))),
        "counters_equal": all(s[k] == o[k] for k in O.COUNTER_NAMES), "counters": {k: s[k] for k in O.COUNTER_NAMES},
        "gpu_ms": s["device_ms_total"], "oracle_seconds": oracle_s,
        "oracle_mrays_s": (o["closest_rays"] + o["shadow_rays"]) / oracle_s / 1e6,
        "hdr_mean": float(img[..., :3].mean()), "hdr_max": float(img[..., :3].max())}
    print("configs[1]", json.dumps(out["config1_1080p_1spp_8b"]), flush=True)

    # ---- configs[2] ------------------------------------------------------------------------------------
    spp = 8 if quick else 64
    params = rf.RenderParameters((w, h), cam, rf.SamplingParams(spp, bounces), rf.Sky(), 0.25)
    ren.set_render_parameters(params)
    ren.set_frame_count(0)
    ren.reset_stats()
    running = {}
    for k in range(spp):
        ren.render()
        if (k + 1) in (1, 2, 4, 8, 16, 32, 64):
            acc, n = ren.read_hdr()
            running[k + 1] = acc[..., :3] / float(n)
    img, n = ren.read_hdr()
    s = ren.stats()
    orc = O.OracleRenderer(pt, w, h, rf.camera_to_array(cam), sky, spp, bounces)
    t0 = time.perf_counter()
    for _ in range(spp):
        orc.render()
    oracle_s = time.perf_counter() - t0
    o = orc.stats()
    final = running[spp]
    out[f"config2_1080p_{spp}spp_8b"] = {
        "accumulated": n, "rmse_sum_buffer_vs_oracle": O.rmse(img, orc.image),
        "rmse_mean_vs_oracle": O.rmse(img / float(n), orc.image / float(n)),
        "counters_equal": all(s[k] == o[k] for k in O.COUNTER_NAMES), "gpu_ms_total": s["device_ms_total"],
        "gpu_ms_per_sample": s["device_ms_total"] / spp, "mrays_s": (s["closest_rays"] + s["shadow_rays"]) / s["device_ms_total"] / 1e3,
        "oracle_seconds": oracle_s,
        "convergence_rmse_of_k_spp_mean_vs_final": {str(k): float(np.sqrt(np.mean((v - final) ** 2))) for k, v in running.items() if k < spp}}
    print("configs[2]", json.dumps(out[f"config2_1080p_{spp}spp_8b"]), flush=True)

    out["config4_sweep_1gpu"] = run_sweep(ren, quick)
    if not quick:
        (ROOT / "profiles" / "r01_configs.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
