#!/usr/bin/env python3
"""Benchmark of the render path: Mrays/s on Sponza 1920x1080, 1 spp, 8 bounces (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU traversal on the host cores

A step = one frame (one sample per pixel, full paths) of the wavefront path tracer over the whole image.
For N > 1 the frame is split by 32x32 tile over the ranks and the HDR buffer is sum-reduced to rank 0 with
NCCL inside the step.  A ray = one rayIntersectBvh (closest hit) or shadowRay (any hit) call of the
reference shader; the kernels count them exactly.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

METRIC = "Mrays/s (Sponza 1080p, 8 bounces)"
DATA_NOTE = "Sponza.pt baked from the reference's assets/Sponza.glb (deterministic asset, not synthetic)"
UNIT = "Mrays/s"
BOUNCES = 8
CONFIGS = {"1080p": (1920, 1080), "4k": (3840, 2160)}


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--config", default="1080p", choices=sorted(CONFIGS),
                   help="1080p = BASELINE.json configs[1] (the metric's configuration); 4k = configs[3] (3840x2160, meant for --gpus 8)")
    p.add_argument("--width", type=int, default=None)
    p.add_argument("--height", type=int, default=None)
    p.add_argument("--bounces", type=int, default=BOUNCES)
    p.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"], help="multi-GPU exchange step (rayfinder_b200/distributed.py)")
    p.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU work per bounded reference sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    args = p.parse_args()
    cw, ch = CONFIGS[args.config]
    args.width = args.width or cw
    args.height = args.height or ch
    return args


# ---- clocks -------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines: list[tuple[float, str]] = []  # (arrival time, csv line)
        self.begin = None

    def start(self):
        # NVML in-process (a sample every 2 ms) when available; the nvidia-smi loop otherwise.
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)  # probe
            self.stop_flag = threading.Event()
            self.proc = "nvml"
            threading.Thread(target=self._pump_nvml, daemon=True).start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _physical_index(self) -> int:
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        if visible:
            ids = [x.strip() for x in visible.split(",") if x.strip()]
            if self.gpu_index < len(ids) and ids[self.gpu_index].isdigit():
                return int(ids[self.gpu_index])
        return self.gpu_index

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.monotonic(), line.strip()))

    def _pump_nvml(self):
        n = self.nvml
        flags = [("hw_slowdown", n.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", n.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", n.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", n.nvmlClocksEventReasonSwPowerCap)]
        try:
            sm_max = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        except Exception:
            sm_max = 0
        while not self.stop_flag.is_set():
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                try:
                    mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                # same 9-column layout as the nvidia-smi query
                cols = [str(self.gpu_index), str(sm), str(sm_max), "", hex(mask)] + ["Active" if mask & bit else "Not Active" for _, bit in flags]
                self.lines.append((time.monotonic(), ",".join(cols)))
            except Exception:
                pass
            time.sleep(0.002)

    def mark_begin(self):
        """The timed region starts now (the sampler was started before the warm-up: nvidia-smi takes longer to come up
        than a multi-GPU timed region lasts)."""
        self.begin = time.monotonic()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        end = time.monotonic()
        if self.proc == "nvml":
            self.stop_flag.set()
        else:
            time.sleep(0.05)
            self.proc.terminate()
        begin = self.begin if self.begin is not None else 0.0
        inside = [line for t, line in self.lines if begin <= t <= end + 0.03]
        note = None
        if not inside:
            # a region shorter than one sampling period: the samples of the warm-up steps right before it (same load)
            inside = [line for t, line in self.lines if begin - 1.0 <= t]
            note = "timed region shorter than the sampling period: samples taken during the warm-up steps just before it"
        sm, sm_max, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in inside:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                sm_max.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        out = {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(sm_max) if sm_max else None,
               "samples": len(sm), "reasons": sorted(reasons), "source": "nvml" if self.proc == "nvml" else "nvidia-smi"}
        if note:
            out["note"] = note
        return out


# ---- the reference's CPU traversal ------------------------------------------------------------------------
class CpuReference:
    """The reference's own CPU implementation of the path's hot loop — nlrs::rayIntersectBvh
    (common/ray_intersection.cpp:138-213), compiled from the reference's sources into oracle/_ref (kind "reference"; the
    oracle's port when that library is absent, kind "port") — on THE SAME WORKLOAD as the GPU arm: the rays the path tracer
    traces for this frame (the closest-hit ray of every bounce and the shadow ray of every hit), generated once, untimed,
    by the oracle's restatement of the WGSL path tracer for every `tile_stride`-th 32x32 tile of the frame.  The reference
    has no CPU any-hit traversal (shadowRay exists in WGSL only), so shadow rays are traced as closest-hit rays too.
    Loads the scene with numpy and builds the camera with the reference's createCamera: nothing of rayfinder_b200 is
    imported or mapped."""

    def __init__(self, width, height, bounces, tile_stride=16):
        import _oracle as O

        self.O = O
        self.pt = O.NumpyPt.load_scene("Sponza")
        if self.pt is None:
            raise SystemExit("bench.py: assets/Sponza.pt[.xz] is missing (bake it with __graft_entry__.build() where the reference is mounted)")
        self.width, self.height, self.bounces, self.tile_stride = width, height, bounces, tile_stride
        self.nodes = np.ascontiguousarray(self.pt.bvh_nodes)
        self.tris = np.ascontiguousarray(self.pt.bvh_position_attributes)
        self.cam = O.fly_camera_array(width, height)
        self.sky = O.default_sky_state()
        self.kind = "reference" if O.have_ref() else "port"
        self.cores = O.num_threads()
        self.rays, kinds = O.frame_rays(self.pt, width, height, self.cam, self.sky, 1, bounces, tile_stride=tile_stride)
        self.shadow_fraction = float(kinds.mean()) if kinds.size else 0.0
        self.t_max = 10000.0  # T_MAX, wgsl:73

    def trace(self, passes: int = 1, threads: int | None = None, limit: int | None = None):
        """`passes` passes over the ray sample (its first `limit` rays); returns (rays, seconds)."""
        fn = self.O.ref_intersect if self.kind == "reference" else self.O.oracle_intersect
        rays = self.rays if limit is None else self.rays[:limit]
        t0 = time.perf_counter()
        for _ in range(passes):
            fn(self.nodes, self.tris, rays, self.t_max, threads=threads or self.cores)
        return passes * len(rays), time.perf_counter() - t0

    def calibrate(self, seconds: float) -> int:
        """Passes over the sample so that one bounded sample is about `seconds` of wall time."""
        n = min(len(self.rays), 200_000)
        rays, secs = self.trace(1, limit=n)
        rate = rays / max(secs, 1e-9)
        return max(1, int(rate * seconds / max(1, len(self.rays))))

    def primary_rays(self, seconds: float):
        """bvh-visualizer's own loop (bvh-visualizer/main.cpp:60-78) over the primary rays of the view, as round 1 reported."""
        fn = self.O.ref_node_counts if self.kind == "reference" else self.O.oracle_node_counts
        rows, rays, secs = 64, 0, 0.0
        while secs < seconds:
            for band in range(8):
                r0 = min(self.height - rows // 8, (band * self.height) // 8)
                _, s = fn(self.nodes, self.tris, self.cam, self.width, self.height, 3.4028234663852886e38, threads=self.cores, rows=(r0, r0 + rows // 8))
                rays, secs = rays + (rows // 8) * self.width, secs + s
            rows = min(rows * 4, self.height)
        return rays / secs / 1e6

    def full_path_port(self):
        """The oracle's restatement of the whole WGSL path tracer (kind "port": the reference has no CPU path tracer) on
        one full frame of the view: (Mrays/s including shading, sample description, exact rays of the frame)."""
        orc = self.O.OracleRenderer(self.pt, self.width, self.height, self.cam, self.sky, 1, self.bounces)
        orc.render()
        st = orc.stats()
        rays = st["closest_rays"] + st["shadow_rays"]
        return rays / orc.seconds / 1e6, f"one {self.width}x{self.height} frame, {self.cores} threads", rays

    def sample_description(self, passes: int, threads: int | None = None, limit: int | None = None) -> str:
        n = len(self.rays) if limit is None else min(limit, len(self.rays))
        return (f"{passes} pass(es) over {n} rays = the closest-hit + shadow rays ({100 * self.shadow_fraction:.0f}% shadow) of every "
                f"{self.tile_stride}th 32x32 tile of the {self.width}x{self.height}, {self.bounces}-bounce frame, each through the reference's "
                f"rayIntersectBvh (closest-hit; the CPU code has no any-hit mode), tmax 10000, {threads or self.cores} thread(s)")


def base_config(width, height, bounces, rays_per_step, paths_per_step, l2, pipeline, partition):
    """The `config` object of both arms: same keys, so the two lines describe the same workload."""
    return {"workload": f"Sponza.pt {width}x{height}, 1 spp, {bounces} bounces, interior fly-camera default view",
            "rays_per_step": rays_per_step, "paths_per_step": paths_per_step, "l2": l2, "pipeline": pipeline, "partition": partition}


def cpu_baseline_block(cpu: CpuReference, seconds: float) -> dict:
    passes = cpu.calibrate(seconds)
    rays, secs = cpu.trace(passes)
    one_n = min(len(cpu.rays), 100_000)
    one_rays, one_secs = cpu.trace(1, threads=1, limit=one_n)
    port_value, port_sample, _ = cpu.full_path_port()
    return {"value": rays / secs / 1e6, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind, "sample": cpu.sample_description(passes),
            "one_core": {"value": one_rays / one_secs / 1e6, "unit": UNIT, "cores": 1, "sample": cpu.sample_description(1, 1, one_n),
                         "note": "the reference's loop is single-threaded (bvh-visualizer/main.cpp:60-78)"},
            "primary_rays_only": {"value": cpu.primary_rays(2.0), "unit": UNIT, "cores": cpu.cores,
                                  "sample": "bvh-visualizer pixel loop over the view's primary rays, tmax FLT_MAX (round 1's figure)"},
            "full_path_port": {"value": port_value, "unit": UNIT, "cores": cpu.cores, "kind": "port", "sample": port_sample}}


def run_reference(args, rank: int):
    if rank != 0:
        return
    cpu = CpuReference(args.width, args.height, args.bounces)
    # K timed steps + W warm-up steps must end within a few minutes: bound each step's sample
    budget = min(args.cpu_seconds, 150.0 / max(1, args.steps + args.warmup))
    passes = cpu.calibrate(budget)
    for _ in range(args.warmup):
        cpu.trace(passes)
    rays = secs = 0.0
    for _ in range(args.steps):
        r, s = cpu.trace(passes)
        rays += r
        secs += s
    value = rays / secs / 1e6
    port_value, port_sample, frame_rays = cpu.full_path_port()  # (also the exact ray count of the frame, as the GPU arm reports it)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, args.steps), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": DATA_NOTE,
        "config": base_config(args.width, args.height, args.bounces, frame_rays, args.width * args.height, "host caches, not flushed",
                              f"reference CPU rayIntersectBvh over a bounded sample of the frame's rays per step, {cpu.cores} host threads",
                              "host only (rank 0)"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind, "sample": cpu.sample_description(passes),
                         "full_path_port": {"value": port_value, "unit": UNIT, "cores": cpu.cores, "kind": "port", "sample": port_sample}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- this repo's CUDA path ----------------------------------------------------------------------------------
def load_scene():
    from rayfinder_b200 import assets as rfa

    if rfa.scene_path("Sponza") is None:
        raise SystemExit("bench.py: assets/Sponza.pt[.xz] is missing (bake it with __graft_entry__.build() where the reference is mounted)")
    return rfa.load_scene("Sponza")


def ncu_reference(world: int, w: int, h: int) -> dict | None:
    """The committed ncu capture this line's roofline refers to: profiles/r02_roofline_ncu.json holds, per frame share
    (key = pixels a GPU owns), the dominant kernel's per-launch DRAM/L2 traffic and its pipe utilisations."""
    path = ROOT / "profiles" / "r02_roofline_ncu.json"
    if not path.exists():
        return None
    table = json.loads(path.read_text())
    return table.get("captures", {}).get(f"{w}x{h}/{world}")


def build_roofline(args, world, stats, stage_stats, clocks, num_sms, w, h):
    """roofline of the dominant kernel (k_trace, rank 0's launches).  Two denominators:
    * the contract's: algorithmic bytes (48 B per node visited + 48 B per triangle tested, counted exactly by
      the kernel) / launch time against the measured HBM copy bandwidth -> `hbm_algorithmic_frac`.  The 28.7 MB node +
      triangle set is L1/L2-resident, so this exceeds 1: HBM is not what bounds the kernel (DRAM traffic per launch is ~1% of
      the algorithmic bytes, `traffic`).
    * the binding one: the SM's L1TEX stage accepts one wavefront (one 128-byte line of one request) per clock.  Every node
      record a lane loads and each of the two loads of a triangle test (LDG.256 + LDG.32) is one wavefront (divergent lanes, distinct
      lines), the kernel counts those loads exactly, so achieved = wavefronts / launch time and peak = SMs x SM clock
      (the clock sampled under load).  This counts only the BVH loads — queue, stack and result traffic add about a quarter on
      top (ncu: l1tex__data_pipe_lsu_wavefronts, `ncu.l1tex_lsu_data_pipe_pct`), so the live figure is a lower bound of the
      pipe's utilisation."""
    peaks = {}
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peaks = json.loads(peaks_path.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    bounces = args.bounces
    persistent = bool(stats.get("persistent_kernel"))
    if persistent:
        # the timed steps ran as ONE persistent launch per frame (k_mega: ray generation, traversal, shading, compaction) plus
        # the accumulation kernel (< 1 % of the step): the dominant kernel's launch time is the step's device time
        src, kernel = stats, "k_mega"
        trace_ms = stats["device_ms_total"]
        launches = args.steps
    else:
        src, kernel = stage_stats, "k_trace"
        trace_ms = stage_stats["device_ms_trace"]
        launches = args.steps * (bounces + 1)
    nodes = src["closest_nodes_visited"] + src["shadow_nodes_visited"]
    tris = src["closest_triangles_tested"] + src["shadow_triangles_tested"]
    node_loads = src.get("node_records_loaded") or nodes
    if trace_ms <= 0:
        return {"bound": "l1tex", "kernel": "k_trace", "achieved": None, "peak": None, "unit": "Gwavefronts/s", "frac": None, "traffic": None}
    seconds = trace_ms * 1e-3
    algorithmic = 48 * (nodes + tris)
    hbm_achieved = algorithmic / seconds / 1e9
    wavefronts = node_loads + 2 * tris
    sm_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0
    l1_peak = num_sms * sm_mhz * 1e6 / 1e9          # Gwavefronts/s
    l1_achieved = wavefronts / seconds / 1e9
    ncu = ncu_reference(world, w, h)
    out = {
        "bound": "l1tex", "kernel": kernel, "achieved": l1_achieved, "peak": l1_peak, "unit": "Gwavefronts/s", "frac": l1_achieved / l1_peak,
        "traffic": ncu.get("dram_bytes_per_launch") if ncu else None,
        "lts_bytes_per_launch": ncu.get("lts_bytes_per_launch") if ncu else None,
        "ncu": ncu,
        "hbm_algorithmic_frac": hbm_achieved / hbm_peak, "hbm_algorithmic_achieved_gbs": hbm_achieved, "hbm_peak_gbs": hbm_peak,
        "hbm_peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6.65 TB/s",
        "note": ("bound = the L1TEX LSU data pipe (1 wavefront / clk / SM): achieved = the kernel's own count of BVH loads (node records loaded + "
                 "2 per triangle tested), a lower bound of the pipe's wavefronts (ncu.l1tex_lsu_data_pipe_pct is the full figure of the "
                 "committed capture); peak = SMs x SM clock under load.  hbm_algorithmic_* is the contract figure (48 B per node visit "
                 "+ 48 B per triangle test over the measured HBM copy bandwidth): it exceeds 1 because the 28.7 MB working set is cache-"
                 "resident — HBM does not bound this kernel (traffic = DRAM bytes per launch from the committed ncu capture)"),
        "algorithmic_bytes_per_launch": algorithmic / launches, "wavefronts_per_launch": wavefronts / launches,
        "avg_launch_ms": trace_ms / launches, "launches": launches,
        "stage_ms_per_step": {k: stage_stats[f"device_ms_{k}"] / args.steps for k in ("trace", "shade", "other")},
        "stage_loop_ms_per_step": stage_stats["device_ms_total"] / args.steps,
        "stage_loop_schedule": "one tile set on one stream, one launch per stage (stage events need the stages back to back); "
                               + ("`value` is measured with the automatic schedule: one persistent launch per frame (k_mega), which the roofline "
                                  "above describes — the stage_* figures are the staged pipeline on the same frame, for comparison" if persistent else
                                  f"`value` is measured with the automatic schedule ({stats['sub_frames']} tile set(s))"),
        "mean_nodes_per_closest_ray": stats["closest_nodes_visited"] / max(1, stats["closest_rays"]),
        "mean_nodes_per_shadow_ray": stats["shadow_nodes_visited"] / max(1, stats["shadow_rays"]),
    }
    return out


def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist

    import rayfinder_b200 as rf
    from rayfinder_b200 import capi
    from rayfinder_b200 import distributed as rfd

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the rayfinder_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    pt = load_scene()
    w, h, bounces = args.width, args.height, args.bounces
    cam = rf.fly_camera(w, h)
    params = rf.RenderParameters((w, h), cam, rf.SamplingParams(1, bounces), rf.Sky(), 0.25)
    ren = rf.ReferencePathTracer(params, (w, h), rf.SceneArrays.from_pt(pt), device=local_rank)
    ren.set_tile_partition(rank, world)
    stream = torch.cuda.current_stream(dev)
    ren.set_stream(stream.cuda_stream)
    exchange = rfd.HdrExchange(ren, w, h, mode=args.exchange)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    host_hdr = torch.empty((h, w, 4), dtype=torch.float32).pin_memory()
    host_np = host_hdr.numpy()
    h2d_bytes = len(bytes(params.to_c()))

    toggle = [0]

    def new_frame():
        # 1 spp per step: any parameter change restarts the accumulation (reference_path_tracer.cpp:556-563), so
        # every step traces the same full frame again.
        toggle[0] ^= 1
        params.exposure = 0.25 if toggle[0] else 0.5
        ren.set_render_parameters(params)

    def step_device():
        new_frame()
        ren.render()
        return exchange()

    # End-to-end step: parameters in, HDR image out, every step.  The device -> host read of frame k runs on a copy stream
    # while frame k + 1 is traced: the image is first copied device -> device into one of two staging buffers on the
    # rendering stream (66 MB of traffic, ~10 us), the copy stream reads that buffer into one of two pinned host buffers.
    # A frame's host image is complete when its copy event has fired; the timed region ends after the last one.
    copy_stream = torch.cuda.Stream(dev)
    staging = [torch.empty((h, w, 4), dtype=torch.float32, device=dev) for _ in range(2)]
    host_bufs = [host_hdr, torch.empty((h, w, 4), dtype=torch.float32).pin_memory()]
    staged = [torch.cuda.Event() for _ in range(2)]
    copied = [None, None]
    e2e_index = [0]

    def step_e2e():
        k = e2e_index[0] & 1
        e2e_index[0] += 1
        new_frame()  # host -> device: the render parameters (uniform block) travel with the launch
        ren.render()
        full = exchange()
        if rank == 0:
            if copied[k] is not None:
                copied[k].synchronize()  # the host buffer (and the staging buffer) of two frames ago is free again
            staging[k].copy_(full)       # rendering stream
            staged[k].record(stream)
            copy_stream.wait_event(staged[k])
            with torch.cuda.stream(copy_stream):
                host_bufs[k].copy_(staging[k], non_blocking=True)  # device -> host: the HDR image (the exchanged frame on a multi-GPU run)
                copied[k] = torch.cuda.Event()
                copied[k].record(copy_stream)

    def finish_e2e():
        for ev in copied:
            if ev is not None:
                ev.synchronize()
        ren.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(3, args.warmup)):
        flush.zero_()
        step_device()
    barrier()

    # ---- device-resident timing (value) ----
    ren.reset_stats()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    sampler.mark_begin()
    for k in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the event pairs)
        e0[k].record(stream)
        step_device()
        e1[k].record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = sum(a.elapsed_time(b) for a, b in zip(e0, e1))
    stats = ren.stats()

    # ---- per-kernel timing for the roofline: the same K steps with CUDA events between the stages.  Stage events
    #      need the stages back to back on one stream, so this loop runs the frame as one tile set (the default
    #      may overlap two tile sets on two streams, which is what `value` measures). ----
    ren.set_pipeline(1)
    ren.set_stage_timing(True)
    step_device()
    barrier()
    ren.reset_stats()
    for k in range(args.steps):
        flush.zero_()
        step_device()
    barrier()
    stage_stats = ren.stats()
    ren.set_stage_timing(False)
    ren.set_pipeline(-1)  # back to the automatic schedule
    barrier()

    # ---- end-to-end timing through the C-ABI with host buffers ----
    step_e2e()
    finish_e2e()
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        step_e2e()
    finish_e2e()  # every step's image has reached host memory
    barrier()
    e2e_seconds = time.perf_counter() - t0

    # ---- parity of the multi-GPU frame: the exchanged image on rank 0 against one untimed single-GPU render of the
    #      same frame on rank 0, bit for bit ----
    parity = None
    if world > 1:
        exchanged = step_device()
        barrier()
        if rank == 0:
            multi = exchanged.cpu().numpy().copy()
        exchange_mode = exchange.mode
        exchange.close()
        if rank == 0:
            ren.set_tile_partition(0, 1)
            new_frame()
            ren.render()
            single, _ = ren.read_hdr()
            parity = {"bit_identical_to_1gpu": bool(np.array_equal(multi.view(np.uint32), single.view(np.uint32))), "mode": exchange_mode,
                      "pixels_compared": int(w * h), "max_abs_diff": float(np.max(np.abs(multi[..., :3] - single[..., :3])))}
        barrier()
    else:
        exchange_mode = exchange.mode
        exchange.close()

    rays_local = stats["closest_rays"] + stats["shadow_rays"]
    agg = torch.tensor([ms_total, e2e_seconds], dtype=torch.float64, device=dev)
    cnt = torch.tensor([rays_local, stats["paths"]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(agg, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_total, e2e_seconds = float(agg[0]), float(agg[1])
    rays_total, paths_total = int(cnt[0]), int(cnt[1])

    if rank == 0:
        num_sms = torch.cuda.get_device_properties(dev).multi_processor_count
        roofline = build_roofline(args, world, stats, stage_stats, clocks, num_sms, w, h)
        value = rays_total / (ms_total * 1e-3) / 1e6
        if stats.get("persistent_kernel"):
            pipeline = ("one persistent launch per frame (k_mega: block-local path loops — ray generation, traversal, shading and compaction in "
                        "one kernel, ready rays in shared-memory rings, lagging paths first, the last rays walked by whole warps) + k_accumulate")
        else:
            pipeline = (f"{stats['sub_frames']} tile set(s) on separate CUDA streams, one launch per stage (raygen, {bounces + 1} x trace, "
                        f"{bounces} x shade, accumulate per set)"
                        + (f"; each trace launch hands warps left with <= {stats['evict_max']} rays to a warp-per-ray tail launch"
                           if stats["evict_max"] else "; once a launch's queue is dry a warp left with one ray walks it with all 32 lanes in place"))
        partition = (f"32x32 tiles, (tx+ty) % {world}; exchange per step: "
                     + ("owned pixels stored into rank 0's double-buffered exchange target over NVLink peer memory by the accumulation "
                        "kernel + a 4-byte all-reduce as frame barrier" if exchange_mode == "p2p"
                        else "one out-of-place NCCL sum-reduce of the HDR buffer")) if world > 1 else "single GPU"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": DATA_NOTE,
            "config": base_config(w, h, bounces, rays_total // args.steps, paths_total // args.steps,
                                  "flushed between timed iterations (256 MiB memset)", pipeline, partition),
            "clocks": clocks,
            "e2e": {"value": rays_total / e2e_seconds / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": w * h * 16, "ms_per_step": 1e3 * e2e_seconds / args.steps},
            "gpu_launches": int(stats["kernel_launches"]),
            "roofline": roofline,
            "library": {"path": str(capi.LIB_PATH.relative_to(ROOT)), "build": capi.lib().rf_build_info().decode()},
        }
        if parity is not None:
            line["parity"] = parity
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_block(CpuReference(w, h, bounces), args.cpu_seconds)
        print(json.dumps(line), flush=True)

    ren.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
