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
UNIT = "Mrays/s"
WIDTH, HEIGHT, BOUNCES = 1920, 1080, 8
KERNELS_PER_STEP = 3 + 2 * BOUNCES  # raygen + trace + (shade, trace) per bounce + accumulate


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--width", type=int, default=WIDTH)
    p.add_argument("--height", type=int, default=HEIGHT)
    p.add_argument("--bounces", type=int, default=BOUNCES)
    p.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work per bounded reference sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    return p.parse_args()


def load_scene():
    import rayfinder_b200 as rf

    from rayfinder_b200 import assets as rfa

    if rfa.scene_path("Sponza") is not None:
        return rfa.load_scene("Sponza"), "Sponza.pt"
    import _oracle as O  # fixture loader only (no oracle code runs)

    return rf.PtFormat.loads(O.duck_pt_bytes()), "Duck.pt (Sponza.pt not baked on this box)"


def workload_name(scene_name, w, h, bounces):
    return f"{scene_name} {w}x{h}, 1 spp, {bounces} bounces, interior fly-camera default view"


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
class CpuTraversal:
    """bvh-visualizer's pixel loop (bvh-visualizer/main.cpp:60-78) over the benchmark view's primary rays, using the
    reference's own compiled rayIntersectBvh (oracle/_ref, kind "reference") or, if that library is absent, the
    oracle port (kind "port"); rows are spread over all host threads."""

    def __init__(self, pt, width, height):
        import _oracle as O
        import rayfinder_b200 as rf

        self.O = O
        self.nodes = np.ascontiguousarray(pt.bvh_nodes)
        self.tris = O.triangles9(pt)
        self.cam = rf.camera_to_array(rf.fly_camera(width, height))
        self.width, self.height = width, height
        self.kind = "reference" if O.have_ref() else "port"
        self.cores = O.num_threads()
        self.t_max = rf.FLT_MAX

    def run_rows(self, rows: int, offset: int = 0):
        """Trace `rows` image rows taken as 8 evenly spaced bands; returns (rays, seconds)."""
        bands = 8
        per = max(1, rows // bands)
        rays, secs = 0, 0.0
        fn = self.O.ref_node_counts if self.kind == "reference" else self.O.oracle_node_counts
        for b in range(bands):
            r0 = min(self.height - per, (b * self.height) // bands + offset % max(1, self.height // bands - per + 1))
            _, s = fn(self.nodes, self.tris, self.cam, self.width, self.height, self.t_max, threads=self.cores, rows=(r0, r0 + per))
            rays += per * self.width
            secs += s
        return rays, secs

    def calibrate(self, seconds: float) -> int:
        """Rows per sample so that one sample is about `seconds` of wall time; more than `height` rows means
        repeated passes over the frame."""
        rays, secs = self.run_rows(64)
        rate = rays / max(secs, 1e-9)
        rows = int(max(8, (rate * seconds) / self.width))
        if rows >= self.height:
            return self.height * max(1, rows // self.height)
        return (rows // 8) * 8

    def run_sample(self, rows: int, offset: int = 0):
        if rows <= self.height:
            return self.run_rows(rows, offset)
        rays = secs = 0.0
        for _ in range(rows // self.height):
            r, s = self.run_rows(self.height)
            rays, secs = rays + r, secs + s
        return rays, secs

    def sample_description(self, rows: int) -> str:
        what = (f"{rows // self.height} full passes over" if rows >= self.height
                else f"{rows} of {self.height} rows (8 evenly spaced bands) of")
        return (f"{what} the {self.width}x{self.height} primary rays, bvh-visualizer loop, tmax=FLT_MAX, "
                f"{self.cores} threads")


def run_reference(args, rank: int):
    if rank != 0:
        return
    import rayfinder_b200 as rf  # noqa: F401  (host-side .pt loader only)

    pt, scene_name = load_scene()
    cpu = CpuTraversal(pt, args.width, args.height)
    # K timed steps + W warm-up steps must end within a few minutes: bound each step's sample
    budget = min(args.cpu_seconds, 150.0 / max(1, args.steps + args.warmup))
    rows = cpu.calibrate(budget)
    for k in range(args.warmup):
        cpu.run_sample(rows, offset=k)
    rays = secs = 0.0
    for k in range(args.steps):
        r, s = cpu.run_sample(rows, offset=k)
        rays += r
        secs += s
    value = rays / secs / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, args.steps), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "Sponza.pt baked from the reference's assets/Sponza.glb",
        "config": {"workload": workload_name(scene_name, args.width, args.height, args.bounces),
                   "note": "reference CPU path = bvh-visualizer traversal of the primary rays (the reference has no CPU path tracer)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind, "sample": cpu.sample_description(rows)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- this repo's CUDA path ----------------------------------------------------------------------------------
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

    pt, scene_name = load_scene()
    w, h, bounces = args.width, args.height, args.bounces
    cam = rf.fly_camera(w, h) if scene_name.startswith("Sponza") else rf.bvh_visualizer_camera(pt.bvh_nodes, w, h)
    params = rf.RenderParameters((w, h), cam, rf.SamplingParams(1, bounces), rf.Sky(), 0.25)
    ren = rf.ReferencePathTracer(params, (w, h), rf.SceneArrays.from_pt(pt), device=local_rank)
    ren.set_tile_partition(rank, world)
    stream = torch.cuda.current_stream(dev)
    ren.set_stream(stream.cuda_stream)
    exchange = rfd.HdrExchange(ren, w, h, mode=os.environ.get("RF_EXCHANGE", "auto"))
    hdr = exchange.hdr
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
        exchange()

    def step_e2e():
        new_frame()  # host -> device: the render parameters (uniform block) travel with the launch
        ren.render()
        exchange()
        if rank == 0:
            ren.read_hdr(host_np)  # device -> host: the HDR image
        else:
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
    #      overlaps two tile sets on two streams, which is what `value` measures). ----
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
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        step_e2e()
    barrier()
    e2e_seconds = time.perf_counter() - t0

    rays_local = stats["closest_rays"] + stats["shadow_rays"]
    agg = torch.tensor([ms_total, e2e_seconds], dtype=torch.float64, device=dev)
    cnt = torch.tensor([rays_local, stats["paths"]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(agg, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_total, e2e_seconds = float(agg[0]), float(agg[1])
    rays_total, paths_total = int(cnt[0]), int(cnt[1])

    if rank == 0:
        peaks = {}
        peaks_path = ROOT / "MEASURED_PEAKS.json"
        if peaks_path.exists():
            peaks = json.loads(peaks_path.read_text())
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # dominant kernel: k_trace (rank 0's launches): closest-hit + shadow rays share the traversal launches.
        # Algorithmic bytes: 48 B per node visited + 48 B per triangle tested.
        nodes = stage_stats["closest_nodes_visited"] + stage_stats["shadow_nodes_visited"]
        tris = stage_stats["closest_triangles_tested"] + stage_stats["shadow_triangles_tested"]
        trace_bytes = 48 * (nodes + tris)
        trace_ms = stage_stats["device_ms_trace"]
        launches = args.steps * (bounces + 1)
        achieved = trace_bytes / (trace_ms * 1e-3) / 1e9 if trace_ms > 0 else None
        traffic = None
        traffic_files = sorted((ROOT / "profiles").glob("r*_k_trace_traffic.json"))
        if traffic_files:  # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu capture
            traffic = json.loads(traffic_files[-1].read_text()).get("dram_bytes_per_launch_mean")
        packed = 32 * nodes + 48 * tris
        roofline = {
            "bound": "hbm", "kernel": "k_trace", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": (achieved / peak) if achieved else None, "traffic": traffic,
            "frac_packed_layout": (packed / (trace_ms * 1e-3) / 1e9 / peak) if achieved else None,
            "note": ("algorithmic bytes = 48 B per node visit + 48 B per triangle test (SURVEY.md 8(d)); the 28.7 MB "
                     "node+triangle set is L1/L2-resident, so DRAM traffic is ~1% of the algorithmic bytes and the binding "
                     "resources are SM issue slots and the L1 tag stage (profiles/)"),
            "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6.65 TB/s",
            "algorithmic_bytes_per_launch": trace_bytes / launches, "avg_launch_ms": trace_ms / launches,
            "launches": launches, "packed_bytes_per_launch": packed / launches,
            "stage_ms_per_step": {k: stage_stats[f"device_ms_{k}"] / args.steps for k in ("trace", "shade", "other")},
            "stage_loop_ms_per_step": stage_stats["device_ms_total"] / args.steps,
            "mean_nodes_per_closest_ray": stats["closest_nodes_visited"] / max(1, stats["closest_rays"]),
            "mean_nodes_per_shadow_ray": stats["shadow_nodes_visited"] / max(1, stats["shadow_rays"]),
        }
        value = rays_total / (ms_total * 1e-3) / 1e6
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "Sponza.pt baked from the reference's assets/Sponza.glb (deterministic asset, not synthetic)",
            "config": {"workload": workload_name(scene_name, w, h, bounces), "rays_per_step": rays_total // args.steps,
                       "paths_per_step": paths_total // args.steps, "l2": "flushed between timed iterations (256 MiB memset)",
                       "pipeline": (f"{stats['sub_frames']} tile set(s) on separate CUDA streams, one launch per stage (raygen, {bounces + 1} x trace, "
                                    f"{bounces} x shade, accumulate per set)"
                                    + (f"; each trace launch hands warps left with <= {stats['evict_max']} rays to a warp-per-ray tail launch"
                                       if stats["evict_max"] else "")),
                       "partition": (f"32x32 tiles, (tx+ty) % {world}; exchange per step: "
                                     + ("owned pixels stored into rank 0's HDR buffer over NVLink peer memory by the accumulation kernel + a 4-byte "
                                        "all-reduce as frame barrier" if exchange.mode == "p2p" else "one NCCL sum-reduce of the HDR buffer"))
                       if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": rays_total / e2e_seconds / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": w * h * 16, "ms_per_step": 1e3 * e2e_seconds / args.steps},
            "gpu_launches": int(stats["kernel_launches"]),
            "roofline": roofline,
            "library": {"path": str(capi.LIB_PATH.relative_to(ROOT)), "build": capi.lib().rf_build_info().decode()},
        }
        if world == 1 and not args.no_cpu_baseline:
            cpu = CpuTraversal(pt, w, h)
            rows = cpu.calibrate(args.cpu_seconds)
            r, s = cpu.run_sample(rows)
            line["cpu_baseline"] = {"value": r / s / 1e6, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind,
                                    "sample": cpu.sample_description(rows)}
        print(json.dumps(line), flush=True)

    exchange.close()
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
