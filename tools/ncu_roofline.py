#!/usr/bin/env python3
"""Turn an `ncu --set full` capture of the traversal launches into the entry bench.py's `roofline` object refers to.

    ncu -i capture.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_roofline.py raw.csv KEY "how the capture was taken" [profiles/r02_roofline_ncu.json]

KEY = "<frame width>x<frame height>/<number of GPUs the frame is split over>", e.g. "1920x1080/1"; "1920x1080/8" is
captured on one GPU tracing the 672x384 frame a rank of 8 owns.  Per-launch means over the captured launches.
"""
import csv
import json
import sys
from pathlib import Path

PER_LAUNCH = {
    "dram_bytes_per_launch": (("dram__bytes_read.sum", "dram__bytes_write.sum"), 1.0),
    "lts_bytes_per_launch": (("lts__t_sectors.sum",), 32.0),
    "l1tex_lsu_wavefronts_per_sm_per_launch": (("SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg",), 1.0),
    "l1tex_lsu_wavefronts_shared_per_launch": (("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",), 1.0),
    "warp_instructions_per_launch": (("smsp__inst_executed.sum",), 1.0),
    "launch_ms_under_ncu": (("gpu__time_duration.sum",), 1.0),
}
RATIOS = {
    "l1tex_lsu_data_pipe_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex_throughput_pct": "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "lts_throughput_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm_issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "lanes_per_instruction": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "long_scoreboard_stall_per_issue": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "l1_hit_rate_pct": "l1tex__t_sector_hit_rate.pct",
    "l2_hit_rate_pct": "lts__t_sector_hit_rate.pct",
    "registers_per_thread": "launch__registers_per_thread",
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}


def main():
    raw, key, how = sys.argv[1], sys.argv[2], sys.argv[3]
    out = Path(sys.argv[4] if len(sys.argv) > 4 else Path(__file__).resolve().parent.parent / "profiles" / "r02_roofline_ncu.json")
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]

    def column(name):
        i = hdr.index(name)
        scale = UNIT_SCALE.get(units[i], 1.0)
        return [float(d[i].replace(",", "")) * scale for d in data]

    entry = {"how": how, "launches": len(data), "kernel": data[0][hdr.index("Kernel Name")].split("(")[0].replace("void ", "")}
    for name, (metrics, factor) in PER_LAUNCH.items():
        per_launch = [sum(vals) * factor for vals in zip(*(column(m) for m in metrics))]
        entry[name] = sum(per_launch) / len(per_launch)
    for name, metric in RATIOS.items():
        vals = column(metric)
        entry[name] = sum(vals) / len(vals)
    table = json.loads(out.read_text()) if out.exists() else {"captures": {}}
    table.setdefault("note", "per-launch means of `ncu --set full --clock-control none` captures of the traversal kernel; written by tools/ncu_roofline.py")
    table["captures"][key] = entry
    out.write_text(json.dumps(table, indent=1) + "\n")
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
