#!/usr/bin/env python3
"""Hot SASS lines of one kernel from `ncu -i X.ncu-rep --page source --csv` (share of executed warp
instructions, share of stall samples, average active threads).

    ncu -i prof.ncu-rep --page source --csv --kernel-name regex:k_closest --launch-count 1 > src.csv
    python tools/ncu_hot_sass.py src.csv [min_share] [--all]
"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    thresh = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("-") else 0.004
    show_all = "--all" in sys.argv
    hdr = next(r for r in rows if r and r[0] == "Address")
    data, seen = [], set()
    for r in rows:
        if r and r[0].startswith("0x") and len(r) >= len(hdr) - 2 and r[0] not in seen:  # the export repeats the listing
            seen.add(r[0])
            data.append(r)
    ia, isrc = hdr.index("Address"), hdr.index("Source")
    ismp, iex, ith = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
    tot_ex = sum(int(r[iex]) for r in data)
    tot_s = sum(int(r[ismp]) for r in data)
    thr_w = sum(int(r[iex]) * float(r[ith]) for r in data) / max(1, tot_ex)
    print(f"total warp instructions {tot_ex}, stall samples {tot_s}, SASS lines {len(data)}, avg active threads {thr_w:.2f}")
    base = int(data[0][ia], 16)
    # opcode histogram weighted by executions
    hist = {}
    for r in data:
        op = r[isrc].strip().split()
        op = op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "?")
        op = op.split(".")[0]
        hist[op] = hist.get(op, 0) + int(r[iex])
    print("opcode mix:", ", ".join(f"{k} {100 * v / tot_ex:.1f}%" for k, v in sorted(hist.items(), key=lambda kv: -kv[1])[:18]))
    for r in data:
        ex, sm = int(r[iex]), int(r[ismp])
        if show_all or ex > tot_ex * thresh or sm > tot_s * thresh * 2:
            print(f"{int(r[ia], 16) - base:5x} {r[isrc].strip()[:66]:66s} ex={100 * ex / tot_ex:5.2f}% smp={100 * sm / tot_s:5.2f}% thr={r[ith]}")


if __name__ == "__main__":
    main()
