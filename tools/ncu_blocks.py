#!/usr/bin/env python3
"""Basic-block summary of one kernel from `ncu --page source --csv` (SASS view): consecutive instructions
with the same execution count are merged into one line: share of issued warp instructions, active threads,
stall-sample share, instruction count, first opcodes."""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
    hdr = next(r for r in rows if r and r[0] == "Address")
    data, seen = [], set()
    for r in rows:
        if r and r[0].startswith("0x") and len(r) >= len(hdr) - 2 and r[0] not in seen:
            seen.add(r[0])
            data.append(r)
    isrc, ismp, iex, ith = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
    tot_ex = sum(int(r[iex]) for r in data)
    tot_s = sum(int(r[ismp]) for r in data)
    lanes = sum(int(r[iex]) * float(r[ith]) for r in data)
    print(f"warp instructions {tot_ex}, avg active threads {lanes / tot_ex:.2f}, stall samples {tot_s}")
    base = int(data[0][0], 16)
    blocks = []
    for r in data:
        ex, thr, smp = int(r[iex]), r[ith], int(r[ismp])
        op = r[isrc].strip().split()
        op = (op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "?")).split(".")[0]
        if blocks and blocks[-1]["ex"] == ex and blocks[-1]["thr"] == thr:
            b = blocks[-1]
            b["n"] += 1
            b["smp"] += smp
            b["ops"].append(op)
        else:
            blocks.append({"addr": int(r[0], 16) - base, "ex": ex, "thr": thr, "n": 1, "smp": smp, "ops": [op]})
    for b in blocks:
        share = 100.0 * b["ex"] * b["n"] / tot_ex
        if share >= min_share:
            ops = " ".join(b["ops"][:14]) + (" ..." if len(b["ops"]) > 14 else "")
            print(f"{b['addr']:5x} n={b['n']:3d} share={share:5.2f}% thr={b['thr']:>5s} smp={100.0 * b['smp'] / tot_s:5.2f}%  {ops}")


if __name__ == "__main__":
    main()
