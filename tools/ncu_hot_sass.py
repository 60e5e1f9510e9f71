#!/usr/bin/env python3
"""Hot SASS lines of one kernel from `ncu -i X.ncu-rep --page source --csv` (share of executed warp
instructions, share of stall samples, average active threads)."""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0015
    hdr = next(r for r in rows if r and r[0] == "Address")
    data = [r for r in rows if r and r[0].startswith("0x") and len(r) >= len(hdr) - 2]
    ia, isrc = hdr.index("Address"), hdr.index("Source")
    ismp, iex, ith = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
    tot_ex = sum(int(r[iex]) for r in data)
    tot_s = sum(int(r[ismp]) for r in data)
    print("total warp instructions", tot_ex, "stall samples", tot_s, "SASS lines", len(data))
    base = int(data[0][ia], 16)
    for r in data:
        ex, sm = int(r[iex]), int(r[ismp])
        if ex > tot_ex * thresh or sm > tot_s * thresh * 2:
            print(f"{int(r[ia], 16) - base:5x} {r[isrc].strip()[:66]:66s} ex={100 * ex / tot_ex:5.2f}% smp={100 * sm / tot_s:5.2f}% thr={r[ith]}")


if __name__ == "__main__":
    main()
