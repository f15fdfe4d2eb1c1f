#!/usr/bin/env python
"""Where a kernel's issued instructions go: groups the SASS of `ncu --page source --csv` output
into runs of similar execution count.   usage: tools/instr_profile.py report.ncu-rep kernel_regex"""
import csv
import subprocess
import sys


def main(rep, kern):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    hi = his[0]
    end = his[1] - 1 if len(his) > 1 else len(rows)
    hdr = rows[hi]
    data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
    ia, isrc, ie = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
    tot = sum(int(r[ie]) for r in data)
    print(f"{rows[hi - 1][1][:60]}: {tot / 1e6:.1f} M warp instructions")
    base = int(data[0][ia], 16)
    runs, start, acc, cur = [], 0, 0, None
    for i, r in enumerate(data):
        c = int(r[ie])
        if cur is not None and abs(c - cur) > 0.25 * max(c, cur) and max(c, cur) > tot * 2e-5:
            runs.append((start, i - 1, acc))
            start, acc = i, 0
        cur = c
        acc += c
    runs.append((start, len(data) - 1, acc))
    for s, e, a in runs:
        if a > 0.004 * tot:
            print(f"{int(data[s][ia], 16) - base:05x}-{int(data[e][ia], 16) - base:05x} n={e - s + 1:4d} "
                  f"{a / 1e6:9.2f}M ({100 * a / tot:5.1f}%) per-instr={a / (e - s + 1) / 1e6:8.3f}M  {data[s][isrc].strip()[:56]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
