#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

  tools/summarize_ncu.py launches gpurun_out/launches_rN.csv  profiles/rN_launches_summary.txt
  tools/summarize_ncu.py full     gpurun_out/blend_rN.ncu-rep profiles/rN_blend_ncu.txt
  tools/summarize_ncu.py frame    gpurun_out/frame_rN.ncu-rep profiles/rN_frame_kernels.txt   (one row per launch)
"""
import collections
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct",
        "sm__throughput.avg.pct", "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__warps_active.avg.pct", "launch__registers_per_thread", "launch__occupancy_limit",
        "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
        "sm__inst_executed_pipe_fma.avg.pct", "sm__inst_executed_pipe_alu.avg.pct",
        "sm__inst_executed_pipe_lsu.avg.pct", "sm__inst_executed_pipe_xu.avg.pct",
        "sm__pipe_fmaheavy_cycles_active.avg", "sm__pipe_fmalite_cycles_active.avg", "sm__pipe_tensor",
        "smsp__average_warp", "smsp__warp_issue_stalled", "smsp__average_warps_issue_stalled",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate", "smsp__sass_thread_inst_executed_op_f"]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    n = 0
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        ms = v / 1e6 if r[ui].startswith("ns") else (v / 1e3 if r[ui].startswith("us") else v)
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
        n += 1
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# source: {src} (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised:\n"
                f"# compare SHARES, not absolutes).  {n} launches, {tot:.3f} ms total.\n")
        f.write(f"{'kernel':44s} {'launches':>8s} {'total_ms':>10s} {'share_%':>8s} {'avg_ms':>9s}\n")
        for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:44]:44s} {c:8d} {ms:10.3f} {ms / tot * 100:8.2f} {ms / c:9.4f}\n")
    print(open(dst).read())


def frame(src, dst, peak_gbs=None):
    """One row per captured launch: duration, DRAM bytes, achieved DRAM GB/s against the measured
    copy peak, and the busiest SM-side figures."""
    import json, os
    if peak_gbs is None:
        try:
            peak_gbs = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            peak_gbs = 6650.0
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    def get(rec, name, scale=True):
        if name not in hdr:
            return float("nan")
        i = hdr.index(name)
        try:
            v = float(rec[i].replace(",", ""))
        except ValueError:
            return float("nan")
        u = units[i]
        if scale:
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
        return v
    cols = [("sm%", "sm__throughput.avg.pct_of_peak_sustained_elapsed"), ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            ("adu%", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_elapsed"), ("lsu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed"),
            ("fma%", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"), ("L2hit%", "lts__t_sector_hit_rate.pct"),
            ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("regs", "launch__registers_per_thread")]
    with open(dst, "w") as f:
        f.write(f"# source: {src} (ncu --set full --clock-control none; per-launch, cold-cache, serialised)\n")
        f.write(f"# DRAM GB/s = (dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration; peak = {peak_gbs:.1f} GB/s measured copy\n")
        f.write(f"{'kernel':28s} {'grid':>7s} {'us':>8s} {'rd_MB':>8s} {'wr_MB':>8s} {'GB/s':>7s} {'%peak':>6s} " + " ".join(f"{c[0]:>6s}" for c in cols) + "\n")
        for rec in rows[2:]:
            name = rec[hdr.index("Kernel Name")].split("(")[0].replace("splat::", "")
            us, rd, wr = get(rec, "gpu__time_duration.sum"), get(rec, "dram__bytes_read.sum"), get(rec, "dram__bytes_write.sum")
            gbs = (rd + wr) / (us * 1e-6) / 1e9
            f.write(f"{name[:28]:28s} {int(get(rec, 'launch__grid_size', False)):7d} {us:8.1f} {rd / 1e6:8.1f} {wr / 1e6:8.1f} {gbs:7.0f} {100 * gbs / peak_gbs:6.1f} "
                    + " ".join(f"{get(rec, c[1], False):6.1f}" for c in cols) + "\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# source: {src} (ncu --set full --clock-control none --import-source on)\n")
        for rec in rows[2:]:
            name = rec[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write(f"\n## {name.split('(')[0]}\n")
            for h, u, v in zip(hdr, units, rec):
                if any(k in h for k in KEEP):
                    f.write(f"{h:95s} {u:14s} {v}\n")
    print(open(dst).read()[:6000])


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "frame":
    frame(sys.argv[2], sys.argv[3])
    sys.exit(0)
if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
