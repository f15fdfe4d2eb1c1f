#!/usr/bin/env python3
"""Static evidence for the shipped library, produced without a GPU: per-kernel registers / spills / shared
memory (ptxas -v) and counts of the SASS mnemonics that show which hardware paths a kernel uses (packed
f32x2 arithmetic, mbarrier = SYNCS, bulk async copies = UBLKCP, ...).  Usage: tools/static_report.py > profiles/<name>.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "splat_b200", "csrc")
LIB = os.path.join(ROOT, "splat_b200", "libsplat_b200.so")
WATCH = ["FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "MUFU", "SYNCS", "UBLKCP", "LDG", "STG", "LDS", "STS", "ATOM", "RED", "SHFL", "VOTE",
         "MATCH", "BAR", "LDL", "STL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(sig):
    s = re.sub(r"\(anonymous namespace\)::", "", sig)
    s = re.sub(r"^void ", "", s)
    return re.sub(r"\(.*$", "", s).replace("splat::", "")


def main():
    subprocess.check_call(["make", "-s", "-C", CSRC])
    v = subprocess.run(["make", "-s", "-C", CSRC, "ptxas-info"], capture_output=True, text=True)
    text = v.stdout + v.stderr
    info, cur = {}, None
    for ln in text.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", ln)
        if m:
            cur = m.group(1); info[cur] = {}
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
        if m:
            info[cur].update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
        m = re.search(r"Used (\d+) registers", ln)
        if m:
            info[cur]["regs"] = int(m.group(1))
            s = re.search(r"(\d+) bytes smem", ln)
            info[cur]["smem"] = int(s.group(1)) if s else 0
            b = re.search(r"used (\d+) barriers", ln)
            info[cur]["bars"] = int(b.group(1)) if b else 0
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, fn = collections.defaultdict(collections.Counter), None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            fn = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m and fn:
            op = m.group(1)
            counts[fn]["_total"] += 1
            counts[fn][op] += 1
    names = demangle(sorted(set(info) | set(counts)))
    print("# static report of splat_b200/libsplat_b200.so (nvcc", subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout.split("release ")[-1].split(",")[0] + ", sm_100a, --fmad=false)")
    print("# produced by tools/static_report.py on a box without a GPU: ptxas -v + cuobjdump -sass\n")
    print(f"{'kernel':34s} {'regs':>4s} {'spill B':>7s} {'smem B':>7s} {'SASS':>6s}  notable mnemonics")
    for k in sorted(info, key=lambda k: short(names[k])):
        i, c = info[k], counts.get(k, {})
        notable = " ".join(f"{op}:{c[op]}" for op in WATCH if c.get(op))
        print(f"{short(names[k]):34s} {i.get('regs', 0):4d} {i.get('spill_st', 0):7d} {i.get('smem', 0):7d} {c.get('_total', 0):6d}  {notable}")
    print("\nFFMA2/FMUL2/FADD2 = Blackwell packed f32x2 arithmetic; SYNCS = mbarrier operations; UBLKCP = cp.async.bulk (TMA 1-D);")
    print("LDL/STL = local-memory (spill) traffic.  blend_kernel's dynamic shared memory (ring + barriers) is requested at launch, not listed here.")


if __name__ == "__main__":
    sys.exit(main())
