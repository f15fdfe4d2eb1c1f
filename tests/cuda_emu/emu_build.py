#!/usr/bin/env python3
"""TEST INFRASTRUCTURE.  Builds tests/cuda_emu/_build/libsplat_b200_emu.so: the product's own sources
(splat_b200/csrc/splat_api.cu and the *.cuh kernels) compiled for the HOST against tests/cuda_emu/cuda_emu.h.

The sources are not edited by hand; three mechanical rewrites make them valid C++:
  1. kernel<<<grid, block, smem, stream>>>(args)   ->  emu::run_grid(emu::cfg(grid, block, smem, stream), [&] { kernel(args); })
  2. extern __shared__ T name[];                   ->  T *name = reinterpret_cast<T *>(emu::dyn_smem());
  3. every inline-PTX statement                    ->  a call of the emu:: function of the same meaning (table below;
                                                       an asm statement the table does not know stops the build)
Nothing here is reachable from the product: splat_b200/_lib.py loads libsplat_b200.so and nothing else."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "splat_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
OUT = os.path.join(BUILD, "libsplat_b200_emu.so")
NCCL = os.path.join(BUILD, "libnccl_emu.so")     # in-process stand-in for libnccl (fake_nccl.cpp); the emulated build dlopens THIS


# ------------------------------------------------------------------------------------------------ small C++ scanners
def match_paren(s, i):
    """s[i] == '(' -> index of the matching ')', skipping string and char literals"""
    depth, j, n = 0, i, len(s)
    while j < n:
        c = s[j]
        if c == '"' or c == "'":
            q = c
            j += 1
            while s[j] != q:
                j += 2 if s[j] == "\\" else 1
        elif c == "(":
            depth += 1
        elif c == ")":
            depth -= 1
            if depth == 0:
                return j
        j += 1
    raise ValueError("unbalanced parenthesis")


def split_top(s, sep):
    """split s at top-level occurrences of the one-character separator (outside (), strings)"""
    out, depth, cur, j, n = [], 0, [], 0, len(s)
    while j < n:
        c = s[j]
        if c == '"':
            k = j + 1
            while s[k] != '"':
                k += 2 if s[k] == "\\" else 1
            cur.append(s[j:k + 1])
            j = k + 1
            continue
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == sep and depth == 0:
            out.append("".join(cur))
            cur = []
        else:
            cur.append(c)
        j += 1
    out.append("".join(cur))
    return out


def operands(section):
    """'"=l"(r), "f"(lo)' -> ['r', 'lo']"""
    ops = []
    for item in split_top(section, ","):
        item = item.strip()
        if not item:
            continue
        m = re.match(r'"[^"]*"\s*\(', item)
        if not m:
            raise ValueError(f"cannot parse asm operand {item!r}")
        ops.append(item[m.end():match_paren(item, m.end() - 1)].strip())
    return ops


def unshared(expr):
    m = re.fullmatch(r"smem_u32\((.*)\)", expr.strip(), re.S)
    return m.group(1) if m else expr


# ------------------------------------------------------------------------------------------------ inline PTX -> emu calls
def translate_asm(template, outs, ins):
    t = " ".join(template.split())
    bar = lambda e: f"reinterpret_cast<uint64_t *>(const_cast<void *>(static_cast<const volatile void *>({unshared(e)})))"
    if "mov.b64 %0, {%1, %2}" in t:
        return f"{outs[0]} = emu::ptx::pack({ins[0]}, {ins[1]});"
    if "mov.b64 {%0, %1}, %2" in t:
        return f"emu::ptx::unpack({ins[0]}, {outs[0]}, {outs[1]});"
    for op, nin in (("fma.rn.f32x2", 3), ("mul.rn.f32x2", 2), ("add.rn.f32x2", 2), ("sub.rn.f32x2", 2), ("add.rz.f32x2", 2)):
        if t.startswith(op):
            return f"{outs[0]} = emu::ptx::{op.replace('.', '_')}({', '.join(ins[:nin])});"
    if t.startswith("mbarrier.init"):
        return f"emu::mb_init({bar(ins[0])}, {ins[1]});"
    if "mbarrier.arrive.expect_tx" in t:
        return f"emu::mb_update({bar(ins[0])}, -1, (long long)({ins[1]}));"
    if "mbarrier.arrive.release" in t:
        return f"emu::mb_update({bar(ins[0])}, -1, 0);"
    if "WAIT_%=" in t and ("mbarrier.test_wait" in t or "mbarrier.try_wait" in t):
        return f"emu::mb_spin({bar(ins[0])}, {ins[1]});"
    if "mbarrier.test_wait" in t and "selp" in t:
        return f"{outs[0]} = emu::mb_test({bar(ins[0])}, {ins[1]}) ? 1u : 0u;"
    if t.startswith("mbarrier.inval"):
        return "/* mbarrier.inval */;"
    if t.startswith("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes"):
        return f"emu::ptx::cp_async_bulk({unshared(ins[0])}, {ins[1]}, {ins[2]}, {bar(ins[3])});"
    m = re.match(r"bar\.sync (\d+), %0", t)
    if m:
        return f"emu_bar_sync({m.group(1)}, {ins[0]});"
    if t.startswith("fence.mbarrier_init"):
        return "__threadfence();"
    raise ValueError(f"emu_build: no emulation for inline PTX {t!r}")


def rewrite_asm(src):
    out, i = [], 0
    for m in re.finditer(r"\basm\s*(?:volatile\s*)?\(", src):
        if m.start() < i:
            continue
        end = match_paren(src, m.end() - 1)
        body = src[m.end():end]
        semi = src.index(";", end)
        parts = split_top(body, ":")
        template = "".join(re.findall(r'"((?:[^"\\]|\\.)*)"', parts[0])).replace("\\n", " ").replace("\\t", " ")
        outs = operands(parts[1]) if len(parts) > 1 else []
        ins = operands(parts[2]) if len(parts) > 2 else []
        out.append(src[i:m.start()])
        out.append(translate_asm(template, outs, ins))
        i = semi + 1
    out.append(src[i:])
    return "".join(out)


# ------------------------------------------------------------------------------------------------ the other two rewrites
def rewrite_extern_shared(src):
    return re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([A-Za-z_][\w ]*?)\s+(\w+)\[\];",
                  r"\1 *\2 = reinterpret_cast<\1 *>(emu::dyn_smem());", src)


def rewrite_launches(src):
    out, i = [], 0
    pat = re.compile(r"([A-Za-z_]\w*(?:<[^<>;(){}]*>)?)\s*<<<")
    while True:
        m = pat.search(src, i)
        if not m:
            break
        close = src.index(">>>", m.end())
        lp = close + 3
        while src[lp].isspace():
            lp += 1
        assert src[lp] == "(", src[m.start():lp + 20]
        rp = match_paren(src, lp)
        out.append(src[i:m.start()])
        # a direct call inside the lambda keeps default arguments and template deduction; the launch is synchronous,
        # so capturing the host's variables by reference is safe
        out.append(f"emu::run_grid(emu::cfg({src[m.end():close]}), [&] {{ {m.group(1)}({src[lp + 1:rp]}); }}, \"{m.group(1)}\")")
        i = rp + 1
    out.append(src[i:])
    return "".join(out)


def preprocess(name):
    src = open(os.path.join(CSRC, name)).read()
    src = rewrite_asm(src)
    src = rewrite_extern_shared(src)
    src = rewrite_launches(src)
    src = src.replace('#include "../../include/splat.h"', f'#include "{os.path.join(ROOT, "include", "splat.h")}"')
    if name == "comm.cuh":
        assert src.count('"libnccl.so.2", "libnccl.so"') == 1
        src = src.replace('"libnccl.so.2", "libnccl.so"', f'"{NCCL}"')
    assert "<<<" not in src and not re.search(r"\basm\b", src), name
    return src


def build(force=False, asan=None):
    """asan (default: env EMU_ASAN=1): AddressSanitizer build, _build/libsplat_b200_emu_asan.so; load it into a python that
    runs with LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0"""
    asan = os.environ.get("EMU_ASAN") == "1" if asan is None else asan
    out = OUT.replace(".so", "_asan.so") if asan else OUT
    defines = os.environ.get("EMU_DEFINES", "").split()        # e.g. EMU_DEFINES="-DSPLAT_TMA_STAGE=1": a compile-time variant
    if defines:
        out = out.replace(".so", "_" + "".join(ch if ch.isalnum() else "_" for ch in " ".join(defines)) + ".so")
    return _build(force, asan, out, defines)


def _build(force, asan, OUT, defines=()):
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cuh") or f == "splat_api.cu")
    deps = [os.path.join(CSRC, f) for f in srcs] + [os.path.join(HERE, f) for f in ("cuda_emu.h", "emu_build.py", "nccl.h", "cuda_runtime.h", "fake_nccl.cpp")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    gen = os.path.join(BUILD, "src")
    os.makedirs(gen, exist_ok=True)
    for f in srcs:
        with open(os.path.join(gen, f.replace(".cu", ".cpp") if f.endswith(".cu") else f), "w") as o:
            o.write(f"// GENERATED by tests/cuda_emu/emu_build.py from splat_b200/csrc/{f} -- do not edit\n" + preprocess(f))
    cxx = os.environ.get("CXX", "g++")
    if asan and os.path.exists("/usr/bin/g++"):
        cxx = "/usr/bin/g++"                      # the distribution's compiler ships libasan
    cmd = [cxx, "-std=c++20", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-frounding-math",
           "-fno-strict-aliasing", "-w", *(["-fsanitize=address", "-fno-omit-frame-pointer"] if asan else []), *defines, "-I", HERE, "-I", gen, "-o", OUT, os.path.join(gen, "splat_api.cpp"), "-ldl"]
    subprocess.check_call(cmd)
    subprocess.check_call([os.environ.get("CXX", "g++"), "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-I", HERE,
                           "-o", NCCL, os.path.join(HERE, "fake_nccl.cpp")])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
