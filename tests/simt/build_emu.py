"""TEST INFRASTRUCTURE.  Builds tests/simt/_build/liblsdreg_emu.so: the product's CUDA sources (lidar-slam-detection_b200/csrc)
translated for the host and compiled by g++ against the SIMT emulator tests/simt/simt.h — same C ABI, kernels executed by
fibers on one CPU thread.  Used to run kernels against the oracle when no GPU is at hand (tests/test_emu_*.py); it is
loaded only when a test asks for it by path.  The translation is textual and minimal:
  kernel<<<grid, block, smem, stream>>>(args)   ->  simt::launch(grid, block, smem, [&]() { kernel(args); })
  extern __shared__ T name[];                   ->  T* name = reinterpret_cast<T*>(simt::dyn_smem());
Everything else compiles as it stands (qualifiers come from cuda_runtime.h; built-ins and the runtime API from simt.h).

    python tests/simt/build_emu.py [--force]
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "lidar-slam-detection_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "liblsdreg_emu.so")
CUDA_INC = os.environ.get("CUDA_INC", "/usr/local/cuda/include")


def _match_back_template(s, i):
    """s[i] == '>' : index of the matching '<'"""
    depth = 0
    while i >= 0:
        if s[i] == '>':
            depth += 1
        elif s[i] == '<':
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced template arguments before <<<")


def _split_top(s):
    parts, depth, cur = [], 0, ""
    for c in s:
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == "," and depth == 0:
            parts.append(cur.strip()); cur = ""
        else:
            cur += c
    parts.append(cur.strip())
    return parts


def translate(src: str) -> str:
    out, pos = "", 0
    while True:
        k = src.find("<<<", pos)
        if k < 0:
            return out + src[pos:]
        # kernel expression, scanning backwards
        i = k - 1
        while src[i].isspace():
            i -= 1
        if src[i] == '>':
            i = _match_back_template(src, i) - 1
        while i >= 0 and (src[i].isalnum() or src[i] in "_:"):
            i -= 1
        start = i + 1
        kern = src[start:k].strip()
        e = src.index(">>>", k)
        cfg = _split_top(src[k + 3:e])
        a = src.index("(", e)
        depth, j = 0, a
        while True:
            if src[j] == '(':
                depth += 1
            elif src[j] == ')':
                depth -= 1
                if depth == 0:
                    break
            j += 1
        args = src[a + 1:j]
        grid, block = cfg[0], cfg[1]
        smem = cfg[2] if len(cfg) > 2 else "0"
        out += src[pos:start] + f"simt::launch(simt::to_dim3({grid}), simt::to_dim3({block}), (size_t)({smem}), [&]() {{ {kern}({args}); }})"
        pos = j + 1


_EXT = re.compile(r"extern\s+__shared__\s+([A-Za-z_][\w:<>\s\*]*?)\s+(\w+)\s*\[\s*\]\s*;")


def build(force=False, asan=None):
    """asan (default: env LSDREG_EMU_ASAN): also instrument with AddressSanitizer — out-of-bounds accesses to "device" memory
    (heap) and to shared-memory arrays (static storage, red-zoned like any global) abort with a report.  Run python with
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0."""
    global LIB
    asan = bool(os.environ.get("LSDREG_EMU_ASAN")) if asan is None else asan
    if asan:
        LIB = os.path.join(OUT, "liblsdreg_emu_asan.so")
    os.makedirs(OUT, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".hpp")))
    deps = [os.path.join(CSRC, f) for f in srcs] + [os.path.join(HERE, "simt.h"), os.path.abspath(__file__),
                                                    os.path.join(ROOT, "include", "lsdreg.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    tr = os.path.join(OUT, "csrc")
    os.makedirs(tr, exist_ok=True)
    for f in srcs:
        s = open(os.path.join(CSRC, f)).read()
        s = translate(s)
        s = _EXT.sub(lambda m: f"{m.group(1)}* {m.group(2)} = reinterpret_cast<{m.group(1)}*>(simt::dyn_smem());", s)
        s = s.replace('#include "../../include/lsdreg.h"', f'#include "{os.path.join(ROOT, "include", "lsdreg.h")}"')
        open(os.path.join(tr, f), "w").write(s)
    main = os.path.join(OUT, "emu_main.cpp")
    open(main, "w").write(f'#include "{os.path.join(HERE, "simt.h")}"\n#include "csrc/lsdreg.cu"\n'
                          'extern "C" long long* simt_stats_export() { return simt_stats(); }\n')
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-ffp-contract=off", "-fno-strict-aliasing", "-w",
           "-x", "c++", main, "-o", LIB, "-I" + CUDA_INC, "-I" + OUT, "-lpthread"]
    if asan:
        cmd[1:1] = ["-fsanitize=address", "-fno-omit-frame-pointer"]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=OUT)
    if r.returncode:
        sys.stderr.write(r.stdout[-6000:] + r.stderr[-12000:])
        raise RuntimeError("g++ failed building liblsdreg_emu.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
