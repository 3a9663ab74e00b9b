#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY: compile the column-physics sources of the product (isca_b200/csrc/physics*.cu, unchanged) for the HOST.
Every `kernel<<<grid, block, smem, stream>>>(args);` launch is rewritten to `ISCA_CPU_LAUNCH(kernel, grid, block, args);`, the
result is compiled by g++ -fopenmp against tests/host/cuda_on_cpu/cuda_runtime.h (device memory = host memory, a launch = an
OpenMP loop) into tests/host/_build/libisca_phys_cpu.so with the same C ABI as the product's column-physics entry points.
Used by tests/test_phys_cpu.py and by bench.py's CPU baseline; never loaded by the isca_b200 package."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "isca_b200", "csrc")
OUT = os.path.join(HERE, "_build")
FILES = ["physics.cu", "physics_conv.cu", "physics_bm.cu", "physics_dry.cu", "physics_surface.cu", "physics_turb.cu", "physics_diff.cu"]
LIB = os.path.join(OUT, "libisca_phys_cpu.so")


def split_top(s):
    """split on the commas that are not inside parentheses"""
    parts, depth, cur = [], 0, ""
    for ch in s:
        depth += ch in "([{"
        depth -= ch in ")]}"
        if ch == "," and depth == 0:
            parts.append(cur.strip()); cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def rewrite_launches(text):
    out, pos = [], 0
    while True:
        i = text.find("<<<", pos)
        if i < 0:
            out.append(text[pos:])
            return "".join(out)
        m = re.search(r"([A-Za-z_][\w:]*)\s*$", text[pos:i])
        name_start = pos + m.start(1)
        j = text.index(">>>", i)
        cfg = split_top(text[i + 3:j])
        k = text.index("(", j)
        depth, e = 0, k
        while True:
            depth += text[e] == "("
            depth -= text[e] == ")"
            if depth == 0:
                break
            e += 1
        out.append(text[pos:name_start])
        out.append(f"ISCA_CPU_LAUNCH({m.group(1)}, {cfg[0]}, {cfg[1]}, {text[k + 1:e]})")
        pos = e + 1


def build(force=False):
    os.makedirs(OUT, exist_ok=True)
    deps = [os.path.join(SRC, f) for f in os.listdir(SRC)] + [os.path.join(HERE, "cuda_on_cpu", "cuda_runtime.h"), os.path.abspath(__file__)]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in deps):
        return LIB
    gen = []
    for f in FILES:
        text = open(os.path.join(SRC, f)).read()
        cpp = os.path.join(OUT, "cpu_" + f.replace(".cu", ".cpp"))
        open(cpp, "w").write(f'#line 1 "{os.path.join(SRC, f)}"\n' + rewrite_launches(text))
        gen.append(cpp)
    glue = os.path.join(OUT, "cpu_glue.cpp")
    open(glue, "w").write('#include <cuda_runtime.h>\nthread_local IscaCpuDim3 threadIdx, blockIdx, blockDim, gridDim;\n'
                          'namespace isca_cpu { double kernel_ms = 0.0; }\n'
                          'extern "C" double isca_cpu_kernel_ms(int reset) { double v = isca_cpu::kernel_ms; if (reset) isca_cpu::kernel_ms = 0.0; return v; }\n')
    cmd = ["g++", "-O2", "-std=c++17", "-fopenmp", "-shared", "-fPIC", "-w", "-I", os.path.join(HERE, "cuda_on_cpu"), "-I", SRC,
           "-I", os.path.join(ROOT, "include"), "-o", LIB] + gen + [glue]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("host build of the column physics failed:\n" + r.stderr[-4000:])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
