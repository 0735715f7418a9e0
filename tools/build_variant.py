#!/usr/bin/env python
"""Build a variant of libb747_b200.so with extra nvcc flags on the f32 kernel unit (perf experiments).
usage: tools/build_variant.py <name> [flags...]   ->  b747_rl_ctrl_b200/lib/variants/lib_<name>.so"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b747_rl_ctrl_b200 import build as B

name, flags = sys.argv[1], sys.argv[2:]
unit = "f32"
if flags and flags[0] in ("--f64", "--f32"):
    unit, flags = flags[0][2:], flags[1:]
B.build_native()
out_dir = os.path.join(B.LIBDIR, "variants")
os.makedirs(out_dir, exist_ok=True)
src = f"b747_kernels_{unit}.cu"
obj = os.path.join(B.OBJDIR, f"{unit}_{name}.o")
subprocess.run([B._nvcc()] + B.ARCH + B.COMMON + B.UNITS[src] + flags + (["-Xptxas", "-v"] if os.environ.get("B747_PTXAS_V") else []) + ["-c", os.path.join(B.CSRC, src), "-o", obj],
               check=True)
other = "b747_kernels_f64.o" if unit == "f32" else "b747_kernels_f32.o"
objs = [os.path.join(B.OBJDIR, other), obj, os.path.join(B.OBJDIR, "b747_capi.o")]
lib = os.path.join(out_dir, f"lib_{name}.so")
subprocess.run([B._nvcc()] + B.ARCH + ["-shared", "-o", lib] + objs +
               ["-Xlinker", "-Bsymbolic", "-lcudart_static", "-lpthread", "-ldl", "-lrt"], check=True)
print(lib)
