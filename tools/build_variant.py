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
B.build_native()
out_dir = os.path.join(B.LIBDIR, "variants")
os.makedirs(out_dir, exist_ok=True)
obj = os.path.join(B.OBJDIR, f"f32_{name}.o")
subprocess.run([B._nvcc()] + B.ARCH + B.COMMON + B.UNITS["b747_kernels_f32.cu"] + flags +
               ["-c", os.path.join(B.CSRC, "b747_kernels_f32.cu"), "-o", obj], check=True)
objs = [os.path.join(B.OBJDIR, "b747_kernels_f64.o"), obj, os.path.join(B.OBJDIR, "b747_capi.o")]
lib = os.path.join(out_dir, f"lib_{name}.so")
subprocess.run([B._nvcc()] + B.ARCH + ["-shared", "-o", lib] + objs +
               ["-Xlinker", "-Bsymbolic", "-lcudart_static", "-lpthread", "-ldl", "-lrt"], check=True)
print(lib)
