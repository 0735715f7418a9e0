#!/usr/bin/env python
"""Steady-state kernel time (tools/steady_run.py: phases spread, auto-resets in every launch) for every library variant.
usage: tools/variant_steady.py [K ...]"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Ks = [int(k) for k in sys.argv[1:]] or [10]
libs = [os.path.join(ROOT, "b747_rl_ctrl_b200", "lib", "libb747_b200.so")] + sorted(
    glob.glob(os.path.join(ROOT, "b747_rl_ctrl_b200", "lib", "variants", "*.so")))
for lib in libs:
    print(os.path.basename(lib), flush=True)
    for K in Ks:
        for rep in range(2):
            subprocess.run([sys.executable, os.path.join(ROOT, "tools", "steady_run.py"), str(K), "100"],
                           env=dict(os.environ, B747_LIB_PATH=lib))
