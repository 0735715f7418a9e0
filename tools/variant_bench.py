#!/usr/bin/env python
"""Kernel time of the bench workload (1 Mi envs, f32, K=10 and K=1) for every library variant.
usage: tools/variant_bench.py [lib.so ...]   (default: the in-tree library + lib/variants/*.so)"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from b747_rl_ctrl_b200 import engine as E
n = 1 << 20
for K, steps in ((10, 300), (1, 200)):
    eng = E.BatchEngine(n_envs=n, dtype=E.F32, sample_time=K * 0.01, seed=1, auto_reset=True)
    s = torch.cuda.current_stream(); eng.use_stream(s.cuda_stream)
    act, obs, rew, done = eng.alloc_io(); eng.reset(obs)
    pool = [torch.empty(n, device="cuda").uniform_(-1, 1) for _ in range(8)]
    for i in range(5): eng.step(pool[i %% 8], obs, rew, done)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(steps): eng.step(pool[i %% 8], obs, rew, done)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    st = eng.episode_stats()
    print(f"  K={K}: {ms:.4f} ms/step  {n / ms * 1e-6:.3f} G env-steps/s  ep_rew_mean {st[1] / max(st[0], 1):.6f}", flush=True)
    eng.close()
''' % ROOT

libs = sys.argv[1:] or [os.path.join(ROOT, "b747_rl_ctrl_b200", "lib", "libb747_b200.so")] + sorted(
    glob.glob(os.path.join(ROOT, "b747_rl_ctrl_b200", "lib", "variants", "*.so")))
for lib in libs:
    print(os.path.basename(lib), flush=True)
    env = dict(os.environ, B747_LIB_PATH=lib)
    subprocess.run([sys.executable, "-c", CHILD], env=env)
