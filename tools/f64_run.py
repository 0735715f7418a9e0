"""Runs the f64 parity step kernel (for ncu): python tools/f64_run.py [n_envs] [K] [steps]"""
import sys
import torch
sys.path.insert(0, ".")
from b747_rl_ctrl_b200 import engine as E
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
K = int(sys.argv[2]) if len(sys.argv) > 2 else 5
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
eng = E.BatchEngine(n_envs=n, dtype=E.F64, sample_time=K * 0.01, seed=1, auto_reset=True)
act, obs, rew, done = eng.alloc_io(); eng.reset(obs)
act.uniform_(-1, 1)
for i in range(steps):
    eng.step(act, obs, rew, done)
eng.synchronize()
print("ok")
