"""Runs only the K-substep f32 step kernel (for ncu): python tools/k1_run.py [K] [steps]"""
import sys
import torch
sys.path.insert(0, ".")
from b747_rl_ctrl_b200 import engine as E
K = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
n = 1 << 20
eng = E.BatchEngine(n_envs=n, dtype=E.F32, sample_time=K * 0.01, seed=1, auto_reset=True)
s = torch.cuda.current_stream(); eng.use_stream(s.cuda_stream)
act, obs, rew, done = eng.alloc_io(); eng.reset(obs)
act.uniform_(-1, 1)
for i in range(steps):
    eng.step(act, obs, rew, done)
torch.cuda.synchronize()
print("ok", eng.episode_stats())
