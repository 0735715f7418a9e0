"""Runs the tier-2 (altitude loop) f32 step kernel for ncu: HYBRID reset, SEMI_MANUAL."""
import sys, torch
sys.path.insert(0, ".")
from b747_rl_ctrl_b200 import engine as E
n = 1 << 20
eng = E.BatchEngine(n_envs=n, dtype=E.F32, sample_time=0.1, seed=1, auto_reset=True, reset_ref_mode=E.RESET_HYBRID, ctrl_type=E.CTRL_SEMI_MANUAL)
act, obs, rew, done = eng.alloc_io(); eng.reset(obs)
act.uniform_(-1, 1)
for i in range(12):
    eng.step(act, obs, rew, done)
eng.synchronize()
print("ok")
