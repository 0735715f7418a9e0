"""How the f32 env-step kernel's time varies over the episode phase, and how many environments sit in the
regimes that leave the kernel's fast paths (polynomial ranges, cached table intervals).

All environments are reset together, so the bench's launches sweep the 200-step episode in phase; the
numbers here say which phase costs what and why.  Usage: python tools/phase_probe.py [K] [sticky]
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from b747_rl_ctrl_b200 import engine as E

K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
sticky = len(sys.argv) > 2 and sys.argv[2] == "sticky"
n = 1 << 20
dev = torch.device("cuda", 0)
eng = E.BatchEngine(n_envs=n, dtype=E.F32, sample_time=K * 0.01, seed=1)
eng.use_stream(torch.cuda.current_stream().cuda_stream)
act, obs, rew, done = eng.alloc_io()
eng.reset(obs)
gen = torch.Generator(device=dev).manual_seed(1234)
pool = [torch.empty(n, device=dev).uniform_(-1, 1, generator=gen) for _ in range(8)]
ep = int(round(20.0 / (K * 0.01)))
steps = 2 * ep + 20
for i in range(5):
    eng.step(pool[0] if sticky else pool[i % 8], obs, rew, done)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
torch.cuda.synchronize()
ev[0].record()
for i in range(steps):
    eng.step(pool[0] if sticky else pool[i % 8], obs, rew, done)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(steps)])
print(f"K={K} sticky={sticky}: mean {ms.mean():.4f} ms/step -> {n / ms.mean() * 1e3:.4e} env-steps/s; "
      f"min {ms.min():.4f} max {ms.max():.4f}")
bins = 10
for b in range(0, steps, bins):
    print(f"  steps {b + 5:4d}..{b + 5 + bins - 1:4d} (phase {(b + 5) % ep:3d}): {ms[b:b + bins].mean():.4f} ms")
eng.close()

# regime census on a smaller batch with the DLL's signals exported
m = 1 << 16
eng = E.BatchEngine(n_envs=m, dtype=E.F32, sample_time=K * 0.01, seed=1, export_signals=True)
eng.use_stream(torch.cuda.current_stream().cuda_stream)
act, obs, rew, done = eng.alloc_io()
eng.reset(obs)
pool = [torch.empty(m, device=dev).uniform_(-1, 1, generator=gen) for _ in range(8)]
prev = None
print("phase  h>11000  |alpha|>20deg  |tan a|>.75  |theta|>1.2  |theta|>pi/2  V<60  any-axis-interval-change")
for i in range(ep):
    eng.step(pool[0] if sticky else pool[i % 8], obs, rew, done)
    if i % 10 == 9 or i == ep - 2:
        torch.cuda.synchronize()
        h = eng.get("sig_state_y"); al = eng.get("sig_alpha"); th = eng.get("sig_state_vartheta"); V = eng.get("sig_V")
        Ma = eng.get("sig_Mach")
        Vx = eng.get("sig_state_Vx")
        print(f"{i + 1:5d}  h<0:{np.mean(h < 0):.4f} Vx<0:{np.mean(Vx < 0):.4f} nan:{np.mean(~np.isfinite(h)):.4f} {np.mean(h > 11000):7.4f}  {np.mean(np.abs(al) > np.radians(20)):12.5f}  "
              f"{np.mean(np.abs(np.tan(al)) > 0.75):10.5f}  {np.mean(np.abs(th) > 1.2):10.5f}  "
              f"{np.mean(np.abs(th) > np.pi / 2):11.5f}  {np.mean(V < 60):6.4f}  "
              f"Mach[{Ma.min():.3f},{Ma.max():.3f}] alpha_deg[{np.degrees(al.min()):.1f},{np.degrees(al.max()):.1f}]")
eng.close()
