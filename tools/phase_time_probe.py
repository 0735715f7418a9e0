"""Kernel time of the K = 10 step as a function of the flight phase (lock-step envs, one CUDA-event pair per launch over a
whole 200-step episode and the auto-reset step), the steady-state time with spread phases, and the host-side overhead of
a launch + synchronise per step."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from b747_rl_ctrl_b200 import engine as E  # noqa: E402

n, K = 1 << 20, 10
eng = E.BatchEngine(n_envs=n, dtype=E.F32, sample_time=K * 0.01, seed=1, auto_reset=True)
st = torch.cuda.current_stream()
eng.use_stream(st.cuda_stream)
act, obs, rew, done = eng.alloc_io()
eng.reset(obs)
gen = torch.Generator(device="cuda").manual_seed(1)
pool = [torch.empty(n, device="cuda").uniform_(-1, 1, generator=gen) for _ in range(8)]
for rep in range(2):  # second episode: the handle is warm, hints persisted
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(201)]
    ev[0].record()
    for k in range(200):
        eng.step(pool[k % 8], obs, rew, done)
        ev[k + 1].record()
    torch.cuda.synchronize()
    ms = np.array([ev[k].elapsed_time(ev[k + 1]) for k in range(200)])
    print(f"episode {rep}: per-step kernel ms by phase (10-step means):", np.round(ms.reshape(20, 10).mean(axis=1), 4).tolist())
    print(f"   step 0: {ms[0]:.4f}  step 199 (all envs auto-reset): {ms[199]:.4f}  mean {ms.mean():.4f}")
# zero actions: how much of the phase dependence is the random elevator
eng.reset(obs)
z = torch.zeros(n, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(201)]
ev[0].record()
for k in range(200):
    eng.step(z, obs, rew, done)
    ev[k + 1].record()
torch.cuda.synchronize()
ms = np.array([ev[k].elapsed_time(ev[k + 1]) for k in range(200)])
print("zero actions: per-step kernel ms by phase:", np.round(ms.reshape(20, 10).mean(axis=1), 4).tolist())
# spread phases without auto-reset cost: desync, then time
ids = torch.arange(n, device="cuda")
eng.reset(obs)
for t in range(200):
    eng.step(pool[t % 8], obs, rew, done)
    eng.reset(mask=((ids % 200) == t).to(torch.uint8))
for order, name in ((ids % 200, "id % 200 (every warp holds 32 phases)"),):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
    ev[0].record()
    for k in range(40):
        eng.step(pool[k % 8], obs, rew, done)
        ev[k + 1].record()
    torch.cuda.synchronize()
    ms = np.array([ev[k].elapsed_time(ev[k + 1]) for k in range(40)])
    print(f"steady state, phases {name}: mean {ms.mean():.4f} min {ms.min():.4f} max {ms.max():.4f}; episodes {eng.episode_stats()[0]:.0f}")
# warp-coherent phases: whole 32-env tiles share a phase ((id // 32) % 200)
eng.reset(obs)
for t in range(200):
    eng.step(pool[t % 8], obs, rew, done)
    eng.reset(mask=(((ids // 32) % 200) == t).to(torch.uint8))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
ev[0].record()
for k in range(40):
    eng.step(pool[k % 8], obs, rew, done)
    ev[k + 1].record()
torch.cuda.synchronize()
ms = np.array([ev[k].elapsed_time(ev[k + 1]) for k in range(40)])
print(f"steady state, phases (id // 32) % 200 (a warp shares one phase): mean {ms.mean():.4f}; episodes {eng.episode_stats()[0]:.0f}")
# host overhead of launch + sync per step
for label, fn in (("step + synchronize", lambda k: (eng.step(pool[k % 8], obs, rew, done), eng.synchronize())),):
    for k in range(5):
        fn(k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for k in range(50):
        fn(k)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 50 * 1e3
    print(f"{label}: wall {wall:.4f} ms/step, GPU span {e0.elapsed_time(e1) / 50:.4f} ms/step")
