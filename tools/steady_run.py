"""Driver for ncu captures of the step kernel in steady state: 1 Mi envs, K substeps, episode phases spread uniformly
(200 / 400 desync launches), then `steps` launches.  usage: steady_run.py [K] [steps] [config kwargs as k=v ...]"""
import sys

import torch

sys.path.insert(0, ".")
from b747_rl_ctrl_b200 import engine as E  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
kw = {}
for a in sys.argv[3:]:
    k, v = a.split("=")
    kw[k] = int(v)
n = 1 << 20
eng = E.BatchEngine(n_envs=n, dtype=E.F32, sample_time=K * 0.01, seed=1, auto_reset=True, **kw)
eng.use_stream(torch.cuda.current_stream().cuda_stream)
act, obs, rew, done = eng.alloc_io()
eng.reset(obs)
gen = torch.Generator(device="cuda").manual_seed(1)
pool = [torch.empty(n, device="cuda").uniform_(-1, 1, generator=gen) for _ in range(8)]
ep_len = 2000 // K
ids = torch.arange(n, device="cuda")
for t in range(ep_len):
    eng.step(pool[t % 8], obs, rew, done)
    eng.reset(mask=((ids % ep_len) == t).to(torch.uint8))
import os
packed = os.environ.get("B747_PACKED", "0") != "0"
out4 = torch.empty(n, eng.record_floats, device="cuda")
bits = torch.empty((n + 31) // 32, dtype=torch.int32, device="cuda")
for k in range(5):
    eng.step_packed(pool[k % 8], out4, bits) if packed else eng.step(pool[k % 8], obs, rew, done)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(steps):
    eng.step_packed(pool[k % 8], out4, bits) if packed else eng.step(pool[k % 8], obs, rew, done)
e1.record()
torch.cuda.synchronize()
print(f"K={K} steady state{' (packed outputs)' if packed else ''}: {e0.elapsed_time(e1) / steps:.4f} ms/step, "
      f"episodes {eng.episode_stats()[0]:.0f}")
