"""Time the f32 env-step kernel for the library named by $B747_LIB_PATH (kernel-only, CUDA events)."""
import sys
import torch
sys.path.insert(0, ".")
from b747_rl_ctrl_b200 import engine as E

n = 1 << 20
for K in (10, 5, 1, 10):
    eng = E.BatchEngine(n_envs=n, dtype=E.F32, sample_time=K * 0.01)
    eng.use_stream(torch.cuda.current_stream().cuda_stream)
    act, obs, rew, done = eng.alloc_io()
    eng.reset(obs)
    act.uniform_(-1, 1)
    for _ in range(300):
        eng.step(act, obs, rew, done)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 600
    e0.record()
    for _ in range(iters):
        eng.step(act, obs, rew, done)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"K={K}: {ms:.4f} ms/step  {n / ms * 1e3:.4e} env-steps/s", flush=True)
    eng.close()
