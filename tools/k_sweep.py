"""Kernel time of the canonical f32 step vs K (substeps) for the library named by $B747_LIB_PATH."""
import sys, torch
sys.path.insert(0, ".")
from b747_rl_ctrl_b200 import engine as E
n = 1 << 20
for K in (1, 2, 3, 4, 5, 10):
    eng = E.BatchEngine(n_envs=n, dtype=E.F32, sample_time=K * 0.01, seed=1, auto_reset=True)
    s = torch.cuda.current_stream(); eng.use_stream(s.cuda_stream)
    act, obs, rew, done = eng.alloc_io(); eng.reset(obs)
    pool = [torch.empty(n, device="cuda").uniform_(-1, 1) for _ in range(8)]
    for i in range(5): eng.step(pool[i % 8], obs, rew, done)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(200): eng.step(pool[i % 8], obs, rew, done)
    e1.record(); torch.cuda.synchronize()
    print(f"K={K:2d}: {e0.elapsed_time(e1) / 200:.4f} ms", flush=True)
    eng.close()
