"""End-to-end env-steps/s of the host-buffer entry points at 1 Mi envs, K = 10 (pinned buffers):
b747_step_host (obs / rew / done arrays, chunked copy pipeline as a CUDA graph) against b747_step_host_packed in its
three host modes (0 staged copies, 1 zero-copy records, 2 zero-copy actions + records)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from b747_rl_ctrl_b200 import engine as E  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = int(sys.argv[3]) if len(sys.argv) > 3 else 0
torch.cuda.set_device(dev)
eng = E.BatchEngine(n_envs=n, dtype=E.F32, device=dev, sample_time=K * 0.01, seed=1, auto_reset=True)
eng.reset()
pin = lambda *s, dt=torch.float32: torch.empty(*s, dtype=dt).pin_memory()
acts = [pin(n).uniform_(-1, 1) for _ in range(4)]
obs, rew, done = pin(n, 3), pin(n), pin(n, dt=torch.uint8)
out4, bits = pin(n, 4), pin((n + 31) // 32, dt=torch.int32)


def rate(fn, steps=30):
    for i in range(4):
        fn(i)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for i in range(steps):
        fn(i)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / steps
    return dt * 1e3, n / dt


ms, r = rate(lambda i: eng.step_host(acts[i % 4].numpy(), obs.numpy(), rew.numpy(), done.numpy()))
print(f"step_host (3 arrays, 4-chunk graph pipeline): {ms:.3f} ms/step  {r:.3e} env-steps/s", flush=True)
for mode in (0, 1, 2):
    eng.set_host_mode(mode)
    ms, r = rate(lambda i: eng.step_host_packed(acts[i % 4].numpy(), out4.numpy(), bits.numpy()))
    print(f"step_host_packed mode {mode}: {ms:.3f} ms/step  {r:.3e} env-steps/s", flush=True)
# device-resident reference point
a_d = torch.empty(n, device="cuda").uniform_(-1, 1)
o_d, b_d = torch.empty(n, 4, device="cuda"), torch.empty((n + 31) // 32, dtype=torch.int32, device="cuda")


def dev_step(i):
    eng.step_packed(a_d, o_d, b_d)
    eng.synchronize()


ms, r = rate(dev_step)
print(f"step_packed (device buffers, sync per step): {ms:.3f} ms/step  {r:.3e} env-steps/s", flush=True)
