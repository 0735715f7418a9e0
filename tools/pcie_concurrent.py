"""Host-link probe with N GPUs busy AT THE SAME TIME (one rank per GPU, torchrun): what limits the end-to-end
(host-buffer) step as the GPU count grows.  Every phase starts on a barrier and runs for a fixed number of repeats;
each rank prints its own rate and rank 0 the aggregate.

  d2h / h2d      16 MiB pinned cudaMemcpyAsync (copy engines)
  packed m2/m0   b747_step_host_packed, 1 Mi envs, K = 10: zero-copy (mode 2) and staged copies (mode 0)
  kernel         the same step with device buffers (no host traffic): the GPU-side ceiling
usage: torchrun --nproc-per-node N tools/pcie_concurrent.py"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from b747_rl_ctrl_b200 import engine as E  # noqa: E402
import bench  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
aff = bench._bind_near_gpu(torch, lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def agg(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t)
    return float(t.item())


def timed(fn, reps):
    for i in range(3):
        fn(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(reps):
        fn(i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    return dt / reps


nb = 16 << 20
h = torch.empty(nb, dtype=torch.uint8).pin_memory()
d = torch.empty(nb, dtype=torch.uint8, device=dev)
res = {}
res["d2h_GBs"] = nb / timed(lambda i: (h.copy_(d, non_blocking=True), torch.cuda.synchronize()), 40) / 1e9
res["h2d_GBs"] = nb / timed(lambda i: (d.copy_(h, non_blocking=True), torch.cuda.synchronize()), 40) / 1e9
n, K = 1 << 20, 10
eng = E.BatchEngine(n_envs=n, dtype=E.F32, device=lr, sample_time=K * 0.01, seed=1, env_id_offset=rank * n, auto_reset=True)
eng.reset()
pin = lambda *s, dt=torch.float32: torch.empty(*s, dtype=dt).pin_memory()
acts = [pin(n).uniform_(-1, 1) for _ in range(4)]
out4, bits = pin(n, 4), pin((n + 31) // 32, dt=torch.int32)
obs, rew, done = pin(n, 3), pin(n), pin(n, dt=torch.uint8)
a_d = torch.empty(n, device=dev).uniform_(-1, 1)
o_d, b_d = torch.empty(n, 4, device=dev), torch.empty((n + 31) // 32, dtype=torch.int32, device=dev)
for mode in (2, 1, 0):
    eng.set_host_mode(mode)
    res[f"packed_m{mode}_Gsteps"] = n / timed(lambda i: eng.step_host_packed(acts[i % 4].numpy(), out4.numpy(), bits.numpy()), 40) / 1e9
res["step_host_Gsteps"] = n / timed(lambda i: eng.step_host(acts[i % 4].numpy(), obs.numpy(), rew.numpy(), done.numpy()), 40) / 1e9
res["kernel_Gsteps"] = n / timed(lambda i: (eng.step_packed(a_d, o_d, b_d), eng.synchronize()), 40) / 1e9
line = f"rank {rank}/{world} gpu {lr} affinity [{aff}]: " + "  ".join(f"{k} {v:.2f}" for k, v in res.items())
tot = {k: agg(v) for k, v in res.items()}
for r in range(world):
    if r == rank:
        print(line, flush=True)
    if world > 1:
        dist.barrier()
if rank == 0:
    print(f"N={world} aggregate: " + "  ".join(f"{k} {v:.2f}" for k, v in tot.items()), flush=True)
    print(f"N={world} host traffic at packed_m2: {tot['packed_m2_Gsteps'] * 20.125:.1f} GB/s "
          f"(4 B action + 16.125 B results per env-step)", flush=True)
if world > 1:
    dist.destroy_process_group()
