"""PCIe copy bandwidth on this box (pinned buffers) and step_host throughput vs chunk count."""
import sys, time
import torch
sys.path.insert(0, ".")
dev = torch.device("cuda")
def bw(nbytes, direction, reps=20):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
    e1.record(); torch.cuda.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
for nb in (128 << 10, 512 << 10, 2 << 20, 16 << 20, 64 << 20):
    print(f"{nb >> 10:8d} KiB  h2d {bw(nb, 'h2d'):6.1f} GB/s   d2h {bw(nb, 'd2h'):6.1f} GB/s", flush=True)
# duplex
nb = 16 << 20
h1 = torch.empty(nb, dtype=torch.uint8).pin_memory(); d1 = torch.empty(nb, dtype=torch.uint8, device=dev)
h2 = torch.empty(nb, dtype=torch.uint8).pin_memory(); d2 = torch.empty(nb, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(20):
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"duplex 16 MiB each way: {nb * 20 / dt / 1e9:.1f} GB/s per direction", flush=True)

from b747_rl_ctrl_b200 import engine as E
n = 1 << 20
eng = E.BatchEngine(n_envs=n, dtype=E.F32, sample_time=0.1, seed=1, auto_reset=True)
eng.reset()
ha = torch.empty(n).uniform_(-1, 1).pin_memory(); ho = torch.empty(n, 3).pin_memory(); hr = torch.empty(n).pin_memory()
hd = torch.empty(n, dtype=torch.uint8).pin_memory()
for ch in (1, 2, 4, 8, 16, 32):
    eng.set_host_chunks(ch)
    for _ in range(3): eng.step_host(ha.numpy(), ho.numpy(), hr.numpy(), hd.numpy())
    t = time.perf_counter()
    for _ in range(20): eng.step_host(ha.numpy(), ho.numpy(), hr.numpy(), hd.numpy())
    dt = (time.perf_counter() - t) / 20
    print(f"chunks {ch:2d}: {dt * 1e3:.3f} ms/step  {n / dt / 1e9:.3f} G env-steps/s", flush=True)
