"""Find where a 4096-env slice handle departs from the same global env ids of a 1 Mi-env handle."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from b747_rl_ctrl_b200 import engine as E

n, K = 1 << 20, 10
kw = dict(sample_time=K * 0.01)
eng = E.BatchEngine(n_envs=n, dtype=E.F32, seed=1, **kw)
eng.use_stream(torch.cuda.current_stream().cuda_stream)
act, obs, rew, done = eng.alloc_io()
eng.reset(obs)
lo = 777 * 128
sl = E.BatchEngine(n_envs=4096, dtype=E.F32, seed=1, env_id_offset=lo, **kw)
sl.use_stream(torch.cuda.current_stream().cuda_stream)
a2, o2, r2, d2 = sl.alloc_io()
sl.reset(o2)
gen = torch.Generator(device="cuda").manual_seed(0)
names = ["h", "th", "Vx", "Vy", "wz", "ssi", "ssf", "dvi", "itse", "d1_u", "vref", "df_x", "df_y", "rl_prev", "deltaz", "uh0", "uh1", "uh2", "uh3", "d2_u", "tick"]
avail = set(eng.field_names())
print("fields:", sorted(avail)[:80])
for k in range(205):
    act.uniform_(-1, 1, generator=gen)
    eng.step(act, obs, rew, done)
    a2.copy_(act[lo:lo + 4096])
    sl.step(a2, o2, r2, d2)
    torch.cuda.synchronize()
    bad = (obs[lo:lo + 4096] != o2).any(dim=1)
    if bad.any():
        idx = bad.nonzero().flatten().cpu().numpy()
        print(f"step {k}: {len(idx)} envs differ; first {idx[:10]}; lanes {idx[:10] % 32}")
        j = int(idx[0])
        print(" big :", obs[lo + j].cpu().numpy(), float(rew[lo + j]))
        print(" slice:", o2[j].cpu().numpy(), float(r2[j]))
        for nm in names:
            if nm in avail:
                try:
                    b = eng.get(nm)[lo + j]; s = sl.get(nm)[j]
                    if b != s:
                        print(f"   {nm}: big {b!r} slice {s!r}")
                except Exception as e:
                    print("   ", nm, e)
        break
else:
    print("no difference in 205 steps")
