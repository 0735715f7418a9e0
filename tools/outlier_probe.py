"""Trace single environments of a configuration family: f32 engine against the oracle, step by step, with the
oracle's internal signals around the step where the deviation takes off.
usage: python tools/outlier_probe.py family env [family env ...]"""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from b747_rl_ctrl_b200 import engine as E  # noqa: E402
from oracle import oracle as O  # noqa: E402
from test_gpu_parity import VARIANTS  # noqa: E402

O.build()
SIGS = ["sim_time", "dvartheta", "dvartheta_dt", "U_com_PID", "U_com", "deltaz_RP", "alpha", "V", "Mach", "CYa", "vartheta_zh"]
args = sys.argv[1:]
for name, env in zip(args[0::2], map(int, args[1::2])):
    kw = dict(VARIANTS.get(name, {}))
    n, steps, seed = 512, 420, 9
    cfg = O.make_cfg(seed=seed, **kw)
    eng = E.BatchEngine(n_envs=n, dtype=E.F32, seed=seed, auto_reset=True, **kw)
    e64 = E.BatchEngine(n_envs=n, dtype=E.F64, seed=seed, auto_reset=True, **kw)
    ob = O.OracleBatch(cfg, n)
    eng.reset(); e64.reset(); ob.reset()
    rng = np.random.default_rng(seed)
    amax = 1.0 if cfg.norm_act else cfg.action_max
    od = eng.obs_dim
    t32, t64 = np.zeros((n, od), np.float32), np.zeros((n, od))
    rows = []
    for s in range(steps):
        a = rng.uniform(-amax, amax, n).astype(np.float32)
        _, r32, d32, t32 = eng.step_host(a, terminal_obs=t32)
        _, r64, d64, t64 = e64.step_host(a.astype(np.float64), terminal_obs=t64)
        _, r_o, d_o, t_o = ob.step(a.astype(np.float64))
        dev = np.abs(t32[env].astype(np.float64) - t_o[env]) / (1 + np.abs(t_o[env]))
        rows.append((s, dev.max(), int(dev.argmax()), abs(r32[env] - r_o[env]), np.abs(t64[env] - t_o[env]).max(), bool(d_o[env]),
                     [ob.gather(k)[env] for k in SIGS], ob.gather("state", 4)[env], ob.gather("state", 1)[env], a[env],
                     float(eng.get("th")[env]), float(eng.get("h")[env])))
    devs = np.array([r[1] for r in rows])
    on = int(np.argmax(devs > 20 * np.median(devs[:50]) + 1e-7))
    print(f"== {name} env {env}: max dev {devs.max():.2e} at step {devs.argmax()}, onset ~ step {on}; f64 engine max dev "
          f"{max(r[4] for r in rows):.1e}")
    print("step  dev_obs(k) dev_rew  done |", " ".join(f"{k:>11s}" for k in SIGS), "| theta h action | th32 h32")
    for r in rows[max(0, on - 6):on + 12] + rows[int(devs.argmax()) - 1:int(devs.argmax()) + 2]:
        print(f"{r[0]:4d} {r[1]:.2e}({r[2]}) {r[3]:.1e} {int(r[5])} |", " ".join(f"{v:11.4g}" for v in r[6]),
              f"| {r[7]:.5f} {r[8]:.2f} {r[9]:+.3f} | {r[10]:.5f} {r[11]:.2f}")
    eng.close(); e64.close()
