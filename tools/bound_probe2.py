import sys
import numpy as np
sys.path.insert(0, ".")
from b747_rl_ctrl_b200 import engine as E
from oracle import oracle as O
O.build()
K, n, seed = 10, 4096, 21
kw = dict(sample_time=K * 0.01)
eng = E.BatchEngine(n_envs=n, dtype=E.F32, seed=seed, auto_reset=True, export_signals=True, **kw)
ob = O.OracleBatch(O.make_cfg(seed=seed, **kw), n)
eng.reset(); ob.reset()
rng = np.random.default_rng(seed)
first = {}
for k in range(200):
    a = rng.uniform(-1, 1, n).astype(np.float32)
    obs, rew, done = eng.step_host(a)
    o_o, r_o, d_o, _ = ob.step(a.astype(np.float64))
    e = np.abs(obs - o_o).max(axis=1)
    for j in np.nonzero(e > 1e-6)[0]:
        if j not in first:
            first[j] = k
            print(f"env {j} step {k}: dev {e[j]:.2e} alpha {eng.get('sig_alpha')[j]*57.3:.1f} deg theta {eng.get('sig_state_vartheta')[j]*57.3:.1f} wz {eng.get('sig_state_wz')[j]:.2f} V {eng.get('sig_V')[j]:.0f} Mach {eng.get('sig_Mach')[j]:.3f} h {eng.get('sig_state_y')[j]:.0f} CYa {eng.get('sig_CYa')[j]:.3f}")
print(len(first), "envs above 1e-6 in the first episode")
