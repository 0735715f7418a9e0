import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from b747_rl_ctrl_b200 import engine as E
from oracle import oracle as O
O.build()
K, n, hold, seed = 10, 512, 200, 33
kw = dict(sample_time=K * 0.01)
cfg_o = O.make_cfg(seed=seed, **kw)
eng = E.BatchEngine(n_envs=n, dtype=E.F32, seed=seed, auto_reset=True, export_signals=True, **kw)
ob = O.OracleBatch(cfg_o, n)
eng.reset(); ob.reset()
rng = np.random.default_rng(seed)
hist = []
for k in range(420):
    if not (hold and k % hold):
        a = rng.uniform(-1, 1, n).astype(np.float32)
    obs, rew, done = eng.step_host(a)
    o_o, r_o, d_o, t_o = ob.step(a.astype(np.float64))
    er = np.abs(rew - r_o)
    j = int(er.argmax())
    hist.append((k, j, er[j], rew[j], r_o[j], obs[j].copy(), o_o[j].copy()))
    if er[j] > 0.05:
        print("step", k, "env", j, "rew gpu", rew[j], "oracle", r_o[j], "obs gpu", obs[j], "oracle", o_o[j])
        for nm in ("sig_state_vartheta", "sig_state_wz", "sig_alpha", "sig_V", "sig_dvartheta", "sig_dvartheta_dt", "sig_dvartheta_dt_dt", "sig_ITSE", "sig_U_com_PID", "sig_state_Vx", "sig_state_Vy", "sig_state_y"):
            print("   ", nm, eng.get(nm)[j])
        break
