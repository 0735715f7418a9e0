"""Small end-to-end exercise of every kernel tier and host path, a quick functional pass (every tier, chunked host path, trace handles)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from b747_rl_ctrl_b200 import engine as E
rng = np.random.default_rng(0)
CASES = [dict(), dict(obs_type=E.OBS_SPEED_MODE), dict(reset_ref_mode=E.RESET_OSCILLATING),
         dict(reset_ref_mode=E.RESET_HYBRID, ctrl_type=E.CTRL_SEMI_MANUAL), dict(disturbance_mode=E.DIST_AERO, obs_type=E.OBS_MODEL_STATE)]
for dtype in (E.F32, E.F64):
    for kw in CASES:
        n = 1000 if dtype == E.F32 else 300
        eng = E.BatchEngine(n_envs=n, dtype=dtype, seed=3, sample_time=0.05, tk=0.3, **kw)
        eng.set_host_chunks(3)
        eng.reset()
        for k in range(8):
            a = rng.uniform(-1, 1, n)
            term = np.zeros((n, eng.obs_dim), eng.np_dtype)
            eng.step_host(a, terminal_obs=term)
        st = eng.episode_stats()
        assert st[0] == n, st
        eng.close()
eng = E.BatchEngine(n_envs=77, dtype=E.F32, seed=3, track_transfer=True, record_capacity=64, export_signals=True)
eng.reset()
for k in range(5):
    eng.step_host(rng.uniform(-1, 1, 77))
eng.close()
print("all tiers ok")
