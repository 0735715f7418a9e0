"""Throughput of the f32 step across env configurations (which kernel template each one runs is decided by f32_is_lean)."""
import sys, time
import torch
sys.path.insert(0, ".")
from b747_rl_ctrl_b200 import engine as E
n = 1 << 20
CFGS = {
    "canonical (LEAN)": {},
    "SPEED_MODE obs": dict(obs_type=E.OBS_SPEED_MODE),
    "PID_SPEED_AERO obs": dict(obs_type=E.OBS_PID_SPEED_AERO),
    "OSCILLATING ref": dict(reset_ref_mode=E.RESET_OSCILLATING),
    "HYBRID ref, SEMI_MANUAL": dict(reset_ref_mode=E.RESET_HYBRID, ctrl_type=E.CTRL_SEMI_MANUAL),
    "AERO disturbance": dict(disturbance_mode=E.DIST_AERO),
    "ADD_PROC mode": dict(ctrl_mode=E.MODE_ADD_PROC, action_max=1.0),
    "f64 canonical": dict(_dtype=E.F64, _n=1 << 18),
}
for name, kw in CFGS.items():
    kw = dict(kw)
    dtype = kw.pop("_dtype", E.F32); nn = kw.pop("_n", n)
    for K in (10, 5):
        try:
            eng = E.BatchEngine(n_envs=nn, dtype=dtype, sample_time=K * 0.01, seed=1, auto_reset=True, **kw)
        except Exception as e:
            print(name, "->", e); break
        s = torch.cuda.current_stream(); eng.use_stream(s.cuda_stream)
        act, obs, rew, done = eng.alloc_io(); eng.reset(obs)
        act.uniform_(-1, 1)
        for _ in range(3): eng.step(act, obs, rew, done)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 100 if dtype == E.F32 else 20
        torch.cuda.synchronize(); e0.record()
        for _ in range(steps): eng.step(act, obs, rew, done)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(f"{name:28s} K={K:2d}: {ms:.4f} ms/step  {nn / ms * 1e-6:.3f} G env-steps/s", flush=True)
        eng.close()
