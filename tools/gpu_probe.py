"""Scratch GPU probe: error statistics of the CUDA paths against the oracle (prints, no asserts)."""
import math
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from b747_rl_ctrl_b200 import engine as E
from oracle import oracle as O


def run(dtype, n, steps, K, seed=3, **kw):
    cfg_o = O.make_cfg(sample_time=K * 0.01, seed=seed, **kw)
    ob = O.OracleBatch(cfg_o, n)
    eng = E.BatchEngine(n_envs=n, dtype=dtype, sample_time=K * 0.01, seed=seed, auto_reset=True, **kw)
    act, obs, rew, done = eng.alloc_io()
    eng.reset(obs)
    o0 = ob.reset()
    eng.synchronize()
    print("reset obs max abs", float(obs.abs().max()), float(np.abs(o0).max()))
    rng = np.random.default_rng(0)
    worst = np.zeros(3)
    worst_obs = np.zeros(eng.obs_dim)
    done_mismatch = 0
    for k in range(steps):
        a = rng.uniform(-1, 1, n)
        act.copy_(torch.from_numpy(a).to(act.dtype))
        eng.step(act, obs, rew, done)
        o_o, r_o, d_o, _ = ob.step(a if dtype == E.F64 else a.astype(np.float32).astype(np.float64))
        eng.synchronize()
        og, rg, dg = obs.cpu().numpy().astype(np.float64), rew.cpu().numpy().astype(np.float64), done.cpu().numpy().astype(bool)
        done_mismatch += int((dg != d_o).sum())
        eo = np.abs(og - o_o).max(axis=0)
        worst_obs = np.maximum(worst_obs, eo)
        worst[0] = max(worst[0], np.abs(rg - r_o).max())
        rel = np.abs(og - o_o) / np.maximum(np.abs(o_o), 1e-6)
        worst[1] = max(worst[1], rel.max())
        if k in (0, 1, 9, 99, 399, steps - 1):
            print(f" step {k+1}: max|dobs| {eo}, max|drew| {np.abs(rg - r_o).max():.3e}, done {dg.sum()}/{d_o.sum()}")
    print(f"dtype={'f64' if dtype == E.F64 else 'f32'} n={n} steps={steps} K={K}: worst |dobs|={worst_obs}, |drew|={worst[0]:.3e}, "
          f"rel obs {worst[1]:.3e}, done mismatches {done_mismatch}")
    print("stats", eng.episode_stats())


def bench(dtype, n, K, iters=20):
    eng = E.BatchEngine(n_envs=n, dtype=dtype, sample_time=K * 0.01)
    act, obs, rew, done = eng.alloc_io()
    eng.reset(obs)
    act.uniform_(-1, 1)
    for _ in range(3):
        eng.step(act, obs, rew, done)
    eng.synchronize()
    t = time.time()
    for _ in range(iters):
        eng.step(act, obs, rew, done)
    eng.synchronize()
    dt = (time.time() - t) / iters
    print(f"bench dtype={'f64' if dtype == E.F64 else 'f32'} n={n} K={K}: {dt*1e3:.3f} ms/step -> {n/dt:.3e} env-steps/s")


def variants():
    """Worst f32-vs-oracle deviations per configuration family (to set the stated bounds)."""
    sys.path.insert(0, "tests")
    from tests.test_gpu_parity import VARIANTS
    for name, kw in sorted(VARIANTS.items()):
        steps = 320 if name == "K1_tk3" else 420
        n = 512
        cfg_o = O.make_cfg(seed=9, **kw)
        eng = E.BatchEngine(n_envs=n, dtype=E.F32, seed=9, auto_reset=True, **kw)
        ob = O.OracleBatch(cfg_o, n)
        eng.reset(); ob.reset()
        rng = np.random.default_rng(9)
        amax = 1.0 if cfg_o.norm_act else cfg_o.action_max
        term = np.zeros((n, eng.obs_dim), np.float32)
        wo = np.zeros(eng.obs_dim); wrel = np.zeros(eng.obs_dim); wr = 0.0; bad = 0
        for k in range(steps):
            a = rng.uniform(-amax, amax, n).astype(np.float32)
            obs, rew, done, term = eng.step_host(a, terminal_obs=term)
            o_o, r_o, d_o, t_o = ob.step(a.astype(np.float64))
            bad += int((done.astype(bool) != d_o).sum())
            e = np.abs(term.astype(np.float64) - t_o)
            wo = np.maximum(wo, e.max(axis=0))
            wrel = np.maximum(wrel, (e / np.maximum(np.abs(t_o), 1e-2)).max(axis=0))
            wr = max(wr, np.abs(rew - r_o).max())
        print(f"{name:28s} done-mismatch {bad} |drew| {wr:.2e} |dobs| {np.array2string(wo, precision=1)} rel(floor 1e-2) {np.array2string(wrel, precision=1)}", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    if len(sys.argv) > 1 and sys.argv[1] == "variants":
        variants()
        run(E.F32, 512, 1000, 5)
        run(E.F32, 256, 1000, 10)
        sys.exit(0)
    run(E.F64, 256, 420, 5)
    run(E.F32, 256, 420, 5)
    run(E.F64, 64, 60, 5, obs_type=O.OBS_PID_SPEED_AERO, reset_ref_mode=O.RESET_OSCILLATING)
    run(E.F32, 64, 60, 5, obs_type=O.OBS_PID_SPEED_AERO, reset_ref_mode=O.RESET_OSCILLATING)
    run(E.F64, 64, 60, 5, obs_type=O.OBS_MODEL_STATE, reset_ref_mode=O.RESET_HYBRID, ctrl_mode=O.MODE_ADD_PROC, action_max=1.0)
    run(E.F32, 64, 60, 5, obs_type=O.OBS_MODEL_STATE, reset_ref_mode=O.RESET_HYBRID, ctrl_mode=O.MODE_ADD_PROC, action_max=1.0)
    bench(E.F64, 4096, 5)
    bench(E.F64, 1 << 18, 5)
    bench(E.F32, 1 << 20, 10)
    bench(E.F32, 1 << 20, 5)
    bench(E.F32, 1 << 20, 1)
