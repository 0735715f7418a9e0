"""Per-environment deviation of the f32 path from the float64 oracle, next to the oracle's OWN sensitivity to a tiny
perturbation of the action stream, for every configuration family of tests/test_gpu_parity.py (+ the canonical
config).  The per-family bounds asserted on every env by test_f32_variants_within_bound come from this data
(profiles/r2_family_probe.md).

For each family (512 envs, the test's seed and step count):
  dev_obs[e], dev_rew[e]  worst |f32 engine - oracle| over the trajectory
  sens_X[e]               worst |oracle(a (1 + X)) - oracle(a)| for X in 2^-23 (one float32 ulp), 1e-6, 1e-5: how much
                          env e amplifies an input perturbation of the size of float32 rounding
  dll[e]  (first 64 envs) worst |restatement - reference DLL| on the same episodes and actions (where oracle/_ref is
                          built): the float64 pair diverges on the same envs
usage: python tools/family_probe.py out.npz [n_envs]
"""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from b747_rl_ctrl_b200 import engine as E  # noqa: E402
from oracle import dllref, oracle as O  # noqa: E402
from test_gpu_parity import VARIANTS  # noqa: E402

PERT = {"ulp": 2.0 ** -23, "1e-6": 1e-6, "1e-5": 1e-5}


def probe(name, kw, n, steps, seed=9):
    cfg = O.make_cfg(seed=seed, **kw)
    eng = E.BatchEngine(n_envs=n, dtype=E.F32, seed=seed, auto_reset=True, **kw)
    ob = O.OracleBatch(cfg, n)
    pert = {k: O.OracleBatch(cfg, n) for k in PERT}
    eng.reset(); ob.reset()
    for p in pert.values():
        p.reset()
    rng = np.random.default_rng(seed)
    amax = 1.0 if cfg.norm_act else cfg.action_max
    od = eng.obs_dim
    out = {"dev_obs": np.zeros(n), "dev_rew": np.zeros(n), "obs_scale": np.zeros(n)}
    for k in PERT:
        out["sens_obs_" + k] = np.zeros(n)
        out["sens_rew_" + k] = np.zeros(n)
    term = np.zeros((n, od), np.float32)
    acts = []
    for s in range(steps):
        a = rng.uniform(-amax, amax, n).astype(np.float32)
        acts.append(a)
        obs, rew, done, term = eng.step_host(a, terminal_obs=term)
        o_o, r_o, d_o, t_o = ob.step(a.astype(np.float64))
        assert np.array_equal(done.astype(bool), d_o), (name, s)
        scale = 1.0 + np.abs(t_o)  # deviations relative to 1 + |obs| (un-normalised layouts carry raw magnitudes)
        out["dev_obs"] = np.maximum(out["dev_obs"], (np.abs(term.astype(np.float64) - t_o) / scale).max(axis=1))
        out["dev_rew"] = np.maximum(out["dev_rew"], np.abs(rew.astype(np.float64) - r_o))
        out["obs_scale"] = np.maximum(out["obs_scale"], np.abs(t_o).max(axis=1))
        for k, eps in PERT.items():
            _, r_p, _, t_p = pert[k].step(a.astype(np.float64) * (1.0 + eps))
            out["sens_obs_" + k] = np.maximum(out["sens_obs_" + k], (np.abs(t_p - t_o) / scale).max(axis=1))
            out["sens_rew_" + k] = np.maximum(out["sens_rew_" + k], np.abs(r_p - r_o))
    eng.close()
    if dllref.available():
        m = min(64, n)
        acts = np.stack(acts, axis=1).astype(np.float64)  # [n, steps]
        ob2 = O.OracleBatch(cfg, m)
        ob2.reset()
        o_seq = np.zeros((m, steps, od)); r_seq = np.zeros((m, steps))
        for s in range(steps):
            _, r, _, t = ob2.step(np.concatenate([acts[:m, s]]))
            o_seq[:, s], r_seq[:, s] = t, r
        dll = np.zeros(m)
        for e in range(m):
            env = O.RefEnv(cfg, env_id=e)
            env.reset()
            o_d, r_d, _ = env.rollout(acts[e], auto_reset=True)
            dll[e] = (np.abs(o_d - o_seq[e]) / (1.0 + np.abs(o_seq[e]))).max()
        out["dll"] = dll
    return out


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/family_probe.npz"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    O.build()
    fams = dict(VARIANTS)
    fams["canonical_K5"] = dict()
    res = {}
    for name, kw in sorted(fams.items()):
        steps = 320 if name == "K1_tk3" else 420
        r = probe(name, kw, n, steps)
        for k, v in r.items():
            res[f"{name}/{k}"] = v
        q = lambda x: np.quantile(x, [0.5, 0.95, 1.0])
        print(f"{name}: dev_obs med/p95/max {q(r['dev_obs'])}, dev_rew {q(r['dev_rew'])}, sens_ulp {q(r['sens_obs_ulp'])}, "
              f"sens_1e-5 {q(r['sens_obs_1e-5'])}" + (f", dll {q(r['dll'])}" if "dll" in r else ""), flush=True)
    np.savez_compressed(path, **res)


if __name__ == "__main__":
    main()
