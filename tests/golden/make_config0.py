#!/usr/bin/env python
"""BASELINE.json configs[0] golden: ONE environment, 10 000 env steps (K = 5: 50 000 model steps) of fixed-seed random
elevator actions, produced by EXECUTING THE REFERENCE DLL (oracle/_ref) under the C env layer -- once episodic (tk = 20:
25 episodes, exercises done / reset) and once as a single long flight (tk = 1e9, pure dynamics).
Stored compactly (checkpoints + sums) in config0_golden.json; the action stream is regenerated from the seed.

    make -C oracle ref && python tests/golden/make_config0.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import dllref  # noqa: E402
from oracle import oracle as O  # noqa: E402

N_STEPS, SEED, ACT_SEED = 10000, 11, 2022
CASES = {"episodic": dict(), "long_flight": dict(tk=1.0e9)}


def actions():
    return np.random.default_rng(ACT_SEED).uniform(-1.0, 1.0, N_STEPS)


def digest(obs, rew, done, every):
    idx = list(range(every - 1, N_STEPS, every))
    ends = np.nonzero(done)[0]
    starts = np.concatenate([[0], ends[:-1] + 1]) if len(ends) else np.array([], dtype=int)
    return dict(every=every, obs=[[float(x) for x in obs[k]] for k in idx], rew=[float(rew[k]) for k in idx],
                n_done=int(done.sum()), done_steps=[int(k) for k in ends[:40]],
                episode_returns=[float(rew[a:b + 1].sum()) for a, b in zip(starts, ends)],
                sum_rew=float(rew.sum()), sum_obs=[float(x) for x in obs.sum(axis=0)],
                sum_abs_obs=[float(x) for x in np.abs(obs).sum(axis=0)])


def run(make_env):
    out = {}
    a = actions()
    for name, kw in CASES.items():
        env = make_env(O.make_cfg(seed=SEED, **kw))
        env.reset()
        obs, rew, done = env.rollout(a, auto_reset=True)
        out[name] = digest(obs, rew, done, 400 if name == "episodic" else 500)
    return out


if __name__ == "__main__":
    if not dllref.available():
        sys.exit("oracle/_ref/libb747_ref.so missing: run `make -C oracle ref` where /root/reference is mounted")
    g = dict(n_steps=N_STEPS, seed=SEED, action_seed=ACT_SEED, cases=run(lambda cfg: O.RefEnv(cfg, env_id=0)))
    with open(os.path.join(HERE, "config0_golden.json"), "w") as f:
        json.dump(g, f, indent=0)
    print({k: (v["n_done"], v["sum_rew"]) for k, v in g["cases"].items()})
