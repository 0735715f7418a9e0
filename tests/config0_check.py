"""Shared checker of the BASELINE configs[0] trajectories (tests/golden/make_config0.py)."""
import numpy as np


def check_config0(case, obs, rew, done, obs_tol, rew_tol):
    """Compare a 10 000-step single-env trajectory with the DLL digest of tests/golden/make_config0.py.
    `obs` holds what the DLL run recorded: the observation BEFORE an auto-reset (terminal observation)."""
    ev = case["every"]
    idx = np.arange(ev - 1, len(rew), ev)
    assert int(done.sum()) == case["n_done"] and list(np.nonzero(done)[0][:40]) == case["done_steps"]
    assert np.abs(obs[idx] - np.array(case["obs"])).max() <= obs_tol
    assert np.abs(rew[idx] - np.array(case["rew"])).max() <= rew_tol
    ends = np.nonzero(done)[0]
    starts = np.concatenate([[0], ends[:-1] + 1]) if len(ends) else np.array([], dtype=int)
    rets = np.array([rew[a:b + 1].sum() for a, b in zip(starts, ends)])
    if len(rets):
        assert np.abs(rets - np.array(case["episode_returns"])).max() <= 400 * rew_tol
    assert abs(rew.sum() - case["sum_rew"]) <= len(rew) * rew_tol
    assert np.abs(obs.sum(axis=0) - np.array(case["sum_obs"])).max() <= len(rew) * obs_tol
