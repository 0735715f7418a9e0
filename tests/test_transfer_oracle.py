"""Step-response metrics and the Storage recorder on the CPU side (no GPU): the oracle's restatement of
Controller._post_step / calc_stepinfo (core/controller.py:209-228, tools/general.py:46-61) against the reference DLL's
K6 episodes and the numbers the reference published (tensorboard.xlsx transfer_custom/*), the committed
golden file, and the host helper the Controller facade uses for arbitrary recordings."""
import json
import math
import os
import random

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
DEG = math.pi / 180
GOLD = json.load(open(os.path.join(HERE, "golden", "transfer_golden.json")))
# SURVEY.md section 6, per-case values from executing the DLL: overshoot %, settling s, quality
K6 = {5: (8.919, 10.450, 0.8327), -5: (10.805, 12.130, 0.6950), 10: (8.360, 10.850, 0.7944), -10: (8.968, 11.770, 0.6885)}


def _episode(O, env, deg):
    env.enable_storage(2000)
    env.reset_to(O.episode([0, 11000, 250, 0, 0, 0], vref=deg * DEG))


def _cfg(O):
    return O.make_cfg(reset_ref_mode=O.RESET_NONE, ctrl_mode=O.MODE_ADD_PROC, action_max=1.0, rew_type=O.REW_QUALITY)


def test_dll_transfer_metrics_match_published(dllref, oracle):
    O = oracle
    over, settle, qual = [], [], []
    for deg, (o_ref, t_ref, q_ref) in K6.items():
        env = O.RefEnv(_cfg(O))
        _episode(O, env, deg)
        _, rew, done = env.rollout(np.zeros(400), auto_reset=False)
        info = env.stepinfo_SS()
        assert info["overshoot"] == pytest.approx(o_ref, abs=6e-4)
        assert info["settling_time"] == pytest.approx(t_ref, abs=1e-9)
        assert rew[-1] == pytest.approx(q_ref, abs=6e-5)
        g = GOLD[str(deg)]
        assert info == g["stepinfo"] and env.stepinfo_CS() == g["stepinfo_CS"] and rew[-1] == g["quality"]
        st = env.storage
        assert len(st["t"]) == g["n"] == 2000
        for k, v in g["samples"].items():
            assert list(st[k][49::50]) == v, k
        over.append(abs(info["overshoot"])); settle.append(info["settling_time"]); qual.append(rew[-1])
    # the reference's first logged transfer_custom/* point (policy ~ 0 => pure PID): 9.26 %, 11.30 s, 0.753
    assert np.mean(over) == pytest.approx(9.263, abs=1e-3)
    assert np.mean(settle) == pytest.approx(11.300, abs=1e-9)
    assert np.mean(qual) == pytest.approx(0.7527, abs=5e-5)


def test_restatement_storage_tracks_golden(oracle):
    """The C restatement (what the GPU tests compare with) against the DLL-generated golden recording."""
    O = oracle
    ob = O.OracleBatch(_cfg(O), 4)
    degs = (5, -5, 10, -10)
    views = [ob.env(i) for i in range(4)]
    for v in views:
        v.enable_storage(2000)
    ob.reset_to([O.episode([0, 11000, 250, 0, 0, 0], vref=d * DEG) for d in degs])
    for _ in range(400):
        _, rew, done, _ = ob.step(np.zeros(4), auto_reset=False)
    assert done.all()
    for i, d in enumerate(degs):
        g = GOLD[str(d)]
        info = views[i].stepinfo_SS()
        for k in ("overshoot", "rise_time", "settling_time", "static_error"):
            assert info[k] == pytest.approx(g["stepinfo"][k], rel=1e-8, abs=1e-10), (d, k)
        assert rew[i] == pytest.approx(g["quality"], rel=1e-9)
        st = views[i].storage
        for k, v in g["samples"].items():
            assert np.allclose(st[k][49::50], v, rtol=1e-8, atol=1e-9), (d, k)


def test_host_calc_stepinfo_agrees_with_oracle(oracle):
    """b747_rl_ctrl_b200.tools.general.calc_stepinfo (used for backed-up / non-constant recordings) vs the oracle's
    independent statement, incl. the cases where the reference returns None."""
    from b747_rl_ctrl_b200.tools.general import calc_stepinfo
    rng = random.Random(3)
    for case in range(200):
        n = rng.randint(2, 60)
        base = rng.choice([-1, 1]) * rng.uniform(0.5, 10)
        kind = case % 4
        if kind == 0:    # well-behaved second-order-like response
            ys = [base * (1 - math.exp(-0.2 * k) * math.cos(0.5 * k)) for k in range(n)]
        elif kind == 1:  # never rises
            ys = [base * 0.1 * rng.random() for _ in range(n)]
        elif kind == 2:  # random walk
            ys = list(np.cumsum([rng.uniform(-1, 1) for _ in range(n)]))
        else:            # reaches the band only at the very last sample
            ys = [0.0] * (n - 1) + [base]
        ys[0] = ys[0] if ys[0] != base else 0.0
        ts = [0.01 * (k + 1) for k in range(n)]
        a = calc_stepinfo(ys, base, ts=ts)
        b = oracle.stepinfo(ys, base, ts)
        assert a == b, (case, a, b)
    assert calc_stepinfo([1.0, 2.0], 0.0, ts=[0.0, 1.0])["overshoot"] is None
