"""GPU: the in-kernel Storage recorder and step-response tracker and the batched control test (N3)
against the oracle and against the reference DLL's K6 episodes / published transfer_custom numbers."""
import json
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
DEG = math.pi / 180
GOLD = json.load(open(os.path.join(HERE, "golden", "transfer_golden.json")))
DEGS = (5, -5, 10, -10)
NAMES = ("overshoot", "rise_time", "settling_time", "static_error")


@pytest.fixture(scope="module")
def E():
    import torch
    assert torch.cuda.is_available()
    from b747_rl_ctrl_b200 import engine
    return engine


def test_control_test_reproduces_published_transfer_numbers(E):
    """ControlTestCallback's procedure with a zero policy in ADD_PROC mode == the pure PID loop: the reference's first
    logged point (overshoot 9.26 %, settling 11.30 s, quality 0.753) and the DLL-generated golden values."""
    from b747_rl_ctrl_b200.control_test import ControlTest, run_control_test
    r = run_control_test(None, [d * DEG for d in DEGS], ctrl_mode=E.MODE_ADD_PROC, action_max=1.0, record=True)
    for i, d in enumerate(DEGS):
        g = GOLD[str(d)]
        for k in NAMES:
            assert r[k][i] == pytest.approx(g["stepinfo"][k], rel=1e-8, abs=1e-10), (d, k)
        assert r["quality"][i] == pytest.approx(g["quality"], rel=1e-9)
        assert r["length"][i] == 400
        st = r["storage"][i]
        assert len(st["t"]) == 2000
        for k, v in g["samples"].items():
            assert np.allclose(st[k][49::50], v, rtol=1e-8, atol=1e-9), (d, k)
    ct = ControlTest([d * DEG for d in DEGS], ctrl_mode=E.MODE_ADD_PROC, action_max=1.0)
    log = ct.evaluate(None)
    assert log["transfer_custom/overshoot"] == pytest.approx(9.263, abs=1e-3)
    assert log["transfer_custom/settling_time"] == pytest.approx(11.300, abs=1e-9)
    assert log["transfer_custom/quality"] == pytest.approx(0.7527, abs=5e-5)
    assert log["improved"] and ct.best_mean_quality == ct.mean_quality


def test_control_test_f32_mode_within_bound(E):
    from b747_rl_ctrl_b200.control_test import run_control_test
    r = run_control_test(None, [d * DEG for d in DEGS], ctrl_mode=E.MODE_ADD_PROC, action_max=1.0, dtype=E.F32)
    for i, d in enumerate(DEGS):
        g = GOLD[str(d)]
        assert r["overshoot"][i] == pytest.approx(g["stepinfo"]["overshoot"], abs=2e-3)       # % of the reference
        assert r["static_error"][i] == pytest.approx(g["stepinfo"]["static_error"], abs=1e-4)  # deg
        for k in ("rise_time", "settling_time"):  # a band crossing may move by one 10 ms sample
            assert abs(r[k][i] - g["stepinfo"][k]) <= 0.0100001, (d, k)
        assert r["quality"][i] == pytest.approx(g["quality"], abs=1e-5)


@pytest.mark.parametrize("dtype_name", ["F64", "F32"])
def test_tracker_matches_oracle_on_random_episodes(E, oracle, dtype_name):
    """64 environments, random constant references and initial states (Controller.reset's draws), random actions,
    both altitude and pitch trackers, running vs finished snapshot, auto-reset clearing."""
    O = oracle
    dtype = getattr(E, dtype_name)
    n, K = 64, 5
    kw = dict(sample_time=0.05, tk=6.0)
    eng = E.BatchEngine(n_envs=n, dtype=dtype, seed=17, auto_reset=True, track_transfer=True, record_capacity=700, **kw)
    ob = O.OracleBatch(O.make_cfg(seed=17, **kw), n)
    views = [ob.env(i) for i in range(n)]
    for v in views:
        v.enable_storage(700)
    eng.reset(); ob.reset()
    rng = np.random.default_rng(2)
    steps = 120  # tk = 6 s at K = 5
    tol = dict(rel=1e-8, abs=1e-9) if dtype == E.F64 else dict(rel=2e-4, abs=2e-4)
    for k in range(steps):
        a = rng.uniform(-1, 1, n).astype(eng.np_dtype)
        if k == steps - 1:  # last step of the episode: compare the RUNNING tracker and the recorder first
            pass
        obs, rew, done = eng.step_host(a)
        o_o, r_o, d_o, _ = ob.step(a.astype(np.float64), auto_reset=False)
        assert np.array_equal(done.astype(bool), d_o)
        if k == 60:
            run = eng.transfer_metrics("SS", finished=False)
            for i in (0, 7, 33):
                info = views[i].stepinfo_SS()
                exp = [np.nan if info[x] is None else info[x] for x in NAMES]
                _cmp(run[i, :4], exp, tol, dtype == E.F64)
            rec = eng.recorder_read(5)
            st = views[5].storage
            assert len(rec["t"]) == len(st["t"]) == 61 * K
            for name in O.REC_FIELDS:
                assert np.allclose(rec[name], st[name], rtol=tol["rel"], atol=10 * tol["abs"]), name
    assert done.all()
    for which in ("SS", "CS"):
        fin = eng.transfer_metrics(which, finished=True)
        for i in range(n):
            info = views[i].stepinfo_SS() if which == "SS" else views[i].stepinfo_CS()
            exp = [np.nan if info[x] is None else info[x] for x in NAMES]
            _cmp(fin[i, :4], exp, tol, dtype == E.F64)
    # the auto-reset cleared the running tracker: a fresh episode starts from sample 0
    a = rng.uniform(-1, 1, n).astype(eng.np_dtype)
    eng.step_host(a)
    run = eng.transfer_metrics("SS", finished=False)
    assert np.isnan(run[:, 1]).all()                 # nothing can have risen within 5 samples of a new episode
    assert (np.abs(run[:, 2] - 0.04) < 1e-12).all()  # last sample outside the band = sample 4, 0.04 s after sample 0
    eng.close()


def _cmp(got, exp, tol, exact_times):
    for j, name in enumerate(NAMES):
        g, e = got[j], exp[j]
        if e != e:
            assert g != g, (name, g, e)
        elif name in ("rise_time", "settling_time") and not exact_times:
            assert abs(g - e) <= 0.0100001, (name, g, e)
        else:
            assert g == pytest.approx(e, **tol), (name, g, e)


def test_controller_facade_storage_and_stepinfo(E):
    """Controller(use_storage=True) / stepinfo_SS / quality as neural/callbacks.py:61-100 uses them."""
    from b747_rl_ctrl_b200.core.controller import CtrlMode, CtrlType
    from b747_rl_ctrl_b200.env.ctrl_env import ControllerEnv, ObservationType, RewardType
    env = ControllerEnv(ObservationType.PID_LIKE, RewardType.CLASSIC, True, True, CtrlType.MANUAL, CtrlMode.ADD_PROC_CONTROL,
                        tk=20, sample_time=0.05, action_max=1.0)
    ctrl = env.ctrl
    with pytest.raises(ValueError):
        ctrl.stepinfo_SS()
    ctrl.use_storage = True          # flipped after construction, like the callback does
    ctrl.vartheta_func = lambda _: 5 * DEG
    obs = env.reset(np.array([0, 11000, 250, 0, 0, 0.]))
    done = False
    while not done:
        obs, _, done, _ = env.step(np.zeros(1, np.float32))
    info = ctrl.stepinfo_SS()
    g = GOLD["5"]
    for k in NAMES:
        assert info[k] == pytest.approx(g["stepinfo"][k], rel=1e-8, abs=1e-10), k
    assert ctrl.quality() == pytest.approx(g["quality"], rel=1e-9)
    st = ctrl.storage.storage
    assert set(st) == {"t", "U_com", "U_PID", "deltaz", "hzh", "vartheta_ref", "U_RL", "x", "y", "Vx", "Vy", "vartheta", "wz"}
    assert len(st["t"]) == 2000 and st["t"][0] == 0.01 and st["vartheta_ref"][-1] == pytest.approx(5.0)
    env.reset()                      # the finished episode moves to storage_backup (core/controller.py:195-199)
    assert len(ctrl.storage_backup.storage["t"]) == 2000 and "t" not in ctrl.storage.storage
    info_b = ctrl.stepinfo_SS(use_backup=True)
    for k in NAMES:
        assert info_b[k] == pytest.approx(g["stepinfo"][k], rel=1e-8, abs=1e-10), k
    env.close()


def test_agent_test_tables_pid_beside_policy(E, oracle):
    """ControllerAgent.test (neural/agent.py:268-409): for every reference value a table with the PID baseline (СС PID in
    the loop, sample_time = dt) beside the policy, and the mean table over the references.  The PID rows are checked
    against the oracle running the same loop (ctrl_type AUTO, K = 1) with the reference's recorder + calc_stepinfo; a
    zero policy in ADD_PROC mode must reproduce the published transfer numbers."""
    from b747_rl_ctrl_b200.control_test import run_agent_test
    refs = [d * DEG for d in DEGS]
    out = run_agent_test(refs, {"zero_addproc": lambda obs: np.zeros(len(obs))}, ctrl_mode=E.MODE_ADD_PROC, action_max=1.0,
                         pid_coefs=[[-5.9151, -1.2404, -6.6927, 58.0826], [-4.0, -1.0, -5.0, 58.0826]])
    assert [r["Устройство"] for r in out["tables"][refs[0]]] == ["СС ПИД [1]", "СС ПИД [2]", "zero_addproc"]
    cfg = oracle.make_cfg(ctrl_type=oracle.CTRL_AUTO, reset_ref_mode=oracle.RESET_NONE, sample_time=None, norm_act=True)
    ob = oracle.OracleBatch(cfg, len(refs))
    ob.reset_to([oracle.episode([0, 11000, 250, 0, 0, 0], vref=r) for r in refs])
    views = [ob.env(i) for i in range(len(refs))]
    for v in views:
        v.enable_storage(2001)
    for k in range(2000):
        _, _, d, _ = ob.step(np.zeros(len(refs)), auto_reset=False)
    assert d.all()
    for j, ref in enumerate(refs):
        info = views[j].stepinfo_SS()
        row = out["tables"][ref][0]                     # default coefficients
        assert row["σ, [%]"] == pytest.approx(info["overshoot"], rel=1e-8)
        assert row["tпп, [с]"] == pytest.approx(info["settling_time"], abs=1e-9)
        assert row["tв, [с]"] == pytest.approx(info["rise_time"], abs=1e-9)
        assert row["Δ, [град]"] == pytest.approx(info["static_error"], rel=1e-7, abs=1e-10)
        assert 0 < row["Q, [-]"] < 1
        nn = out["tables"][ref][2]
        g = GOLD[str(DEGS[j])]
        assert nn["σ, [%]"] == pytest.approx(g["stepinfo"]["overshoot"], rel=1e-8) and nn["Q, [-]"] == pytest.approx(g["quality"], rel=1e-9)
        assert out["tables"][ref][1]["σ, [%]"] != row["σ, [%]"]   # the second coefficient set really is another controller
    mean = {m["Устройство"]: m for m in out["mean"]}
    assert mean["zero_addproc"]["σ, [%]"] == pytest.approx(9.263, abs=1e-3)
    assert mean["zero_addproc"]["Q, [-]"] == pytest.approx(0.7527, abs=5e-5)
    assert mean["СС ПИД [1]"]["σ, [%]"] == pytest.approx(np.mean([abs(out["tables"][r][0]["σ, [%]"]) for r in refs]))
    # altitude references through the СУ PID (stepinfo_CS): FULL_AUTO baseline, no policy
    alt = run_agent_test([10500.0, 11500.0], ctrl_type=E.CTRL_SEMI_MANUAL, no_neural=True, tk=60.0)
    assert alt["unit"] == "Δ, [м]" and [r["Устройство"] for r in alt["tables"][10500.0]] == ["CУ ПИД"]
    for ref in (10500.0, 11500.0):
        r = alt["tables"][ref][0]
        assert r["Δ, [м]"] is not None and r["Δ, [м]"] < 60.0 and r["tв, [с]"] is not None
