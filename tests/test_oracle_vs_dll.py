"""Pins the oracle: the plain-C restatement (oracle/b747_model_ref.c, b747_env_ref.c) against the
reference DLL's own machine code (oracle/_ref/libb747_ref.so) and against the known answers the
survey obtained from that DLL.  Skipped where oracle/_ref is not built."""
import math
import random

import numpy as np
import pytest

DEG = math.pi / 180
ALL_SIG = None


def _close(a, b, rel, abs_):
    return abs(a - b) <= abs_ + rel * abs(b)


def _run_pair(D, O, params, steps, elevator, check_every=1):
    d = D.DllModel(); c = O.CModel()
    for m in (d, c):
        for k, v in params.items():
            m.set(k, v)
        m.initialize()
    rng = random.Random(0)
    worst = {}
    for k in range(steps):
        if elevator and k % 5 == 0:
            a = rng.uniform(-0.2967, 0.2967)
            d.set("deltaz", a); c.set("deltaz", a)
        d.step(); c.step()
        if k % check_every:
            continue
        for name, n in D.SIGNALS.items():
            x, y = d.get(name), c.get(name)
            xs, ys = (x, y) if n > 1 else ([x], [y])
            for u, v in zip(xs, ys):
                scale = {"dvartheta_dt_dt": 1e-2, "dvartheta_dt": 1e-3}.get(name, 0.0)
                err = abs(u - v) / max(abs(u), scale, 1e-12)
                worst[name] = max(worst.get(name, 0.0), err)
    return d, c, worst


def test_k1_defaults(dllref):
    m = dllref.DllModel()
    m.set("use_PID_SS", 1.0)
    m.initialize()
    assert m.get("state") == [0.0] * 6 and m.get("sim_time") == 0.0
    m.step()
    assert m.get("sim_time") == 0.01
    assert m.get("state") == [2.5916868693309953, 10999.999628217964, 259.17067385352044, -0.07435406285977463,
                              4.379386114678432e-06, 0.000875851931227687]
    assert m.get("dvartheta") == 0.0872620832136018
    m.step(4)
    assert m.get("state") == [12.958832615739883, 10999.990761242996, 259.18663005749744, -0.3679509230649856,
                              0.00010888949718821038, 0.00433851607891542]


K_FINAL = {
    "K2": [24310.62757377837, 10568.191363641768, 264.9689242497395, -27.41210230550044, -0.041197934774639434, 0.006091114314126824],
    "K3": [4970.894113753458, 11109.746835924929, 247.26247896615428, 6.192514013202871, 0.08730266472912393, -0.0001027859971435344],
    "K4": [15559.412815750316, 10499.808644172148, 262.2367041062174, 0.02189940183906506, 0.05278167044515464, -2.0898738561887323e-05],
    "K5": [6027.154604646006, 3479.6121888905004, 128.05734724813468, 30.63844719055044, 0.3799725555680438, 0.025450960379116174],
}
K_CASES = {
    "K2": (dict(state0=[0, 11000, 250, 0, 0, 0], vartheta=5 * DEG), 10000, True),
    "K3": (dict(state0=[0, 11000, 250, 0, 0, 0], vartheta=5 * DEG, use_PID_SS=1.0), 2000, False),
    "K4": (dict(state0=[0, 11000, 250, 0, 0, 0], vartheta=5 * DEG, use_PID_SS=1.0, use_PID_CS=1.0, h_zh=10500.0), 6000, False),
    "K5": (dict(aero_err=[-0.1, 0.1, -0.1, -0.1, 0.1], state0=[0, 3000, 150, 5, 0.02, 0.001], vartheta=5 * DEG), 4000, True),
}


@pytest.mark.parametrize("name", sorted(K_CASES))
def test_model_restatement_tracks_dll(dllref, oracle, name):
    params, steps, elev = K_CASES[name]
    d, c, worst = _run_pair(dllref, oracle, params, steps, elev, check_every=3)
    # the DLL itself reproduces the survey's known answers bit for bit
    assert d.get("state") == K_FINAL[name]
    # restatement: libm differences only (glibc vs the DLL's static UCRT)
    for u, v in zip(d.get("state"), c.get("state")):
        assert _close(v, u, 1e-10, 1e-12)
    for sig, err in worst.items():
        tol = 2e-6 if sig == "dvartheta_dt_dt" else 1e-8
        assert err < tol, (sig, err)


def test_env_layer_k7_on_dll(dllref, oracle):
    """the env layer over the DLL reproduces the line-by-line Python emulation."""
    O = oracle
    cfg = O.make_cfg(reset_ref_mode=O.RESET_NONE)
    ep = O.episode([0, 11000, 250, 0, 0, 0], vref=5 * DEG)
    env = O.RefEnv(cfg)
    assert list(env.reset_to(ep)) == [0.0, 0.0, 0.0]
    rng = random.Random(0)
    acts = [rng.uniform(-1, 1) for _ in range(400)]
    assert acts[:3] == [0.6888437030500962, 0.515908805880605, -0.15885683833831]
    ret = 0.0
    exp = {1: ([2.314023880182406e-05, 0.027749475655376003, -0.0010158120326844735], 0.6157932351868455),
           2: ([4.623699101073092e-05, 0.02767976252279455, -0.0014885835516942074], 0.6771185939077887),
           3: ([6.927546055580163e-05, 0.027616845651624037, -0.0010452751904213616], 0.6650545720474278),
           400: ([0.0049432886184918815, 0.007467549251812716, 0.01122935958087855], 0.5912528354158655)}
    for k, a in enumerate(acts, 1):
        obs, r, done = env.step(a)
        ret += r
        if k in exp:
            assert list(obs) == exp[k][0] and r == exp[k][1]
        assert done == (k == 400)
    assert ret == 218.4897443782656
    env = O.RefEnv(cfg); env.reset_to(ep)
    obs, rew, done = env.rollout(np.zeros(400), auto_reset=False)
    assert math.fsum(rew) == pytest.approx(299.9338252856275, abs=1e-11) and rew[-1] == 0.7388463138290949
    assert list(obs[-1]) == [0.002650412572448381, 0.0060902751961632, -0.0002978783942516373]
    cfg2 = O.make_cfg(reset_ref_mode=O.RESET_NONE, ctrl_mode=O.MODE_ADD_PROC, action_max=1.0)
    env = O.RefEnv(cfg2); env.reset_to(ep)
    _, rew, done = env.rollout(np.zeros(400), auto_reset=False)
    assert sum(rew) == pytest.approx(365.092282070656, abs=1e-10) and done[-1] and not done[:-1].any()


def test_transfer_metrics_match_published_xlsx(dllref, oracle):
    """K6: the four PID transfer episodes (ADD_PROC, a=0) -> quality exp(-6 ITSE/(tk vref^2)); the mean
    0.7527 equals the reference's published tensorboard value 0.753."""
    O = oracle
    itse_exp = {5: 4.647529e-03, -5: 9.235935e-03, 10: 2.336604e-02, -10: 3.789869e-02}
    qs = []
    for deg, itse_ref in itse_exp.items():
        cfg = O.make_cfg(reset_ref_mode=O.RESET_NONE, ctrl_mode=O.MODE_ADD_PROC, action_max=1.0, rew_type=O.REW_QUALITY)
        env = O.RefEnv(cfg)
        env.reset_to(O.episode([0, 11000, 250, 0, 0, 0], vref=deg * DEG))
        _, rew, done = env.rollout(np.zeros(400), auto_reset=False)
        q = rew[-1]
        itse = -math.log(q) * 20 * (deg * DEG) ** 2 / 6
        assert itse == pytest.approx(itse_ref, rel=2e-6)
        qs.append(q)
    assert np.mean(qs) == pytest.approx(0.7527, abs=5e-5)


def test_env_layer_restatement_tracks_dll(dllref, oracle):
    """Same env-layer code over both back-ends, every configuration family, with auto-reset."""
    O = oracle
    from tests.golden.make_golden import ENV_CASES
    for name, (kw, n, steps) in ENV_CASES.items():
        cfg = O.make_cfg(seed=11, **kw)
        rng = np.random.default_rng(5)
        amax = 1.0 if cfg.norm_act else cfg.action_max
        acts = rng.uniform(-amax, amax, size=(2, steps))
        ob = O.OracleBatch(cfg, 2)
        ob.reset()
        o_c = np.zeros((2, steps, O.OBS_DIM[cfg.obs_type])); r_c = np.zeros((2, steps)); d_c = np.zeros((2, steps), bool)
        for k in range(steps):
            o, r, d, term = ob.step(acts[:, k])
            o_c[:, k], r_c[:, k], d_c[:, k] = term, r, d  # rollout() records the pre-reset observation
        env = O.RefEnv(cfg, env_id=1); env.reset()
        o_d, r_d, d_d = env.rollout(acts[1], auto_reset=True)
        assert (d_d == d_c[1]).all(), name
        assert np.allclose(o_c[1], o_d, rtol=1e-8, atol=2e-9), name
        assert np.allclose(r_c[1], r_d, rtol=0, atol=2e-7 if cfg.rew_type == O.REW_CLASSIC else 1e-9), name


def test_dll_host_never_maps_writable_and_executable(dllref, oracle):
    """pe_host.c lays the image out read-write and then re-protects per section: no page of the hosted DLL (or of the
    ELF face the reference's Python loads) is writable and executable at once."""
    env = oracle.RefEnv(oracle.make_cfg(), env_id=0)
    env.reset()
    env.step(0.1)
    rwx = [ln for ln in open("/proc/self/maps") if ln.split()[1].startswith("rwx")]
    assert not rwx, rwx


def _sandbox_child(q):
    import numpy as np
    from oracle import dllref as D, oracle as O
    env = O.RefEnv(O.make_cfg(), env_id=0)
    env.reset()
    res = {"installed": D.sandbox()}
    for what, fn in (("write", lambda: open("/tmp/b747_sandbox_probe", "w")),
                     ("socket", lambda: __import__("socket").socket()),
                     ("exec", lambda: __import__("os").execv("/bin/true", ["true"]))):
        try:
            fn()
            res[what] = "allowed"
        except OSError as e:
            res[what] = e.errno
    _, r, _ = env.rollout(np.zeros(400))
    res["ret"] = float(r.sum())
    q.put(res)


def test_worker_sandbox_denies_io_but_runs_the_dll(dllref):
    """bench.py's reference-arm workers call dllref.sandbox() before they run the DLL in bulk: file writes, sockets and
    exec are refused (EPERM) from then on, the numeric path is unaffected (K7: a = 0 return 299.93)."""
    import errno
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_sandbox_child, args=(q,))
    p.start()
    res = q.get(timeout=120)
    p.join(30)
    if not res["installed"]:
        pytest.skip("seccomp filters cannot be installed in this container")
    assert res["write"] == errno.EPERM and res["socket"] == errno.EPERM and res["exec"] == errno.EPERM
    assert res["ret"] > 100
