"""`Model` pinned to the reference's OWN class: tests/golden/model_refpy.json is what /root/reference/core/model.py's
`Model` returns -- every public property -- along three call sequences (its `__main__` smoke, a manual tour with aero
errors and a mid-flight re-initialisation, both PIDs) when run over the DLL (tests/golden/make_model_refpy.py).  The GPU
test replays the same calls on b747_rl_ctrl_b200.core.model.Model; the CPU test replays them on the oracle's restatement
with the Python-side semantics of core/model.py:238-244 spelled out."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
R = json.load(open(os.path.join(HERE, "golden", "model_refpy.json")))
TOL = {"dvartheta_dt_dt": (1e-6, 1e-9), "dvartheta_dt": (1e-8, 1e-11)}


def _check(tag, got, want):
    for k, w in want.items():
        if k == "state_dict":
            for kk, ww in w.items():
                assert np.isclose(got(k)[kk], ww, rtol=1e-9, atol=1e-12), (tag, k, kk)
            continue
        rtol, atol = TOL.get(k, (1e-9, 1e-12))
        assert np.allclose(np.asarray(got(k), dtype=float), w, rtol=rtol, atol=atol), (tag, k, got(k), w)


def _replay(make):
    """The three call sequences of make_model_refpy.py on `make(**ctor kwargs)`."""
    c = R["main_smoke"]
    m = make(use_PID_CS=False, initial_state=np.array([100, 1000, 300, 0, 0, 0]))
    m.hzh = 2000
    m.P = 300000
    m.vartheta_zh = -0.1
    _check("main/ctor", lambda k: getattr(m, k), c["after_ctor_and_writes"])
    for n in range(1, 601):
        m.step()
        if str(n) in c["snaps"]:
            _check(f"main/{n}", lambda k: getattr(m, k), c["snaps"][str(n)])
    c = R["manual_tour"]
    m = make(use_PID_SS=False, use_PID_CS=False, initial_state=np.array([0, 5000, 180, 2, 0.01, 0.0005]))
    m.aero_err = np.array([-0.1, 0.1, -0.1, -0.1, 0.1])
    _check("tour/ctor", lambda k: getattr(m, k), c["after_ctor_and_writes"])
    rng = np.random.default_rng(3)
    for n in range(1, 301):
        if (n - 1) % 5 == 0:
            m.deltaz = float(rng.uniform(-0.25, 0.25))
        m.step()
        if str(n) in c["snaps"]:
            _check(f"tour/{n}", lambda k: getattr(m, k), c["snaps"][str(n)])
    m.set_initial(np.array([10, 7000, 220, -3, 0.02, 0.0]))
    m.initialize()
    _check("tour/reinit", lambda k: getattr(m, k), c["after_reinitialize"])
    m.deltaz = 0.05
    for n in range(40):
        m.step()
    _check("tour/reinit+40", lambda k: getattr(m, k), c["snaps"]["reinit+40"])
    c = R["both_pids"]
    m = make(use_PID_SS=True, use_PID_CS=True, initial_state=np.array([0, 11000, 250, 0, 0, 0]))
    m.hzh = 10500
    for n in range(1, 1001):
        m.step()
        if str(n) in c["snaps"]:
            _check(f"pids/{n}", lambda k: getattr(m, k), c["snaps"][str(n)])


@pytest.mark.gpu
def test_model_facade_replays_reference_model_class():
    from b747_rl_ctrl_b200.core.model import Model
    _replay(Model)


class _RestatedModel:
    """core/model.py's Model semantics over the oracle's restatement: the name traps (vartheta_ref -> signal vartheta_zh,
    vartheta_zh -> parameter vartheta, deltaz_ref / deltaz_com / deltaz_real), nan_to_num on `state`, and initialize()
    zeroing deltaz / vartheta_zh and setting step_num = -1 (core/model.py:129-164, 238-250)."""
    _sig = {"time": "sim_time", "vartheta_ref": "vartheta_zh", "deltaz_ref": "U_com_PID", "deltaz_com": "U_com",
            "deltaz_real": "deltaz_RP", "Kalpha": "K_alpha"}
    _par = {"hzh": "h_zh", "vartheta_zh": "vartheta"}
    labels = ['x', 'y', 'Vx', 'Vy', 'vartheta', 'wz']

    def __init__(self, oracle, use_PID_SS=True, use_PID_CS=True, initial_state=None, use_RP=True):
        object.__setattr__(self, "m", oracle.CModel())
        object.__setattr__(self, "step_num", -1)
        if initial_state is not None:
            self.m.set("state0", initial_state)
        self.m.set("use_RP", float(use_RP)); self.m.set("use_PID_CS", float(use_PID_CS)); self.m.set("use_PID_SS", float(use_PID_SS))
        self.initialize()

    def initialize(self):
        self.m.initialize()
        object.__setattr__(self, "step_num", -1)
        self.m.set("deltaz", 0.0); self.m.set("vartheta", 0.0)

    def step(self):
        self.m.step()
        object.__setattr__(self, "step_num", self.step_num + 1)

    def set_initial(self, s):
        self.m.set("state0", s)

    def __getattr__(self, k):
        if k == "state":
            return np.nan_to_num(np.array(self.m.get("state")))
        if k == "state_dict":
            return dict(zip(self.labels, self.state))
        return self.m.get(self._sig.get(k, self._par.get(k, k)))

    def __setattr__(self, k, v):
        self.m.set(self._par.get(k, k), v)


def test_restatement_replays_reference_model_class(oracle):
    _replay(lambda **kw: _RestatedModel(oracle, **kw))
