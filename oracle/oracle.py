"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE -- not product code).

Two interchangeable back-ends behind one env-layer implementation (b747_env_ref.c):
  * `OracleBatch`  -- the plain-C float64 restatement (oracle/_build/liboracle.so), many envs;
  * `RefEnv`       -- the reference DLL's own machine code (oracle/_ref/libb747_ref.so), one
                      private DLL instance per env.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
import ctypes
import math
import os
import subprocess

import numpy as np

from . import dllref

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "_build", "liboracle.so")

# enum values == the reference's Enum values (core/controller.py:14-36, env/ctrl_env.py:16-30)
CTRL_FULL_AUTO, CTRL_AUTO, CTRL_SEMI_MANUAL, CTRL_MANUAL = 0, 1, 2, 3
MODE_DIRECT, MODE_ADD_PROC, MODE_ANG_VEL, MODE_ADD_DIRECT = 0, 1, 2, 3
MODE_NONE = -1
RESET_NONE, RESET_CONST, RESET_OSCILLATING, RESET_HYBRID = -1, 0, 1, 2
DIST_NONE, DIST_AERO = -1, 0
OBS_PID_LIKE, OBS_SPEED_MODE, OBS_PID_AERO, OBS_PID_SPEED_AERO, OBS_MODEL_STATE = 0, 1, 2, 3, 4
REW_CLASSIC, REW_PID_LIKE, REW_QUALITY, REW_MINIMAL, REW_TF_REFERENCE = 0, 1, 2, 3, 4
OBS_DIM = {0: 3, 1: 5, 2: 8, 3: 10, 4: 7}


class EnvCfg(ctypes.Structure):
    _fields_ = [
        ("obs_type", ctypes.c_int32), ("rew_type", ctypes.c_int32), ("ctrl_type", ctypes.c_int32),
        ("ctrl_mode", ctypes.c_int32), ("reset_ref_mode", ctypes.c_int32), ("disturbance_mode", ctypes.c_int32),
        ("norm_obs", ctypes.c_int32), ("norm_act", ctypes.c_int32), ("use_limiter", ctypes.c_int32),
        ("substeps", ctypes.c_int32), ("done_tick", ctypes.c_int64),
        ("tk", ctypes.c_double), ("action_max", ctypes.c_double), ("vartheta_max", ctypes.c_double),
        ("sample_time", ctypes.c_double), ("rew", ctypes.c_double * 8),
        ("fixed_aero_err", ctypes.c_double * 5), ("has_fixed_aero_err", ctypes.c_int32),
        ("seed", ctypes.c_uint64),
    ]


class Episode(ctypes.Structure):
    _fields_ = [
        ("state0", ctypes.c_double * 6), ("use_ctrl", ctypes.c_int32), ("vref_const", ctypes.c_double),
        ("osc_A", ctypes.c_double * 3), ("osc_f", ctypes.c_double * 3), ("oscillating", ctypes.c_int32),
        ("h_ref", ctypes.c_double), ("aero_err", ctypes.c_double * 5),
    ]


def reward_constants(rew_type, reward_config=None):
    """The numbers ControllerEnv._get_reward_def closes over (env/ctrl_env.py:109-192)."""
    rc = dict(reward_config or {})
    out = [0.0] * 8
    if rew_type == REW_CLASSIC:
        k1, k2, k3 = rc.get("k1", 2), rc.get("k2", 2), rc.get("k3", 1)
        kf, kITSE = rc.get("kf", 0.1), rc.get("kITSE", 0.3)
        kt = -math.log(0.8) / 10        # calc_exp_k(0.8, 10), tools/general.py:32-33
        ko = -math.log(0.75) / 0.15     # calc_exp_k(0.75, 0.15)
        k0 = rc.get("k0", 2)
        s = k1 + k2 + k3
        out = [k1 / s, k2 / s, k3 / s, k0, kITSE, kf, kt, ko]
    elif rew_type == REW_PID_LIKE:
        out[0] = rc.get("k", 10)
    elif rew_type == REW_TF_REFERENCE:
        out[0], out[1], out[2] = rc.get("overshoot_ref", 2), rc.get("tp_ref", 5), rc.get("k", 0.1)
    return [float(x) for x in out]


def substeps_of(sample_time, dt=0.01):
    """K of Controller.step's loop: round(sample_time/dt), Python (banker's) rounding
    (core/controller.py:110,261)."""
    st = sample_time if sample_time else dt
    return max(1, round(st / dt))


def done_tick_of(tk):
    """Smallest tick with fl(tick*0.01) >= tk (Controller.is_done, core/controller.py:316-319)."""
    if math.isinf(tk) or tk != tk:
        return 2 ** 62
    if tk <= 0:
        return 0
    n = max(0, int(math.floor(tk / 0.01)) - 2)
    while not (n * 0.01 >= tk):
        n += 1
    return n


def make_cfg(obs_type=OBS_PID_LIKE, rew_type=REW_CLASSIC, ctrl_type=CTRL_MANUAL, ctrl_mode=MODE_DIRECT,
             reset_ref_mode=RESET_CONST, disturbance_mode=DIST_NONE, norm_obs=True, norm_act=True,
             use_limiter=False, tk=20.0, sample_time=0.05, action_max=17 * math.pi / 180,
             vartheta_max=10 * math.pi / 180, reward_config=None, aero_err=None, seed=1):
    """Defaults = the canonical configuration main.py:88-121 trains."""
    c = EnvCfg()
    c.obs_type, c.rew_type, c.ctrl_type, c.ctrl_mode = obs_type, rew_type, ctrl_type, ctrl_mode
    c.reset_ref_mode, c.disturbance_mode = reset_ref_mode, disturbance_mode
    c.norm_obs, c.norm_act, c.use_limiter = int(norm_obs), int(norm_act), int(use_limiter)
    c.substeps = substeps_of(sample_time)
    c.done_tick = done_tick_of(tk)
    c.tk, c.action_max, c.vartheta_max = float(tk), float(action_max), float(vartheta_max)
    c.sample_time = float(sample_time if sample_time else 0.01)
    for i, x in enumerate(reward_constants(rew_type, reward_config)):
        c.rew[i] = x
    if aero_err is not None:
        c.has_fixed_aero_err = 1
        for i in range(5):
            c.fixed_aero_err[i] = float(aero_err[i])
    c.seed = seed
    return c


def episode(state0, vref=0.0, h_ref=11000.0, use_ctrl=False, osc=None, aero_err=None):
    e = Episode()
    for i in range(6):
        e.state0[i] = float(state0[i])
    e.vref_const, e.h_ref, e.use_ctrl = float(vref), float(h_ref), int(use_ctrl)
    if osc is not None:
        e.oscillating = 1
        for i in range(3):
            e.osc_A[i], e.osc_f[i] = float(osc[0][i]), float(osc[1][i])
    if aero_err is not None:
        for i in range(5):
            e.aero_err[i] = float(aero_err[i])
    return e


def build(force=False):
    """Compile the oracle libraries (checker only).  `ref` needs /root/reference."""
    targets = ["oracle"]
    if os.path.exists("/root/reference/core/model_simple_win64.dll"):
        targets.append("ref")
    cmd = ["make", "-C", _HERE] + (["-B"] if force else []) + targets
    subprocess.run(cmd, check=True, capture_output=True)


_olib = None


REC_FIELDS = ("t", "U_com", "U_PID", "deltaz", "hzh", "vartheta_ref", "U_RL", "x", "y", "Vx", "Vy", "vartheta", "wz")


def stepinfo(ys, y_base, ts, error_band=0.05):
    """Step-response figures of a recorded signal, as tools/general.py:46-61 defines them: overshoot in % of the
    reference (peak on the reference's side), rise time = first sample (the last one excluded) whose normalised
    response reaches 1 - band, settling time = LAST sample outside the +-band, both counted from the first
    sample's time stamp; static error = |last - reference|.  None where the reference returns None."""
    ys = [float(v) for v in ys]
    n = len(ys)
    norm = [(y - ys[0]) / (y_base - ys[0]) for y in ys]
    peak = max(ys) if y_base > 0 else min(ys)
    info = {"overshoot": (peak - y_base) / y_base * 100 if y_base != 0 else None, "rise_time": None,
            "settling_time": None, "static_error": abs(ys[-1] - y_base)}
    for i in range(n - 1):
        if norm[i] >= 1 - error_band:
            info["rise_time"] = ts[i] - ts[0]
            break
    for i in range(n - 1, -1, -1):
        if norm[i] <= 1 - error_band or norm[i] >= 1 + error_band:
            info["settling_time"] = ts[i] - ts[0]
            break
    return info


class _Recorder:
    """Mixin: Controller(use_storage=True) for one oracle env (pointer in self._e)."""

    def enable_storage(self, capacity):
        L = self._L
        L.b747o_env_set_recorder.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]
        L.b747o_env_recorded.restype = ctypes.c_int32
        L.b747o_env_recorded.argtypes = [ctypes.c_void_p]
        self._rec = np.zeros((int(capacity), len(REC_FIELDS)))
        L.b747o_env_set_recorder(self._e, _dptr(self._rec), int(capacity))

    @property
    def storage(self):
        n = self._L.b747o_env_recorded(self._e)
        return {nm: self._rec[:n, k].copy() for k, nm in enumerate(REC_FIELDS)}

    def stepinfo_SS(self):
        st = self.storage
        return stepinfo(st["vartheta"], st["vartheta_ref"][-1], st["t"])

    def stepinfo_CS(self):
        st = self.storage
        return stepinfo(st["y"], st["hzh"][-1], st["t"])


def _proto_env_api(L):
    L.b747o_env_step.restype = ctypes.c_int
    L.b747o_env_step.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.POINTER(ctypes.c_double),
                                 ctypes.POINTER(ctypes.c_double)]
    L.b747o_env_reset.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)]
    L.b747o_env_reset_to.argtypes = [ctypes.c_void_p, ctypes.POINTER(Episode), ctypes.POINTER(ctypes.c_double)]
    L.b747o_env_draw_episode.argtypes = [ctypes.POINTER(EnvCfg), ctypes.c_uint64, ctypes.c_uint64,
                                         ctypes.POINTER(Episode)]
    L.b747o_uniform53.restype = ctypes.c_double
    L.b747o_uniform53.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32]
    L.b747o_done_tick.restype = ctypes.c_int64
    L.b747o_done_tick.argtypes = [ctypes.c_double]


def olib():
    global _olib
    if _olib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        L = ctypes.CDLL(ORACLE_SO)
        _proto_env_api(L)
        L.b747o_model_new.restype = ctypes.c_void_p
        L.b747o_model_free.argtypes = [ctypes.c_void_p]
        L.b747o_model_ptr.restype = ctypes.c_void_p
        L.b747o_model_ptr.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.b747o_model_initialize.argtypes = [ctypes.c_void_p]
        L.b747o_model_step.argtypes = [ctypes.c_void_p]
        L.b747o_model_step_n.argtypes = [ctypes.c_void_p, ctypes.c_long]
        L.b747o_model_tick.restype = ctypes.c_uint32
        L.b747o_model_tick.argtypes = [ctypes.c_void_p]
        L.b747o_batch_create.restype = ctypes.c_void_p
        L.b747o_batch_create.argtypes = [ctypes.POINTER(EnvCfg), ctypes.c_int64, ctypes.c_uint64]
        L.b747o_batch_destroy.argtypes = [ctypes.c_void_p]
        L.b747o_batch_reset.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.b747o_batch_reset_to.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.b747o_batch_step.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 5 + [ctypes.c_int]
        L.b747o_batch_env.restype = ctypes.c_void_p
        L.b747o_batch_env.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        L.b747o_batch_model.restype = ctypes.c_void_p
        L.b747o_batch_model.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        L.b747o_philox4x32.argtypes = [ctypes.POINTER(ctypes.c_uint32)] * 3
        L.b747o_batch_gather.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p]
        L.b747o_batch_ticks.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        _olib = L
    return _olib


_ALL = {**dllref.SIGNALS, **dllref.PARAMS, "X": 18, "dX": 18, "t": 1, "df_x": 1, "df_y": 1,
        "rl_prev": 1, "rl_t": 1, "td": 1}


class CModel:
    """One instance of the C restatement with the DllModel get/set interface."""

    def __init__(self, handle=None):
        self._L = olib()
        self._own = handle is None
        self._h = self._L.b747o_model_new() if handle is None else handle

    def __del__(self):
        try:
            if self._own and self._h:
                self._L.b747o_model_free(self._h)
                self._h = None
        except Exception:
            pass

    def _arr(self, name):
        p = self._L.b747o_model_ptr(self._h, name.encode())
        if not p:
            raise KeyError(name)
        return (ctypes.c_double * _ALL[name]).from_address(p)

    def get(self, name):
        a = self._arr(name)
        return a[0] if len(a) == 1 else list(a)

    def set(self, name, value):
        a = self._arr(name)
        if len(a) == 1:
            a[0] = float(value)
        else:
            for i, x in enumerate(value):
                a[i] = float(x)

    def initialize(self):
        self._L.b747o_model_initialize(self._h)

    def step(self, n=1):
        self._L.b747o_model_step_n(self._h, n)

    @property
    def tick(self):
        return self._L.b747o_model_tick(self._h)


def _dptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class OracleBatch:
    """N independent ControllerEnv equivalents over the C restatement."""

    def __init__(self, cfg, n_envs, env_id_offset=0):
        self._L = olib()
        self.cfg, self.n = cfg, int(n_envs)
        self.obs_dim = OBS_DIM[cfg.obs_type]
        self._h = self._L.b747o_batch_create(ctypes.byref(cfg), self.n, env_id_offset)

    def __del__(self):
        try:
            if self._h:
                self._L.b747o_batch_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def reset(self):
        obs = np.zeros((self.n, self.obs_dim))
        self._L.b747o_batch_reset(self._h, _dptr(obs))
        return obs

    def reset_to(self, episodes):
        arr = (Episode * self.n)(*episodes)
        obs = np.zeros((self.n, self.obs_dim))
        self._L.b747o_batch_reset_to(self._h, ctypes.cast(arr, ctypes.c_void_p), _dptr(obs))
        return obs

    def step(self, actions, auto_reset=True):
        a = np.ascontiguousarray(np.asarray(actions, dtype=np.float64).reshape(self.n))
        obs = np.zeros((self.n, self.obs_dim)); rew = np.zeros(self.n); done = np.zeros(self.n, dtype=np.uint8)
        term = np.zeros((self.n, self.obs_dim))
        self._L.b747o_batch_step(self._h, _dptr(a), _dptr(obs), _dptr(rew), _dptr(done), _dptr(term), int(auto_reset))
        return obs, rew, done.astype(bool), term

    def model(self, i):
        return CModel(self._L.b747o_batch_model(self._h, i))

    def gather(self, name, idx=0):
        """Field `name`[idx] of every env's model (the DLL-global view), as one float64 array."""
        out = np.empty(self.n)
        if self._L.b747o_batch_gather(self._h, name.encode(), int(idx), _dptr(out)):
            raise KeyError(name)
        return out

    def ticks(self):
        out = np.empty(self.n, np.int64)
        self._L.b747o_batch_ticks(self._h, _dptr(out))
        return out

    def env(self, i):
        """Recorder view of env i (enable_storage / storage / stepinfo_SS / stepinfo_CS)."""
        v = _Recorder()
        v._L, v._e = self._L, self._L.b747o_batch_env(self._h, i)
        return v


_rlib = None


def rlib():
    global _rlib
    if _rlib is None:
        L = dllref.lib()
        _proto_env_api(L)
        L.b747ref_env_create.restype = ctypes.c_void_p
        L.b747ref_env_create.argtypes = [ctypes.POINTER(EnvCfg), ctypes.c_uint64]
        L.b747ref_env_destroy.argtypes = [ctypes.c_void_p]
        L.b747ref_env_get.restype = ctypes.c_void_p
        L.b747ref_env_get.argtypes = [ctypes.c_void_p]
        L.b747ref_env_rollout.restype = ctypes.c_long
        L.b747ref_env_rollout.argtypes = [ctypes.c_void_p, ctypes.c_long] + [ctypes.c_void_p] * 4 + [ctypes.c_int]
        _rlib = L
    return _rlib


class RefEnv(_Recorder):
    """ControllerEnv equivalent driving a private instance of the reference DLL."""

    def __init__(self, cfg, env_id=0):
        self._L = rlib()
        self.cfg = cfg
        self.obs_dim = OBS_DIM[cfg.obs_type]
        self._h = self._L.b747ref_env_create(ctypes.byref(cfg), env_id)
        if not self._h:
            raise RuntimeError("b747ref_env_create failed")
        self._e = self._L.b747ref_env_get(self._h)

    def __del__(self):
        try:
            if self._h:
                self._L.b747ref_env_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def reset(self):
        obs = (ctypes.c_double * self.obs_dim)()
        self._L.b747o_env_reset(self._e, obs)
        return np.array(obs)

    def reset_to(self, ep):
        obs = (ctypes.c_double * self.obs_dim)()
        self._L.b747o_env_reset_to(self._e, ctypes.byref(ep), obs)
        return np.array(obs)

    def step(self, action):
        obs = (ctypes.c_double * self.obs_dim)()
        r = ctypes.c_double()
        d = self._L.b747o_env_step(self._e, float(action), obs, ctypes.byref(r))
        return np.array(obs), r.value, bool(d)

    def rollout(self, actions, auto_reset=True, record=True):
        a = np.ascontiguousarray(np.asarray(actions, dtype=np.float64))
        n = a.shape[0]
        if record:
            obs = np.zeros((n, self.obs_dim)); rew = np.zeros(n); done = np.zeros(n, dtype=np.uint8)
            self._L.b747ref_env_rollout(self._h, n, _dptr(a), _dptr(obs), _dptr(rew), _dptr(done), int(auto_reset))
            return obs, rew, done.astype(bool)
        self._L.b747ref_env_rollout(self._h, n, _dptr(a), None, None, None, int(auto_reset))
        return None


def draw_episode(cfg, env_id, episode_idx, lib_=None):
    L = lib_ or olib()
    ep = Episode()
    L.b747o_env_draw_episode(ctypes.byref(cfg), env_id, episode_idx, ctypes.byref(ep))
    return ep
