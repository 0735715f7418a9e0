"""Controller -- drop-in for core/controller.py (enums + Controller class) over the CUDA engine.

The reference's Controller.step runs the reference update, the action law and a K-substep loop of
ctypes calls (core/controller.py:231-264); here one call of the batched C ABI (n_envs = 1, float64)
does all of it in a single kernel launch.  The same engine also computes the observation/reward
that ControllerEnv returns, so Controller and ControllerEnv share one handle.
"""
import math
from enum import Enum
from math import exp, pi

import numpy as np

from .. import engine as E
from ..tools.general import Storage, calc_err, calc_stepinfo
from .model import ModelView


class CtrlType(Enum):  # core/controller.py:14-19
    FULL_AUTO = 0
    AUTO = 1
    SEMI_MANUAL = 2
    MANUAL = 3


class CtrlMode(Enum):  # core/controller.py:21-26
    DIRECT_CONTROL = 0
    ADD_PROC_CONTROL = 1
    ANG_VEL_CONTROL = 2
    ADD_DIRECT_CONTROL = 3


class ResetRefMode(Enum):  # core/controller.py:28-32
    CONST = 0
    OSCILLATING = 1
    HYBRID = 2


class DisturbanceMode(Enum):  # core/controller.py:34-36
    AERO_DISTURBANCE = 0


def _const_value(func, what, require=False):
    """The constant a reference function returns, or None if it varies with time.  The batched kernels hold the
    reference constant over an episode (or use the built-in 3-sine reference); a time-varying function is evaluated on
    the host before every env step by the single-env Controller (core/controller.py:233-239 does the same: it writes
    func(model.time) into the model before the K substeps).  require=True (vectorised views): constants only."""
    if func is None:
        return None
    a, b, c = func(0.0), func(1.2345), func(17.0)
    if not (a == b == c):
        if require:
            raise NotImplementedError(f"{what}: only constant reference functions run in-kernel "
                                      "(use ResetRefMode.OSCILLATING for the 3-sine reference)")
        return None
    return float(a)


class Controller:
    """core/controller.py:43-344."""

    def __init__(self, ctrl_type, ctrl_mode, reset_ref_mode=None, disturbance_mode=None, tk=60, sample_time=None,
                 h_func=None, vartheta_func=None, use_storage=False, action_max=17 * pi / 180,
                 vartheta_max=10 * pi / 180, use_limiter=False, logging_path=None, aero_err=None,
                 _env=None, device=0, dtype=E.F64, seed=1):
        self.ctrl_type = ctrl_type
        self.ctrl_mode = ctrl_mode
        self.reset_ref_mode = reset_ref_mode
        self.disturbance_mode = disturbance_mode
        self.aero_err = aero_err
        assert self.ctrl_mode is not None or (self.ctrl_mode is None and self.ctrl_type in [CtrlType.AUTO, CtrlType.FULL_AUTO]), \
            'Режим управления не может быть None при наличии СС НС.'
        self.use_ctrl = ctrl_type in [CtrlType.SEMI_MANUAL, CtrlType.FULL_AUTO]
        self.manual_stab = ctrl_type in [CtrlType.MANUAL, CtrlType.SEMI_MANUAL]
        self.sample_time = sample_time if sample_time else 0.01
        assert self.sample_time >= 0.01, "Шаг интегрирования не может превышать шаг взаимодействия."
        self.no_correct = True
        self.tk = tk
        self._h_func = h_func
        self._vartheta_func = vartheta_func
        # time-varying reference functions are evaluated on the host before every env step (step())
        self._vf_dynamic = vartheta_func is not None and _const_value(vartheta_func, "vartheta_func") is None
        self._hf_dynamic = h_func is not None and _const_value(h_func, "h_func") is None
        self.action_max = action_max
        self.vartheta_max = vartheta_max
        self.use_limiter = use_limiter
        # Storage: the kernel records what Controller._post_step records, after every model step
        # (core/controller.py:209-228), and evaluates calc_stepinfo online; `use_storage` only gates the reads, so
        # callers may flip it after construction like neural/callbacks.py:66 does.
        self.use_storage = use_storage
        self.storage_backup = Storage()
        self._K = E._lib.substeps_of(sample_time)
        env = _env or {}
        self._engine = E.BatchEngine(
            n_envs=1, dtype=dtype, device=device,
            obs_type=env.get("obs_type", E.OBS_PID_LIKE), rew_type=env.get("rew_type", E.REW_CLASSIC),
            norm_obs=env.get("norm_obs", True), norm_act=False,  # ControllerEnv scales the action itself, in place
            reward_config=env.get("reward_config"),
            ctrl_type=ctrl_type.value, ctrl_mode=(ctrl_mode.value if ctrl_mode is not None else E.MODE_NONE),
            reset_ref_mode=(reset_ref_mode.value if reset_ref_mode is not None else E.RESET_NONE),
            disturbance_mode=(disturbance_mode.value if disturbance_mode is not None else E.DIST_NONE),
            use_limiter=use_limiter, tk=tk, sample_time=sample_time, action_max=action_max, vartheta_max=vartheta_max,
            aero_err=aero_err, seed=seed, auto_reset=False, env_layer=True, export_signals=True,
            track_transfer=True, record_capacity=self._rec_capacity(tk))
        self.model = ModelView(self._engine, 0)
        self.state_backup = np.zeros(6)
        self._state0 = None
        self._last = None  # (obs, reward, done) of the last step, produced by the kernel

    # reference functions: assigning a new (constant) function re-targets the running episode, as
    # neural/callbacks.py does with `env.ctrl.vartheta_func = lambda _: ref`
    @property
    def vartheta_func(self):
        return self._vartheta_func

    @vartheta_func.setter
    def vartheta_func(self, f):
        self._vartheta_func = f
        v = _const_value(f, "vartheta_func")
        self._vf_dynamic = f is not None and v is None
        if v is not None:
            self._engine.set("vref", v)

    @property
    def h_func(self):
        return self._h_func

    @h_func.setter
    def h_func(self, f):
        self._h_func = f
        v = _const_value(f, "h_func")
        self._hf_dynamic = f is not None and v is None
        if v is not None:
            self._engine.set("href", v)

    def _rec_capacity(self, tk):
        n = E._lib.done_tick_of(tk)
        return int(n + self._K) if n < 10 ** 6 else 0  # an unbounded episode (tk = inf) is not recorded

    @property
    def storage(self):
        """The running episode's records, one entry per model step (empty unless use_storage)."""
        if not self.use_storage or not self._engine.cfg.record_capacity:
            return Storage()
        return Storage({k: list(v) for k, v in self._engine.recorder_read(0).items() if len(v)})

    def reset(self, state0=None):
        """core/controller.py:134-201."""
        eng = self._engine
        if self.use_storage:  # core/controller.py:195-199
            self.storage_backup = self.storage
        if state0 is not None and len(state0) > 0:
            state0 = np.asarray(state0, dtype=np.float64)
            assert state0.shape == (6,), "Размерности заданного вектора состояния state0 и вектора состояния модели не совпадают."
            assert self.reset_ref_mode is None, "Попытка случайного сброса при наличии начального вектора состояния."
            self._state0 = state0
        if self.reset_ref_mode is not None:
            assert self.ctrl_type in [CtrlType.SEMI_MANUAL, CtrlType.MANUAL], \
                "Случайный сброс не поддерживается при отсутствии СС НС в контуре."
            eng.reset()
        else:
            s0 = self._state0 if self._state0 is not None else self.model.state0
            # a time-varying function is written before every step; the episode starts from its value at t = 0
            vref = (self._vartheta_func(0.0) if self._vf_dynamic else _const_value(self._vartheta_func, "vartheta_func")) or 0.0
            href = self._h_func(0.0) if self._hf_dynamic else _const_value(self._h_func, "h_func")
            aero_err = self.aero_err
            if self.disturbance_mode == DisturbanceMode.AERO_DISTURBANCE and aero_err is None:
                # core/controller.py:181-191: a fresh Gaussian error on every reset, drawn from numpy's global stream
                # in the reference's order (ControllerEnv.__init__ seeds it with np.random.seed(0))
                aero_err = np.array([np.random.normal(m, 0.5, size=None) for m in (-0.1, 0.1, -0.1, -0.1, 0.1)])
            ep = E.episode(s0, vref=float(vref), h_ref=(float(href) if href is not None else 11000.0), use_ctrl=self.use_ctrl,
                           aero_err=aero_err)
            eng.reset_to([ep])
        self._last = None

    def step(self, action=None):
        """core/controller.py:231-264: one kernel launch = reference + action law + K model steps."""
        self.state_backup = self.model.state
        # core/controller.py:233-239: the reference is a function of model.time, written before the K substeps
        if self._vf_dynamic and not self.use_ctrl:
            self._engine.set("vref", float(self._vartheta_func(self.model.time)))
        if self._hf_dynamic and self.use_ctrl:
            self._engine.set("href", float(self._h_func(self.model.time)))
        a = 0.0
        if action is not None and len(action) > 0:
            a = float(action[-1])
        obs, rew, done = self._engine.step_host(np.array([a]))
        self._last = (obs[0].astype(np.float64), float(rew[0]), bool(done[0]))

    # ---- properties (core/controller.py:267-344) ----------------------------------------------
    @property
    def vartheta_ref(self):
        return self.model.vartheta_ref if self.model.use_PID_CS else self.model.vartheta_zh

    @property
    def dstate(self):
        return self.model.state - self.state_backup

    @property
    def dstate_dict(self):
        d = self.dstate
        return dict((self.model.labels[i], d[i]) for i in range(6))

    @property
    def err_vartheta(self):
        return self.vartheta_ref - self.model.state_dict['vartheta']

    @property
    def err_vartheta_rel(self):
        return self.err_vartheta if self.vartheta_ref == 0 else self.err_vartheta / self.vartheta_ref

    @property
    def err_h(self):
        return self.model.hzh - self.model.state_dict['y']

    @property
    def is_limit_err(self):
        return self.use_limiter and (abs(self.model.state_dict['vartheta']) > 5 * pi / 180 + self.vartheta_max
                                     or self.model.deltaz > self.action_max)

    @property
    def is_nan_err(self):
        return bool(np.isnan(np.sum(self.model.state)))

    @property
    def is_done(self):
        return self.model.time >= self.tk

    def calc_SS_err(self):
        return calc_err(self.model.state_dict['vartheta'], self.vartheta_ref)

    def calc_CS_err(self):
        return calc_err(self.model.state_dict['y'], self.model.hzh)

    def stepinfo_SS(self, use_backup=False):
        """core/controller.py:346-352.  The running episode is answered by the in-kernel tracker (calc_stepinfo
        evaluated online, constant reference); a backed-up episode from its recorded arrays."""
        return self._stepinfo("SS", 'vartheta', 'vartheta_ref', use_backup)

    def stepinfo_CS(self, use_backup=False):
        """core/controller.py:354-360."""
        return self._stepinfo("CS", 'y', 'hzh', use_backup)

    def _stepinfo(self, which, sig, ref, use_backup):
        st = (self.storage_backup if use_backup else self.storage).storage
        if not self.use_storage or sig not in st or 't' not in st:
            raise ValueError('Вычисление хар-к ПП недоступно: ошибка хранилища.')
        const_ref = len(set(st[ref])) == 1
        if use_backup or not const_ref:
            return calc_stepinfo(st[sig], st[ref][-1], ts=st['t'])
        m = self._engine.transfer_metrics(which, finished=False)[0]
        val = lambda x: None if x != x else float(x)
        return {'overshoot': val(m[0]), 'static_error': val(m[3]), 'rise_time': val(m[1]), 'settling_time': val(m[2])}

    def quality(self):
        """core/controller.py:334-336 (float64 host evaluation of two signals; the in-kernel QUALITY
        reward computes the same expression on the device)."""
        return exp(-60 * 0.1 * self.model.ITSE / (self.tk * self.vartheta_ref ** 2))
