"""Control tests on the GPU: the reference's acceptance procedure, batched.

`neural/callbacks.py:46-120` (ControlTestCallback.calc_stepinfo) and `neural/agent.py:235-409` (ControllerAgent.test)
evaluate a controller on deterministic episodes -- state0 = [0, 11000, 250, 0, 0, 0], constant pitch references
(+-5, +-10 deg in main.py:112-121), CtrlType.MANUAL, no random reset -- one episode after the other through Python,
recording every model step and then scanning the recording (`calc_stepinfo`).  Here every (reference, controller)
pair is one environment of one handle: the policy is queried once per env step for the whole batch, the step-response
figures are accumulated in-kernel after every model step (b747_common.cuh `trk_update`), and only five numbers per
episode come back.  The published `transfer_custom/*` numbers are reproduced by
tests/test_gpu_transfer.py.
"""
import math

import numpy as np

from . import engine as E

DEG = math.pi / 180
DEFAULT_STATE0 = (0.0, 11000.0, 250.0, 0.0, 0.0, 0.0)   # main.py:113
DEFAULT_REFS = (5 * DEG, -5 * DEG, 10 * DEG, -10 * DEG)  # main.py:112


def _enum(x, default=None):
    if x is None:
        return default
    return x.value if hasattr(x, "value") else int(x)


def run_control_test(policy=None, vartheta_ref=DEFAULT_REFS, state0=DEFAULT_STATE0, *, observation_type=E.OBS_PID_LIKE,
                     reward_type=E.REW_CLASSIC, norm_obs=True, norm_act=True, ctrl_mode=E.MODE_DIRECT, ctrl_type=E.CTRL_MANUAL,
                     tk=20.0, sample_time=0.05, action_max=17 * DEG, vartheta_max=10 * DEG, use_limiter=False,
                     aero_err=None, disturbance_mode=None, reward_config=None, pid_ss=None, pid_cs=None, h_ref=None,
                     dtype=E.F64, device=0, record=False, max_steps=None):
    """One deterministic episode per reference value, all in one batch.

    policy: callable(obs[n, obs_dim] float array) -> actions[n] (or [n, 1]); None -> zero actions, i.e. the pure
            PID loop in ADD_PROC/ADD_DIRECT modes or with ctrl_type AUTO (what agent.test calls `no_neural`).
    h_ref : per-episode altitude references instead of pitch references (ctrl types that close the altitude loop).
    Returns {"overshoot", "rise_time", "settling_time", "static_error", "quality", "return", "length"}: arrays with
    one entry per reference (NaN where the reference yields None) [+ "storage": per-episode recordings if record].
    """
    refs = [float(v) for v in (vartheta_ref if h_ref is None else h_ref)]
    n = len(refs)
    K = E._lib.substeps_of(sample_time)
    n_model = E._lib.done_tick_of(tk)
    eng = E.BatchEngine(n_envs=n, dtype=dtype, device=device, obs_type=_enum(observation_type), rew_type=_enum(reward_type),
                        norm_obs=norm_obs, norm_act=norm_act, ctrl_type=_enum(ctrl_type), ctrl_mode=_enum(ctrl_mode, E.MODE_NONE),
                        reset_ref_mode=E.RESET_NONE, disturbance_mode=_enum(disturbance_mode, E.DIST_NONE),
                        use_limiter=use_limiter, tk=tk, sample_time=sample_time, action_max=action_max,
                        vartheta_max=vartheta_max, aero_err=aero_err, reward_config=reward_config, auto_reset=False,
                        track_transfer=True, record_capacity=(n_model + K if record else 0))
    try:
        if pid_ss is not None:
            eng.set_param("PID_SS", np.asarray(pid_ss, np.float64))
        if pid_cs is not None:
            eng.set_param("PID_CS", np.asarray(pid_cs, np.float64))
        use_ctrl = _enum(ctrl_type) in (E.CTRL_SEMI_MANUAL, E.CTRL_FULL_AUTO)
        eps = [E.episode(state0, vref=(0.0 if h_ref is not None else r), h_ref=(r if h_ref is not None else 11000.0),
                         use_ctrl=use_ctrl, aero_err=aero_err) for r in refs]
        eng.reset_to(eps)
        obs = np.zeros((n, eng.obs_dim), eng.np_dtype)
        ret = np.zeros(n)
        length = np.zeros(n, np.int64)
        alive = np.ones(n, bool)
        limit = max_steps if max_steps is not None else -(-n_model // K) + 1
        for _ in range(limit):
            a = np.zeros(n, eng.np_dtype) if policy is None else np.asarray(policy(obs), eng.np_dtype).reshape(n)
            obs, rew, done = eng.step_host(a)
            ret += np.where(alive, rew, 0.0)
            length += alive
            alive &= ~done.astype(bool)
            if not alive.any():
                break
        which = "CS" if h_ref is not None else "SS"
        m = eng.transfer_metrics(which, finished=True)
        out = {k: m[:, j].copy() for j, k in enumerate(eng.METRIC_NAMES)}
        out["return"], out["length"] = ret, length
        if record:
            out["storage"] = [eng.recorder_read(i) for i in range(n)]
        return out
    finally:
        eng.close()


class ControlTest:
    """The bookkeeping of ControlTestCallback (neural/callbacks.py:46-120) without the SB3 base class: every call of
    `evaluate(policy)` runs the test episodes, appends the means over the references to a sliding window
    (`window_length`, default 30) and exposes the windowed means that the callback logs as
    transfer_custom/{settling_time, overshoot, quality}; `best_mean_quality` tracks the best windowed quality."""

    def __init__(self, vartheta_ref=DEFAULT_REFS, state0=DEFAULT_STATE0, window_length=30, **env_kwargs):
        self.vartheta_ref = list(vartheta_ref) if isinstance(vartheta_ref, (list, tuple)) else [vartheta_ref]
        self.state0 = state0
        self.window_length = window_length
        self.env_kwargs = env_kwargs
        self.infos = {'settling_time': [], 'overshoot': [], 'quality': []}
        self.best_mean_quality = self.mean_quality = 0
        self.last = None

    def evaluate(self, policy=None):
        r = run_control_test(policy, self.vartheta_ref, self.state0, **self.env_kwargs)
        self.last = r
        self.infos['settling_time'].append(float(np.mean(r["settling_time"])))
        self.infos['overshoot'].append(float(np.mean(np.abs(r["overshoot"]))))
        self.infos['quality'].append(float(np.mean(r["quality"])))
        for k in self.infos:
            self.infos[k] = self.infos[k][-self.window_length:]
        self.mean_quality = float(np.mean(self.infos['quality']))
        log = {'transfer_custom/settling_time': float(np.mean(self.infos['settling_time'])),
               'transfer_custom/overshoot': float(np.mean(self.infos['overshoot'])),
               'transfer_custom/quality': self.mean_quality}
        improved = self.mean_quality > self.best_mean_quality
        if improved:
            self.best_mean_quality = self.mean_quality
        log['improved'] = improved
        return log


# ---- ControllerAgent.test (neural/agent.py:235-409): PID baselines beside the trained policies -------------------------
TABLE_COLUMNS = ('Устройство', 'σ, [%]', 'tпп, [с]', 'tв, [с]', 'Δ', 'Q, [-]')   # the reference's DataFrame columns


def run_agent_test(ref_values, policies=None, state0=DEFAULT_STATE0, *, ctrl_type=E.CTRL_MANUAL, pid_coefs=(),
                   no_neural=False, **env_kwargs):
    """The comparison `ControllerAgent.test` prints and writes to xlsx, as plain tables, every episode of a device in one
    batch.

    For every reference value (pitch references in rad for CtrlType.MANUAL, altitude references in m for SEMI_MANUAL --
    the reference switches on `env.ctrl.use_ctrl`, agent.py:250-253) one row per device:
      * the PID baseline(s): the same loop with the СС PID in place of the network -- ctrl_type AUTO (from MANUAL) or
        FULL_AUTO (from SEMI_MANUAL), no action law, sample_time = dt, tk / dt interactions (agent.py:300-307, 342-344);
        `pid_coefs` = alternative coefficient sets, written to PID_SS or PID_CS (agent.py:292-297);
      * every policy of `policies` ({name: callable(obs[n, obs_dim]) -> actions[n]}), deterministic, on its own env
        configuration (`env_kwargs`, shared by all policies here), tk / sample_time interactions.
    A row = overshoot [%], settling time, rise time, static error (calc_stepinfo of the pitch angle or the altitude,
    stepinfo_SS / stepinfo_CS) and Controller.quality().  Returns {"tables": {ref: [rows]}, "mean": [rows]}: `mean` is
    the reference's data_*_info_mean table -- |overshoot| first, then the mean over the reference values per device.
    """
    ctrl_type = _enum(ctrl_type)
    use_ctrl = ctrl_type in (E.CTRL_SEMI_MANUAL, E.CTRL_FULL_AUTO)
    pid_type = E.CTRL_FULL_AUTO if use_ctrl else E.CTRL_AUTO
    refs = [float(v) for v in ref_values]
    base_name = "CУ ПИД" if use_ctrl else "СС ПИД"
    coef_sets = [np.asarray(c, np.float64) for c in pid_coefs] or [None]
    which = dict(h_ref=refs) if use_ctrl else dict(vartheta_ref=refs)
    devices = []   # (name, result dict with one entry per reference)
    pid_kw = {k: v for k, v in env_kwargs.items() if k not in ("ctrl_mode", "sample_time", "observation_type", "reward_type")}
    for i, coefs in enumerate(coef_sets):
        name = base_name + (f" [{i + 1}]" if len(coef_sets) > 1 else "")
        kw = dict(pid_kw, ctrl_type=pid_type, ctrl_mode=None, sample_time=None)   # agent.py:301: env_PID.ctrl.ctrl_mode = None
        if coefs is not None:
            kw["pid_cs" if use_ctrl else "pid_ss"] = coefs
        devices.append((name, run_control_test(None, state0=state0, **which, **kw)))
    if not no_neural:
        for name, policy in (policies or {}).items():
            devices.append((name, run_control_test(policy, state0=state0, ctrl_type=ctrl_type, **which, **env_kwargs)))
    unit = 'Δ, [м]' if use_ctrl else 'Δ, [град]'

    def row(name, r, j):
        val = lambda x: None if x != x else float(x)
        return {'Устройство': name, 'σ, [%]': val(r["overshoot"][j]), 'tпп, [с]': val(r["settling_time"][j]),
                'tв, [с]': val(r["rise_time"][j]), unit: val(r["static_error"][j]), 'Q, [-]': val(r["quality"][j])}
    tables = {ref: [row(name, r, j) for name, r in devices] for j, ref in enumerate(refs)}
    mean = []
    for name, r in devices:
        m = {'Устройство': name}
        for col, key, absolute in (('σ, [%]', "overshoot", True), ('tпп, [с]', "settling_time", False),
                                   ('tв, [с]', "rise_time", False), (unit, "static_error", False), ('Q, [-]', "quality", False)):
            v = np.abs(r[key]) if absolute else r[key]
            m[col] = float(np.nanmean(v)) if np.isfinite(v).any() else None
        mean.append(m)
    return {"tables": tables, "mean": mean, "unit": unit}
