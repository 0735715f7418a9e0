"""B747VecEnv -- the batched GPU environment with the stable-baselines3 VecEnv contract that
neural/agent.py:63-81 consumes (num_envs, observation_space, action_space, reset, step_async /
step_wait, close, seed, get_attr / set_attr / env_method / env_is_wrapped), replacing
SubprocVecEnv([ControllerEnv]*4) + VecMonitor: on done the returned observation is the reset
observation, infos[i]["terminal_observation"] holds the last one and infos[i]["episode"] the
{"r", "l", "t"} record VecMonitor would add.

Two I/O modes:
  * numpy (default; what SB3 passes): actions are staged in a pinned host buffer and obs/rew/done land in pinned
    buffers, so b747_step_host runs its chunked copy/step/copy pipeline as one CUDA-graph launch per step;
  * torch device tensors (`device_tensors=True`): obs/rew/done stay in HBM, for GPU-resident policies.

When stable-baselines3 is importable the class derives from its `VecEnv`, so `PPO('MlpPolicy', env)` takes it as is
(`BaseAlgorithm._wrap_env` wraps anything that is not a `VecEnv` instance into a DummyVecEnv).
"""
import time

import numpy as np

from . import engine as E
from .env.ctrl_env import ObservationType, RewardType, make_spaces
from .core.controller import CtrlMode, CtrlType


try:  # a real SB3 VecEnv where SB3 exists (not in this image: no network), a duck type otherwise
    from stable_baselines3.common.vec_env import VecEnv as _VecEnvBase
except Exception:
    _VecEnvBase = object


def _pinned(shape, dtype):
    """Page-locked numpy array (a view of a pinned torch tensor, kept alive by the returned pair); pageable if torch
    cannot pin here."""
    try:
        import torch
        t = torch.zeros(shape, dtype=getattr(torch, np.dtype(dtype).name)).pin_memory()
        return t.numpy(), t
    except Exception:
        return np.zeros(shape, dtype), None


def th_index(t, idx):
    import torch
    return torch.as_tensor(idx, device=t.device, dtype=torch.long)


def _val(x, default=None):
    if x is None:
        return default
    return x.value if hasattr(x, "value") else int(x)


class B747VecEnv(_VecEnvBase):
    def __init__(self, num_envs, observation_type=ObservationType.PID_LIKE, reward_type=RewardType.CLASSIC,
                 norm_obs=True, norm_act=True, ctrl_type=CtrlType.MANUAL, ctrl_mode=CtrlMode.DIRECT_CONTROL,
                 reset_ref_mode=None, disturbance_mode=None, tk=20, sample_time=0.05, action_max=17 * np.pi / 180,
                 vartheta_max=10 * np.pi / 180, use_limiter=False, aero_err=None, reward_config=None, seed=1,
                 dtype=E.F32, device=0, device_tensors=False, env_id_offset=0, monitor=True):
        from .core.controller import ResetRefMode
        if reset_ref_mode is None:
            reset_ref_mode = ResetRefMode.CONST  # a VecEnv auto-resets, which needs the random reset
        self.observation_type, self.reward_type = observation_type, reward_type
        self.engine = E.BatchEngine(
            n_envs=num_envs, dtype=dtype, device=device, obs_type=_val(observation_type), rew_type=_val(reward_type),
            norm_obs=norm_obs, norm_act=norm_act, ctrl_type=_val(ctrl_type), ctrl_mode=_val(ctrl_mode, 0),
            reset_ref_mode=_val(reset_ref_mode), disturbance_mode=_val(disturbance_mode, E.DIST_NONE),
            use_limiter=use_limiter, tk=tk, sample_time=sample_time, action_max=action_max, vartheta_max=vartheta_max,
            aero_err=aero_err, reward_config=reward_config, seed=seed, auto_reset=True, env_layer=True,
            env_id_offset=env_id_offset)
        self.num_envs = int(num_envs)
        self.observation_space, self.action_space = make_spaces(observation_type, norm_obs, norm_act, action_max)
        if _VecEnvBase is not object:
            _VecEnvBase.__init__(self, self.num_envs, self.observation_space, self.action_space)
        self.device_tensors = bool(device_tensors)
        self.monitor = bool(monitor)
        self.tk, self.sample_time, self.action_max, self.vartheta_max = tk, sample_time, action_max, vartheta_max
        self._t0 = time.time()
        self._actions = None
        # SB3 wants one info dict per env and step; building 10^5..10^6 dicts per step would dominate the step, so the
        # empty ones are shared between steps and only the environments that finished get a fresh dict
        self._no_infos = [{} for _ in range(self.num_envs)]
        od = self.engine.obs_dim
        if self.device_tensors:
            self._act_d, self._obs_d, self._rew_d, self._done_d, self._term_d = self.engine.alloc_io(terminal_obs=True)
        else:
            dt = self.engine.np_dtype
            self._keep = []
            for name, shape, d in (("_act_h", (self.num_envs,), dt), ("_obs", (self.num_envs, od), dt),
                                   ("_rew", (self.num_envs,), dt), ("_done", (self.num_envs,), np.uint8),
                                   ("_term", (self.num_envs, od), dt)):
                arr, owner = _pinned(shape, d)
                setattr(self, name, arr)
                self._keep.append(owner)

    # ---- VecEnv API ------------------------------------------------------------------------
    def reset(self):
        if self.device_tensors:
            self.engine.reset(self._obs_d)
            self.engine.synchronize()
            return self._obs_d
        self.engine.reset()
        self.engine.synchronize()
        self._obs[:] = 0  # every exported signal is zero after initialize (env/ctrl_env.py:273-278)
        return self._obs.copy()

    def step_async(self, actions):
        self._actions = actions

    def step_wait(self):
        a = self._actions
        if self.device_tensors:
            self._act_d.copy_(a.reshape(self.num_envs))
            # run on the handle's stream after torch's current stream has produced the actions
            th = self.engine._th()
            th.cuda.current_stream().synchronize()
            self.engine.step(self._act_d, self._obs_d, self._rew_d, self._done_d, self._term_d)
            self.engine.synchronize()
            done_any = bool(self._done_d.any().item())
            infos = self._infos(self._done_d.cpu().numpy(), self._term_d) if done_any else self._no_infos
            return self._obs_d, self._rew_d, self._done_d.bool(), infos
        np.copyto(self._act_h, np.asarray(a).reshape(self.num_envs), casting="unsafe")
        self.engine.step_host(self._act_h, self._obs, self._rew, self._done, self._term)
        done = self._done.astype(bool)
        infos = self._infos(self._done, self._term) if done.any() else self._no_infos
        return self._obs.copy(), self._rew.copy(), done, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def _infos(self, done, term):
        infos = list(self._no_infos)
        idx = np.nonzero(done)[0]
        if len(idx):
            ret, ln = self.engine.last_episode() if self.monitor else (None, None)
            t = round(time.time() - self._t0, 6)
            dev = hasattr(term, "cpu")
            rows = term[idx] if not dev else term[th_index(term, idx)]
            for j, i in enumerate(idx):
                info = {"terminal_observation": rows[j].clone() if dev else rows[j].copy()}
                if self.monitor:
                    info["episode"] = {"r": float(ret[i]), "l": int(ln[i]), "t": t}
                infos[i] = info
        return infos

    def close(self):
        self.engine.close()

    def seed(self, seed=None):
        # the Philox key is fixed at construction (b747_cfg.seed); SB3 calls env.seed(1) (neural/agent.py:80)
        return [seed] * self.num_envs

    def get_attr(self, attr_name, indices=None):
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [getattr(self, attr_name)] * n

    def set_attr(self, attr_name, value, indices=None):
        setattr(self, attr_name, value)

    def env_method(self, method_name, *args, indices=None, **kwargs):
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [getattr(self, method_name)(*args, **kwargs)] * n

    def env_is_wrapped(self, wrapper_class, indices=None):
        n = self.num_envs if indices is None else len(np.atleast_1d(indices))
        return [False] * n

    def episode_stats(self):
        """(episodes, sum of returns, sum of lengths, sum of squared returns) since the last call."""
        return self.engine.episode_stats()

    @property
    def unwrapped(self):
        return self

    def render(self, mode="human"):
        pass
