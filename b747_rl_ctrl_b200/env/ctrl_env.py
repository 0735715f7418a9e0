"""ControllerEnv -- drop-in for env/ctrl_env.py over the CUDA engine (single environment, gym 0.19 API:
reset() -> obs, step(a) -> (obs, reward, done, info)).  Observation, reward and done come from the
kernel; Python only keeps the reference's API quirks (in-place `action *= action_max`, zero reset
observation, reward function stored on the class, old 4-tuple API).  For throughput use B747VecEnv.
"""
from enum import Enum
from math import pi

import numpy as np

from .. import engine as E
from ..core.controller import Controller, CtrlMode


class ObservationType(Enum):  # env/ctrl_env.py:16-22
    PID_LIKE = 0
    SPEED_MODE = 1
    PID_AERO = 2
    PID_SPEED_AERO = 3
    MODEL_STATE = 4


class RewardType(Enum):  # env/ctrl_env.py:24-30
    CLASSIC = 0
    PID_LIKE = 1
    QUALITY = 2
    MINIMAL = 3
    TF_REFERENCE = 4


try:  # gym is optional: the reference pins gym 0.19, which is not installed everywhere
    from gym import spaces as _spaces
    Box = _spaces.Box
except Exception:  # minimal stand-in with the attributes SB3-style code reads
    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.shape = tuple(shape) if shape is not None else np.shape(low)
            self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape).copy()
            self.dtype = np.dtype(dtype)

        def sample(self):
            return np.random.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


OBS_MAX = {  # env/ctrl_env.py:200-214
    ObservationType.PID_LIKE: np.array([60 * pi, pi, pi]),
    ObservationType.SPEED_MODE: np.array([60 * pi, pi, pi, 500, 100]),
    ObservationType.PID_SPEED_AERO: np.array([60 * pi, pi, pi, 500, 100, 0.5, 2, 0.6, 0.05, 1.]),
    ObservationType.PID_AERO: np.array([60 * pi, pi, pi, 0.5, 2, 0.6, 0.05, 1.]),
    ObservationType.MODEL_STATE: np.array([10 * pi / 180, 12000, 15000, 500, 100, pi, pi]),
}


def make_spaces(observation_type, norm_obs, norm_act, action_max):
    obs_max = OBS_MAX[observation_type]
    acts_high = np.array([action_max])
    if norm_act:
        action_space = Box(low=-1, high=1, shape=acts_high.shape)
    else:
        action_space = Box(low=-acts_high, high=acts_high, shape=acts_high.shape)
    if norm_obs:
        observation_space = Box(low=-1, high=1, shape=obs_max.shape)
    else:
        observation_space = Box(low=-obs_max, high=obs_max, shape=obs_max.shape)
    return observation_space, action_space


class ControllerEnv:
    """env/ctrl_env.py:61-282."""
    metadata = {'render.modes': ['human']}

    def __init__(self, observation_type, reward_type, norm_obs, norm_act, *ctrl_args, **ctrl_kwargs):
        self.observation_type = observation_type
        self.reward_type = reward_type
        self.norm_obs = norm_obs
        self.norm_act = norm_act
        self.reward_range = (0, 1)
        self._reward_config = {}
        self._ctrl_args, self._ctrl_kwargs = ctrl_args, dict(ctrl_kwargs)
        self._build()
        self.observation_space, self.action_space = make_spaces(observation_type, norm_obs, norm_act,
                                                                self.ctrl.action_max)
        self.state_box = np.zeros(self.observation_space.shape)
        self._last_reward = 0.0
        self.reset()

    def _build(self):
        env_cfg = dict(obs_type=self.observation_type.value, rew_type=self.reward_type.value, norm_obs=self.norm_obs,
                       reward_config=self._reward_config)
        self.ctrl = Controller(*self._ctrl_args, _env=env_cfg, **self._ctrl_kwargs)

    def _get_action_def(self):
        return np.array([-self.ctrl.action_max]), np.array([self.ctrl.action_max])

    def _get_obs_def(self):
        m = OBS_MAX[self.observation_type]
        return -m, m

    def set_rew_config(self, rew_config):
        """env/ctrl_env.py:250-252: new reward constants (rebuilds the handle's configuration)."""
        self._reward_config = dict(rew_config)
        vfunc, hfunc = self.ctrl.vartheta_func, self.ctrl.h_func
        self._build()
        self.ctrl.vartheta_func, self.ctrl.h_func = vfunc, hfunc
        self.reset()

    def get_reward(self, action=None):
        return self._last_reward

    def is_done(self):
        return self.ctrl.is_done or self.ctrl.is_nan_err or self.ctrl.is_limit_err

    def step(self, action):
        """env/ctrl_env.py:260-270."""
        if self.norm_act and action is not None:
            _, action_max = self._get_action_def()
            action *= action_max  # in place on the caller's array, like the reference
        self.ctrl.step(action)
        obs, rew, done = self.ctrl._last
        self.state_box = obs
        self._last_reward = rew
        return self.state_box, rew, done, {}

    def reset(self, state0=None):
        """env/ctrl_env.py:273-278: all exported signals are zero after initialize, so is the observation."""
        self.ctrl.reset(state0=state0)
        self.state_box = np.zeros(self.observation_space.shape)
        return self.state_box

    def render(self, mode='human'):
        pass

    def seed(self, seed=None):
        return [seed]

    def close(self):
        self.ctrl._engine.close()
