"""Minimal in-tree stand-ins for the parts of stable-baselines3 1.4.0 that neural/agent.py:63-81 puts around the
vectorised env (SB3 is not installable here: no network): the `VecEnv` abstract interface, `VecMonitor`'s episode
bookkeeping, and the order of calls of `OnPolicyAlgorithm.collect_rollouts`.  Written from SB3's documented contract
(method names, argument meaning, return shapes); test infrastructure only."""
import inspect
import time
from abc import ABC, abstractmethod

import numpy as np


class VecEnv(ABC):
    """The abstract methods every SB3 VecEnv implements (stable_baselines3.common.vec_env.base_vec_env.VecEnv)."""

    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs, self.observation_space, self.action_space = num_envs, observation_space, action_space

    @abstractmethod
    def reset(self): ...
    @abstractmethod
    def step_async(self, actions): ...
    @abstractmethod
    def step_wait(self): ...
    @abstractmethod
    def close(self): ...
    @abstractmethod
    def get_attr(self, attr_name, indices=None): ...
    @abstractmethod
    def set_attr(self, attr_name, value, indices=None): ...
    @abstractmethod
    def env_method(self, method_name, *method_args, indices=None, **method_kwargs): ...
    @abstractmethod
    def env_is_wrapped(self, wrapper_class, indices=None): ...
    @abstractmethod
    def seed(self, seed=None): ...

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()


def conforms(obj):
    """Names of the VecEnv abstract methods `obj` lacks or implements with an incompatible signature."""
    bad = []
    for name in VecEnv.__abstractmethods__:
        f = getattr(obj, name, None)
        if f is None or not callable(f):
            bad.append(name)
            continue
        want = [p for p in inspect.signature(getattr(VecEnv, name)).parameters.values() if p.name != "self"]
        have = inspect.signature(f).parameters
        for p in want:
            if p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD):
                continue
            if p.name not in have:
                bad.append(f"{name}({p.name})")
    for attr in ("num_envs", "observation_space", "action_space"):
        if not hasattr(obj, attr):
            bad.append(attr)
    return bad


class VecMonitor:
    """VecMonitor's bookkeeping: per-env return / length accumulators; on done the info dict gets
    info["episode"] = {"r", "l", "t"} (a copy of the env's own info), accumulators are zeroed."""

    def __init__(self, venv):
        self.venv = venv
        self.num_envs = venv.num_envs
        self.observation_space, self.action_space = venv.observation_space, venv.action_space
        self.episode_returns = self.episode_lengths = None
        self.t_start = time.time()

    def reset(self):
        obs = self.venv.reset()
        self.episode_returns = np.zeros(self.num_envs, dtype=np.float32)
        self.episode_lengths = np.zeros(self.num_envs, dtype=np.int32)
        return obs

    def step_async(self, actions):
        self.venv.step_async(actions)

    def step_wait(self):
        obs, rewards, dones, infos = self.venv.step_wait()
        self.episode_returns += rewards
        self.episode_lengths += 1
        new_infos = list(infos[:])
        for i in np.flatnonzero(dones):
            info = dict(infos[i])
            info["episode"] = {"r": float(self.episode_returns[i]), "l": int(self.episode_lengths[i]),
                               "t": round(time.time() - self.t_start, 6)}
            self.episode_returns[i] = 0
            self.episode_lengths[i] = 0
            new_infos[i] = info
        return obs, rewards, dones, new_infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def seed(self, seed=None):
        return self.venv.seed(seed)

    def get_attr(self, name, indices=None):
        return self.venv.get_attr(name, indices)

    def env_method(self, name, *a, indices=None, **k):
        return self.venv.env_method(name, *a, indices=indices, **k)

    def close(self):
        return self.venv.close()


def collect_rollouts(env, policy, n_steps, last_obs):
    """The call order of OnPolicyAlgorithm.collect_rollouts: the observation that produced an action is stored AFTER the
    following env.step returned (so arrays returned by one step must survive the next one); actions are float32
    [num_envs, 1], clipped to the action space; a finished env's value bootstrap reads infos[i]["terminal_observation"]."""
    buf_obs, buf_act, buf_rew, buf_done, terminal = [], [], [], [], []
    for _ in range(n_steps):
        actions = np.asarray(policy(last_obs), dtype=np.float32).reshape(env.num_envs, 1)
        clipped = np.clip(actions, env.action_space.low, env.action_space.high)
        new_obs, rewards, dones, infos = env.step(clipped)
        for i in np.flatnonzero(dones):
            assert infos[i].get("terminal_observation") is not None
            terminal.append((i, np.array(infos[i]["terminal_observation"]), infos[i]["episode"]))
        buf_obs.append(np.array(last_obs, copy=True))   # rollout_buffer.add(self._last_obs, ...)
        buf_act.append(actions); buf_rew.append(np.array(rewards, copy=True)); buf_done.append(np.array(dones, copy=True))
        last_obs = new_obs
    return dict(obs=np.stack(buf_obs), act=np.stack(buf_act), rew=np.stack(buf_rew), done=np.stack(buf_done),
                terminal=terminal, last_obs=last_obs)
