#!/usr/bin/env python
"""Generate the golden vectors in tests/golden/ by EXECUTING THE REFERENCE DLL
(core/model_simple_win64.dll hosted by oracle/_ref/libb747_ref.so; see oracle/ref_dll/pe_host.c).

Run in the build container where /root/reference is mounted:
    make -C oracle ref && python tests/golden/make_golden.py
Outputs (committed; they travel to the GPU box where the reference does not exist):
    model_kats.json   model-level known answers K1-K5: parameters, elevator stream,
                      final `state`, and all exported signals at a few step counts
    transfer_golden.json  K6: the four PID transfer episodes of the reference's control test (ADD_PROC, a = 0, refs
                      +-5, +-10 deg): calc_stepinfo figures, Controller.quality and every 50th Storage record
    env_golden.npz    env-level trajectories: for each named case the config, per-env episode
                      descriptors, action streams and the (obs, reward, done) the DLL-backed
                      ControllerEnv equivalent produced
The env layer driving the DLL is oracle/b747_env_ref.c, which reproduces SURVEY.md's K7 values
(obtained from a line-by-line emulation of env/ctrl_env.py + core/controller.py) exactly.
"""
import json
import math
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import dllref  # noqa: E402
from oracle import oracle as O  # noqa: E402

DEG = math.pi / 180


def model_kats():
    cases = {
        "K1": dict(params={"use_PID_SS": 1.0}, steps=5, elevator=False, snaps=[1, 5]),
        "K2": dict(params={"state0": [0, 11000, 250, 0, 0, 0], "vartheta": 5 * DEG}, steps=10000, elevator=True,
                   snaps=[1, 2, 7, 50, 333, 1000, 10000]),
        "K3": dict(params={"state0": [0, 11000, 250, 0, 0, 0], "vartheta": 5 * DEG, "use_PID_SS": 1.0}, steps=2000,
                   elevator=False, snaps=[1, 10, 100, 2000]),
        "K4": dict(params={"state0": [0, 11000, 250, 0, 0, 0], "vartheta": 5 * DEG, "use_PID_SS": 1.0,
                           "use_PID_CS": 1.0, "h_zh": 10500.0}, steps=6000, elevator=False, snaps=[1, 100, 1500, 6000]),
        "K5": dict(params={"aero_err": [-0.1, 0.1, -0.1, -0.1, 0.1], "state0": [0, 3000, 150, 5, 0.02, 0.001],
                           "vartheta": 5 * DEG}, steps=4000, elevator=True, snaps=[1, 3, 4, 5, 6, 1200, 4000]),
    }
    out = {}
    for name, c in cases.items():
        m = dllref.DllModel()
        for k, v in c["params"].items():
            m.set(k, v)
        m.initialize()
        rng = random.Random(0)
        snaps = {}
        for k in range(c["steps"]):
            if c["elevator"] and k % 5 == 0:
                m.set("deltaz", rng.uniform(-0.2967, 0.2967))
            m.step()
            if (k + 1) in c["snaps"]:
                snaps[str(k + 1)] = {s: m.get(s) for s in dllref.SIGNALS}
        out[name] = dict(params=c["params"], steps=c["steps"], elevator=c["elevator"], snaps=snaps)
    return out


ENV_CASES = {
    # name: (cfg kwargs, n_envs, n_steps)
    "canonical_K5": (dict(), 6, 450),
    "canonical_K10": (dict(sample_time=0.10), 3, 230),
    "canonical_K1": (dict(sample_time=None, tk=3.0), 2, 350),
    "speed_addproc": (dict(obs_type=O.OBS_SPEED_MODE, ctrl_mode=O.MODE_ADD_PROC, action_max=1.0), 3, 420),
    "aero_adddirect_osc": (dict(obs_type=O.OBS_PID_SPEED_AERO, ctrl_mode=O.MODE_ADD_DIRECT, action_max=10 * DEG,
                                reset_ref_mode=O.RESET_OSCILLATING), 3, 420),
    "state_angvel_hybrid_dist": (dict(obs_type=O.OBS_MODEL_STATE, ctrl_mode=O.MODE_ANG_VEL, action_max=2 * DEG,
                                      reset_ref_mode=O.RESET_HYBRID, disturbance_mode=O.DIST_AERO), 6, 420),
    "pidaero_pidlike_limiter": (dict(obs_type=O.OBS_PID_AERO, rew_type=O.REW_PID_LIKE, use_limiter=True), 3, 420),
    "quality_semimanual": (dict(rew_type=O.REW_QUALITY, ctrl_type=O.CTRL_SEMI_MANUAL, reset_ref_mode=O.RESET_HYBRID), 3, 420),
    "tfref_unnormalised": (dict(rew_type=O.REW_TF_REFERENCE, norm_obs=False, norm_act=False), 2, 420),
}


def env_golden():
    out = {}
    meta = {}
    for name, (kw, n, steps) in ENV_CASES.items():
        cfg = O.make_cfg(seed=7, **kw)
        od = O.OBS_DIM[cfg.obs_type]
        rng = np.random.default_rng(sum(map(ord, name)))
        amax = 1.0 if cfg.norm_act else cfg.action_max
        acts = rng.uniform(-amax, amax, size=(n, steps))
        obs = np.zeros((n, steps, od)); rew = np.zeros((n, steps)); done = np.zeros((n, steps), dtype=np.uint8)
        reset_obs = np.zeros((n, od))
        for i in range(n):
            env = O.RefEnv(cfg, env_id=i)
            reset_obs[i] = env.reset()
            o, r, d = env.rollout(acts[i], auto_reset=True)
            obs[i], rew[i], done[i] = o, r, d
        out[name + "/actions"] = acts
        out[name + "/obs"] = obs
        out[name + "/rew"] = rew
        out[name + "/done"] = done
        out[name + "/reset_obs"] = reset_obs
        meta[name] = dict(kw={k: (None if v is None else v) for k, v in kw.items()}, n=n, steps=steps, seed=7)
    out["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    return out


def k7():
    """canonical config, reset(state0=[0,11000,250,0,0,0]), constant ref +5 deg."""
    cfg = O.make_cfg(reset_ref_mode=O.RESET_NONE)
    ep = O.episode([0, 11000, 250, 0, 0, 0], vref=5 * DEG)
    env = O.RefEnv(cfg)
    env.reset_to(ep)
    rng = random.Random(0)
    acts = np.array([rng.uniform(-1, 1) for _ in range(400)])
    obs, rew, done = env.rollout(acts, auto_reset=False)
    env0 = O.RefEnv(cfg)
    env0.reset_to(ep)
    o0, r0, d0 = env0.rollout(np.zeros(400), auto_reset=False)
    cfg2 = O.make_cfg(reset_ref_mode=O.RESET_NONE, ctrl_mode=O.MODE_ADD_PROC, action_max=1.0)
    env2 = O.RefEnv(cfg2)
    env2.reset_to(ep)
    o2, r2, d2 = env2.rollout(np.zeros(400), auto_reset=False)
    return {"k7/actions": acts, "k7/obs": obs, "k7/rew": rew, "k7/done": done.astype(np.uint8),
            "k7zero/obs": o0, "k7zero/rew": r0, "k7addproc/rew": r2}


def transfer_golden():
    """SURVEY.md 8c K6 / BASELINE.md: overshoot 9.263 %, settling 11.300 s, quality 0.7527 (means over the refs)."""
    out = {}
    for deg in (5, -5, 10, -10):
        cfg = O.make_cfg(reset_ref_mode=O.RESET_NONE, ctrl_mode=O.MODE_ADD_PROC, action_max=1.0, rew_type=O.REW_QUALITY)
        env = O.RefEnv(cfg)
        env.enable_storage(2000)
        env.reset_to(O.episode([0, 11000, 250, 0, 0, 0], vref=deg * DEG))
        _, rew, done = env.rollout(np.zeros(400), auto_reset=False)
        st = env.storage
        out[str(deg)] = dict(stepinfo=env.stepinfo_SS(), stepinfo_CS=env.stepinfo_CS(), quality=float(rew[-1]),
                             n=len(st["t"]), samples={k: [float(x) for x in v[49::50]] for k, v in st.items()})
    return out


if __name__ == "__main__":
    if not dllref.available():
        sys.exit("oracle/_ref/libb747_ref.so missing: run `make -C oracle ref` where /root/reference is mounted")
    with open(os.path.join(HERE, "model_kats.json"), "w") as f:
        json.dump(model_kats(), f, indent=0)
    with open(os.path.join(HERE, "transfer_golden.json"), "w") as f:
        json.dump(transfer_golden(), f, indent=0)
    g = env_golden()
    g.update(k7())
    np.savez_compressed(os.path.join(HERE, "env_golden.npz"), **g)
    print("wrote", os.listdir(HERE))
