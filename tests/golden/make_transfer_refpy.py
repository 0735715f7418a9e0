#!/usr/bin/env python
"""The reference's acceptance procedure run by the reference's OWN Python (oracle/refpy.py): ControlTestCallback's
episode loop (neural/callbacks.py:61-100) on env/ctrl_env.py + core/controller.py + tools/general.py over the DLL --
`ctrl.use_storage = True`, `ctrl.vartheta_func = lambda _: vref`, `env.reset(state0)`, step to done, then the reference's
`ctrl.stepinfo_SS()` (its own `calc_stepinfo`) and `ctrl.quality()` -- plus `calc_stepinfo` / `calc_err` on random arrays.
Output: tests/golden/transfer_refpy.json.  Needs /root/reference (build container only)."""
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refpy  # noqa: E402

DEG = math.pi / 180


def main():
    ref = refpy.load()
    CE, C = ref.ctrl_env, ref.controller
    import tools.general as G   # the reference's own tools/general.py (imported by core/controller.py)
    assert os.path.realpath(G.__file__).startswith(os.path.realpath(refpy.REFERENCE))
    out = {"episodes": {}, "calc_stepinfo": [], "calc_err": []}
    for mode, amax in ((C.CtrlMode.ADD_PROC_CONTROL, 1.0), (C.CtrlMode.DIRECT_CONTROL, 17 * DEG)):
        for deg in (5, -5, 10, -10):
            env = CE.ControllerEnv(CE.ObservationType.PID_LIKE, CE.RewardType.CLASSIC, True, True, C.CtrlType.MANUAL, mode,
                                   tk=20, reset_ref_mode=None, sample_time=0.05, use_limiter=False, action_max=amax,
                                   vartheta_func=lambda _: 0.0)
            ctrl = env.ctrl
            ctrl.use_storage = True
            ctrl.vartheta_func = lambda _, v=deg * DEG: v
            ctrl._init_model()
            obs = env.reset(np.array([0, 11000, 250, 0, 0, 0], dtype=float))
            done, ret, n = False, 0.0, 0
            rng = np.random.default_rng(deg + 100)
            while not done:
                a = np.array([0.0]) if mode == C.CtrlMode.ADD_PROC_CONTROL else np.array([float(np.float32(rng.uniform(-0.3, 0.3)))])
                obs, r, done, _ = env.step(a)
                ret += r
                n += 1
            info = ctrl.stepinfo_SS(use_backup=False)
            st = ctrl.storage.storage
            out["episodes"][f"{mode.name}/{deg}"] = {
                "stepinfo": info, "quality": ctrl.quality(), "return": ret, "length": n, "records": len(st["t"]),
                "err_vartheta": ctrl.err_vartheta, "calc_SS_err": ctrl.calc_SS_err(), "err_h": ctrl.err_h,
                "samples": {k: [float(x) for x in st[k][49::250]] for k in ("t", "U_com", "U_PID", "deltaz", "vartheta", "y", "Vx")},
                "actions_seed": deg + 100}
            print(mode.name, deg, info, ctrl.quality())
    rng = np.random.default_rng(0)
    for k in range(40):
        n = int(rng.integers(5, 60))
        ts = np.cumsum(rng.uniform(0.01, 0.1, n))
        base = float(rng.choice([-1, 1]) * rng.uniform(0.5, 5))
        ys = base * (1 - np.exp(-ts * rng.uniform(0.5, 4)) * np.cos(ts * rng.uniform(0, 6))) + rng.normal(0, 0.02, n)
        if k % 7 == 0:
            base = 0.0
        info = G.calc_stepinfo(list(ys), base, ts=list(ts))
        out["calc_stepinfo"].append({"ys": list(map(float, ys)), "ts": list(map(float, ts)), "y_base": base, "info": info})
    for a, b in ((1.0, 2.0), (0.0, 2.0), (3.0, 0.0), (0.0, 0.0), (-1.5, 0.5)):
        out["calc_err"].append([a, b, G.calc_err(a, b)])
    out["calc_exp_k"] = [G.calc_exp_k(0.8, 10), G.calc_exp_k(0.75, 0.15)]
    out["_provenance"] = ("/root/reference/env/ctrl_env.py + core/controller.py + tools/general.py (unmodified, imported in place) "
                          "over core/model_simple_win64.dll's machine code; see oracle/refpy.py")
    json.dump(out, open(os.path.join(HERE, "transfer_refpy.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
