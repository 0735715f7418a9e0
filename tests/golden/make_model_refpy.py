#!/usr/bin/env python
"""The reference's own `Model` class (core/model.py, unmodified, over the DLL; oracle/refpy.py) driven through the call
sequence of its `__main__` smoke (core/model.py:270-282) and a parameter tour; every public property recorded at
checkpoints.  Output tests/golden/model_refpy.json: what b747_rl_ctrl_b200.core.model.Model must reproduce."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refpy  # noqa: E402

PROPS = ["time", "vartheta_ref", "deltaz_ref", "deltaz_com", "deltaz_real", "CXa", "CYa", "mz", "Kalpha", "dCm_ddeltaz",
         "dvartheta", "dvartheta_int", "dvartheta_dt", "dvartheta_dt_dt", "TAE", "ITAE", "TSE", "ITSE", "AE", "IAE", "SE",
         "ISE", "hzh", "use_RP", "use_PID_SS", "use_PID_CS", "deltaz", "vartheta_zh", "P", "step_num"]
ARRS = ["state", "state0", "PID_SS", "PID_CS", "aero_err"]


def snap(m):
    d = {k: float(getattr(m, k)) for k in PROPS}
    d.update({k: [float(x) for x in getattr(m, k)] for k in ARRS})
    d["state_dict"] = {k: float(v) for k, v in m.state_dict.items()}
    return d


def main():
    M = refpy.load().model.Model
    out = {}
    # 1. core/model.py main()
    m = M(use_PID_CS=False, initial_state=np.array([100, 1000, 300, 0, 0, 0]))
    m.hzh = 2000
    m.P = 300000
    m.vartheta_zh = -0.1
    case = {"after_ctor_and_writes": snap(m), "snaps": {}}
    for n in range(1, 601):
        m.step()
        if n in (1, 2, 5, 50, 200, 600):
            case["snaps"][str(n)] = snap(m)
    out["main_smoke"] = case
    # 2. manual elevator, aero errors, re-initialisation mid-flight
    m = M(use_PID_SS=False, use_PID_CS=False, initial_state=np.array([0, 5000, 180, 2, 0.01, 0.0005]))
    m.aero_err = np.array([-0.1, 0.1, -0.1, -0.1, 0.1])
    case = {"after_ctor_and_writes": snap(m), "snaps": {}}
    rng = np.random.default_rng(3)
    for n in range(1, 301):
        if (n - 1) % 5 == 0:
            m.deltaz = float(rng.uniform(-0.25, 0.25))
        m.step()
        if n in (1, 7, 150, 300):
            case["snaps"][str(n)] = snap(m)
    m.set_initial(np.array([10, 7000, 220, -3, 0.02, 0.0]))
    m.initialize()                       # zeroes signals, time, deltaz and vartheta_zh (core/model.py:238-244)
    case["after_reinitialize"] = snap(m)
    m.deltaz = 0.05
    for n in range(40):
        m.step()
    case["snaps"]["reinit+40"] = snap(m)
    out["manual_tour"] = case
    # 3. both PIDs, altitude reference
    m = M(use_PID_SS=True, use_PID_CS=True, initial_state=np.array([0, 11000, 250, 0, 0, 0]))
    m.hzh = 10500
    case = {"after_ctor_and_writes": snap(m), "snaps": {}}
    for n in range(1, 1001):
        m.step()
        if n in (1, 100, 1000):
            case["snaps"][str(n)] = snap(m)
    out["both_pids"] = case
    out["_provenance"] = "/root/reference/core/model.py Model (unmodified, imported in place) over the DLL's machine code"
    json.dump(out, open(os.path.join(HERE, "model_refpy.json"), "w"), indent=0)
    print({k: v["snaps"][list(v["snaps"])[-1]]["state"] for k, v in out.items() if not k.startswith("_")})


if __name__ == "__main__":
    main()
