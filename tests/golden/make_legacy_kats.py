#!/usr/bin/env python
"""Known answers of the LEGACY dynamics library (core/model_win64.dll), produced by executing its own machine code
(oracle/legacy.py).  Needs /root/reference at build time of oracle/_ref; output tests/golden/legacy_kats.json."""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import legacy  # noqa: E402

CASES = {
    "L1_defaults": dict(params={}, steps=5, every=1, elevator=False),
    "L2_manual_random_elevator": dict(params={"state0": [0, 11000, 0, 250, 0, 0]}, steps=2000, every=100, elevator=True),
    "L3_ss_pid": dict(params={"state0": [0, 11000, 0, 250, 0, 0], "vartheta": 0.08726646259971647, "use_PID_SS": 1.0},
                      steps=2000, every=100, elevator=False),
    "L4_cs_ss_pid": dict(params={"state0": [0, 11000, 0, 250, 0, 0], "use_PID_SS": 1.0, "use_PID_CS": 1.0, "h_zh": 10500.0},
                         steps=3000, every=150, elevator=False),
    "L5_aero_err": dict(params={"state0": [0, 3000, 0, 150, 5, 0], "aero_err": [-0.1, 0.1, -0.1, -0.1, 0.1]},
                        steps=1500, every=100, elevator=True),
}


def main():
    out = {}
    for name, c in CASES.items():
        m = legacy.LegacyDll()
        for k, v in c["params"].items():
            m.set(k, v)
        m.initialize()
        after_init = {k: m.get(k) for k in legacy.SIGNALS}
        rng = random.Random(0)
        snaps = {}
        for n in range(1, c["steps"] + 1):
            if c["elevator"] and (n - 1) % 5 == 0:
                m.set("deltaz", rng.uniform(-0.2967, 0.2967))
            m.step()
            if n % c["every"] == 0:
                snaps[str(n)] = {k: m.get(k) for k in legacy.SIGNALS}
        out[name] = {"params": c["params"], "steps": c["steps"], "elevator": c["elevator"], "after_initialize": after_init,
                     "snaps": snaps}
        print(name, "final state", snaps[str(c["steps"])]["state"])
    out["_provenance"] = ("exported double globals of /root/reference/core/model_win64.dll after model_initialize / model_step, "
                          "its own machine code executed natively (oracle/_ref/libb747_legacy.so)")
    json.dump(out, open(os.path.join(HERE, "legacy_kats.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
