"""env-steps/s of the reference's OWN Python path -- env/ctrl_env.py -> core/controller.py -> core/model.py (unmodified, imported
in place; oracle/refpy.py) over the DLL's machine code -- on one core of THIS container (needs /root/reference, so it cannot
run on the GPU box; bench.py's CPU arms time the DLL + C env layer there).  The canonical env of main.py:88-121."""
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refpy  # noqa: E402

ref = refpy.load()
CE, C = ref.ctrl_env, ref.controller
DEG = math.pi / 180
for K in (5, 10):
    env = CE.ControllerEnv(CE.ObservationType.PID_LIKE, CE.RewardType.CLASSIC, True, True, C.CtrlType.MANUAL,
                           C.CtrlMode.DIRECT_CONTROL, reset_ref_mode=C.ResetRefMode.CONST, tk=20, sample_time=K * 0.01,
                           use_limiter=False, action_max=17 * DEG)
    env.reset()
    rng = np.random.default_rng(0)
    n = 3000
    acts = rng.uniform(-1, 1, (n, 1))
    t = time.perf_counter()
    for k in range(n):
        _, _, done, _ = env.step(acts[k].copy())
        if done:
            env.reset()
    dt = time.perf_counter() - t
    print(f"reference Python ControllerEnv.step, K={K}: {n / dt:.0f} env-steps/s on one core ({n} steps in {dt:.2f} s)")
