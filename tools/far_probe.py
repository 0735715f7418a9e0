"""Distribution of the f32 path's deviation from the oracle on the far-envelope workload (tests/test_gpu_parity.py)."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from b747_rl_ctrl_b200 import engine as E
from oracle import oracle as O
from test_gpu_parity import _rollout_compare
O.build()
for K, n, hold in ((10, 512, 200), (5, 512, 40)):
    wo, wr, nd, eng = _rollout_compare(E, O, E.F32, n, 420, dict(sample_time=K * 0.01), 33, (1e9, 0.0), 1e9, sticky=hold)
    q = np.quantile(eng.env_worst, [0.5, 0.9, 0.99])
    srt = np.sort(eng.env_worst)[-6:]
    print(f"K={K} hold={hold}: max|dobs|={wo:.2e} max|drew|={wr:.2e}; median {q[0]:.1e} p90 {q[1]:.1e} p99 {q[2]:.1e}; top {srt}")
