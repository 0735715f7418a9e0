"""Distribution over environments of the f32 path's worst deviation from the float64 oracle (canonical config)."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from b747_rl_ctrl_b200 import engine as E
from oracle import oracle as O
from test_gpu_parity import _rollout_compare
O.build()
for K, n in ((5, 4096), (10, 4096), (10, 16384)):
    wo, wr, nd, eng = _rollout_compare(E, O, E.F32, n, 1000, dict(sample_time=K * 0.01), 21, (1e9, 0.0), 1e9)
    w = eng.env_worst
    q = np.quantile(w, [0.5, 0.9, 0.99, 0.999])
    print(f"K={K} n={n}: max|dobs|={wo:.2e} max|drew|={wr:.2e}; per-env worst |dobs| median {q[0]:.1e} p90 {q[1]:.1e} "
          f"p99 {q[2]:.1e} p99.9 {q[3]:.1e}; envs > 1e-5: {(w > 1e-5).sum()}  > 1e-4: {(w > 1e-4).sum()}", flush=True)
