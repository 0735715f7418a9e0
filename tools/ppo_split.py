"""PPO throughput on the 65536-env GPU VecEnv: CUDA-graph replay vs eager, TF32 vs fp32 GEMMs."""
import sys
sys.path.insert(0, ".")
import torch
from b747_rl_ctrl_b200 import ppo
for g, tf in ((True, True), (True, False), (False, True)):
    torch.backends.cuda.matmul.allow_tf32 = False
    r = ppo.train(n_envs=65536, threshold=None, total_steps=65536 * 32 * 10, max_seconds=120, seed=1, use_graphs=g, tf32=tf)
    print("graphs" if r["graphs"] else "eager", "tf32" if tf else "fp32", ": 10 updates", round(r["seconds"], 3), "s", f"{r['steps_per_s']:.3e}", "steps/s", "final ep_rew_mean", r["final_ep_rew_mean"])
