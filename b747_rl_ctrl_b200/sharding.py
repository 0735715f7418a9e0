"""Env-parallel sharding across the GPUs of one box.

Environments are independent, so the step path has NO collective: rank g owns the global env range
[g*N/G, (g+1)*N/G) (Philox streams are keyed by the global env id, so results do not depend on G).
The only quantity that crosses GPUs is the episode-statistics reduction -- four scalars per rollout --
done here with one all-reduce (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import numpy as np


def shard_range(n_total, world_size, rank):
    """Contiguous global env-id range [lo, hi) of `rank`."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    lo = (n_total * rank) // world_size
    hi = (n_total * (rank + 1)) // world_size
    return lo, hi


def reduce_episode_stats(stats, group=None):
    """Sum (episodes, sum_return, sum_length, sum_return_sq) over ranks.  `stats` is a length-4 array;
    returns a float64 numpy array.  A no-op without an initialised process group."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.asarray(stats, dtype=np.float64))
    if dist.is_available() and dist.is_initialized():
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def summarize(stats):
    """mean return, mean length, return std from the reduced sums."""
    n, sr, sl, sq = [float(x) for x in stats]
    if n <= 0:
        return dict(episodes=0, ep_rew_mean=float("nan"), ep_len_mean=float("nan"), ep_rew_std=float("nan"))
    mean = sr / n
    var = max(sq / n - mean * mean, 0.0)
    return dict(episodes=int(n), ep_rew_mean=mean, ep_len_mean=sl / n, ep_rew_std=var ** 0.5)
