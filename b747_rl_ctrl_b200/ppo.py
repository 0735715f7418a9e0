"""End-to-end PPO on the GPU environment (BASELINE.json configs[4]: 65536-env GPU VecEnv, torch MLP policy).

The reference trains `PPO('MlpPolicy', env)` from stable-baselines3 1.4.0 with default hyper-parameters on
`SubprocVecEnv([env] * 4)` (neural/agent.py:46-81; setups.py's dict never matches, so the defaults apply): 2 x 64 tanh
MLPs for policy and value, state-independent log-std, Adam(3e-4, eps 1e-5), gamma 0.99, GAE lambda 0.95, clip 0.2,
vf_coef 0.5, max_grad_norm 0.5, advantage normalisation, n_steps 2048, batch 64, 10 epochs.  stable-baselines3 is not
installed in this image, so the algorithm is restated here in plain torch.  The loss, the optimiser and the network
are SB3's defaults; the ROLLOUT GEOMETRY is not, and `train()`'s own defaults say so: with 65536 environments one
update already holds n_steps x n_envs = 32 x 65536 = 2 Mi samples (SB3: 2048 x 4 = 8192), cut into 16 minibatches of
128 Ki (SB3: 128 minibatches of 64) for 4 epochs (SB3: 10).  Every call reports the values it used under "hyper" in
its result, and `ep_rew_mean` is the mean over the episodes that finished during the last update (episode phases are
spread first), so "steps / seconds to ep_rew_mean 225" is a statement about this geometry, not a like-for-like replay
of the reference's 4-env run.

Observations, actions, rewards and dones never leave HBM: the environment step is one launch of k_env_step32 on
torch's current stream, the policy is a torch module on the same device.  Episodes end by the time limit and are
treated as terminal, as SB3 does for an env that sets no `TimeLimit.truncated` (the reference's info dict is empty).

    python -m b747_rl_ctrl_b200.ppo --envs 65536 --threshold 225
prints one JSON line: env-steps/s (rollout + update, like SB3's time/fps) and wall-clock to the reward threshold
(the reference run reached ep_rew_mean 225.7 after 98 304 steps at ~340 steps/s).
"""
import argparse
import json
import math
import time

import torch
import torch.nn as nn

from . import engine as E


class ActorCritic(nn.Module):
    """SB3 MlpPolicy defaults: separate 2 x 64 tanh networks, orthogonal init (gain sqrt2 / 0.01 / 1), log_std = 0."""

    def __init__(self, obs_dim, hidden=64):
        super().__init__()
        def mlp(out_gain):
            layers = [nn.Linear(obs_dim, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh(), nn.Linear(hidden, 1)]
            for m in layers[:-1]:
                if isinstance(m, nn.Linear):
                    nn.init.orthogonal_(m.weight, math.sqrt(2)); nn.init.zeros_(m.bias)
            nn.init.orthogonal_(layers[-1].weight, out_gain); nn.init.zeros_(layers[-1].bias)
            return nn.Sequential(*layers)
        self.pi = mlp(0.01)
        self.vf = mlp(1.0)
        self.log_std = nn.Parameter(torch.zeros(1))

    def dist(self, obs):
        return torch.distributions.Normal(self.pi(obs).squeeze(-1), self.log_std.exp())

    def value(self, obs):
        return self.vf(obs).squeeze(-1)


def _log_prob(a, mean, log_std):
    # Normal(mean, exp(log_std)).log_prob(a), written out (no distribution objects inside captured graphs)
    return -0.5 * ((a - mean) * torch.exp(-log_std)) ** 2 - log_std - 0.5 * math.log(2 * math.pi)


def train(n_envs=65536, total_steps=None, threshold=225.0, max_seconds=600.0, n_steps=32, n_minibatches=16, n_epochs=4,
          lr=3e-4, gamma=0.99, gae_lambda=0.95, clip=0.2, vf_coef=0.5, ent_coef=0.0, max_grad_norm=0.5, seed=1, device=0,
          env_kwargs=None, log=None, desync=True, use_graphs=True, tf32=True):
    """PPO with SB3's loss / optimiser / network defaults on a rollout geometry sized for 10^4..10^5 environments (module
    docstring; the values used are returned under "hyper").  use_graphs: the T-step rollout (policy, sampling, env kernel, GAE) and one epoch of
    minibatch updates are each captured ONCE as a CUDA graph and replayed -- the eager form spends 98 % of a rollout step
    in launch overhead of ~30 tiny policy kernels around a 0.02 ms environment step."""
    torch.manual_seed(seed)
    dev = torch.device("cuda", device)
    torch.cuda.set_device(dev)
    if tf32:  # the 64-wide MLP GEMMs on the tensor cores (fp32 accumulate); SB3's own default on Ampere+ GPUs as well
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
    kw = dict(sample_time=0.05, tk=20.0)  # main.py:18, 95-96: K = 5, 400-step episodes
    kw.update(env_kwargs or {})
    eng = E.BatchEngine(n_envs=n_envs, dtype=E.F32, device=device, seed=seed, auto_reset=True, **kw)
    eng.use_stream(torch.cuda.current_stream().cuda_stream)
    act, obs, rew, done = eng.alloc_io()
    eng.reset(obs)
    od = eng.obs_dim
    # Spread the episode phases: a VecEnv whose environments all start together keeps them in lock-step (every
    # episode lasts exactly tk / sample_time steps), so each rollout would see a single slice of the episode.
    ep_len = int(round(kw["tk"] / (kw["sample_time"] or 0.01)))
    if desync and ep_len > 1:
        ids = torch.arange(n_envs, device=dev)
        act.zero_()
        for t in range(ep_len):
            eng.step(act, obs, rew, done)
            mask = ((ids % ep_len) == t).to(torch.uint8)
            eng.reset(mask=mask)
        eng.step(act, obs, rew, done)
        eng.episode_stats()  # discard the warm-up episodes
    net = ActorCritic(od).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=lr, eps=1e-5, capturable=use_graphs)
    T, N = n_steps, n_envs
    b_obs = torch.empty(T, N, od, device=dev); b_act = torch.empty(T, N, device=dev); b_logp = torch.empty(T, N, device=dev)
    b_val = torch.empty(T, N, device=dev); b_rew = torch.empty(T, N, device=dev); b_done = torch.empty(T, N, device=dev)
    adv = torch.empty(T, N, device=dev); ret = torch.empty(T, N, device=dev)
    mb = T * N // n_minibatches
    perm = torch.empty(T * N, dtype=torch.long, device=dev)
    f_obs, f_act, f_logp = b_obs.view(-1, od), b_act.view(-1), b_logp.view(-1)
    f_adv, f_ret = adv.view(-1), ret.view(-1)
    params = list(net.parameters())

    def rollout():
        with torch.no_grad():
            for t in range(T):
                mean = net.pi(obs).squeeze(-1)
                a = mean + torch.exp(net.log_std) * torch.randn_like(mean)
                b_obs[t] = obs; b_act[t] = a; b_logp[t] = _log_prob(a, mean, net.log_std); b_val[t] = net.value(obs)
                torch.clamp(a, -1.0, 1.0, out=act)   # SB3 clips the action to the Box before env.step
                eng.step(act, obs, rew, done)        # one kernel launch; obs is the reset observation where done
                b_rew[t] = rew; b_done[t] = done
            last_val = net.value(obs)
            gae = torch.zeros(N, device=dev)
            for t in reversed(range(T)):
                nonterm = 1.0 - b_done[t]
                nxt = last_val if t == T - 1 else b_val[t + 1]
                delta = b_rew[t] + gamma * nxt * nonterm - b_val[t]
                gae = delta + gamma * gae_lambda * nonterm * gae
                adv[t] = gae
            torch.add(adv, b_val, out=ret)

    def minibatch(k):
        idx = perm[k * mb:(k + 1) * mb]
        o_mb = f_obs[idx]
        logp = _log_prob(f_act[idx], net.pi(o_mb).squeeze(-1), net.log_std)
        a_mb = f_adv[idx]
        a_mb = (a_mb - a_mb.mean()) / (a_mb.std() + 1e-8)
        ratio = (logp - f_logp[idx]).exp()
        pg = -torch.min(ratio * a_mb, ratio.clamp(1 - clip, 1 + clip) * a_mb).mean()
        v_loss = ((net.value(o_mb) - f_ret[idx]) ** 2).mean()
        entropy = (0.5 + 0.5 * math.log(2 * math.pi) + net.log_std).mean()
        loss = pg + vf_coef * v_loss - ent_coef * entropy
        opt.zero_grad(set_to_none=False)
        loss.backward()
        # clip_grad_norm_ without a host read
        tot = torch.sqrt(sum((p.grad.detach() ** 2).sum() for p in params))
        scale = torch.clamp(max_grad_norm / (tot + 1e-6), max=1.0)
        for p in params:
            p.grad.mul_(scale)
        opt.step()

    def epoch():
        for k in range(n_minibatches):
            minibatch(k)

    g_roll = g_epoch = None
    if use_graphs:
        try:
            # warm-up on a side stream (allocator, autograd and optimizer state), then capture each phase once
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                eng.use_stream(side.cuda_stream)
                perm.copy_(torch.randperm(T * N, device=dev))
                rollout()
                minibatch(0)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g_roll, g_epoch = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            # the env kernel is launched through the C ABI on the handle's stream: point it at the capture stream first
            # (b747_set_stream synchronises, which is not allowed once the capture has begun)
            with torch.cuda.graph(g_roll, stream=side):
                rollout()
            with torch.cuda.graph(g_epoch, stream=side):
                epoch()
            eng.use_stream(torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            eng.episode_stats()  # the warm-up / capture steps do not count
        except Exception as e:  # capture not available: eager
            if log:
                log(f"CUDA-graph capture failed ({type(e).__name__}: {e}); running eagerly")
            g_roll = g_epoch = None
            eng.use_stream(torch.cuda.current_stream().cuda_stream)
    steps_done, updates = 0, 0
    history = []
    t_hit = None
    ep_n = ep_sum = 0.0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    while True:
        if g_roll is not None:
            g_roll.replay()
        else:
            rollout()
        for _ in range(n_epochs):
            perm.copy_(torch.randperm(T * N, device=dev))
            if g_epoch is not None:
                g_epoch.replay()
            else:
                epoch()
        steps_done += T * N
        updates += 1
        st = eng.episode_stats()  # in-kernel episode statistics since the last call (the only host read per update)
        now = time.perf_counter() - t0
        if st[0] > 0:
            ep_n, ep_sum = st[0], st[1]
            mean_r = ep_sum / ep_n
            history.append((steps_done, now, mean_r))
            if log:
                log(f"update {updates:4d}  steps {steps_done:.3e}  {steps_done / now:.3e} steps/s  ep_rew_mean {mean_r:8.2f}")
            if t_hit is None and threshold is not None and mean_r >= threshold:
                t_hit = (now, steps_done, mean_r)
                break
        if (total_steps and steps_done >= total_steps) or now > max_seconds:
            break
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    eng.close()
    hyper = dict(n_envs=n_envs, n_steps=n_steps, n_minibatches=n_minibatches, minibatch_size=mb, n_epochs=n_epochs, lr=lr,
                 gamma=gamma, gae_lambda=gae_lambda, clip=clip, vf_coef=vf_coef, ent_coef=ent_coef, max_grad_norm=max_grad_norm,
                 sb3_defaults=dict(n_envs=4, n_steps=2048, minibatch_size=64, n_epochs=10), tf32=tf32, desync=desync)
    return dict(steps=steps_done, seconds=wall, steps_per_s=steps_done / wall, updates=updates, history=history, hyper=hyper,
                threshold=threshold, reached=(None if t_hit is None else dict(seconds=t_hit[0], steps=t_hit[1], ep_rew_mean=t_hit[2])),
                final_ep_rew_mean=(history[-1][2] if history else None), net=net, graphs=g_roll is not None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--threshold", type=float, default=225.0)
    ap.add_argument("--max-seconds", type=float, default=300.0)
    ap.add_argument("--n-steps", type=int, default=32)
    ap.add_argument("--minibatches", type=int, default=16)
    ap.add_argument("--epochs", type=int, default=4)
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--quiet", action="store_true")
    ap.add_argument("--eager", action="store_true", help="no CUDA-graph capture of the rollout / update phases")
    a = ap.parse_args()
    r = train(n_envs=a.envs, threshold=a.threshold, max_seconds=a.max_seconds, n_steps=a.n_steps, n_minibatches=a.minibatches,
              n_epochs=a.epochs, lr=a.lr, log=None if a.quiet else print, use_graphs=not a.eager)
    r.pop("net"); hist = r.pop("history")
    r["history_tail"] = hist[-5:]
    r["config"] = dict(envs=a.envs, n_steps=a.n_steps, minibatches=a.minibatches, epochs=a.epochs, lr=a.lr,
                       env="PID_LIKE obs, CLASSIC reward, MANUAL/DIRECT_CONTROL, CONST reference, K=5, tk=20 (main.py:88-121)",
                       policy="2x64 tanh MLP actor + critic (SB3 MlpPolicy defaults)")
    print(json.dumps(r))


if __name__ == "__main__":
    main()
