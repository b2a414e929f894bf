#!/usr/bin/env python
"""PPO with the B200 batched env in the loop (BASELINE.json configs[4]).

stable_baselines3 / gymnasium are not installed in this image, so this is a small in-repo PPO (same structure as SB3's:
rollout buffer, GAE(lambda), clipped surrogate, value loss, entropy bonus; MultiInput obs = concat of graph_embeddings and
discrete_features; net_arch [256,128,64] as agents/config/algo_config.yaml:6).  Everything stays on the GPU: the policy's
actions go straight into ``BatchedCyberBattleEnv.step`` (zero copy) and the observation cache is read in place.  With SB3
installed, ``CyberBattleVecEnv`` (ccbs_b200/vec_env.py) plugs into ``PPO("MultiInputPolicy", venv)`` instead.

    python examples/train_ppo_b200.py --envs 4096 --updates 10
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ccbs_b200 as cb                                    # noqa: E402
from ccbs_b200 import constants as C                      # noqa: E402
from ccbs_b200.dist import reduce_episode_stats           # noqa: E402


class ActorCritic(nn.Module):
    def __init__(self, obs_dim=C.OBS_DIM + 2, act_dim=C.ACTION_DIM, arch=(256, 128, 64)):
        super().__init__()
        def mlp():
            layers, d = [], obs_dim
            for h in arch:
                layers += [nn.Linear(d, h), nn.Tanh()]
                d = h
            return nn.Sequential(*layers), d
        self.pi, d = mlp()
        self.vf, _ = mlp()
        self.mu = nn.Linear(d, act_dim)
        self.v = nn.Linear(d, 1)
        self.log_std = nn.Parameter(torch.zeros(act_dim))
        self.register_buffer("obs_scale", torch.cat([torch.ones(C.OBS_DIM), torch.full((2,), 1 / 32.0)]))

    def dist(self, obs):
        h = self.pi(obs * self.obs_scale)
        return torch.distributions.Normal(self.mu(h), self.log_std.exp())

    def value(self, obs):
        return self.v(self.vf(obs * self.obs_scale)).squeeze(-1)


def main(argv=None, quiet=False):
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--nodes", type=int, default=20)
    ap.add_argument("--scenarios", type=int, default=16)
    ap.add_argument("--n-steps", type=int, default=32)
    ap.add_argument("--updates", type=int, default=10)
    ap.add_argument("--epochs", type=int, default=4)
    ap.add_argument("--minibatch", type=int, default=16384)
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args(argv)
    torch.manual_seed(args.seed)
    dev = torch.device("cuda", 0)
    pool = cb.synthetic_vuln_pool(1234, 200)
    specs = [cb.synthetic_spec(100 + k, args.nodes, pool=pool) for k in range(args.scenarios)]
    env = cb.BatchedCyberBattleEnv(specs, cb.GaeWeights.random(0), cb.EnvConfig(), num_envs=args.envs, seed=args.seed,
                                   switch_interval=5)
    B, T = args.envs, args.n_steps
    ac = ActorCritic().to(dev)
    opt = torch.optim.Adam(ac.parameters(), lr=args.lr)
    gamma, lam, clip, vf_coef, ent_coef = 0.99, 0.95, 0.2, 0.5, 0.0
    obs_buf = torch.empty(T, B, C.OBS_DIM + 2, device=dev)
    act_buf = torch.empty(T, B, C.ACTION_DIM, device=dev)
    logp_buf, rew_buf, val_buf = (torch.empty(T, B, device=dev) for _ in range(3))
    done_buf = torch.empty(T, B, device=dev)
    obs = env.reset().clone()
    log = []
    for upd in range(args.updates):
        env.reset_stat_accum()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad():
            for t in range(T):
                d = ac.dist(obs)
                a = d.sample()
                obs_buf[t], act_buf[t] = obs, a
                logp_buf[t] = d.log_prob(a).sum(-1)
                val_buf[t] = ac.value(obs)
                nobs, rew, done, _ = env.step(a.clamp(-4.0, 4.0), None, want_info=False)   # clip to the Box like SB3
                rew_buf[t], done_buf[t] = rew * 1e-3, done.float()                          # reward scale ~ winning_reward
                obs = nobs.clone()
            last_val = ac.value(obs)
        torch.cuda.synchronize()
        t_roll = time.perf_counter() - t0
        # GAE(lambda); an episode end (goal, loss or cut-off) is terminal, as in the reference (compressed:451)
        adv = torch.zeros_like(rew_buf)
        gae = torch.zeros(B, device=dev)
        for t in reversed(range(T)):
            nv = last_val if t == T - 1 else val_buf[t + 1]
            nonterm = 1.0 - done_buf[t]
            delta = rew_buf[t] + gamma * nv * nonterm - val_buf[t]
            gae = delta + gamma * lam * nonterm * gae
            adv[t] = gae
        ret = adv + val_buf
        flat = lambda x: x.reshape(T * B, *x.shape[2:])   # noqa: E731
        O, A, LP, ADV, RET = flat(obs_buf), flat(act_buf), flat(logp_buf), flat(adv), flat(ret)
        ADV = (ADV - ADV.mean()) / (ADV.std() + 1e-8)
        for _ in range(args.epochs):
            perm = torch.randperm(T * B, device=dev)
            for i in range(0, T * B, args.minibatch):
                idx = perm[i:i + args.minibatch]
                d = ac.dist(O[idx])
                lp = d.log_prob(A[idx]).sum(-1)
                ratio = (lp - LP[idx]).exp()
                pg = -torch.min(ratio * ADV[idx], ratio.clamp(1 - clip, 1 + clip) * ADV[idx]).mean()
                vl = (ac.value(O[idx]) - RET[idx]).pow(2).mean()
                loss = pg + vf_coef * vl - ent_coef * d.entropy().sum(-1).mean()
                opt.zero_grad(set_to_none=True)
                loss.backward()
                nn.utils.clip_grad_norm_(ac.parameters(), 0.5)
                opt.step()
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
        st = reduce_episode_stats(env.stat_accum_tensor())
        row = dict(update=upd, rollout_env_steps_per_s=T * B / t_roll, train_env_steps_per_s=T * B / t_all,
                   seconds_per_update=t_all, ep_rew_mean=st["ep_rew_mean"], ep_len_mean=st["ep_len_mean"],
                   episodes=st["episodes"], owned_mean=st["stat0"] / max(st["episodes"], 1), win_rate=st["win_rate"])
        log.append(row)
        if not quiet:
            print(json.dumps(row))
    env.sync()
    env.close()
    return log


if __name__ == "__main__":
    main()
