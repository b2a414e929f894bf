"""VecNormalize on the device.

The reference wraps its VecEnv in Stable-Baselines3's ``VecNormalize`` (agents/train_agent.py:116,
agents/multi_env/train_agent_multi_env.py:171; keys ``norm_obs`` / ``norm_reward`` of agents/config/train_config.yaml:11-12).
With the batched env the observations and rewards already live in HBM, so the running statistics are kept there too:
same update rule (parallel-variance merge of batch moments), same normalisation, clipping and discounted-return
bookkeeping as SB3 2.3.2's ``VecNormalize`` / ``RunningMeanStd`` — restated, not imported (SB3 is not installed here).
Works on any torch device (the CPU tests run it against a numpy restatement)."""
from __future__ import annotations

import torch


class RunningMeanStd:
    def __init__(self, shape=(), device="cpu", epsilon: float = 1e-4):
        self.mean = torch.zeros(shape, dtype=torch.float64, device=device)
        self.var = torch.ones(shape, dtype=torch.float64, device=device)
        self.count = epsilon

    def update(self, x: torch.Tensor) -> None:
        x = x.to(torch.float64)
        batch_mean, batch_var, batch_count = x.mean(dim=0), x.var(dim=0, unbiased=False), x.shape[0]
        delta = batch_mean - self.mean
        tot = self.count + batch_count
        m_a, m_b = self.var * self.count, batch_var * batch_count
        self.mean = self.mean + delta * batch_count / tot
        self.var = (m_a + m_b + delta.square() * self.count * batch_count / tot) / tot
        self.count = tot


class DeviceVecNormalize:
    """Normalises the [B,194] observation tensor (graph_embeddings and discrete_features have separate statistics, as
    with SB3's per-key treatment of Dict observations) and the reward vector of a BatchedCyberBattleEnv."""

    def __init__(self, env, norm_obs: bool = True, norm_reward: bool = True, clip_obs: float = 10.0, clip_reward: float = 10.0,
                 gamma: float = 0.99, epsilon: float = 1e-8, training: bool = True, obs_split: int = None):
        self.env, self.norm_obs, self.norm_reward = env, norm_obs, norm_reward
        self.clip_obs, self.clip_reward, self.gamma, self.epsilon, self.training = clip_obs, clip_reward, gamma, epsilon, training
        dev = env.device
        obs_split = env.obs.shape[1] - 2 if obs_split is None else obs_split
        self.split = obs_split
        self.obs_rms = {"graph_embeddings": RunningMeanStd((obs_split,), dev),
                        "discrete_features": RunningMeanStd((env.obs.shape[1] - obs_split,), dev)}
        self.ret_rms = RunningMeanStd((), dev)
        self.returns = torch.zeros(env.num_envs, dtype=torch.float64, device=dev)
        self.num_envs = env.num_envs

    def _norm_obs(self, obs: torch.Tensor, update: bool) -> torch.Tensor:
        if not self.norm_obs:
            return obs.clone()
        parts = []
        for key, sl in (("graph_embeddings", slice(0, self.split)), ("discrete_features", slice(self.split, None))):
            x = obs[:, sl]
            rms = self.obs_rms[key]
            if update and self.training:
                rms.update(x)
            parts.append(torch.clamp((x.to(torch.float64) - rms.mean) / torch.sqrt(rms.var + self.epsilon),
                                     -self.clip_obs, self.clip_obs))
        return torch.cat(parts, dim=1).to(torch.float32)

    def reset(self) -> torch.Tensor:
        obs = self.env.reset()
        self.returns.zero_()
        return self._norm_obs(obs, update=True)

    def step(self, actions: torch.Tensor, uniforms=None, want_info: bool = False):
        obs, rew, done, info = self.env.step(actions, uniforms, want_info=want_info)
        rew64 = rew.to(torch.float64)
        self.returns = self.returns * self.gamma + rew64
        if self.training and self.norm_reward:
            self.ret_rms.update(self.returns)
        out_rew = rew64
        if self.norm_reward:
            out_rew = torch.clamp(rew64 / torch.sqrt(self.ret_rms.var + self.epsilon), -self.clip_reward, self.clip_reward)
        self.returns = torch.where(done.bool(), torch.zeros_like(self.returns), self.returns)
        return self._norm_obs(obs, update=True), out_rew.to(torch.float32), done, info
