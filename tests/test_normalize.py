"""DeviceVecNormalize against a numpy restatement of SB3 2.3.2 VecNormalize / RunningMeanStd (CPU tensors)."""
import numpy as np
import torch

from ccbs_b200.normalize import DeviceVecNormalize


class _NpRms:
    def __init__(self, shape, eps=1e-4):
        self.mean, self.var, self.count = np.zeros(shape), np.ones(shape), eps

    def update(self, x):
        bm, bv, bc = x.mean(0), x.var(0), x.shape[0]
        d = bm - self.mean
        tot = self.count + bc
        self.mean, self.var, self.count = (self.mean + d * bc / tot,
                                           (self.var * self.count + bv * bc + d ** 2 * self.count * bc / tot) / tot, tot)


class _FakeEnv:
    def __init__(self, B, rng):
        self.num_envs, self.device, self.rng = B, torch.device("cpu"), rng
        self.obs = torch.zeros(B, 194)

    def _draw(self):
        self.obs = torch.from_numpy(self.rng.normal(1.0, 3.0, size=(self.num_envs, 194)).astype(np.float32))
        return self.obs

    def reset(self):
        return self._draw()

    def step(self, a, u, want_info=False):
        rew = torch.from_numpy(self.rng.normal(-5, 40, size=self.num_envs).astype(np.float32))
        done = torch.from_numpy((self.rng.random(self.num_envs) < 0.1).astype(np.uint8))
        return self._draw(), rew, done, None


def test_matches_numpy_restatement():
    B = 64
    env = _FakeEnv(B, np.random.default_rng(0))
    vn = DeviceVecNormalize(env, gamma=0.97)
    ref_env = _FakeEnv(B, np.random.default_rng(0))
    rms_g, rms_d, rms_r, returns = _NpRms((192,)), _NpRms((2,)), _NpRms(()), np.zeros(B)

    def np_obs(o):
        o = o.numpy().astype(np.float64)
        rms_g.update(o[:, :192]); rms_d.update(o[:, 192:])
        return np.concatenate([np.clip((o[:, :192] - rms_g.mean) / np.sqrt(rms_g.var + 1e-8), -10, 10),
                               np.clip((o[:, 192:] - rms_d.mean) / np.sqrt(rms_d.var + 1e-8), -10, 10)], axis=1)
    got = vn.reset()
    want = np_obs(ref_env.reset())
    np.testing.assert_allclose(got.numpy(), want, rtol=1e-5, atol=1e-6)
    for _ in range(30):
        o, r, d, _ = vn.step(None)
        ro, rr, rd, _ = ref_env.step(None, None)
        returns = returns * 0.97 + rr.numpy().astype(np.float64)
        rms_r.update(returns)
        want_r = np.clip(rr.numpy().astype(np.float64) / np.sqrt(rms_r.var + 1e-8), -10, 10)
        returns[rd.numpy().astype(bool)] = 0
        np.testing.assert_allclose(r.numpy(), want_r, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(o.numpy(), np_obs(ro), rtol=1e-5, atol=1e-6)
        assert np.array_equal(d.numpy(), rd.numpy())
