"""-m gpu: the decode contraction VT = A_v x Vemb^T (tcgen05 TF32 path and SIMT float32 path) against float64."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def check_products(gemm, tol, num_envs, pool, prestaged=False):
    import torch
    import ccbs_b200 as cb
    from ccbs_b200.batched_env import BatchedCyberBattleEnv
    from ccbs_b200.gae import GaeWeights
    p = cb.synthetic_vuln_pool(7, pool)
    specs = [cb.synthetic_spec(300 + k, 14, pool=p, vulns_per_service_range=(6, 14)) for k in range(12)]
    env = BatchedCyberBattleEnv(specs, GaeWeights.random(0), cb.EnvConfig(), num_envs=num_envs, decode_gemm=gemm)
    if gemm == 0:
        assert env.tensor_core_decode, "tcgen05 path not active on this device"
    if prestaged:
        env.set_actions_prestaged(True)
    env.reset()
    rng = np.random.default_rng(1)
    a = rng.uniform(-4, 4, size=(num_envs, 905)).astype(np.float32)
    env.decode(torch.from_numpy(a).to(env.device))
    env.sync()
    Ug = env.tables.vemb32.shape[0]
    got = env.vt()[:, :Ug].astype(np.float64)
    want = a[:, 128:896].astype(np.float64) @ env.tables.vemb64.T
    err = np.abs(got - want)
    print("Ug", Ug, "max abs err", err.max(), "rel fro", np.linalg.norm(got - want) / np.linalg.norm(want))
    assert err.max() < tol
    env.close()


@pytest.mark.parametrize("gemm,tol", [(1, 1e-3), (0, 0.35)], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("num_envs,pool", [(300, 40), (1024, 200), (130, 330)])
def test_decode_products(gemm, tol, num_envs, pool):
    check_products(gemm, tol, num_envs, pool)


def test_decode_products_prestaged_actions():
    """cbs_set_actions_prestaged: the contraction reads its actions before waiting for the previous kernel on the stream."""
    check_products(0, 0.35, 1024, 200, prestaged=True)


def test_decode_products_f16_variant():
    """The opt-in half-precision contraction (CBS_GEMM_F16=1, read once per process: run in a child)."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r); from tests.test_gpu_decode_gemm import check_products\n"
            "for n, p in ((300, 40), (1024, 200), (130, 330)): check_products(0, 0.35, n, p)\nprint('f16 ok')" % root)
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, CBS_GEMM_F16="1"), capture_output=True, text=True, timeout=300)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "f16 ok" in r.stdout
