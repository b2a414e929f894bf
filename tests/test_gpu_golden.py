"""-m gpu: the CUDA path (through the C ABI) replays the reference's golden traces.
Integer state bit-exact; reward / distance / observation within rtol 1e-5 (float32 GAE, float64 decode)."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# g100_10k (BASELINE configs[0] at its full 10 000 steps) is an oracle-only fixture: replaying 10 k single steps through the split
# calls with a host round trip each would take minutes of GPU time; the same scenario class runs here as g100_control.
# s* (sample_subset_samples), e* (ExternalRandomEvents defender) and x* (precise_action_space_positions under a defender) are
# oracle-only too: not on the CUDA path yet.
CASES = sorted(n for n in (os.path.splitext(os.path.basename(p))[0]
                           for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
               if n != "g100_10k" and not n.startswith(("s", "e", "x")))


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("gemm", [1, 0], ids=["simt", "default"])
def test_cuda_replays_reference_trace(name, gemm, golden_dir):
    from ccbs_b200.batched_env import BatchedCyberBattleEnv
    from oracle import gen_golden as gg, trace as tr
    from oracle.cbs_oracle import OracleEnv
    from tests.gpu_harness import replay, TieFollower
    case = gg.load_case(os.path.join(golden_dir, name + ".npz"))
    if gemm == 1 and case["cfg"].distance_metric != "cosine":
        pytest.skip("the l1 / l2 / inf decode does not use the contraction the gemm switch selects")
    B = 5
    interest = None if case["interest"] is None else [case["interest"]]
    env = BatchedCyberBattleEnv([case["spec"]], case["weights"], case["cfg"], num_envs=B, auto_reset=True,
                                decode_gemm=gemm, interest_nodes=interest)
    env.set_starter_queue(np.tile(case["starters"][None, :], (B, 1)))
    follower = TieFollower(OracleEnv(case["spec"], case["weights"], case["cfg"], interest_node=case["interest"]), case["spec"],
                           case["starters"], golden_sel=case["trace"]["sel"])
    rec, consistent = replay(env, case["actions"], case["uniforms"], case["spec"].num_nodes, check_env=B - 1,
                             follower=follower, policy_rows=case["policy_rows"], defender_draws=case["defender_draws"])
    env.close()
    assert consistent, "envs fed identical inputs diverged"
    report = tr.compare(rec, case["trace"], rtol=1e-5, atol=2e-5, label=f"{name}/gemm{gemm}")
    print(name, report, "near-tie flips:", follower.flips, "oracle-vs-record flips:", follower.oracle_flips, "max gap", follower.max_gap)
    # near-ties decided by float32 rounding must stay rare: < 0.5 % of the steps.  l1 sums the rounding differences of two
    # mathematically equal source embeddings instead of letting them cancel, so a scripted attacker that keeps two such sources
    # owned sees them more often (p6_l1 on the GPU box: 3 CUDA-vs-oracle + 4 oracle-vs-record in 400 steps, all < 2e-5): < 5 %
    per = 20 if case["cfg"].distance_metric == "l1" else 200
    assert follower.flips + follower.oracle_flips <= max(1, len(case["actions"]) // per)
