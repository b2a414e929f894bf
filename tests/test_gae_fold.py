"""fold_gae: the factorised encoder (what the observe kernel evaluates) against the plain GAE forward of the oracle
(torch fp32 restatement of gae/model.py:70-82) on random visible graphs.  rtol 1e-5 / atol 2e-5 (float32)."""
import numpy as np
import torch

import ccbs_b200 as cb
import ccbs_b200.constants as C
from ccbs_b200.gae import GaeWeights, fold_gae, node_feature_static
from oracle.cbs_oracle import GaeOracle


def folded_forward(ft, node_rows, dyn, edges, edge_m):
    """numpy restatement of csrc/k_observe.cu::encode_env on the folded tables."""
    n = len(node_rows)
    y = np.zeros((n, 64), np.float32)
    T = np.zeros((n, 18, 64), np.float32)
    for i, g in enumerate(node_rows):
        vis, x = dyn[i][0], dyn[i][1:]
        T[i] = ft.node_static[g, int(vis)] + np.tensordot(x, ft.dyn_proj, axes=1)
        y[i] = T[i, 17]
    deg = np.ones(n, np.float32)
    for (s, d), m in zip(edges, edge_m):
        hb = np.concatenate([np.maximum(m + ft.nn0_b, 0), [1.0]]).astype(np.float32)
        y[d] += hb @ T[s, :17]
        if s != d:
            deg[d] += 1
    h1 = np.maximum(y * ft.bn1_scale + ft.bn1_shift, 0)
    g = h1 @ ft.gcn_wt
    dinv = 1 / np.sqrt(deg)
    out = np.zeros_like(g)
    for s, d in edges:
        if s != d:
            out[d] += dinv[s] * dinv[d] * g[s]
    out += (dinv * dinv)[:, None] * g
    return np.maximum(out * ft.bn2_scale + ft.bn2_shift, 0)


def test_folded_encoder_matches_plain_forward():
    rng = np.random.default_rng(0)
    spec = cb.synthetic_spec(5, 12)
    tables = cb.compile_scenarios([spec])
    w = GaeWeights.random(2)
    ft = fold_gae(tables, w)
    oracle = GaeOracle(w)
    for trial in range(5):
        n = int(rng.integers(1, 12))
        nodes = list(rng.permutation(12)[:n])
        dyn = []
        X = np.zeros((n, C.NODE_FEAT_DIM), np.float32)
        for i, j in enumerate(nodes):
            xa, xv = node_feature_static(spec.nodes[j], spec.vuln_emb)
            vis = float(rng.integers(0, 2))
            x = np.array([rng.integers(0, 2), rng.integers(0, 2), rng.integers(0, 2), rng.integers(0, 2),
                          rng.choice([0, 1, 3]), rng.integers(0, 2)], np.float32)
            dyn.append(np.concatenate([[vis], x]).astype(np.float32))
            full = xa + vis * xv
            for f, v in zip(C.DYN_FEATURES, x):
                full[f] = v
            X[i] = full.astype(np.float32)
        E = int(rng.integers(0, 2 * n + 1))
        pairs = set()
        while len(pairs) < min(E, n * n):
            pairs.add((int(rng.integers(n)), int(rng.integers(n))))
        edges = sorted(pairs)
        vids = list(spec.vuln_emb)
        attrs, edge_m = [], []
        for _ in edges:
            ks = rng.choice(len(vids), size=int(rng.integers(1, 4)), replace=False)
            attrs.append(np.mean([spec.vuln_emb[vids[k]] for k in ks], axis=0))
            gl = [tables.global_vuln_ids.index(vids[k]) for k in ks]
            edge_m.append(ft.vuln_h[gl].mean(axis=0))
        if edges:
            ei = torch.tensor(np.array(edges).T, dtype=torch.long)
            ea = torch.from_numpy(np.array(attrs)).float()
        else:
            ei = torch.zeros(2, 0, dtype=torch.long)
            ea = torch.zeros(C.VULN_EMB_DIM)
        want = oracle.forward(torch.from_numpy(X), ei, ea).numpy()
        got = folded_forward(ft, nodes, dyn, edges, edge_m)
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=2e-5)


def test_state_dict_roundtrip():
    w = GaeWeights.random(4)
    w2 = GaeWeights.from_state_dict(w.state_dict())
    assert np.array_equal(w.nn2_w, w2.nn2_w) and np.array_equal(w.bn2["running_var"], w2.bn2["running_var"])
