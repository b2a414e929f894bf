"""Graph auto-encoder (encoder only) weights and their folding into per-scenario tables.

Reference: gae/model.py:25-82 with the default layer config gae/config/train_config.yaml:4-12 —
``NNConv(1576->64, edge nn 768->16->1576*64) -> BatchNorm -> ReLU -> GCNConv(64->64) -> BatchNorm -> ReLU``,
run in eval mode (agents/train_agent.py:333).

The observation kernel never forms the 1576-wide node feature vector.  NNConv's message
``x_j @ reshape(W2 h_e + b2, [1576, 64])`` is bilinear in ``x_j`` and ``[h_e; 1]``, and ``x_j`` is a fixed
vector per scenario node plus a `visible`-gated block plus six scalars that change during an episode
(constants.DYN_FEATURES).  :func:`fold_gae` therefore precomputes, per scenario node, the 17x64
message tensor and the 64-wide root term for both parts, and per unique vulnerability the
16-wide first-layer projection of its embedding (the edge attribute is a *mean* of such embeddings,
and the first layer is linear before its ReLU).  All folding is done in float64 on the float32
parameters and rounded once.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

from . import constants as C
from .scenario import ScenarioTables

D_IN, D_H, D_E, D_NN = C.NODE_FEAT_DIM, C.NODE_EMB_DIM, C.VULN_EMB_DIM, C.NN_CHANNELS


@dataclass
class GaeWeights:
    """Parameters of the default encoder, float32, reference state-dict names in comments."""
    nn0_w: np.ndarray    # [16, 768]        layers.0.nn.0.weight
    nn0_b: np.ndarray    # [16]             layers.0.nn.0.bias
    nn2_w: np.ndarray    # [1576*64, 16]    layers.0.nn.2.weight
    nn2_b: np.ndarray    # [1576*64]        layers.0.nn.2.bias
    root_w: np.ndarray   # [64, 1576]       layers.0.lin.weight
    conv1_b: np.ndarray  # [64]             layers.0.bias
    bn1: Dict[str, np.ndarray]   # weight, bias, running_mean, running_var   layers.1.module.*
    gcn_w: np.ndarray    # [64, 64]         layers.3.lin.weight
    gcn_b: np.ndarray    # [64]             layers.3.bias
    bn2: Dict[str, np.ndarray]   # layers.4.module.*
    bn_eps: float = 1e-5

    @classmethod
    def random(cls, seed: int = 0) -> "GaeWeights":
        """Seeded random init with torch-like scales; BatchNorm running statistics are randomised
        so that the eval-mode folding is actually exercised."""
        r = np.random.default_rng(seed)
        u = lambda shape, b: r.uniform(-b, b, size=shape).astype(np.float32)  # noqa: E731
        bn = lambda: dict(weight=r.uniform(0.5, 1.5, D_H).astype(np.float32),  # noqa: E731
                          bias=r.uniform(-0.5, 0.5, D_H).astype(np.float32),
                          running_mean=r.normal(0, 0.5, D_H).astype(np.float32),
                          running_var=r.uniform(0.5, 2.0, D_H).astype(np.float32))
        return cls(nn0_w=u((D_NN, D_E), 1 / np.sqrt(D_E)), nn0_b=u((D_NN,), 1 / np.sqrt(D_E)),
                   nn2_w=u((D_IN * D_H, D_NN), 0.02), nn2_b=u((D_IN * D_H,), 0.02),
                   root_w=u((D_H, D_IN), 1 / np.sqrt(D_IN)), conv1_b=u((D_H,), 0.1), bn1=bn(),
                   gcn_w=u((D_H, D_H), np.sqrt(6 / (2 * D_H))), gcn_b=u((D_H,), 0.1), bn2=bn())

    @classmethod
    def from_state_dict(cls, sd) -> "GaeWeights":
        """Accepts the reference's ``encoder.pth`` state dict (torch tensors or arrays)."""
        g = lambda k: np.asarray(sd[k].detach().cpu().numpy() if hasattr(sd[k], "detach") else sd[k], dtype=np.float32)  # noqa: E731
        bn = lambda p: dict(weight=g(p + ".weight"), bias=g(p + ".bias"), running_mean=g(p + ".running_mean"),  # noqa: E731
                            running_var=g(p + ".running_var"))
        w = cls(nn0_w=g("layers.0.nn.0.weight"), nn0_b=g("layers.0.nn.0.bias"), nn2_w=g("layers.0.nn.2.weight"),
                nn2_b=g("layers.0.nn.2.bias"), root_w=g("layers.0.lin.weight"), conv1_b=g("layers.0.bias"),
                bn1=bn("layers.1.module"), gcn_w=g("layers.3.lin.weight"), gcn_b=g("layers.3.bias"),
                bn2=bn("layers.4.module"))
        if w.nn2_w.shape != (D_IN * D_H, D_NN) or w.gcn_w.shape != (D_H, D_H):
            raise ValueError("only the default GAE layer config (NNConv 1576->64 / 16, GCNConv 64->64) is supported")
        return w

    def state_dict(self):
        """Reference-named state dict of torch tensors (for loading into the reference GAEEncoder)."""
        import torch
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))  # noqa: E731
        sd = {"layers.0.nn.0.weight": t(self.nn0_w), "layers.0.nn.0.bias": t(self.nn0_b),
              "layers.0.nn.2.weight": t(self.nn2_w), "layers.0.nn.2.bias": t(self.nn2_b),
              "layers.0.lin.weight": t(self.root_w), "layers.0.bias": t(self.conv1_b),
              "layers.3.lin.weight": t(self.gcn_w), "layers.3.bias": t(self.gcn_b)}
        for p, bn in (("layers.1.module", self.bn1), ("layers.4.module", self.bn2)):
            for k, v in bn.items():
                sd[f"{p}.{k}"] = t(v)
            sd[f"{p}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
        return sd


@dataclass
class GaeTables:
    """Folded tables (host numpy, float32)."""
    node_static: np.ndarray   # [Nn, 2, 18, 64]: [.,0] = the node while NOT visible, [.,1] = while visible (always part + visible-gated part, added in float32); rows 0..16 = T (k'=0..15 hidden, 16 = bias slot), row 17 = root term
    dyn_proj: np.ndarray      # [6, 18, 64]  same layout for the six dynamic scalar features
    vuln_h: np.ndarray        # [Ug, 16]     W1 @ emb_u (no bias)
    nn0_b: np.ndarray         # [16]
    bn1_scale: np.ndarray     # [64]  gamma / sqrt(var + eps)
    bn1_shift: np.ndarray     # [64]  (conv1_b - mean) * scale + beta      (conv bias folded in)
    gcn_wt: np.ndarray        # [64(in), 64(out)]  transposed for coalesced reads
    bn2_scale: np.ndarray     # [64]
    bn2_shift: np.ndarray     # [64]  (gcn_b - mean) * scale + beta
    ev_proj: np.ndarray = None  # [30, 18, 64] the firewall-in | firewall-out | service-running feature columns (events defender)


def node_feature_static(spec_node, vuln_emb) -> tuple:
    """Static split of convert_node_info_to_observation (compressed:319-380): returns
    (x_always, x_visible) as float32-rounded float64 vectors of length 1576; the live feature vector is
    x_always + visible * x_visible + dynamic scalars."""
    xa = np.zeros(D_IN, dtype=np.float64)
    xv = np.zeros(D_IN, dtype=np.float64)
    M = C.MAX_SERVICES
    ports = [s.port for s in spec_node.services]

    def svc_index(port):            # cyberbattle_env.py:547-551 get_service_index (first match)
        return ports.index(port) if port in ports else -1
    for port, perm in spec_node.fw_in:
        i = svc_index(port)
        if i != -1 and i < M:
            xv[C.F_FW_IN + i] = perm
    for port, perm in spec_node.fw_out:
        i = svc_index(port)
        if i != -1 and i < M:
            xv[C.F_FW_OUT + i] = perm
    fv = np.zeros(D_E, dtype=np.float64)
    for i, s in enumerate(spec_node.services):
        if i >= M:
            break
        xv[C.F_SVC_RUNNING + i] = int(s.running)
        fv = fv + np.asarray(s.fv, dtype=np.float64)
    if len(spec_node.services) > 0:
        fv = fv / len(spec_node.services)
    xv[C.F_SVC_FV:C.F_SVC_FV + D_E] = fv
    xv[C.F_VISIBLE] = 1.0
    xa[C.F_REIMAGEABLE] = int(spec_node.reimageable)
    xa[C.F_VALUE] = spec_node.value
    xa[C.F_SLA] = spec_node.sla_weight
    mean = np.zeros(D_E, dtype=np.float64)
    if spec_node.vulns:
        for v in spec_node.vulns:
            mean = mean + np.asarray(vuln_emb[v.vid], dtype=np.float64)
        mean = mean / len(spec_node.vulns)
    xa[C.F_VULN_MEAN:C.F_VULN_MEAN + D_E] = mean
    # get_node_feature_vector casts the flattened list to float32 (compressed:201-202)
    return xa.astype(np.float32).astype(np.float64), xv.astype(np.float32).astype(np.float64)


def fold_gae(tables: ScenarioTables, w: GaeWeights, chunk: int = 64) -> GaeTables:
    Nn = int(tables.sc_node_off[-1])
    # P[f, k', c]: k' < 16 from nn2_w, k' = 16 from nn2_b, k' = 17 root weight
    P = np.empty((D_IN, 18, D_H), dtype=np.float64)
    P[:, :16, :] = w.nn2_w.astype(np.float64).reshape(D_IN, D_H, D_NN).transpose(0, 2, 1)
    P[:, 16, :] = w.nn2_b.astype(np.float64).reshape(D_IN, D_H)
    P[:, 17, :] = w.root_w.astype(np.float64).T
    Pf = P.reshape(D_IN, 18 * D_H)
    node_static = np.empty((Nn, 2, 18, D_H), dtype=np.float32)
    g = 0
    for spec in tables.specs:
        X = np.empty((2 * spec.num_nodes, D_IN), dtype=np.float64)
        for j, nd in enumerate(spec.nodes):
            X[2 * j], X[2 * j + 1] = node_feature_static(nd, spec.vuln_emb)
        out = (X @ Pf).reshape(spec.num_nodes, 2, 18, D_H)
        out = out.astype(np.float32)
        out[:, 1] = out[:, 0] + out[:, 1]       # the kernel picks one of the two variants by the node's `visible` bit
        node_static[g:g + spec.num_nodes] = out
        g += spec.num_nodes
    dyn = np.stack([P[f] for f in C.DYN_FEATURES]).astype(np.float32)
    ev_proj = np.ascontiguousarray(P[C.F_FW_IN:C.F_SVC_RUNNING + C.MAX_SERVICES].astype(np.float32))     # columns 0..29
    vuln_h = (tables.vemb64.astype(np.float32).astype(np.float64) @ w.nn0_w.astype(np.float64).T).astype(np.float32)

    def bn_fold(bn, conv_bias):
        scale = bn["weight"].astype(np.float64) / np.sqrt(bn["running_var"].astype(np.float64) + w.bn_eps)
        shift = (conv_bias.astype(np.float64) - bn["running_mean"].astype(np.float64)) * scale + bn["bias"].astype(np.float64)
        return scale.astype(np.float32), shift.astype(np.float32)
    s1, h1 = bn_fold(w.bn1, w.conv1_b)
    s2, h2 = bn_fold(w.bn2, w.gcn_b)
    return GaeTables(node_static=node_static, dyn_proj=dyn, vuln_h=vuln_h, nn0_b=w.nn0_b.astype(np.float32),
                     bn1_scale=s1, bn1_shift=h1, gcn_wt=np.ascontiguousarray(w.gcn_w.T.astype(np.float32)),
                     bn2_scale=s2, bn2_shift=h2, ev_proj=ev_proj)
