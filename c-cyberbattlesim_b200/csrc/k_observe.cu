// k_observe.cu — evolving visible graph, graph auto-encoder forward, action-table maintenance and reset.
//
// One warp per env.  Follows CyberBattleCompressedEnv.step after the inner transition
// (_env/cyberbattle_env_compressed.py:399-428): update_evolving_visible_graph_after_step (:465-484),
// add_edge_evolving_visible_graph (:214-246), encode (:249-306) with GAEEncoder.forward (gae/model.py:70-82),
// create_continuous_action_space (:487-523), and reset (:158-189 + _env/cyberbattle_env.py:134-186).
//
// The 1576-wide node feature vector is never materialised: ccbs_b200.gae.fold_gae pre-multiplies the static
// part of every scenario node with the NNConv edge-network output basis (17 x 64 per node) and the root
// weight (64 per node); here only the `visible` gate and six dynamic scalars are applied.  Per edge the
// message is a 17-term combination of those rows with [ReLU(W1 e + b1); 1].  GCNConv is a 64x64 projection
// (weights staged once per CTA in shared memory) followed by the degree-normalised neighbour sum.  Node
// embeddings of one env live in shared memory (<= 32 nodes) or in an L2-resident scratch slab (<= 128 nodes).
#include "cbs_device.cuh"
#include "philox.cuh"
#include "subset.cuh"
#include <type_traits>

namespace cbs {

#ifndef CBS_OBS_WARPS
#define CBS_OBS_WARPS 8
#endif
#ifndef CBS_OBS_MINB
#define CBS_OBS_MINB 1
#endif
constexpr int OBS_WARPS = CBS_OBS_WARPS;          // warps per CTA (one CTA per SM)
constexpr int SMEM_NODES = CBS_OBS_SMEM_NODES;    // graphs up to this many nodes keep their embeddings in shared memory

struct SharedWeights {
  float gcn[NODE_EMB * NODE_EMB];               // [in][out]
  float dyn[NUM_DYN * PROJ_ROWS * NODE_EMB];    // [d][row][c]  (reading it through L1 instead of staging it was measured slower)
  float bn1s[NODE_EMB], bn1h[NODE_EMB], bn2s[NODE_EMB], bn2h[NODE_EMB];
  float nn0b[NN_CH];
  double accum[N_ACCUM];   // this CTA's share of the episode sums (flushed to State::accum by its last warp)
  int warps_done;
};

struct WarpScratch {
  float* y;           // [n][64]
  float* g;           // [n][64]
  float* dinv;        // [ncap]
  float* ysm;         // this warp's shared-memory buffers (2 x [32][64])
  uint8_t* pos;       // [MAX_NODES] node id -> position in discovered order
  uint8_t* ord;       // [MAX_NODES] graph node ids by position (discovered order, then the interest node of *_node goals)
  uint8_t* dynb;      // [MAX_NODES] per position: visible | persistence<<1 | collected<<2 | exfiltrated<<3 | evasion<<4 | privilege<<5 | running<<7
  uint8_t* xst;       // [MAX_NODES] per position: MachineStatus value in the node's CACHED feature vector (differs from the live
                      //             status only under a defender: the vector is rebuilt only when the node is a successful target)
};

// the dynamic part of a node's feature vector (compressed:365-380), packed once per encode by lane-per-node
__device__ __forceinline__ uint8_t pack_dyn(const State& S, const Params& P, int b, int node) {
  const int w = node >> 5, sh = node & 31;
  auto bit = [&](int plane) -> uint32_t { return (ld_mask(S, P, plane, w, b) >> sh) & 1u; };
  const uint32_t priv = bit(M_PRIV_ROOT) ? 3u : bit(M_PRIV_USER);
  const uint32_t running = (bit(M_STOPPED) | (P.defender ? bit(M_IMAGING) : 0u)) ^ 1u;
  return (uint8_t)(bit(M_VISIBLE) | (bit(M_PERSISTENCE) << 1) | (bit(M_COLLECTED) << 2) | (bit(M_EXFILTRATED) << 3) |
                   (bit(M_EVASION) << 4) | (priv << 5) | (running << 7));
}
// MachineStatus in the cached feature vector: Stopped 0, Running 1, Imaging 2 (model.py:287-291)
__device__ __forceinline__ uint8_t pack_xstatus(const State& S, const Params& P, int b, int node) {
  if (bit_of(S, P, M_STOPPED, node, b)) return 0;
  return (P.defender && bit_of(S, P, M_X_IMAGING, node, b)) ? 2 : 1;
}
__device__ __forceinline__ void node_dyn(uint8_t d, uint8_t xstatus, float& vis, float x[NUM_DYN]) {
  vis = (float)(d & 1);
  x[0] = (float)((d >> 1) & 1);
  x[1] = (float)((d >> 2) & 1);
  x[2] = (float)((d >> 3) & 1);
  x[3] = (float)((d >> 4) & 1);
  x[4] = (float)((d >> 5) & 3);     // privilege level 0 / 1 / 3
  x[5] = (float)xstatus;
}

// ---- ExternalRandomEvents defender: the firewall-in / firewall-out / service-running columns of a VISIBLE node's cached feature
//      vector (compressed:365-370) are per-env state; node_static holds the scenario's initial values, so the columns that
//      differ are added (or removed) here.  `evx` / `init` = the node's { running, incoming BLOCK, outgoing BLOCK } bit sets as
//      cached in the graph / as compiled; only the first MAX_SERVICES slots are features.  Returns the correction of channel c
//      of folded row k. ----
__device__ __forceinline__ float ev_delta(const float* __restrict__ ev_proj, const uint16_t* evx, const uint16_t* init, int k, int c) {
  float acc = 0.f;
  const int col0[3] = {20, 0, 10};          // word 0 running -> F_SVC_RUNNING, word 1 incoming -> F_FW_IN, word 2 outgoing -> F_FW_OUT
#pragma unroll
  for (int w = 0; w < 3; ++w) {
    uint32_t diff = (uint32_t)(evx[w] ^ init[w]) & 0x3FFu;
    while (diff) {
      const int i = __ffs(diff) - 1;
      diff &= diff - 1;
      const float v = ev_proj[((size_t)(col0[w] + i) * PROJ_ROWS + k) * NODE_EMB + c];
      acc += ((evx[w] >> i) & 1) ? v : -v;
    }
  }
  return acc;
}

// ---- add_edge_evolving_visible_graph (compressed:214-246), mean aggregation, in the W1-projected space ----
__device__ void edge_update(const Tables& T, const Params& P, const State& S, int b, int lane) {
  const int4 sl = reinterpret_cast<const int4*>(S.sel)[b];
  const int s = sl.x, t = sl.y, u = sl.z;
  const int sc = scalar(S, P, S_SCENARIO, b);
  const int gv = T.uvuln_global[T.sc_uvuln_off[sc] + u];
  const float p = lane < NN_CH ? T.vuln_h[(size_t)gv * NN_CH + lane] : 0.f;
  int E = scalar(S, P, S_N_EDGES, b);
  uint8_t* es = S.edge_src + (size_t)b * P.ecap;
  uint8_t* ed = S.edge_dst + (size_t)b * P.ecap;
  int32_t* ec = S.edge_cnt + (size_t)b * P.ecap;
  int found = -1;
  for (int base = 0; base < E; base += 32) {
    const int e = base + lane;
    const bool hit = e < E && es[e] == s && ed[e] == t;
    const unsigned m = __ballot_sync(0xFFFFFFFFu, hit);
    if (m) { found = base + __ffs(m) - 1; break; }
  }
  if (found >= 0) {
    const int cnt = ec[found];
    float* sum = S.edge_sum + ((size_t)b * P.ecap + found) * NN_CH;
    if (lane < NN_CH) {
      const float v = (cnt == 0 ? 0.f : sum[lane]) + p;   // :226-228 (a wiped accumulator restarts)
      sum[lane] = v;
      S.edge_m[((size_t)b * P.ecap + found) * NN_CH + lane] = v / (float)(cnt + 1);
    }
    __syncwarp();
    if (lane == 0) ec[found] = cnt + 1;
  } else {
    if (E >= P.ecap) { if (lane == 0) atomicExch(S.errflag, 2); return; }
    for (int base = 0; base < E; base += 32) {             // :237 wipes every accumulator of this source
      const int e = base + lane;
      if (e < E && es[e] == s) ec[e] = 0;
    }
    if (lane < NN_CH) {
      S.edge_sum[((size_t)b * P.ecap + E) * NN_CH + lane] = p;
      S.edge_m[((size_t)b * P.ecap + E) * NN_CH + lane] = p;
    }
    if (lane == 0) { es[E] = (uint8_t)s; ed[E] = (uint8_t)t; ec[E] = 1; scalar(S, P, S_N_EDGES, b) = E + 1; }
  }
  __syncwarp();
}

// ---- encode (compressed:249-306) : returns node embeddings z in W.y (position-major) and writes S.obs ----
// EV: the ExternalRandomEvents defender is configured (compile-time: its feature corrections sit inside the unrolled hot loops,
// and the default instance must not carry their code)
template <bool EV>
__device__ void encode_env(const Tables& T, const Params& P, const State& S, const SharedWeights& SW, WarpScratch& W, int b,
                           int lane) {
  const int sc = scalar(S, P, S_SCENARIO, b);
  const int node_off = T.sc_node_off[sc];
  const int n_disc = scalar(S, P, S_N_DISC, b);
  const int E = scalar(S, P, S_N_EDGES, b);
  const uint8_t* disc_order = S.disc_order + (size_t)b * P.ncap;
  const int c0 = lane, c1 = lane + 32;
  constexpr int ROW = PROJ_ROWS * NODE_EMB;   // floats per (node, part)
  // *_node goals: once an encode has added the interest node to the live graph (compressed:254-256) it is part of
  // every later encode, discovered or not
  const int interest = is_node_goal(P) ? T.sc_interest[sc] : -1;
  const bool interest_known = interest >= 0 && bit_of(S, P, M_DISCOVERED, interest, b);
  const bool extra = interest >= 0 && !interest_known && (scalar(S, P, S_FLAGS, b) & FL_INTEREST_IN_GRAPH);
  const int n = n_disc + (extra ? 1 : 0);
  uint8_t* order = W.ord;

  for (int i = lane; i < n; i += 32) {
    const int node = i < n_disc ? disc_order[i] : interest;
    W.ord[i] = (uint8_t)node;
    W.pos[node] = (uint8_t)i;
    W.dinv[i] = 1.f;
    W.dynb[i] = pack_dyn(S, P, b, node);
    W.xst[i] = pack_xstatus(S, P, b, node);
  }
  __syncwarp();

  // root term x_i W_root (+ conv bias folded into bn1 shift); ROOT_TILE nodes per pass so that their table rows are in flight
  // together (a pass is one L2 round trip; four per pass made a 9-node graph wait three times)
  constexpr int ROOT_TILE = 12;
  for (int i0 = 0; i0 < n; i0 += ROOT_TILE) {
    float r0[ROOT_TILE], r1[ROOT_TILE];
#pragma unroll
    for (int j = 0; j < ROOT_TILE; ++j) {
      const int i = i0 + j;
      r0[j] = r1[j] = 0.f;
      if (i < n) {
        const float* ns = T.node_static + ((size_t)(node_off + order[i]) * 2 + (W.dynb[i] & 1)) * ROW + 17 * NODE_EMB;
        r0[j] = ns[c0];
        r1[j] = ns[c1];
      }
    }
#pragma unroll
    for (int j = 0; j < ROOT_TILE; ++j) {
      const int i = i0 + j;
      if (i >= n) break;
      float vis, x[NUM_DYN];
      node_dyn(W.dynb[i], W.xst[i], vis, x);
      float a0 = r0[j], a1 = r1[j];
#pragma unroll
      for (int d = 0; d < NUM_DYN; ++d) {
        a0 = fmaf(x[d], SW.dyn[(d * PROJ_ROWS + 17) * NODE_EMB + c0], a0);
        a1 = fmaf(x[d], SW.dyn[(d * PROJ_ROWS + 17) * NODE_EMB + c1], a1);
      }
      if (EV && vis != 0.f) {
        const uint16_t* evx = S.ev_x + ((size_t)b * P.ncap + order[i]) * 4;
        const uint16_t* ini = T.nd_ev_init + (size_t)(node_off + order[i]) * 4;
        a0 += ev_delta(T.ev_proj, evx, ini, 17, c0);
        a1 += ev_delta(T.ev_proj, evx, ini, 17, c1);
      }
      W.y[i * NODE_EMB + c0] = a0;
      W.y[i * NODE_EMB + c1] = a1;
    }
  }
  __syncwarp();

  // NNConv messages: y[dst] += [relu(m_e + b1); 1] . T_src
  const uint8_t* es = S.edge_src + (size_t)b * P.ecap;
  const uint8_t* ed = S.edge_dst + (size_t)b * P.ecap;
  // the next edge's end points and attribute are requested while the current edge is combined (one exposed round trip less per edge)
  int js_n = 0, jd_n = 0;
  float hm_n = 0.f;
  if (E > 0) {
    js_n = es[0]; jd_n = ed[0];
    if (lane < NN_CH) hm_n = S.edge_m[((size_t)b * P.ecap) * NN_CH + lane];
  }
  for (int e = 0; e < E; ++e) {
    const int js = js_n, jd = jd_n;
    const float hm = hm_n;
    if (e + 1 < E) {
      js_n = es[e + 1]; jd_n = ed[e + 1];
      if (lane < NN_CH) hm_n = S.edge_m[((size_t)b * P.ecap + e + 1) * NN_CH + lane];
    }
    const int is = W.pos[js], id = W.pos[jd];
    float hl = 0.f;
    if (lane < NN_CH) hl = fmaxf(hm + SW.nn0b[lane], 0.f);
    else if (lane == NN_CH) hl = 1.f;
    float vis, x[NUM_DYN];
    node_dyn(W.dynb[is], W.xst[is], vis, x);
    const float* ns = T.node_static + ((size_t)(node_off + js) * 2 + (vis != 0.f ? 1 : 0)) * ROW;   // visible / not-visible variant
    const bool evd = EV && vis != 0.f;
    const uint16_t* evx = evd ? S.ev_x + ((size_t)b * P.ncap + js) * 4 : nullptr;
    const uint16_t* ini = evd ? T.nd_ev_init + (size_t)(node_off + js) * 4 : nullptr;
    const bool ev_any = evd && (((evx[0] ^ ini[0]) | (evx[1] ^ ini[1]) | (evx[2] ^ ini[2])) & 0x3FF) != 0;
    float m0 = 0.f, m1 = 0.f;
#pragma unroll
    for (int k = 0; k < NN_CH + 1; ++k) {
      const float hk = __shfl_sync(0xFFFFFFFFu, hl, k);
      float t0 = ns[k * NODE_EMB + c0];
      float t1 = ns[k * NODE_EMB + c1];
#pragma unroll
      for (int d = 0; d < NUM_DYN; ++d) {
        t0 = fmaf(x[d], SW.dyn[(d * PROJ_ROWS + k) * NODE_EMB + c0], t0);
        t1 = fmaf(x[d], SW.dyn[(d * PROJ_ROWS + k) * NODE_EMB + c1], t1);
      }
      if (EV && ev_any) { t0 += ev_delta(T.ev_proj, evx, ini, k, c0); t1 += ev_delta(T.ev_proj, evx, ini, k, c1); }
      m0 = fmaf(hk, t0, m0);
      m1 = fmaf(hk, t1, m1);
    }
    W.y[id * NODE_EMB + c0] += m0;
    W.y[id * NODE_EMB + c1] += m1;
    if (lane == 0 && is != id) W.dinv[id] += 1.f;           // GCN in-degree (self loops are replaced, not counted)
    __syncwarp();
  }
  for (int i = lane; i < n; i += 32) W.dinv[i] = rsqrtf(W.dinv[i]);
  // BatchNorm(eval) + ReLU, then the GCN projection G = H1 Wg^T
  for (int i = 0; i < n; ++i) {
    W.y[i * NODE_EMB + c0] = fmaxf(fmaf(W.y[i * NODE_EMB + c0], SW.bn1s[c0], SW.bn1h[c0]), 0.f);
    W.y[i * NODE_EMB + c1] = fmaxf(fmaf(W.y[i * NODE_EMB + c1], SW.bn1s[c1], SW.bn1h[c1]), 0.f);
  }
  __syncwarp();
  // Several nodes per pass: each weight pair feeds two independent FMA chains per node, and a pass's broadcast reads of the
  // nodes' activations are all in flight together (two nodes per pass left the loop a chain of dependent shared-memory reads).
  // Passes of 8, then 4 / 2 / 1 for the remainder, so that no pass computes padding.  Every node's sum runs over k in the same
  // order whatever the tile, so equal inputs still give bit-identical outputs.
  auto gcn_pass = [&](auto tile_c, int i0) {
    constexpr int TILE = decltype(tile_c)::value;
    float acc0[TILE], acc1[TILE];
#pragma unroll
    for (int j = 0; j < TILE; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }
    const float* yrow = W.y + i0 * NODE_EMB;
#pragma unroll 4
    for (int k = 0; k < NODE_EMB; ++k) {
      const float w0 = SW.gcn[k * NODE_EMB + c0], w1 = SW.gcn[k * NODE_EMB + c1];
#pragma unroll
      for (int j = 0; j < TILE; ++j) {
        const float a = yrow[j * NODE_EMB + k];
        acc0[j] = fmaf(a, w0, acc0[j]);
        acc1[j] = fmaf(a, w1, acc1[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < TILE; ++j) {
      W.g[(i0 + j) * NODE_EMB + c0] = acc0[j];
      W.g[(i0 + j) * NODE_EMB + c1] = acc1[j];
    }
  };
  {
    int i0 = 0;
    for (; i0 + 8 <= n; i0 += 8) gcn_pass(std::integral_constant<int, 8>{}, i0);
    if (i0 + 4 <= n) { gcn_pass(std::integral_constant<int, 4>{}, i0); i0 += 4; }
    if (i0 + 2 <= n) { gcn_pass(std::integral_constant<int, 2>{}, i0); i0 += 2; }
    if (i0 < n) gcn_pass(std::integral_constant<int, 1>{}, i0);
  }
  __syncwarp();
  // normalised aggregation, edges first and the self loop last (the order PyG's add_remaining_self_loops +
  // scatter-add gives).  Products and sums are rounded separately (no FMA contraction): a pair of nodes that
  // attack each other then gets bit-identical embeddings, exactly as in the reference, and the exact ties this
  // creates in the action table resolve by insertion order instead of by rounding noise.
  for (int i = 0; i < n; ++i) {
    W.y[i * NODE_EMB + c0] = 0.f;
    W.y[i * NODE_EMB + c1] = 0.f;
  }
  __syncwarp();
  for (int e = 0; e < E; ++e) {
    const int is = W.pos[es[e]], id = W.pos[ed[e]];
    if (is == id) continue;
    const float w = __fmul_rn(W.dinv[is], W.dinv[id]);
    W.y[id * NODE_EMB + c0] = __fadd_rn(W.y[id * NODE_EMB + c0], __fmul_rn(w, W.g[is * NODE_EMB + c0]));
    W.y[id * NODE_EMB + c1] = __fadd_rn(W.y[id * NODE_EMB + c1], __fmul_rn(w, W.g[is * NODE_EMB + c1]));
    __syncwarp();
  }
  for (int i = 0; i < n; ++i) {
    const float d2 = __fmul_rn(W.dinv[i], W.dinv[i]);
    W.y[i * NODE_EMB + c0] = __fadd_rn(W.y[i * NODE_EMB + c0], __fmul_rn(d2, W.g[i * NODE_EMB + c0]));
    W.y[i * NODE_EMB + c1] = __fadd_rn(W.y[i * NODE_EMB + c1], __fmul_rn(d2, W.g[i * NODE_EMB + c1]));
  }
  __syncwarp();
  // BatchNorm + ReLU -> z ; readout over Running nodes (mean | max | min), compressed:266-298
  float s0 = 0.f, s1 = 0.f, mx0 = -INFINITY, mx1 = -INFINITY, mn0 = INFINITY, mn1 = INFINITY;
  int running = 0;
  for (int i = 0; i < n; ++i) {
    const float z0 = fmaxf(fmaf(W.y[i * NODE_EMB + c0], SW.bn2s[c0], SW.bn2h[c0]), 0.f);
    const float z1 = fmaxf(fmaf(W.y[i * NODE_EMB + c1], SW.bn2s[c1], SW.bn2h[c1]), 0.f);
    W.y[i * NODE_EMB + c0] = z0;
    W.y[i * NODE_EMB + c1] = z1;
    if (W.dynb[i] & 0x80) {
      ++running;
      s0 += z0; s1 += z1;
      mx0 = fmaxf(mx0, z0); mx1 = fmaxf(mx1, z1);
      mn0 = fminf(mn0, z0); mn1 = fminf(mn1, z1);
    }
  }
  float* obs = S.obs + (size_t)b * P.obs_dim;
  const bool none_running = running == 0;
  if (none_running) { s0 = s1 = mx0 = mx1 = mn0 = mn1 = 0.f; running = 1; }
  obs[c0] = s0 / (float)running;
  obs[c1] = s1 / (float)running;
  obs[NODE_EMB + c0] = mx0;
  obs[NODE_EMB + c1] = mx1;
  obs[2 * NODE_EMB + c0] = mn0;
  obs[2 * NODE_EMB + c1] = mn1;
  if (interest >= 0) {   // compressed:299-303: the interest node's own embedding, zeros if it is not in the graph / not Running
    const int ip = interest_known ? (int)W.pos[interest] : (extra ? n_disc : -1);
    const bool have = ip >= 0 && !none_running && (W.dynb[ip] & 0x80);
    obs[OBS_GRAPH + c0] = have ? W.y[ip * NODE_EMB + c0] : 0.f;
    obs[OBS_GRAPH + c1] = have ? W.y[ip * NODE_EMB + c1] : 0.f;
  }
  if (lane == 0) {
    obs[P.obs_dim - 2] = (float)n_disc;                     // create_discrete_features, compressed:309-316
    obs[P.obs_dim - 1] = (float)scalar(S, P, P.defender ? S_N_OWNED_RAW : S_N_OWNED, b);   // len(owned_nodes)
    scalar(S, P, S_N_ENCODES, b) += 1;
  }
  __syncwarp();
}

// ---- create_continuous_action_space (compressed:487-523): new (source,target) pairs freeze the embeddings
//      of THIS encode; they are stored once per table-growing encode in a snapshot slot ----
// `refresh`: precise_action_space_positions and this encode follows a step (compressed:419-422,498-506): pairs whose source
//      or target can reach the action's source or target node in the visible graph (nx.has_path on the DiGraph; a node
//      reaches itself) take the CURRENT embeddings — they move to this encode's snapshot slot while pair_epoch keeps their
//      place in the table's insertion order.
// With sample_subset_samples the pairs this build touches are also listed in the warp's scratch and handed to subset_update
// (subset.cuh), which keeps the explicit, sub-sampled table.
// SUBSET: sample_subset_samples is configured (compile-time: the default instances carry none of its code — with the balance step
// inlined behind a run-time test the default observe kernel ran 64 us instead of 43)
template <bool PRECISE, bool SUBSET>
__device__ void build_table(const Tables& T, const Params& P, const State& S, WarpScratch& W, int b, int lane, bool refresh_arg,
                            int skipped_builds = 0) {
  const bool refresh = PRECISE && refresh_arg;
  SubScratch sub;
  int n_newp = 0;
  if constexpr (SUBSET) {
    const bool in_smem = W.g == W.ysm + SMEM_NODES * NODE_EMB;   // the projection buffer of a small graph; free once the encode is done
    sub.carve(reinterpret_cast<unsigned char*>(W.ysm + SMEM_NODES * NODE_EMB), in_smem ? nullptr : S.sub_newp + (size_t)b * P.ncap * P.ncap);
  }
  const int node_off_bt = scalar(S, P, S_NODE_OFF, b);
  int new_rows = 0;
  uint32_t reach[MAX_NODES / 32] = {0u, 0u, 0u, 0u};
  if (PRECISE && refresh) {
    const int4 sl = reinterpret_cast<const int4*>(S.sel)[b];
    // compressed:419-427: around the action's nodes when its desired outcome re-encodes by itself (or precise_graph_encoding),
    // else (the re-encode is the defender's doing) around the nodes the defender changed in this step — possibly none
    const bool own_doing = sl.w == K_LATERAL || sl.w == K_DOS || sl.w == K_RECON || P.precise_graph || !P.defender;
    if (own_doing) {
      reach[sl.x >> 5] |= 1u << (sl.x & 31);
      reach[sl.y >> 5] |= 1u << (sl.y & 31);
    } else {
      for (int w = 0; w < P.words; ++w) reach[w] = S.changed[(size_t)b * P.words + w];
    }
    const int E = scalar(S, P, S_N_EDGES, b);
    const uint8_t* es = S.edge_src + (size_t)b * P.ecap;
    const uint8_t* ed = S.edge_dst + (size_t)b * P.ecap;
    for (bool changed = true; changed;) {       // backward closure over the edge list: at most one pass per graph node
      uint32_t add[MAX_NODES / 32] = {0u, 0u, 0u, 0u};
      for (int e = lane; e < E; e += 32) {
        const int s = es[e], t = ed[e];
        if (((reach[t >> 5] >> (t & 31)) & 1u) && !((reach[s >> 5] >> (s & 31)) & 1u)) add[s >> 5] |= 1u << (s & 31);
      }
      changed = false;
#pragma unroll
      for (int w = 0; w < MAX_NODES / 32; ++w) {
        const uint32_t a = __reduce_or_sync(0xFFFFFFFFu, add[w]);
        changed |= a != 0u;
        reach[w] |= a;
      }
    }
  }
  auto in_reach = [&](int n) { return ((reach[n >> 5] >> (n & 31)) & 1u) != 0u; };
  // sources: env.owned_nodes.  Under a defender that is the exact list (removals, duplicates: the dict comprehension of
  // compressed:491-492 keeps the first occurrence); otherwise the append-only list.
  const int n_disc = scalar(S, P, S_N_DISC, b), n_owned = scalar(S, P, P.defender ? S_N_OWNED_RAW : S_N_OWNED, b);
  const uint8_t* dorder = S.disc_order + (size_t)b * P.ncap;
  const uint8_t* oorder = P.defender ? S.owned_raw + (size_t)b * P.ocap : S.owned_order + (size_t)b * P.ncap;
  uint8_t* ps = S.pair_slot + (size_t)b * P.ncap * P.ncap;
  uint8_t* po = S.pair_opos + (size_t)b * P.ncap * P.ncap;
  uint8_t* pe = S.pair_epoch + (size_t)b * P.ncap * P.ncap;
  const int slot = scalar(S, P, S_N_SLOTS, b);
  bool any_new = false;
  uint32_t seen[MAX_NODES / 32] = {0u, 0u, 0u, 0u};
  for (int op = 0; op < n_owned; ++op) {
    const int s = oorder[op];
    if (P.defender) {
      if ((seen[s >> 5] >> (s & 31)) & 1u) continue;
      seen[s >> 5] |= 1u << (s & 31);
    }
    if (!(W.dynb[W.pos[s]] & 0x80)) continue;               // stopped sources add no rows (compressed:491-492)
    for (int base = 0; base < n_disc; base += 32) {
      const int dp = base + lane;
      bool fresh = false, fresh_refresh_only = false;
      if (dp < n_disc) {
        const int t = dorder[dp];
        const bool is_new = ps[s * P.ncap + t] == 0xFF;
        fresh_refresh_only = !is_new;
        fresh = (W.dynb[dp] & 0x80) && (is_new || (PRECISE && refresh && (in_reach(s) || in_reach(t))));
        if (fresh && slot < P.slots) {
          ps[s * P.ncap + t] = (uint8_t)slot;
          if (is_new) {
            if (P.defender) po[s * P.ncap + t] = (uint8_t)op;   // insertion order inside the slot (exact-tie order of the decode)
            if (PRECISE) pe[s * P.ncap + t] = (uint8_t)slot;
            const int g = node_off_bt + t;
            new_rows += T.nd_row_off[2 * g + 2] - T.nd_row_off[2 * g + (s == t ? 0 : 1)];
          }
        }
      }
      const unsigned fm = __ballot_sync(0xFFFFFFFFu, fresh);
      any_new |= fm != 0u;
      if (SUBSET && fm && slot < P.slots) {
        if (fresh) sub.newp[n_newp + __popc(fm & ((1u << lane) - 1u))] = (uint16_t)(dp | (op << 7) | (fresh_refresh_only ? 0x8000 : 0));
        n_newp += __popc(fm);
      }
    }
  }
  if (!any_new) {
    // the reference still balances (and draws from its generator) at the end of this create_continuous_action_space
    if constexpr (SUBSET) subset_update(T, P, S, sub, b, lane, slot, 0, oorder, n_owned, dorder, W.pos, skipped_builds);
    return;
  }
  if (slot >= P.slots) { if (lane == 0) atomicExch(S.errflag, 1); return; }
  new_rows = (int)warp_sum((float)new_rows);
  if (lane == 0 && !SUBSET) S.work_est[b] += new_rows;
  float* zh = S.z_hist + ((size_t)b * P.slots + slot) * P.ncap * NODE_EMB;
  float* zn = S.zn2_hist + ((size_t)b * P.slots + slot) * P.ncap;
  __half2* zh16 = reinterpret_cast<__half2*>(S.z16_hist + ((size_t)b * P.slots + slot) * P.ncap * NODE_EMB);
  for (int i = 0; i < n_disc; ++i) {
    const int node = dorder[i];
    if (!(W.dynb[i] & 0x80)) continue;                   // only Running nodes have embeddings (compressed:266-280)
    const float2 z = reinterpret_cast<const float2*>(W.y + i * NODE_EMB)[lane];   // channels 2*lane, 2*lane+1
    reinterpret_cast<float2*>(zh + node * NODE_EMB)[lane] = z;
    zh16[node * (NODE_EMB / 2) + lane] = __floats2half2_rn(z.x, z.y);
    if (fmaxf(fabsf(z.x), fabsf(z.y)) > 65504.f) atomicExch(S.errflag, 8);     // beyond half precision: the decode scan reads these copies
    const float n2 = warp_sum(z.x * z.x + z.y * z.y);
    if (lane == 0) zn[node] = n2;
  }
  if (lane == 0) scalar(S, P, S_N_SLOTS, b) = slot + 1;
  __syncwarp();
  if constexpr (SUBSET) subset_update(T, P, S, sub, b, lane, slot, n_newp, oorder, n_owned, dorder, W.pos, skipped_builds);
}

// ---- get_statistics (cyberbattle_env.py:517-524) + episode accumulators ----
__constant__ int kFinishPlanes[11] = {S_N_DISC, S_FLAGS, S_OWNABLE, S_DISCOVERABLE, S_DISRUPTABLE, S_DISC_AMOUNT,
                                      S_DISCOVERABLE_AMOUNT, S_STEPCOUNT, S_N_REIMAGED, S_SCENARIO, S_EPISODES};
// Everything the statistics need is fetched by ONE load per lane (mask words on lanes 0-15, scalars on lanes 16-26, the
// episode return on lane 31) together with the observation to be kept as terminal observation, then exchanged with
// shuffles: a serial version of this function (a dozen dependent round trips) was the largest single stall source of
// the kernel.  The episode sums go to per-CTA shared-memory accumulators, flushed once per CTA.
__device__ void finish_episode(const Tables& T, const Params& P, const State& S, double* __restrict__ cta_accum, int b, int lane) {
  constexpr int MAX_OBS_PER_LANE = (OBS_DIM + NODE_EMB + 31) / 32;
  float ob[MAX_OBS_PER_LANE];
#pragma unroll
  for (int i = 0; i < MAX_OBS_PER_LANE; ++i) {
    const int idx = lane + 32 * i;
    ob[i] = idx < P.obs_dim ? S.obs[(size_t)b * P.obs_dim + idx] : 0.f;
  }
  uint32_t mv = 0;
  int sv = 0;
  double ret = 0.0;
  if (lane < 16) {
    const int w = lane & 3, grp = lane >> 2;
    const int plane = grp == 0 ? M_OWNED : (grp == 1 ? M_DISCOVERED : (grp == 2 ? M_STOPPED : M_IMAGING));
    if (w < P.words && (grp < 3 || P.defender)) mv = ld_mask(S, P, plane, w, b);
  } else if (lane < 27) {
    sv = scalar(S, P, kFinishPlanes[lane - 16], b);
  } else if (lane == 31) ret = *reinterpret_cast<const double*>(&scalar(S, P, S_EP_RETURN, b));
  int owned = 0, disrupted = 0, running = 0;
#pragma unroll
  for (int w = 0; w < MAX_NODES / 32; ++w) {
    const uint32_t own = __shfl_sync(0xFFFFFFFFu, mv, w), disc = __shfl_sync(0xFFFFFFFFu, mv, 4 + w);
    const uint32_t stop = __shfl_sync(0xFFFFFFFFu, mv, 8 + w), img = __shfl_sync(0xFFFFFFFFu, mv, 12 + w);
    owned += __popc(own);
    disrupted += __popc(disc & stop);
    running += __popc(disc & ~stop & ~img);
  }
  auto sc_of = [&](int k) { return __shfl_sync(0xFFFFFFFFu, sv, 16 + k); };
  const int n_disc = sc_of(0), flags = sc_of(1), stepcount = sc_of(7), sc = sc_of(9), episodes = sc_of(10);
  const int reason = (flags >> FL_REASON_SHIFT) & 3;
  const int N = T.sc_num_nodes[sc];
  double st[14];
  st[0] = owned; st[1] = n_disc; st[2] = N - n_disc; st[3] = disrupted; st[4] = N;
  st[5] = sc_of(2); st[6] = sc_of(3); st[7] = sc_of(4);
  st[8] = (double)running / (double)n_disc;
  st[10] = P.defender ? sc_of(8) : 0;                  // num_events (cyberbattle_env.py:419)
  st[9] = P.defender == 1 ? st[10] : 0.0;              // len(overall_reimaged): only the re-imaging defender fills it (:420-422)
  st[11] = sc_of(5); st[12] = sc_of(6);
  // attacker_goal_reached() at the end of an episode: the goal test comes first in the end-of-step chain
  // (cyberbattle_env.py:343-370), so it holds exactly when the episode's last step reported reason 1
  st[13] = reason == 1 ? 1.0 : 0.0;
  const double ep_ret = __shfl_sync(0xFFFFFFFFu, ret, 31);
  double v = 0.0;
  if (lane < 14) {
#pragma unroll
    for (int i = 0; i < 14; ++i) if (i == lane) v = st[i];
    S.last_stats[(size_t)b * 14 + lane] = v;
    atomicAdd(&cta_accum[A_STAT0 + lane], v);
  } else if (lane == 14) atomicAdd(&cta_accum[A_EPISODES], 1.0);
  else if (lane == 15) atomicAdd(&cta_accum[A_RETURN], ep_ret);
  else if (lane == 16) atomicAdd(&cta_accum[A_LENGTH], (double)stepcount);
  else if (lane == 17 && reason >= 1) atomicAdd(&cta_accum[A_WINS + reason - 1], 1.0);
#pragma unroll
  for (int i = 0; i < MAX_OBS_PER_LANE; ++i) {
    const int idx = lane + 32 * i;
    if (idx < P.obs_dim) S.term_obs[(size_t)b * P.obs_dim + idx] = ob[i];
  }
  if (lane == 0) scalar(S, P, S_EPISODES, b) = episodes + 1;
  __syncwarp();
}

// ---- reset (cyberbattle_env.py:134-186,189-296 ; compressed:158-189 ; switch.py:151-167,218-220) ----
__device__ int2 reset_env(const Tables& T, const Params& P, const State& S, int b, int lane) {
  const uint64_t genv = (uint64_t)(P.global_env_offset + b);
  const int episodes = scalar(S, P, S_EPISODES, b);
  int sc = scalar(S, P, S_SCENARIO, b);
  if (P.switch_interval >= 0 && (episodes + 1) % (P.switch_interval + 1) == 0) {   // _check_switch (switch.py:218-220); < 0 = never
    const Philox4 r = philox4x32_10(P.seed, genv, (uint32_t)episodes, 2u);
    sc = (int)(((uint64_t)r.x * (uint64_t)T.num_scenarios) >> 32);
  }
  int starter;
  if (S.starter_queue) starter = S.starter_queue[(size_t)b * P.qlen + (episodes % P.qlen)];
  else {
    const int f0 = T.sc_feasible_off[sc], f1 = T.sc_feasible_off[sc + 1];
    const Philox4 r = philox4x32_10(P.seed, genv, (uint32_t)episodes, 1u);
    starter = (f1 > f0) ? T.feasible_starters[f0 + (int)(((uint64_t)r.x * (uint64_t)(f1 - f0)) >> 32)] : 0;
  }
  const int g = T.sc_node_off[sc] + starter;
  const int laa = T.nd_level_at_access[g];
  const uint32_t sbit = 1u << (starter & 31);
  const int sw = starter >> 5;
  for (int i = lane; i < N_MASKS * P.words; i += 32) {
    const int plane = i / P.words, w = i % P.words;
    uint32_t v = 0;
    if (plane == M_HAS_DATA) v = T.sc_init_has_data[sc * P.words + w];
    else if (plane == M_VISIBLE) v = T.sc_init_visible[sc * P.words + w];
    else if (w == sw && (plane == M_OWNED || plane == M_DISCOVERED || (plane == M_PRIV_USER && laa >= 1) ||
                         (plane == M_PRIV_ROOT && laa == 3) || (plane == M_EVER_OWNED && P.defender))) v = sbit;
    S.masks[(size_t)b * P.mpitch + i] = v;
  }
  uint32_t* ps = reinterpret_cast<uint32_t*>(S.pair_slot + (size_t)b * P.ncap * P.ncap);
  for (int i = lane; i < P.ncap * P.ncap / 4; i += 32) ps[i] = 0xFFFFFFFFu;
  if (P.defender == 2) {      // pristine services / firewall rules (the reference's deepcopy of the scenario, cyberbattle_env.py:145)
    const int N = T.sc_num_nodes[sc];
    const uint2* ini = reinterpret_cast<const uint2*>(T.nd_ev_init + (size_t)T.sc_node_off[sc] * 4);
    uint2* cur = reinterpret_cast<uint2*>(S.ev_cur + (size_t)b * P.ncap * 4);
    uint2* evx = reinterpret_cast<uint2*>(S.ev_x + (size_t)b * P.ncap * 4);
    for (int i = lane; i < N; i += 32) { cur[i] = ini[i]; evx[i] = ini[i]; }
  }
  if (P.subset_k) {   // empty table; the lifetime balance counter ([13]) goes on
    int32_t* meta = S.sub_meta + (size_t)b * SUB_META;
    if (lane < 13) meta[lane] = (lane == 11 || lane == 12) ? -1 : 0;
    if (P.precise_positions) {
      uint32_t* al = S.sub_alive + (size_t)b * P.ncap * P.ncap * (SUB_MAX_ROWS_PER_PAIR / 32);
      for (int i = lane; i < P.ncap * P.ncap * (SUB_MAX_ROWS_PER_PAIR / 32); i += 32) al[i] = 0u;
    }
  }
  if (lane == 0) {
    scalar(S, P, S_SCENARIO, b) = sc;
    scalar(S, P, S_NODE_OFF, b) = T.sc_node_off[sc];
    scalar(S, P, S_STARTER, b) = starter;
    scalar(S, P, S_SCST, b) = (sc << 8) | starter;      // what the transition reads instead of the constant sector
    scalar(S, P, S_STEPCOUNT, b) = 0;
    scalar(S, P, S_NUM_ITER, b) = 0;
    scalar(S, P, S_N_DISC, b) = 1;
    scalar(S, P, S_N_OWNED, b) = 1;
    scalar(S, P, S_DISC_AMOUNT, b) = 0;
    const int own = T.nd_ownable[g], dis = T.nd_discoverable[g], dsr = T.nd_disruptable[g];
    scalar(S, P, S_OWNABLE, b) = own;
    scalar(S, P, S_DISCOVERABLE, b) = dis;
    scalar(S, P, S_DISRUPTABLE, b) = dsr;
    const int bg = base_goal(P);
    scalar(S, P, S_PROP_NODES, b) = bg == GOAL_CONTROL ? own : (bg == GOAL_DISCOVERY ? dis : dsr);
    scalar(S, P, S_DISCOVERABLE_AMOUNT, b) = T.sc_discoverable_amount[sc];
    scalar(S, P, S_N_SLOTS, b) = 0;
    S.work_est[b] = 0;
    scalar(S, P, S_N_EDGES, b) = 0;
    scalar(S, P, S_FLAGS, b) = 0;
    scalar(S, P, S_OUTCOME, b) = -1;
    S.disc_order[(size_t)b * P.ncap] = (uint8_t)starter;
    S.owned_order[(size_t)b * P.ncap] = (uint8_t)starter;
    if (P.defender) {
      S.owned_raw[(size_t)b * P.ocap] = (uint8_t)starter;
      scalar(S, P, S_N_OWNED_RAW, b) = 1;
      scalar(S, P, S_N_REIMAGED, b) = 0;
    }
    *reinterpret_cast<double*>(&scalar(S, P, S_EP_RETURN, b)) = 0.0;
  }
  __syncwarp();
  return make_int2(sc, starter);
}

// ---- first observation of an episode.  The visible graph is the starter alone, so encode + table depend on (scenario,
//      starter) only: the first reset from a pair runs the generic code and publishes its result (observation, the
//      starter's embedding and squared norm); every later one copies the entry and writes the one-pair table directly.
//      The entry holds what the generic path computed, bit for bit, so results do not depend on who filled it. ----
__device__ bool reset_from_cache(const Tables& T, const Params& P, const State& S, int b, int lane, int sc, int starter) {
  const int g = T.sc_node_off[sc] + starter;
  int hit = 0;
  if (lane == 0) hit = *reinterpret_cast<volatile const int32_t*>(S.reset_cache_flag + g);
  hit = __shfl_sync(0xFFFFFFFFu, hit, 0);
  if (!hit) return false;
  const float* rc = S.reset_cache + (size_t)g * RC_PITCH;
  for (int idx = lane; idx < P.obs_dim; idx += 32) S.obs[(size_t)b * P.obs_dim + idx] = __ldcg(rc + idx);
  const float2 z = __ldcg(reinterpret_cast<const float2*>(rc + RC_Z) + lane);
  const size_t slab = (size_t)b * P.slots * P.ncap;         // snapshot slot 0 of this env
  reinterpret_cast<float2*>(S.z_hist + (slab + starter) * NODE_EMB)[lane] = z;
  reinterpret_cast<__half2*>(S.z16_hist + (slab + starter) * NODE_EMB)[lane] = __floats2half2_rn(z.x, z.y);
  if (lane == 0) {
    S.zn2_hist[slab + starter] = __ldcg(rc + RC_N2);
    const size_t pp = (size_t)b * P.ncap * P.ncap + (size_t)starter * P.ncap + starter;   // the (starter, starter) pair
    S.pair_slot[pp] = 0;
    if (P.defender) S.pair_opos[pp] = 0;
    if (P.precise_positions) S.pair_epoch[pp] = 0;
    S.work_est[b] = T.nd_row_off[2 * g + 2] - T.nd_row_off[2 * g];      // local + remote candidate rows of the starter
    scalar(S, P, S_N_SLOTS, b) = 1;
    scalar(S, P, S_N_ENCODES, b) += 1;
  }
  __syncwarp();
  return true;
}
__device__ void reset_cache_publish(const Tables& T, const Params& P, const State& S, const WarpScratch& W, int b, int lane, int sc,
                                    int starter) {
  const int g = T.sc_node_off[sc] + starter;
  float* rc = S.reset_cache + (size_t)g * RC_PITCH;
  __syncwarp();
  for (int idx = lane; idx < P.obs_dim; idx += 32) rc[idx] = S.obs[(size_t)b * P.obs_dim + idx];
  reinterpret_cast<float2*>(rc + RC_Z)[lane] = reinterpret_cast<const float2*>(W.y)[lane];      // position 0 = the starter
  if (lane == 0) rc[RC_N2] = S.zn2_hist[(size_t)b * P.slots * P.ncap + starter];
  __threadfence();
  __syncwarp();
  if (lane == 0) atomicExch(S.reset_cache_flag + g, 1);
}

// BIG_GRAPHS: scenarios with more than 32 nodes exist, so an env's graph may outgrow the shared-memory buffers
template <bool BIG_GRAPHS, bool SUBSET>
__global__ void __launch_bounds__(OBS_WARPS * 32, CBS_OBS_MINB) observe_kernel(Tables T, Params P, State S,
                                                                const uint8_t* __restrict__ reset_mask, int mode,
                                                                long long* __restrict__ trace) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SharedWeights& SW = *reinterpret_cast<SharedWeights*>(smem_raw);
  {   // stage the weights with 128-bit loads, all requests in flight before the first store
    constexpr int NT = OBS_WARPS * 32;
    constexpr int G4 = NODE_EMB * NODE_EMB / 4, D4 = NUM_DYN * PROJ_ROWS * NODE_EMB / 4;   // 1024, 1728 float4
    const float4* gsrc = reinterpret_cast<const float4*>(T.gcn_wt);
    const float4* dsrc = reinterpret_cast<const float4*>(T.dyn_proj);
    float4* gdst = reinterpret_cast<float4*>(SW.gcn);
    float4* ddst = reinterpret_cast<float4*>(SW.dyn);
    float4 tg[(G4 + NT - 1) / NT], td[(D4 + NT - 1) / NT];
#pragma unroll
    for (int i = 0; i < (G4 + NT - 1) / NT; ++i) { const int j = threadIdx.x + i * NT; tg[i] = j < G4 ? gsrc[j] : make_float4(0, 0, 0, 0); }
#pragma unroll
    for (int i = 0; i < (D4 + NT - 1) / NT; ++i) { const int j = threadIdx.x + i * NT; td[i] = j < D4 ? dsrc[j] : make_float4(0, 0, 0, 0); }
#pragma unroll
    for (int i = 0; i < (G4 + NT - 1) / NT; ++i) { const int j = threadIdx.x + i * NT; if (j < G4) gdst[j] = tg[i]; }
#pragma unroll
    for (int i = 0; i < (D4 + NT - 1) / NT; ++i) { const int j = threadIdx.x + i * NT; if (j < D4) ddst[j] = td[i]; }
  }
  if (threadIdx.x < NODE_EMB) {
    SW.bn1s[threadIdx.x] = T.bn1_scale[threadIdx.x];
    SW.bn1h[threadIdx.x] = T.bn1_shift[threadIdx.x];
    SW.bn2s[threadIdx.x] = T.bn2_scale[threadIdx.x];
    SW.bn2h[threadIdx.x] = T.bn2_shift[threadIdx.x];
  }
  if (threadIdx.x < NN_CH) SW.nn0b[threadIdx.x] = T.nn0_b[threadIdx.x];
  if (threadIdx.x < N_ACCUM) SW.accum[threadIdx.x] = 0.0;
  if (threadIdx.x == 0) SW.warps_done = 0;
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * OBS_WARPS + warp, total_warps = gridDim.x * OBS_WARPS;

  // per-warp scratch
  unsigned char* wbase = smem_raw + sizeof(SharedWeights);
  constexpr size_t kWarpBytesSmem = (size_t)2 * SMEM_NODES * NODE_EMB * 4 + MAX_NODES * 4 + 4 * MAX_NODES;
  WarpScratch W;
  {
    unsigned char* p = wbase + (size_t)warp * kWarpBytesSmem;
    W.ysm = reinterpret_cast<float*>(p);
    W.y = W.ysm;
    W.g = W.y + SMEM_NODES * NODE_EMB;
    W.dinv = W.g + SMEM_NODES * NODE_EMB;
    W.pos = reinterpret_cast<uint8_t*>(W.dinv + MAX_NODES);
    W.dynb = W.pos + MAX_NODES;
    W.ord = W.dynb + MAX_NODES;
    W.xst = W.ord + MAX_NODES;
  }

  // mode 1 (cbs_reset): every env, optionally masked.  mode 0 (after a transition): only the envs the
  // transition kernel put on the worklist — the others keep their cached observation untouched.
  // mode 0: the OBS_CLASSES class lists are consumed as one sequence, heaviest class first
  int cls_end[OBS_CLASSES];                     // exclusive end of every class in the concatenated item sequence
  int count = 0;
#pragma unroll
  for (int k = 0; k < OBS_CLASSES; ++k) { count += (mode == 1) ? (k == 0 ? P.B : 0) : S.work_ctr[4 + k]; cls_end[k] = count; }
  int i = gw;                                   // mode 1: static stride.  mode 0: items are claimed one at a time
  for (;;) {
    if (mode == 0) {
      if (lane == 0) i = atomicAdd(&S.work_ctr[2], 1);
      i = __shfl_sync(0xFFFFFFFFu, i, 0);
    }
    if (i >= count) break;
    int b = i;
    if (mode == 0) {
      int k = 0, first = 0;
#pragma unroll
      for (int q = 0; q < OBS_CLASSES - 1; ++q) if (i >= cls_end[q]) { k = q + 1; first = cls_end[q]; }
      b = S.worklist[(size_t)k * P.B + (i - first)];
    }
    // node-embedding buffers: shared memory while the env's visible graph has <= 32 nodes (always, when the
    // scenarios have <= 32 nodes), its slab of the L2-resident scratch otherwise.  A reset shrinks the graph to 1 node.
    if (BIG_GRAPHS) {
      const bool small = scalar(S, P, S_N_DISC, b) <= SMEM_NODES;
      W.y = small ? W.ysm : S.scratch + (size_t)b * 2 * P.ncap * NODE_EMB;
      W.g = small ? W.ysm + SMEM_NODES * NODE_EMB : W.y + (size_t)P.ncap * NODE_EMB;
    }
    const int flags = scalar(S, P, S_FLAGS, b);
    long long t_item = 0, t_ph[5] = {0, 0, 0, 0, 0};
    if (trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_item));
    bool do_reset = false;
    if (mode == 1) {
      do_reset = reset_mask ? (reset_mask[b] != 0) : true;
    } else if (!(flags & FL_NEEDS_RESET)) {
      int keep = flags & ~(FL_ADD_EDGE | FL_REENCODE | FL_FINISHED_THIS_STEP);
      if (flags & FL_ADD_EDGE) edge_update(T, P, S, b, lane);
      if (flags & FL_REENCODE) {
        if (P.defender == 2) encode_env<true>(T, P, S, SW, W, b, lane); else encode_env<false>(T, P, S, SW, W, b, lane);
        if (trace) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_ph[0])); }
        // (an env that finished in this step drops its action table at the reset that follows: only the sub-sampled table is
        // still maintained, because every create_continuous_action_space advances the env's balance counter)
        if (!(flags & FL_FINISHED_THIS_STEP) || SUBSET) {
          const int skipped = SUBSET ? (flags >> FL_PENDING_SHIFT) & 0xFFFF : 0;
          if (P.precise_positions) build_table<true, SUBSET>(T, P, S, W, b, lane, true, skipped);
          else build_table<false, SUBSET>(T, P, S, W, b, lane, false, skipped);
          keep &= 0xFFFF;
        }
        if (trace) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_ph[1])); }
        keep &= ~FL_DIRTY;
        if (is_node_goal(P)) keep |= FL_INTEREST_IN_GRAPH;
      }
      if (flags & FL_FINISHED_THIS_STEP) {
        finish_episode(T, P, S, SW.accum, b, lane);
        if (trace) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_ph[2])); }
        if (P.auto_reset) do_reset = true;
      }
      __syncwarp();
      if (lane == 0) scalar(S, P, S_FLAGS, b) = keep;
      __syncwarp();
    }
    if (do_reset) {
      if constexpr (SUBSET) {   // balance calls of skipped builds that no later build picked up
        const int left = (mode == 1 ? flags : scalar(S, P, S_FLAGS, b)) >> FL_PENDING_SHIFT & 0xFFFF;
        if (left && lane == 0) S.sub_meta[(size_t)b * SUB_META + 13] += left;
        __syncwarp();
      }
      const int2 ss = reset_env(T, P, S, b, lane);
      if (trace) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_ph[3])); }
      if (!reset_from_cache(T, P, S, b, lane, ss.x, ss.y)) {
        W.y = W.ysm;                               // a fresh episode's graph is one node
        W.g = W.ysm + SMEM_NODES * NODE_EMB;
        if (P.defender == 2) encode_env<true>(T, P, S, SW, W, b, lane); else encode_env<false>(T, P, S, SW, W, b, lane);
        if (P.precise_positions) build_table<true, SUBSET>(T, P, S, W, b, lane, false);
        else build_table<false, SUBSET>(T, P, S, W, b, lane, false);
        reset_cache_publish(T, P, S, W, b, lane, ss.x, ss.y);
      } else if constexpr (SUBSET) {
        // the cached first observation is per (scenario, starter); the sub-sampled table is per env: its one pair (starter,
        // starter) goes through the balance step like every build of the reference
        SubScratch sub;
        sub.carve(reinterpret_cast<unsigned char*>(W.ysm + SMEM_NODES * NODE_EMB), nullptr);
        if (lane == 0) { sub.newp[0] = 0; W.pos[ss.y] = 0; }
        __syncwarp();
        subset_update(T, P, S, sub, b, lane, 0, 1, P.defender ? S.owned_raw + (size_t)b * P.ocap : S.owned_order + (size_t)b * P.ncap, 1,
                      S.disc_order + (size_t)b * P.ncap, W.pos, 0);
      }
      if (trace) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_ph[4])); }
      // *_node goals: that first encode put the interest node into the live graph, so the next re-encode differs even
      // if nothing else changes
      if (is_node_goal(P) && lane == 0) scalar(S, P, S_FLAGS, b) = FL_DIRTY | FL_INTEREST_IN_GRAPH;
    }
    __syncwarp();
    if (trace && lane == 0) {   // debug: per item {start ns, duration ns, flags at entry, nodes << 16 | edges}
      long long t_end;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
      long long* tr = trace + (size_t)b * 9;
      tr[0] = t_item; tr[1] = t_end - t_item; tr[2] = flags;
      tr[3] = ((long long)scalar(S, P, S_N_DISC, b) << 16) | scalar(S, P, S_N_EDGES, b);
#pragma unroll
      for (int k = 0; k < 5; ++k) tr[4 + k] = t_ph[k] ? t_ph[k] - t_item : 0;   // phase end times relative to the item start
    }
    if (mode == 1) i += total_warps;
  }
  // the CTA's last warp adds the CTA's episode sums to the global accumulators (one atomic per slot and CTA)
  __syncwarp();
  int last_in_cta = 0;
  if (lane == 0) { __threadfence_block(); last_in_cta = atomicAdd(&SW.warps_done, 1) == OBS_WARPS - 1; }
  last_in_cta = __shfl_sync(0xFFFFFFFFu, last_in_cta, 0);
  if (last_in_cta && lane < N_ACCUM) {
    const double v = SW.accum[lane];
    if (v != 0.0) atomicAdd(&S.accum[lane], v);
  }
  if (mode == 0 && lane == 0) {   // the last warp to run dry clears the counters for the next transition
    __threadfence();
    if (atomicAdd(&S.work_ctr[1], 1) == total_warps - 1) {
      S.work_ctr[1] = 0; S.work_ctr[2] = 0;
#pragma unroll
      for (int k = 0; k < OBS_CLASSES; ++k) S.work_ctr[4 + k] = 0;
      __threadfence();
    }
  }
}

long long* g_obs_trace = nullptr;   // debug: per-env {start ns, duration ns, flags, nodes << 16 | edges} of the env's last item

size_t observe_smem_bytes() {
  const size_t per_warp = (size_t)2 * SMEM_NODES * NODE_EMB * 4 + MAX_NODES * 4 + 4 * MAX_NODES;
  return sizeof(SharedWeights) + OBS_WARPS * per_warp;
}

cudaError_t launch_observe(const Tables& T, const Params& P, const State& S, const uint8_t* reset_mask, int mode,
                           int num_sms, cudaStream_t stream) {
  const bool big = P.ncap > SMEM_NODES;
  const size_t smem = observe_smem_bytes();
  // persistent grid: one CTA of 8 warps per SM (181 KB of shared memory: weights + per-warp node buffers)
  int grid = num_sms;
  const int need = (P.B + OBS_WARPS - 1) / OBS_WARPS;
  if (grid > need) grid = need;
  using KernelFn = void (*)(Tables, Params, State, const uint8_t*, int, long long*);
  const KernelFn kernels[4] = {observe_kernel<false, false>, observe_kernel<true, false>, observe_kernel<false, true>, observe_kernel<true, true>};
  const int which = (big ? 1 : 0) | (P.subset_k ? 2 : 0);
  static bool attr_set[4] = {false, false, false, false};
  if (!attr_set[which]) {
    cudaError_t e = cudaFuncSetAttribute(kernels[which], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set[which] = true;
  }
  kernels[which]<<<grid, OBS_WARPS * 32, smem, stream>>>(T, P, S, reset_mask, mode, g_obs_trace);
  return cudaGetLastError();
}

}  // namespace cbs
