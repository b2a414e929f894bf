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
#include <cstdio>
#include <cstdlib>

namespace cbs {

#ifndef CBS_OBS_WARPS
#define CBS_OBS_WARPS 8
#endif
#ifndef CBS_OBS_MINB
#define CBS_OBS_MINB 1
#endif
#ifndef CBS_OBS_BOUND
#define CBS_OBS_BOUND (CBS_OBS_WARPS * 32)   // threads the register budget is sized for (experiments: > the CTA size caps the registers)
#endif
constexpr int OBS_WARPS = CBS_OBS_WARPS;          // warps per CTA (one CTA per SM)
constexpr int SMEM_NODES = CBS_OBS_SMEM_NODES;    // graphs up to this many nodes keep their embeddings in shared memory

struct SharedWeights {
  float gcn[NODE_EMB * NODE_EMB];               // [in][out]
  float dyn[NUM_DYN * PROJ_ROWS * NODE_EMB];    // [d][row][c]  (reading it through L1 instead of staging it, to fit 12 warps per SM, was measured slower: 57 us against 37)
  float bn1s[NODE_EMB], bn1h[NODE_EMB], bn2s[NODE_EMB], bn2h[NODE_EMB];
  float nn0b[NN_CH];
  double accum[N_ACCUM];   // this CTA's share of the episode sums (flushed to State::accum by its last warp)
  unsigned long long wbar; // mbarrier: the staged weights have landed (one arrival per thread, when its cp.async copies complete)
  int warps_done;
};

constexpr size_t kSharedWeightsBytes = (sizeof(SharedWeights) + 15) & ~(size_t)15;   // the warps' buffers are read 128 bits at a time

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
// Everything that depends on the env index alone is requested first (the selected action, the counters, the first 32 edges and
// their counts): in program order — scenario -> vulnerability offset -> global vulnerability -> its hidden vector, then the edge
// count -> the list -> the hit's accumulator — the function was six dependent round trips, 1.8 us per item and a tenth of the
// kernel's warp time.
__device__ void edge_update(const Tables& T, const Params& P, const State& S, int b, int lane) {
  const int4 sl = reinterpret_cast<const int4*>(S.sel)[b];
  const int s = sl.x, t = sl.y, u = sl.z;
  const int uvoff = scalar(S, P, S_UVULN_OFF, b);
  const int E = scalar(S, P, S_N_EDGES, b);
  uint8_t* es = S.edge_src + (size_t)b * P.ecap;
  uint8_t* ed = S.edge_dst + (size_t)b * P.ecap;
  int32_t* ec = S.edge_cnt + (size_t)b * P.ecap;
  int my_s = -1, my_t = -1, my_c = 0;                 // edge `lane` of the first block of 32
  if (lane < P.ecap) { my_s = es[lane]; my_t = ed[lane]; my_c = ec[lane]; }
  const int gv = T.uvuln_global[uvoff + u];
  const float p = lane < NN_CH ? T.vuln_h[(size_t)gv * NN_CH + lane] : 0.f;
  int found = -1, cnt = 0;
  for (int base = 0; base < E; base += 32) {
    const int e = base + lane;
    int cs = my_s, ct = my_t, cc = my_c;
    if (base) { cs = ct = -1; if (e < E) { cs = es[e]; ct = ed[e]; cc = ec[e]; } }
    const bool hit = e < E && cs == s && ct == t;
    const unsigned m = __ballot_sync(0xFFFFFFFFu, hit);
    if (m) { found = base + __ffs(m) - 1; cnt = __shfl_sync(0xFFFFFFFFu, cc, __ffs(m) - 1); break; }
  }
  if (found >= 0) {
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
      const int cs = base ? (e < E ? (int)es[e] : -1) : my_s;
      if (e < E && cs == s) ec[e] = 0;
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
#ifdef CBS_OBS_SUBTRACE
#define SUBT(k) if (tsub) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tsub[k]))
#else
#define SUBT(k)
#endif
template <bool EV>
__device__ void encode_env(const Tables& T, const Params& P, const State& S, const SharedWeights& SW, WarpScratch& W, int b,
                           int lane, long long* tsub = nullptr) {
  // everything that depends on the env index alone is requested here, in one round trip: the scalars, the discovered order, the
  // first 32 edges, and the two counters the observation's tail needs (loaded where they were used, each was an exposed trip)
  const int sc = scalar(S, P, S_SCENARIO, b);
  const int node_off = scalar(S, P, S_NODE_OFF, b);          // T.sc_node_off[sc], kept per env by the reset
  const int n_disc = scalar(S, P, S_N_DISC, b);
  const int E = scalar(S, P, S_N_EDGES, b);
  const int n_owned_obs = scalar(S, P, P.defender ? S_N_OWNED_RAW : S_N_OWNED, b);
  const int n_encodes = scalar(S, P, S_N_ENCODES, b);
  const uint8_t* disc_order = S.disc_order + (size_t)b * P.ncap;
  const uint8_t* es = S.edge_src + (size_t)b * P.ecap;
  const uint8_t* ed = S.edge_dst + (size_t)b * P.ecap;
  int my_es = 0, my_ed = 0;                                  // edge `lane` of the current block of 32 edges
  if (lane < P.ecap) { my_es = es[lane]; my_ed = ed[lane]; }
  float hm_n = 0.f;                                          // attribute of the next edge (requested one edge ahead)
  if (lane < NN_CH) hm_n = S.edge_m[((size_t)b * P.ecap) * NN_CH + lane];
  // lane l owns channels 2l and 2l+1 of every 64-wide row: one 64-bit access per row and lane
  constexpr int C2 = NODE_EMB / 2;            // float2 per row
  constexpr int ROW = PROJ_ROWS * NODE_EMB;   // floats per (node, part)
  float2* Y = reinterpret_cast<float2*>(W.y);
  float2* G = reinterpret_cast<float2*>(W.g);
  const float2* DYN = reinterpret_cast<const float2*>(SW.dyn);
  // *_node goals: once an encode has added the interest node to the live graph (compressed:254-256) it is part of
  // every later encode, discovered or not
  const int interest = is_node_goal(P) ? T.sc_interest[sc] : -1;
  const bool interest_known = interest >= 0 && bit_of(S, P, M_DISCOVERED, interest, b);
  const bool extra = interest >= 0 && !interest_known && (scalar(S, P, S_FLAGS, b) & FL_INTEREST_IN_GRAPH);
  const int n = n_disc + (extra ? 1 : 0);
  uint8_t* order = W.ord;
#ifdef CBS_OBS_SUBTRACE
  if (tsub && n + E + node_off >= 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tsub[0]));   // the scalars have arrived
#endif

  for (int i = lane; i < n; i += 32) {
    const int node = i < n_disc ? disc_order[i] : interest;
    W.ord[i] = (uint8_t)node;
    W.pos[node] = (uint8_t)i;
    W.dinv[i] = 1.f;
    const uint8_t db = pack_dyn(S, P, b, node);
    W.dynb[i] = db;
    W.xst[i] = pack_xstatus(S, P, b, node);
    // the node's root row (256 bytes of the folded table) on its way into L1 before the tiles below ask for it
    const float* rr = T.node_static + ((size_t)(node_off + node) * 2 + (db & 1)) * ROW + 17 * NODE_EMB;
    prefetch_l1(rr);
    prefetch_l1(rr + 32);
  }
  __syncwarp();
  SUBT(1);

  // per edge of the block: source / target positions and the source's table row (visible / not-visible variant), lane-local
  int my_is = 0, my_id = 0, my_row = 0;
  auto edge_block = [&](int base, bool reload) {
    if (reload) {
      my_es = my_ed = 0;
      if (base + lane < E) { my_es = es[base + lane]; my_ed = ed[base + lane]; }
    }
    my_is = my_id = my_row = 0;
    if (base + lane < E) {
      my_is = W.pos[my_es]; my_id = W.pos[my_ed];
      my_row = (node_off + my_es) * 2 + (W.dynb[my_is] & 1);
    }
  };
  // the 18 rows (4.5 KB) a source contributes to an edge message, requested into L1 one edge ahead
  auto prefetch_rows = [&](int row) {
    const float* base = T.node_static + (size_t)row * ROW;
    prefetch_l1(base + lane * 32);
    if (lane < (ROW * 4 / 128) - 32) prefetch_l1(base + (32 + lane) * 32);
  };
  edge_block(0, false);
  if (E > 0) prefetch_rows(__shfl_sync(0xFFFFFFFFu, my_row, 0));

  // root term x_i W_root (+ conv bias folded into bn1 shift); ROOT_TILE nodes per pass, their rows requested together.  The six
  // weight pairs of the root row stay in registers; a node is twelve straight FMAs.
  float2 wr[NUM_DYN];
#pragma unroll
  for (int d = 0; d < NUM_DYN; ++d) wr[d] = DYN[(d * PROJ_ROWS + 17) * C2 + lane];
  constexpr int ROOT_TILE = 4;      // (2 / 4 / 8 / 12 nodes per pass: 37.6 / 36.7 / 37.2 / 39.8 us for the kernel)
  for (int i0 = 0; i0 < n; i0 += ROOT_TILE) {
    float2 r[ROOT_TILE];
    uint32_t dx[ROOT_TILE];
#pragma unroll
    for (int j = 0; j < ROOT_TILE; ++j) {
      const int i = i0 + j;
      r[j] = make_float2(0.f, 0.f);
      dx[j] = 0;
      if (i < n) {
        dx[j] = (uint32_t)W.dynb[i] | ((uint32_t)W.xst[i] << 8);
        r[j] = reinterpret_cast<const float2*>(T.node_static + ((size_t)(node_off + order[i]) * 2 + (dx[j] & 1)) * ROW + 17 * NODE_EMB)[lane];
      }
    }
#pragma unroll
    for (int j = 0; j < ROOT_TILE; ++j) {
      const int i = i0 + j;
      if (i >= n) break;
      float vis, x[NUM_DYN];
      node_dyn((uint8_t)dx[j], (uint8_t)(dx[j] >> 8), vis, x);
      float2 a = r[j];
#pragma unroll
      for (int d = 0; d < NUM_DYN; ++d) {
        a.x = fmaf(x[d], wr[d].x, a.x);
        a.y = fmaf(x[d], wr[d].y, a.y);
      }
      if (EV && vis != 0.f) {
        const uint16_t* evx = S.ev_x + ((size_t)b * P.ncap + order[i]) * 4;
        const uint16_t* ini = T.nd_ev_init + (size_t)(node_off + order[i]) * 4;
        a.x += ev_delta(T.ev_proj, evx, ini, 17, 2 * lane);
        a.y += ev_delta(T.ev_proj, evx, ini, 17, 2 * lane + 1);
      }
      Y[i * C2 + lane] = a;
    }
  }
  __syncwarp();
  SUBT(2);

  // NNConv messages: y[dst] += [relu(m_e + b1); 1] . T_src
  for (int e = 0; e < E; ++e) {
    if ((e & 31) == 0 && e) edge_block(e, true);
    const int is = __shfl_sync(0xFFFFFFFFu, my_is, e & 31), id = __shfl_sync(0xFFFFFFFFu, my_id, e & 31);
    const int row = __shfl_sync(0xFFFFFFFFu, my_row, e & 31);
    const float hm = hm_n;
    if (e + 1 < E) {
      if (lane < NN_CH) hm_n = S.edge_m[((size_t)b * P.ecap + e + 1) * NN_CH + lane];
      if (((e + 1) & 31) != 0) prefetch_rows(__shfl_sync(0xFFFFFFFFu, my_row, (e + 1) & 31));
    }
    float hl = 0.f;
    if (lane < NN_CH) hl = fmaxf(hm + SW.nn0b[lane], 0.f);
    else if (lane == NN_CH) hl = 1.f;
    float vis, x[NUM_DYN];
    node_dyn(W.dynb[is], W.xst[is], vis, x);
    const float2* ns = reinterpret_cast<const float2*>(T.node_static + (size_t)row * ROW);
    const int js = row / 2 - node_off;
    const bool evd = EV && vis != 0.f;
    const uint16_t* evx = evd ? S.ev_x + ((size_t)b * P.ncap + js) * 4 : nullptr;
    const uint16_t* ini = evd ? T.nd_ev_init + (size_t)(node_off + js) * 4 : nullptr;
    const bool ev_any = evd && (((evx[0] ^ ini[0]) | (evx[1] ^ ini[1]) | (evx[2] ^ ini[2])) & 0x3FF) != 0;
    // t_k = static row k + sum over the NON-ZERO dynamic scalars (ascending d, so every t_k sums in the order it always did;
    // fmaf(0, w, t) == t).  The 18 row loads stay unrolled — they must all be in flight together: with a rolled loop over k (three
    // rows per pass) an edge took 2.9 us instead of 1.0 — but the pass per scalar is rolled: unrolled over the six scalars it was
    // 5 KB of straight-line code that every warp streamed through the instruction cache once per edge (observe 39.8 -> 36.6 us).
    float2 t[NN_CH + 1];
#pragma unroll
    for (int k = 0; k < NN_CH + 1; ++k) t[k] = ns[k * C2 + lane];
    uint32_t nzl = 0;      // the non-zero scalars, ascending: 4 bits each
    int nnz = 0;
#pragma unroll
    for (int d = 0; d < NUM_DYN; ++d) if (x[d] != 0.f) { nzl |= (uint32_t)d << (4 * nnz); ++nnz; }
#pragma unroll 1
    for (int q = 0; q < nnz; ++q) {
      const int d = (nzl >> (4 * q)) & 15;
      const float xd = d < 4 ? 1.f : (d == 4 ? x[4] : x[5]);
      const float2* dw = DYN + (size_t)d * PROJ_ROWS * C2 + lane;
#pragma unroll
      for (int k = 0; k < NN_CH + 1; ++k) {
        const float2 w = dw[k * C2];
        t[k].x = fmaf(xd, w.x, t[k].x);
        t[k].y = fmaf(xd, w.y, t[k].y);
      }
    }
    float2 m = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < NN_CH + 1; ++k) {
      const float hk = __shfl_sync(0xFFFFFFFFu, hl, k);
      if (EV && ev_any) { t[k].x += ev_delta(T.ev_proj, evx, ini, k, 2 * lane); t[k].y += ev_delta(T.ev_proj, evx, ini, k, 2 * lane + 1); }
      m.x = fmaf(hk, t[k].x, m.x);
      m.y = fmaf(hk, t[k].y, m.y);
    }
    float2 yv = Y[id * C2 + lane];
    yv.x += m.x; yv.y += m.y;
    Y[id * C2 + lane] = yv;
    if (lane == 0 && is != id) W.dinv[id] += 1.f;           // GCN in-degree (self loops are replaced, not counted)
    __syncwarp();
  }
  SUBT(3);
  for (int i = lane; i < n; i += 32) W.dinv[i] = rsqrtf(W.dinv[i]);
  // BatchNorm(eval) + ReLU, then the GCN projection G = H1 Wg^T.  Node loops run four nodes per pass, loads before stores: one
  // node per pass is a chain of dependent shared-memory round trips (the compiler must keep a load behind the previous store).
  const float2 b1s = reinterpret_cast<const float2*>(SW.bn1s)[lane], b1h = reinterpret_cast<const float2*>(SW.bn1h)[lane];
  for (int i0 = 0; i0 < n; i0 += 4) {
    float2 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) if (i0 + j < n) v[j] = Y[(i0 + j) * C2 + lane];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (i0 + j < n) Y[(i0 + j) * C2 + lane] = make_float2(fmaxf(fmaf(v[j].x, b1s.x, b1h.x), 0.f), fmaxf(fmaf(v[j].y, b1s.y, b1h.y), 0.f));
  }
  __syncwarp();
  SUBT(4);
  // Several nodes per pass: each weight pair feeds two independent FMA chains per node, and a pass's broadcast reads of the
  // nodes' activations (128 bits = four k at a time) are all in flight together.  Passes of 8, then 4 / 2 / 1 for the remainder,
  // so that no pass computes padding.  Every node's sum runs over k in the same order whatever the tile, so equal inputs still
  // give bit-identical outputs.  A pass ends by clearing its rows of Y for the aggregation that follows.
  auto gcn_pass = [&](auto tile_c, int i0) {
    constexpr int TILE = decltype(tile_c)::value;
    float2 acc[TILE];
#pragma unroll
    for (int j = 0; j < TILE; ++j) acc[j] = make_float2(0.f, 0.f);
    const float4* yrow = reinterpret_cast<const float4*>(W.y + i0 * NODE_EMB);
    const float2* gw = reinterpret_cast<const float2*>(SW.gcn);
#pragma unroll 2
    for (int k4 = 0; k4 < NODE_EMB / 4; ++k4) {
      float2 w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) w[q] = gw[(k4 * 4 + q) * C2 + lane];
      float4 a[TILE];
#pragma unroll
      for (int j = 0; j < TILE; ++j) a[j] = yrow[j * (NODE_EMB / 4) + k4];
#pragma unroll
      for (int j = 0; j < TILE; ++j) {
        acc[j].x = fmaf(a[j].x, w[0].x, acc[j].x); acc[j].y = fmaf(a[j].x, w[0].y, acc[j].y);
        acc[j].x = fmaf(a[j].y, w[1].x, acc[j].x); acc[j].y = fmaf(a[j].y, w[1].y, acc[j].y);
        acc[j].x = fmaf(a[j].z, w[2].x, acc[j].x); acc[j].y = fmaf(a[j].z, w[2].y, acc[j].y);
        acc[j].x = fmaf(a[j].w, w[3].x, acc[j].x); acc[j].y = fmaf(a[j].w, w[3].y, acc[j].y);
      }
    }
    __syncwarp();                       // every lane has read the whole rows before they are cleared
#pragma unroll
    for (int j = 0; j < TILE; ++j) {
      G[(i0 + j) * C2 + lane] = acc[j];
      Y[(i0 + j) * C2 + lane] = make_float2(0.f, 0.f);
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
  SUBT(5);
  // normalised aggregation, edges first and the self loop last (the order PyG's add_remaining_self_loops +
  // scatter-add gives).  Products and sums are rounded separately (no FMA contraction): a pair of nodes that
  // attack each other then gets bit-identical embeddings, exactly as in the reference, and the exact ties this
  // creates in the action table resolve by insertion order instead of by rounding noise.
  if (E > 32) edge_block(0, true);
  for (int e = 0; e < E; ++e) {
    if ((e & 31) == 0 && e) edge_block(e, true);
    const int is = __shfl_sync(0xFFFFFFFFu, my_is, e & 31), id = __shfl_sync(0xFFFFFFFFu, my_id, e & 31);
    if (is == id) continue;
    const float w = __fmul_rn(W.dinv[is], W.dinv[id]);
    const float2 g = G[is * C2 + lane];
    float2 yv = Y[id * C2 + lane];
    yv.x = __fadd_rn(yv.x, __fmul_rn(w, g.x));
    yv.y = __fadd_rn(yv.y, __fmul_rn(w, g.y));
    Y[id * C2 + lane] = yv;
    __syncwarp();
  }
  SUBT(6);
  // self loop, BatchNorm + ReLU -> z ; readout over Running nodes (mean | max | min), compressed:266-298
  const float2 b2s = reinterpret_cast<const float2*>(SW.bn2s)[lane], b2h = reinterpret_cast<const float2*>(SW.bn2h)[lane];
  float s0 = 0.f, s1 = 0.f, mx0 = -INFINITY, mx1 = -INFINITY, mn0 = INFINITY, mn1 = INFINITY;
  int running = 0;
  for (int i0 = 0; i0 < n; i0 += 4) {
    float2 yv[4], gv[4];
    float dv[4];
    uint8_t db[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (i0 + j < n) { yv[j] = Y[(i0 + j) * C2 + lane]; gv[j] = G[(i0 + j) * C2 + lane]; dv[j] = W.dinv[i0 + j]; db[j] = W.dynb[i0 + j]; }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (i0 + j >= n) break;
      const float d2 = __fmul_rn(dv[j], dv[j]);
      const float y0 = __fadd_rn(yv[j].x, __fmul_rn(d2, gv[j].x)), y1 = __fadd_rn(yv[j].y, __fmul_rn(d2, gv[j].y));
      const float z0 = fmaxf(fmaf(y0, b2s.x, b2h.x), 0.f), z1 = fmaxf(fmaf(y1, b2s.y, b2h.y), 0.f);
      Y[(i0 + j) * C2 + lane] = make_float2(z0, z1);
      if (db[j] & 0x80) {
        ++running;
        s0 += z0; s1 += z1;
        mx0 = fmaxf(mx0, z0); mx1 = fmaxf(mx1, z1);
        mn0 = fminf(mn0, z0); mn1 = fminf(mn1, z1);
      }
    }
  }
  SUBT(7);
  float* obs = S.obs + (size_t)b * P.obs_dim;
  const int c0 = 2 * lane, c1 = 2 * lane + 1;
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
    __syncwarp();
    obs[OBS_GRAPH + c0] = have ? W.y[ip * NODE_EMB + c0] : 0.f;
    obs[OBS_GRAPH + c1] = have ? W.y[ip * NODE_EMB + c1] : 0.f;
  }
  if (lane == 0) {
    obs[P.obs_dim - 2] = (float)n_disc;                     // create_discrete_features, compressed:309-316
    obs[P.obs_dim - 1] = (float)n_owned_obs;                // len(owned_nodes)
    scalar(S, P, S_N_ENCODES, b) = n_encodes + 1;
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
                            int skipped_builds = 0, long long* tsub = nullptr) {
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
  // Everything that depends on the env index alone is requested first and together (the scalars, both node orders, the counter
  // updated at the end, the pair plane on its way into L1): read where they were used, the loops below were a chain of dependent
  // round trips, one per source and one per stored node.
  const int n_disc = scalar(S, P, S_N_DISC, b), n_owned = scalar(S, P, P.defender ? S_N_OWNED_RAW : S_N_OWNED, b);
  const uint8_t* dorder = S.disc_order + (size_t)b * P.ncap;
  const uint8_t* oorder = P.defender ? S.owned_raw + (size_t)b * P.ocap : S.owned_order + (size_t)b * P.ncap;
  uint8_t* ps = S.pair_slot + (size_t)b * P.ncap * P.ncap;
  uint8_t* po = S.pair_opos + (size_t)b * P.ncap * P.ncap;
  uint8_t* pe = S.pair_epoch + (size_t)b * P.ncap * P.ncap;
  for (int l = lane; l * 128 < P.ncap * P.ncap; l += 32) prefetch_l1(ps + l * 128);
  const int slot = scalar(S, P, S_N_SLOTS, b);
  int work_est0 = 0;
  if (lane == 0 && !SUBSET) work_est0 = S.work_est[b];
  int my_t = lane < P.ncap ? dorder[lane] : 0;                 // discovered position `lane` (first block of 32 targets)
  int my_o = lane < (P.defender ? P.ocap : P.ncap) ? oorder[lane] : 0;   // owned position `lane` (first block of 32 sources)
  // per target: candidate rows of a remote / a local pair (what a new pair adds to the env's decode work estimate)
  int rows_remote0 = 0, rows_self0 = 0;
  auto target_rows = [&](int t, int& rr, int& rs) {
    const int g = node_off_bt + t;
    const int o0 = T.nd_row_off[2 * g], o1 = T.nd_row_off[2 * g + 1], o2 = T.nd_row_off[2 * g + 2];
    rr = o2 - o1; rs = o2 - o0;
  };
  if (lane < n_disc) target_rows(my_t, rows_remote0, rows_self0);
  bool any_new = false;
  uint32_t seen[MAX_NODES / 32] = {0u, 0u, 0u, 0u};
#ifdef CBS_OBS_SUBTRACE
  if (tsub && n_disc + n_owned + slot + rows_self0 + my_o >= 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tsub[0]));
#endif
  for (int op = 0; op < n_owned; ++op) {
    if ((op & 31) == 0 && op) my_o = op + lane < n_owned ? oorder[op + lane] : 0;
    const int s = __shfl_sync(0xFFFFFFFFu, my_o, op & 31);
    if (P.defender) {
      if ((seen[s >> 5] >> (s & 31)) & 1u) continue;
      seen[s >> 5] |= 1u << (s & 31);
    }
    if (!(W.dynb[W.pos[s]] & 0x80)) continue;               // stopped sources add no rows (compressed:491-492)
    for (int base = 0; base < n_disc; base += 32) {
      const int dp = base + lane;
      bool fresh = false, fresh_refresh_only = false;
      if (dp < n_disc) {
        int t = my_t, rr = rows_remote0, rs = rows_self0;
        if (base) { t = dorder[dp]; target_rows(t, rr, rs); }
        const bool is_new = ps[s * P.ncap + t] == 0xFF;
        fresh_refresh_only = !is_new;
        fresh = (W.dynb[dp] & 0x80) && (is_new || (PRECISE && refresh && (in_reach(s) || in_reach(t))));
        if (fresh && slot < P.slots) {
          ps[s * P.ncap + t] = (uint8_t)slot;
          if (is_new) {
            if (P.defender) po[s * P.ncap + t] = (uint8_t)op;   // insertion order inside the slot (exact-tie order of the decode)
            if (PRECISE) pe[s * P.ncap + t] = (uint8_t)slot;
            new_rows += s == t ? rs : rr;
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
  SUBT(1);
  if (!any_new) {
    // the reference still balances (and draws from its generator) at the end of this create_continuous_action_space
    if constexpr (SUBSET) subset_update(T, P, S, sub, b, lane, slot, 0, oorder, n_owned, dorder, W.pos, skipped_builds);
    return;
  }
  if (slot >= P.slots) { if (lane == 0) atomicExch(S.errflag, 1); return; }
  new_rows = (int)warp_sum((float)new_rows);
  if (lane == 0 && !SUBSET) S.work_est[b] = work_est0 + new_rows;
  float* zh = S.z_hist + ((size_t)b * P.slots + slot) * P.ncap * NODE_EMB;
  float* zn = S.zn2_hist + ((size_t)b * P.slots + slot) * P.ncap;
  __half2* zh16 = reinterpret_cast<__half2*>(S.z16_hist + ((size_t)b * P.slots + slot) * P.ncap * NODE_EMB);
  // four nodes per pass: their squared norms go through the five butterfly levels together (one node per pass was a chain of five
  // dependent shuffles per node, 0.23 us per node; every node's sum still runs in the same order)
  for (int i0 = 0; i0 < n_disc; i0 += 4) {
    float2 z[4];
    float sq[4];
    int node[4];
    bool run[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = i0 + j;
      run[j] = i < n_disc && (W.dynb[i < n_disc ? i : 0] & 0x80);       // only Running nodes have embeddings (compressed:266-280)
      node[j] = i < 32 ? __shfl_sync(0xFFFFFFFFu, my_t, i & 31) : (i < n_disc ? (int)dorder[i] : 0);
      z[j] = run[j] ? reinterpret_cast<const float2*>(W.y + i * NODE_EMB)[lane] : make_float2(0.f, 0.f);   // channels 2*lane, 2*lane+1
      sq[j] = z[j].x * z[j].x + z[j].y * z[j].y;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (!run[j]) continue;
      reinterpret_cast<float2*>(zh + node[j] * NODE_EMB)[lane] = z[j];
      zh16[node[j] * (NODE_EMB / 2) + lane] = __floats2half2_rn(z[j].x, z[j].y);
      if (fmaxf(fabsf(z[j].x), fabsf(z[j].y)) > 65504.f) atomicExch(S.errflag, 8);     // beyond half precision: the decode scan reads these copies
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) sq[j] += __shfl_xor_sync(0xFFFFFFFFu, sq[j], o);
    }
    if (lane < 4) {
      const int nd = lane == 0 ? node[0] : (lane == 1 ? node[1] : (lane == 2 ? node[2] : node[3]));
      const float v = lane == 0 ? sq[0] : (lane == 1 ? sq[1] : (lane == 2 ? sq[2] : sq[3]));
      const bool r = lane == 0 ? run[0] : (lane == 1 ? run[1] : (lane == 2 ? run[2] : run[3]));
      if (r) zn[nd] = v;
    }
  }
  SUBT(2);
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
    scalar(S, P, S_UVULN_OFF, b) = T.sc_uvuln_off[sc];
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
// PRECISE: precise_action_space_positions is configured (compile-time like SUBSET: one build_table instance per kernel)
template <bool BIG_GRAPHS, bool SUBSET, bool PRECISE>
__global__ void __launch_bounds__(CBS_OBS_BOUND, CBS_OBS_MINB) observe_kernel(Tables T, Params P, State S,
                                                                const uint8_t* __restrict__ reset_mask, int mode,
                                                                long long* __restrict__ trace) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SharedWeights& SW = *reinterpret_cast<SharedWeights*>(smem_raw);
  // The weights (44 KB) are staged with asynchronous copies that nobody waits for here: a warp waits on the mbarrier right before its
  // first encode, i.e. behind its first item's claim, flags and edge update — staged with plain loads behind a barrier, the copy
  // was 2 us at the head of a 35 us kernel whose duration is its first (heaviest) items'.
  const uint32_t wbar = (uint32_t)__cvta_generic_to_shared(&SW.wbar);
  if (threadIdx.x < N_ACCUM) SW.accum[threadIdx.x] = 0.0;
  if (threadIdx.x == 0) {
    SW.warps_done = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(wbar), "r"(OBS_WARPS * 32));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  {
    constexpr int NT = OBS_WARPS * 32;
    constexpr int G4 = NODE_EMB * NODE_EMB / 4, D4 = NUM_DYN * PROJ_ROWS * NODE_EMB / 4;   // 1024, 1728 float4
    auto cp16 = [](void* dst, const void* src) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
    };
    auto cp4 = [](void* dst, const void* src) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
    };
    for (int j = threadIdx.x; j < G4; j += NT) cp16(reinterpret_cast<float4*>(SW.gcn) + j, reinterpret_cast<const float4*>(T.gcn_wt) + j);
    for (int j = threadIdx.x; j < D4; j += NT) cp16(reinterpret_cast<float4*>(SW.dyn) + j, reinterpret_cast<const float4*>(T.dyn_proj) + j);
    if (threadIdx.x < NODE_EMB) {
      cp4(&SW.bn1s[threadIdx.x], T.bn1_scale + threadIdx.x);
      cp4(&SW.bn1h[threadIdx.x], T.bn1_shift + threadIdx.x);
      cp4(&SW.bn2s[threadIdx.x], T.bn2_scale + threadIdx.x);
      cp4(&SW.bn2h[threadIdx.x], T.bn2_shift + threadIdx.x);
    }
    if (threadIdx.x < NN_CH) cp4(&SW.nn0b[threadIdx.x], T.nn0_b + threadIdx.x);
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(wbar) : "memory");
  }
  // Programmatic dependent launch (cbs_types.h): up to here only constant tables were read, so this CTA may have started while the
  // transition that feeds it was still running.  The trigger comes AFTER the wait: the next step's contraction may then start its
  // operand pipeline as this kernel's SMs free up, and everything before this kernel is known to be complete by then.
  pdl_wait();
  pdl_trigger();
  bool weights_ready = false;
  auto wait_weights = [&]() {
    if (weights_ready) return;
    weights_ready = true;
    uint32_t ok = 0;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(wbar) : "memory");
    } while (!ok);
  };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * OBS_WARPS + warp, total_warps = gridDim.x * OBS_WARPS;

  // per-warp scratch
  unsigned char* wbase = smem_raw + kSharedWeightsBytes;
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
  // mode 1: static stride.  mode 0: items are claimed one at a time.  The claim for the NEXT item is issued when the current item
  // is nearly done (after its encode; at once for the short edge-only items), so that the atomic's round trip is hidden but no
  // warp sits on a claimed heavy item: claimed a whole item ahead, the kernel ran 51 us instead of 40 (the heaviest items are
  // first in the queue, and a warp busy with one held the next one back).
  int i = gw, pend = 0;
  bool claimed = false;
  auto claim_next = [&]() {
    if (mode != 0 || claimed) return;
    claimed = true;
    if (lane == 0) pend = total_warps + atomicAdd(&S.work_ctr[2], 1);
  };
  auto lookup = [&](int item) {
    int k = 0, first = 0;
#pragma unroll
    for (int q = 0; q < OBS_CLASSES - 1; ++q) if (item >= cls_end[q]) { k = q + 1; first = cls_end[q]; }
    return S.worklist[(size_t)k * P.B + (item - first)];
  };
  // (every warp's first item is its own index: 1184 simultaneous claims of one counter took the last of them microseconds; the
  // counter then hands out the items behind the first round)
  for (;;) {
    if (i >= count) break;
    const int b = mode == 0 ? lookup(i) : i;
    claimed = false;
    // node-embedding buffers: shared memory while the env's visible graph has <= 32 nodes (always, when the
    // scenarios have <= 32 nodes), its slab of the L2-resident scratch otherwise.  A reset shrinks the graph to 1 node.
    if (BIG_GRAPHS) {
      const bool small = scalar(S, P, S_N_DISC, b) <= SMEM_NODES;
      W.y = small ? W.ysm : S.scratch + (size_t)b * 2 * P.ncap * NODE_EMB;
      W.g = small ? W.ysm + SMEM_NODES * NODE_EMB : W.y + (size_t)P.ncap * NODE_EMB;
    }
    const int flags = scalar(S, P, S_FLAGS, b);
#ifdef CBS_OBS_SUBTRACE
    constexpr int TR_PH = 14;      // debug build: trace rows of 4 + 14 values, slots 5.. = the encode's sub-phases
#else
    constexpr int TR_PH = 5;
#endif
    long long t_item = 0, t_ph[TR_PH] = {};
    if (trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_item));
    bool do_reset = false;
    if (mode == 1) {
      do_reset = reset_mask ? (reset_mask[b] != 0) : true;
    } else if (!(flags & FL_NEEDS_RESET)) {
      int keep = flags & ~(FL_ADD_EDGE | FL_REENCODE | FL_FINISHED_THIS_STEP);
      if (!(flags & (FL_REENCODE | FL_FINISHED_THIS_STEP))) claim_next();    // a short item
      if (flags & FL_ADD_EDGE) edge_update(T, P, S, b, lane);
#ifdef CBS_OBS_SUBTRACE
      if (trace) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_ph[13])); }
#endif
      if (flags & FL_REENCODE) {
#ifdef CBS_OBS_SUBTRACE
        wait_weights();
        if (P.defender == 2) encode_env<true>(T, P, S, SW, W, b, lane, trace ? t_ph + 5 : nullptr); else encode_env<false>(T, P, S, SW, W, b, lane, trace ? t_ph + 5 : nullptr);
#else
        wait_weights();
        if (P.defender == 2) encode_env<true>(T, P, S, SW, W, b, lane); else encode_env<false>(T, P, S, SW, W, b, lane);
#endif
        if (trace) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_ph[0])); }
        // (an env that finished in this step drops its action table at the reset that follows: only the sub-sampled table is
        // still maintained, because every create_continuous_action_space advances the env's balance counter)
        if (!(flags & FL_FINISHED_THIS_STEP)) claim_next();
        if (!(flags & FL_FINISHED_THIS_STEP) || SUBSET) {
          const int skipped = SUBSET ? (flags >> FL_PENDING_SHIFT) & 0xFFFF : 0;
#ifdef CBS_OBS_SUBTRACE
          build_table<PRECISE, SUBSET>(T, P, S, W, b, lane, PRECISE, skipped, trace ? t_ph + 2 : nullptr);
#else
          build_table<PRECISE, SUBSET>(T, P, S, W, b, lane, PRECISE, skipped);
#endif
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
      claim_next();
      if (trace) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_ph[3])); }
      if (!reset_from_cache(T, P, S, b, lane, ss.x, ss.y)) {
        W.y = W.ysm;                               // a fresh episode's graph is one node
        W.g = W.ysm + SMEM_NODES * NODE_EMB;
        wait_weights();
        if (P.defender == 2) encode_env<true>(T, P, S, SW, W, b, lane); else encode_env<false>(T, P, S, SW, W, b, lane);
        build_table<PRECISE, SUBSET>(T, P, S, W, b, lane, false);
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
      long long* tr = trace + (size_t)b * (4 + TR_PH);
      tr[0] = t_item; tr[1] = t_end - t_item; tr[2] = (long long)(unsigned)flags | ((long long)(blockIdx.x * OBS_WARPS + warp) << 32);
      tr[3] = ((long long)scalar(S, P, S_N_DISC, b) << 16) | scalar(S, P, S_N_EDGES, b);
#pragma unroll
      for (int k = 0; k < TR_PH; ++k) tr[4 + k] = t_ph[k] ? t_ph[k] - t_item : 0;   // phase end times relative to the item start
    }
    claim_next();
    if (mode == 1) i += total_warps; else i = __shfl_sync(0xFFFFFFFFu, pend, 0);
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
  // the last CTA to run dry clears the counters for the next transition.  No fence: a warp gets here after its last claim has
  // RETURNED (the value ended its loop), so every claim precedes the count that detects the last CTA.
  if (mode == 0 && last_in_cta && lane == 0) {
    if (atomicAdd(&S.work_ctr[1], 1) == (int)gridDim.x - 1) {
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
  return kSharedWeightsBytes + OBS_WARPS * per_warp;
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
#ifdef CBS_OBS_ONLY_DEFAULT   // experiments: build the default instance alone (a full build of this file takes two minutes)
  const KernelFn kernels[8] = {observe_kernel<false, false, false>, observe_kernel<false, false, false>, observe_kernel<false, false, false>,
                               observe_kernel<false, false, false>, observe_kernel<false, false, false>, observe_kernel<false, false, false>,
                               observe_kernel<false, false, false>, observe_kernel<false, false, false>};
  if (big || P.subset_k || P.precise_positions) { fprintf(stderr, "CBS_OBS_ONLY_DEFAULT build\n"); abort(); }
#else
  const KernelFn kernels[8] = {observe_kernel<false, false, false>, observe_kernel<true, false, false>, observe_kernel<false, true, false>,
                               observe_kernel<true, true, false>,   observe_kernel<false, false, true>, observe_kernel<true, false, true>,
                               observe_kernel<false, true, true>,   observe_kernel<true, true, true>};
#endif
  const int which = (big ? 1 : 0) | (P.subset_k ? 2 : 0) | (P.precise_positions ? 4 : 0);
  {
    cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(kernels[which]), smem);
    if (e != cudaSuccess) return e;
  }
  // (pinning the folded node table in L2 with an access-policy window on this launch changed nothing: 37.1 us either way)
  return launch_pdl(kernels[which], dim3(grid), dim3(OBS_WARPS * 32), smem, stream, true, T, P, S, reset_mask, mode, g_obs_trace);
}

}  // namespace cbs
