// k_decode_metric.cu — nearest-action decode for the reference's other metrics (find_closest_action_embedding,
// _env/cyberbattle_env_compressed.py:571-576): 'l1', 'l2', 'inf' = np.linalg.norm(action - rows, ord, axis=1) in float64.
//
// A table row is [emb(src) | emb(tgt) | vuln_emb | onehot] and all three norms separate over those four parts:
//     l1:  |a_s - z_s|_1 + |a_t - z_t|_1 + |a_v - v_u|_1 + |a_o - e_oh|_1
//     l2:  sqrt of the sum of the four squared part norms          inf: max of the four part maxima
// so, like the cosine path, the table stays implicit.  The vulnerability part is one dense [B, Ug] "distance contraction"
// shared by every row (metric_vuln_kernel, float64 on the FP64 pipe: |x - y| accumulations are not a tensor-core shape);
// metric_select_kernel then walks the env's (source, target) pairs in table-insertion order, one warp per env, computes the
// two node parts per pair in float64 from the float32 snapshot embeddings and combines them per row with the cached
// vulnerability and one-hot parts.  Everything is float64 from the same float32 / float64 inputs the reference holds, so only
// the summation order differs from numpy (<= 1e-15 relative); ties and NaN follow np.argmin (first NaN, else first minimum in
// insertion order).  The non-default metrics run decode -> transition as two launches (no fused variant).
#include "transition.cuh"
#include "subset.cuh"

namespace cbs {

template <int METRIC>
__device__ __forceinline__ void macc(double& m, double x) {
  if (METRIC == METRIC_L1) m += fabs(x);
  else if (METRIC == METRIC_L2) m = fma(x, x, m);
  else { const double v = fabs(x); m = (v > m || v != v) ? v : m; }      // np.max propagates NaN
}
template <int METRIC>
__device__ __forceinline__ double mjoin(double a, double b) {
  if (METRIC == METRIC_INF) return (b > a || b != b) ? b : a;
  return a + b;
}
template <int METRIC>
__device__ __forceinline__ double mfinish(double a) { return METRIC == METRIC_L2 ? sqrt(a) : a; }

// ------------------------------------------------------------------------------------------------
// vulnerability part: vt64[b][u] = sum_k |a_v[b][k] - v[u][k]|  (l1) / sum of squares (l2) / max (inf)
// 64 x 64 output tile per CTA, 4 x 4 per thread, K staged 16 at a time in shared memory as float64
// ------------------------------------------------------------------------------------------------
constexpr int MT_M = 64, MT_N = 64, MT_K = 16;

template <int METRIC>
__global__ void __launch_bounds__(256) metric_vuln_kernel(const float* __restrict__ actions, int act_stride, const double* __restrict__ vemb64,
                                                          double* __restrict__ vt64, int B, int Ug) {
  __shared__ double As[MT_K][MT_M + 1];
  __shared__ double Bs[MT_K][MT_N + 1];
  const int m0 = blockIdx.x * MT_M, n0 = blockIdx.y * MT_N;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int k0 = 0; k0 < VULN_EMB; k0 += MT_K) {
    for (int i = threadIdx.x; i < MT_M * MT_K; i += 256) {
      const int r = i / MT_K, k = i % MT_K;
      const int m = m0 + r;
      As[k][r] = m < B ? (double)actions[(size_t)m * act_stride + 2 * NODE_EMB + k0 + k] : 0.0;
    }
    for (int i = threadIdx.x; i < MT_N * MT_K; i += 256) {
      const int r = i / MT_K, k = i % MT_K;
      const int n = n0 + r;
      Bs[k][r] = n < Ug ? vemb64[(size_t)n * VULN_EMB + k0 + k] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < MT_K; ++k) {
      double a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; bb[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) macc<METRIC>(acc[i][j], a[i] - bb[j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= B) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < Ug) vt64[(size_t)m * Ug + n] = acc[i][j];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// pair walk + argmin, one warp per env
// ------------------------------------------------------------------------------------------------
constexpr int MET_WARPS = 4;
constexpr int MET_THREADS = MET_WARPS * 32;

struct MetWarp {
  double d_o[16];             // one-hot part per outcome column
  double p_d[32];             // per staged pair: the two node parts joined
  float a_st[2 * NODE_EMB];   // source | target parts of the action
  float a_o[16];
  int p_r0[32], p_pre[33];    // first candidate row, exclusive prefix of row counts
  uint32_t p_key[32];         // insertion epoch << 24 | source's insertion position << 16 | discovered position << 8 | index in owned_order
  uint8_t oorder[MAX_NODES], dorder[MAX_NODES];
  uint8_t opos[MAX_NODES], dpos[MAX_NODES];   // sample_subset_samples: node -> position in the two order lists (tie order)
};

// np.argmin over distances in insertion order: the first NaN wins if there is one, else the first minimum
__device__ __forceinline__ bool argmin_better(bool have, double bd, unsigned long long bk, double d, unsigned long long k) {
  const bool dn = d != d, bn = bd != bd;
  return !have || (dn ? (!bn || k < bk) : (!bn && (d < bd || (d == bd && k < bk))));
}

template <int METRIC>
__global__ void __launch_bounds__(MET_THREADS) metric_select_kernel(Tables T, Params P, State S, const float* __restrict__ actions,
                                                                    const double* __restrict__ vt64, int Ug, int sched_buf,
                                                                    int32_t* __restrict__ sel_out, double* __restrict__ dist_out) {
  __shared__ MetWarp sh_all[MET_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * MET_WARPS + warp;
  const bool in_range = b < P.B;
  MetWarp& sh = sh_all[warp];
  int flags = FL_NEEDS_RESET, node_off = 0, starter = 0, n_disc = 0, n_owned = 0;
  if (in_range) {
    flags = scalar(S, P, S_FLAGS, b);
    node_off = scalar(S, P, S_NODE_OFF, b);
    starter = scalar(S, P, S_STARTER, b);
    n_disc = scalar(S, P, S_N_DISC, b);
    n_owned = scalar(S, P, S_N_OWNED, b);
    const float* act = actions + (size_t)b * P.act_stride;
    for (int i = lane; i < P.ncap; i += 32) {
      sh.dorder[i] = S.disc_order[(size_t)b * P.ncap + i];
      sh.oorder[i] = S.owned_order[(size_t)b * P.ncap + i];
    }
    for (int i = lane; i < 2 * NODE_EMB; i += 32) sh.a_st[i] = act[i];
    if (lane < OUTCOME_DIM) sh.a_o[lane] = act[2 * NODE_EMB + VULN_EMB + lane];
    __syncwarp();
    if (lane < OUTCOME_DIM) {
      double m = 0.0;
      for (int j = 0; j < OUTCOME_DIM; ++j) macc<METRIC>(m, (double)sh.a_o[j] - (j == lane ? 1.0 : 0.0));
      sh.d_o[lane] = m;
    }
    __syncwarp();
  }
  // a finished env (the reference raises there, cyberbattle_env.py:300-302) decodes to zeros
  const bool active = in_range && !(flags & (FL_DONE | FL_TRUNC | FL_NEEDS_RESET));
  int4 out = make_int4(0, 0, 0, 0);
  double dres = 0.0;
  if (active) {
    const size_t pbase = (size_t)b * P.ncap * P.ncap;
    const double* vt_row = vt64 + (size_t)b * Ug;
    const int interest = is_node_goal(P) ? T.sc_interest[scalar(S, P, S_SCENARIO, b)] : -1;
    bool have = false;
    double best_d = 0.0;
    unsigned long long best_k = ~0ull;
    int best_st = 0, best_r = 0;               // source | target << 8 and candidate row of the best distance so far
    if (P.subset_k) {
      // sample_subset_samples: the table is the env's explicit per-class row lists (subset.cuh); one row per lane, the node parts
      // computed per row in float64 from the float32 snapshot embeddings
      const int K = P.subset_k;
      const uint32_t* lists = S.sub_rows + (size_t)b * SUB_CLASSES * K;
      const int32_t* meta = S.sub_meta + (size_t)b * SUB_META;
      int cls_end[SUB_CLASSES], total = 0;
#pragma unroll
      for (int c = 0; c < SUB_CLASSES; ++c) { total += meta[c]; cls_end[c] = total; }
      const uint32_t rank_lo = (uint32_t)meta[11], rank_hi = (uint32_t)meta[12];
      for (int i = lane; i < n_disc; i += 32) sh.dpos[sh.dorder[i]] = (uint8_t)i;
      if (!P.defender) for (int i = lane; i < n_owned; i += 32) sh.opos[sh.oorder[i]] = (uint8_t)i;
      __syncwarp();
      for (int j = lane; j < total; j += 32) {
        int c = 0, first = 0;
#pragma unroll
        for (int q = 0; q < SUB_CLASSES - 1; ++q) if (j >= cls_end[q]) { c = q + 1; first = cls_end[q]; }
        const uint32_t e = lists[c * K + (j - first)];
        const int s = e & 127, t = (e >> 7) & 127, rip = (e >> 14) & 255, epoch = (int)(e >> 22);
        const int slot = S.pair_slot[pbase + s * P.ncap + t];
        const int g = node_off + t;
        const int r = T.nd_row_off[2 * g + (s == t ? 0 : 1)] + rip;
        const uint32_t packed = T.row_packed[r];
        const size_t zbase = ((size_t)b * P.slots + slot) * P.ncap;
        const float4* zs = reinterpret_cast<const float4*>(S.z_hist + (zbase + s) * NODE_EMB);
        const float4* zt = reinterpret_cast<const float4*>(S.z_hist + (zbase + t) * NODE_EMB);
        double ms = 0.0, mt = 0.0;
        for (int i = 0; i < NODE_EMB / 4; ++i) {
          const float4 x = zs[i], y = zt[i];
          macc<METRIC>(ms, (double)sh.a_st[4 * i + 0] - (double)x.x);
          macc<METRIC>(ms, (double)sh.a_st[4 * i + 1] - (double)x.y);
          macc<METRIC>(ms, (double)sh.a_st[4 * i + 2] - (double)x.z);
          macc<METRIC>(ms, (double)sh.a_st[4 * i + 3] - (double)x.w);
          macc<METRIC>(mt, (double)sh.a_st[NODE_EMB + 4 * i + 0] - (double)y.x);
          macc<METRIC>(mt, (double)sh.a_st[NODE_EMB + 4 * i + 1] - (double)y.y);
          macc<METRIC>(mt, (double)sh.a_st[NODE_EMB + 4 * i + 2] - (double)y.z);
          macc<METRIC>(mt, (double)sh.a_st[NODE_EMB + 4 * i + 3] - (double)y.w);
        }
        const int u = packed & 0xFFFFF, oh = (packed >> 24) & 15;
        const double d = mfinish<METRIC>(mjoin<METRIC>(mjoin<METRIC>(mjoin<METRIC>(ms, mt), vt_row[u]), sh.d_o[oh]));
        const uint32_t rank = ((c < 8 ? rank_lo : rank_hi) >> (4 * (c & 7))) & 15u;
        const int opk = P.defender ? (int)S.pair_opos[pbase + s * P.ncap + t] : (int)sh.opos[s];
        const unsigned long long k64 = ((unsigned long long)rank << 32) | ((unsigned long long)epoch << 24) | ((unsigned long long)opk << 16) |
                                       ((unsigned long long)sh.dpos[t] << 8) | (unsigned long long)rip;
        if (argmin_better(have, best_d, best_k, d, k64)) { have = true; best_d = d; best_k = k64; best_st = s | (t << 8); best_r = r; }
      }
    }
    const int combos = P.subset_k ? 0 : n_owned * n_disc;
    for (int cbase = 0; cbase < combos; cbase += 32) {
      // ---- one (source, target) combination per lane: its two node parts ----
      const int c = cbase + lane;
      bool live = false;
      double pd = 0.0;
      int r0 = 0, cnt = 0;
      uint32_t key = 0;
      if (c < combos) {
        const int op = c / n_disc, dp = c - op * n_disc;
        const int s = sh.oorder[op], t = sh.dorder[dp];
        const int g = node_off + t;
        const int ra = T.nd_row_off[2 * g], rb = T.nd_row_off[2 * g + 1], rc = T.nd_row_off[2 * g + 2];
        const int slot = S.pair_slot[pbase + s * P.ncap + t];
        if (slot != 0xFF) {
          const size_t zbase = ((size_t)b * P.slots + slot) * P.ncap;
          const float4* zs = reinterpret_cast<const float4*>(S.z_hist + (zbase + s) * NODE_EMB);
          const float4* zt = reinterpret_cast<const float4*>(S.z_hist + (zbase + t) * NODE_EMB);
          double ms = 0.0, mt = 0.0;
          for (int i = 0; i < NODE_EMB / 4; ++i) {
            const float4 x = zs[i], y = zt[i];
            macc<METRIC>(ms, (double)sh.a_st[4 * i + 0] - (double)x.x);
            macc<METRIC>(ms, (double)sh.a_st[4 * i + 1] - (double)x.y);
            macc<METRIC>(ms, (double)sh.a_st[4 * i + 2] - (double)x.z);
            macc<METRIC>(ms, (double)sh.a_st[4 * i + 3] - (double)x.w);
            macc<METRIC>(mt, (double)sh.a_st[NODE_EMB + 4 * i + 0] - (double)y.x);
            macc<METRIC>(mt, (double)sh.a_st[NODE_EMB + 4 * i + 1] - (double)y.y);
            macc<METRIC>(mt, (double)sh.a_st[NODE_EMB + 4 * i + 2] - (double)y.z);
            macc<METRIC>(mt, (double)sh.a_st[NODE_EMB + 4 * i + 3] - (double)y.w);
          }
          pd = mjoin<METRIC>(ms, mt);
          r0 = (s == t) ? ra : rb;
          cnt = rc - r0;
          // exact ties resolve in table-insertion order (see decode_select_kernel)
          const int opk = P.defender ? (int)S.pair_opos[pbase + s * P.ncap + t] : op;
          const int epoch = P.precise_positions ? (int)S.pair_epoch[pbase + s * P.ncap + t] : slot;
          key = ((uint32_t)epoch << 24) | ((uint32_t)opk << 16) | ((uint32_t)dp << 8) | (uint32_t)op;
          live = cnt > 0;
        }
      }
      const unsigned lmask = __ballot_sync(0xFFFFFFFFu, live);
      const int npairs = __popc(lmask);
      if (npairs == 0) continue;
      const int idx = __popc(lmask & ((1u << lane) - 1u));
      __syncwarp();
      if (live) { sh.p_d[idx] = pd; sh.p_r0[idx] = r0; sh.p_key[idx] = key; sh.p_pre[idx + 1] = cnt; }
      if (lane == 0) sh.p_pre[0] = 0;
      __syncwarp();
      int run = (lane < npairs) ? sh.p_pre[lane + 1] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xFFFFFFFFu, run, o);
        if (lane >= o) run += v;
      }
      __syncwarp();
      if (lane < npairs) sh.p_pre[lane + 1] = run;
      __syncwarp();
      const int total = sh.p_pre[npairs];
      // ---- the staged pairs' candidate rows, flattened over the lanes ----
      int pi = 0;
      for (int j = lane; j < total; j += 32) {
        while (sh.p_pre[pi + 1] <= j) ++pi;
        const int r = sh.p_r0[pi] + (j - sh.p_pre[pi]);
        const uint32_t packed = T.row_packed[r];
        const int kind = (packed >> 20) & 15;
        const uint32_t k = sh.p_key[pi];
        const int s = sh.oorder[k & 0xFF], t = sh.dorder[(k >> 8) & 0xFF];
        if (row_filtered(P, kind, s, t, starter, interest)) continue;
        const int u = packed & 0xFFFFF, oh = (packed >> 24) & 15;
        const double d = mfinish<METRIC>(mjoin<METRIC>(mjoin<METRIC>(sh.p_d[pi], vt_row[u]), sh.d_o[oh]));
        const unsigned long long k64 = ((unsigned long long)k << 32) | (unsigned long long)(unsigned)r;
        if (argmin_better(have, best_d, best_k, d, k64)) { have = true; best_d = d; best_k = k64; best_st = s | (t << 8); best_r = r; }
      }
      __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const bool oh_ = __shfl_xor_sync(0xFFFFFFFFu, (int)have, o) != 0;
      const double od = __shfl_xor_sync(0xFFFFFFFFu, best_d, o);
      const unsigned long long ok = __shfl_xor_sync(0xFFFFFFFFu, best_k, o);
      const int ost = __shfl_xor_sync(0xFFFFFFFFu, best_st, o), orr = __shfl_xor_sync(0xFFFFFFFFu, best_r, o);
      if (oh_ && argmin_better(have, best_d, best_k, od, ok)) { have = true; best_d = od; best_k = ok; best_st = ost; best_r = orr; }
    }
    if (lane == 0) {
      out = make_int4(starter, starter, 0, 0);
      dres = 1.0;
      if (have) {
        const int r = best_r;
        out = make_int4(best_st & 0xFF, best_st >> 8, T.row_ulocal[r], (int)((T.row_packed[r] >> 20) & 15));
        dres = best_d;
      } else {
        atomicExch(S.errflag, 3);   // empty action table: outside the reference's domain
      }
    }
  }
  if (in_range && lane == 0) {
    reinterpret_cast<int4*>(S.sel)[b] = out;
    S.dist[b] = dres;
    if (sel_out) reinterpret_cast<int4*>(sel_out)[b] = out;
    if (dist_out) dist_out[b] = dres;
  }
  // every warp reports once; the last one clears this step's cost bins for the transition after next (as decode_select does)
  __syncwarp();
  if (lane == 0) {
    int32_t* cnt = S.bin_cnt + sched_buf * (SCHED_BINS + 1);     // (no fence: the warp's own reads of the bins returned long ago)
    if (atomicAdd(&cnt[SCHED_BINS], 1) == (int)(gridDim.x * MET_WARPS) - 1) {
#pragma unroll
      for (int k = 0; k <= SCHED_BINS; ++k) cnt[k] = 0;
    }
  }
}

cudaError_t launch_decode_metric(const Tables& T, const Params& P, const State& S, const float* actions, double* vt64, int Ug,
                                 int sched_buf, int32_t* sel_out, double* dist_out, cudaStream_t stream) {
  const dim3 ggrid((P.B + MT_M - 1) / MT_M, (Ug + MT_N - 1) / MT_N);
  const int sgrid = (P.B + MET_WARPS - 1) / MET_WARPS;
  switch (P.metric) {
    case METRIC_L1:
      metric_vuln_kernel<METRIC_L1><<<ggrid, 256, 0, stream>>>(actions, P.act_stride, T.vemb64, vt64, P.B, Ug);
      metric_select_kernel<METRIC_L1><<<sgrid, MET_THREADS, 0, stream>>>(T, P, S, actions, vt64, Ug, sched_buf, sel_out, dist_out);
      break;
    case METRIC_L2:
      metric_vuln_kernel<METRIC_L2><<<ggrid, 256, 0, stream>>>(actions, P.act_stride, T.vemb64, vt64, P.B, Ug);
      metric_select_kernel<METRIC_L2><<<sgrid, MET_THREADS, 0, stream>>>(T, P, S, actions, vt64, Ug, sched_buf, sel_out, dist_out);
      break;
    case METRIC_INF:
      metric_vuln_kernel<METRIC_INF><<<ggrid, 256, 0, stream>>>(actions, P.act_stride, T.vemb64, vt64, P.B, Ug);
      metric_select_kernel<METRIC_INF><<<sgrid, MET_THREADS, 0, stream>>>(T, P, S, actions, vt64, Ug, sched_buf, sel_out, dist_out);
      break;
    default:
      return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace cbs
