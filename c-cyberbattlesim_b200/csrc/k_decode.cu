// k_decode.cu — nearest-action decode (find_closest_action_embedding, _env/cyberbattle_env_compressed.py:570-590).
//
// The reference keeps an explicit table of 905-float rows [emb(src) | emb(tgt) | vuln_emb | onehot] and takes
// argmin of scipy's cosine distance (float64).  Here the table is implicit: a row is (processed pair (s,t),
// candidate entry of node t); its dot product with the action factorises as
//     a_s.z_s + a_t.z_t + a_v.v_u + a_o[onehot]      and      |row|^2 = |z_s|^2 + |z_t|^2 + |v_u|^2 + 1 ,
// where z_s, z_t are the embeddings frozen when the pair entered the table (snapshot slot) and a_v.v_u comes from
// one dense contraction  VT[B,Ug] = A_v[B,768] x Vemb[Ug,768]^T  shared by every row (decode_gemm_*).
// decode_select scans the rows in float32 (half-precision copies of the frozen embeddings, TF32/FP32 products
// from the contraction), and re-scores every row within `margin` of the best one in float64 from the float32
// embeddings and the float64 vulnerability table with the same formula scipy uses, so the chosen (s,t,vuln,outcome) and the returned distance match a
// float64 evaluation; ties break on insertion order (epoch, source position, target position, row) like
// np.argmin over the reference's insertion-ordered dict.
#include "cbs_device.cuh"

namespace cbs {

// ------------------------------------------------------------------------------------------------
// SIMT float32 contraction (fallback / validation path for the tcgen05 kernel in k_decode_tc.cu)
// ------------------------------------------------------------------------------------------------
constexpr int GT_M = 64, GT_N = 64, GT_K = 16;

__global__ void __launch_bounds__(256) decode_gemm_simt_kernel(const float* __restrict__ actions, int act_stride,
                                                               const float* __restrict__ vemb, float* __restrict__ vt, int B, int Ug,
                                                               int vt_stride) {
  __shared__ float As[GT_K][GT_M + 1];
  __shared__ float Bs[GT_K][GT_N + 1];
  const int m0 = blockIdx.x * GT_M, n0 = blockIdx.y * GT_N;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < VULN_EMB; k0 += GT_K) {
    for (int i = threadIdx.x; i < GT_M * GT_K; i += 256) {
      const int r = i / GT_K, k = i % GT_K;
      const int m = m0 + r;
      As[k][r] = m < B ? actions[(size_t)m * act_stride + 2 * NODE_EMB + k0 + k] : 0.f;
    }
    for (int i = threadIdx.x; i < GT_N * GT_K; i += 256) {
      const int r = i / GT_K, k = i % GT_K;
      const int n = n0 + r;
      Bs[k][r] = n < Ug ? vemb[(size_t)n * VULN_EMB + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GT_K; ++k) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; bb[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
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
      if (n < Ug) vt[(size_t)m * vt_stride + n] = acc[i][j];
    }
  }
}

cudaError_t launch_decode_gemm_simt(const float* actions, int act_stride, const float* vemb, float* vt, int B, int Ug,
                                    int vt_stride, cudaStream_t stream) {
  dim3 grid((B + GT_M - 1) / GT_M, (Ug + GT_N - 1) / GT_N);
  decode_gemm_simt_kernel<<<grid, 256, 0, stream>>>(actions, act_stride, vemb, vt, B, Ug, vt_stride);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// candidate scan + float64 re-score + argmin
// ------------------------------------------------------------------------------------------------
constexpr int SEL_WARPS = 4;

__device__ __forceinline__ bool row_filtered(const Params& P, int kind, int s, int t, int starter) {
  if ((s == t && kind == K_LATERAL) || kind == K_CREDACCESS) return true;                         // compressed:532
  if (P.remove_all && (P.goal == GOAL_CONTROL || P.goal == GOAL_DISCOVERY) && kind == K_DOS) return true;   // :536-538
  if (P.remove_main && kind == K_DOS && t == starter) return true;                                // :541-543
  return false;
}

// Per-lane slice of the action kept in registers for the float64 re-score: lane l owns vulnerability-part elements
// 128*i + 4*l + j (i < 6, j < 4) and channels 2l, 2l+1 of the source / target parts.
struct LaneAction {
  float av[24];
  float2 as, at;
};

// float64 cosine distance of one table row, warp-cooperative; mirrors scipy's cdist 'cosine'
// (1 - u.v / (|u| |v|) on float64 inputs; the action is float32 widened to float64, compressed:582)
__device__ __forceinline__ double exact_distance(const Tables& T, const Params& P, const State& S, const LaneAction& A,
                                                 const float* s_ao, int b, int s, int t, int slot, int r, double na, int lane) {
  const uint32_t packed = T.row_packed[r];
  const int u = packed & 0xFFFFF, oh = (packed >> 24) & 15;
  const size_t zbase = ((size_t)b * P.slots + slot) * P.ncap;
  const float2 zs = reinterpret_cast<const float2*>(S.z_hist + (zbase + s) * NODE_EMB)[lane];
  const float2 zt = reinterpret_cast<const float2*>(S.z_hist + (zbase + t) * NODE_EMB)[lane];
  const double2* v = reinterpret_cast<const double2*>(T.vemb64 + (size_t)u * VULN_EMB) + 2 * lane;
  double2 vv[12];
#pragma unroll
  for (int i = 0; i < 6; ++i) { vv[2 * i] = v[64 * i]; vv[2 * i + 1] = v[64 * i + 1]; }
  double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    d0 = fma((double)A.av[4 * i + 0], vv[2 * i].x, d0);
    d1 = fma((double)A.av[4 * i + 1], vv[2 * i].y, d1);
    d2 = fma((double)A.av[4 * i + 2], vv[2 * i + 1].x, d2);
    d3 = fma((double)A.av[4 * i + 3], vv[2 * i + 1].y, d3);
  }
  d0 = fma((double)A.as.x, (double)zs.x, d0);
  d1 = fma((double)A.as.y, (double)zs.y, d1);
  d2 = fma((double)A.at.x, (double)zt.x, d2);
  d3 = fma((double)A.at.y, (double)zt.y, d3);
  double ne2 = fma((double)zs.x, (double)zs.x, (double)zs.y * (double)zs.y) +
               fma((double)zt.x, (double)zt.x, (double)zt.y * (double)zt.y);
  double dot = warp_sum((d0 + d1) + (d2 + d3));
  ne2 = warp_sum(ne2);
  dot += (double)s_ao[oh];
  ne2 += T.vnorm2[u] + 1.0;
  double c = dot / (na * sqrt(ne2));
  if (fabs(c) > 1.0) c = copysign(1.0, c);
  return 1.0 - c;
}

constexpr int CAND_CAP = 8;

struct SelShared {
  float a_st[2 * NODE_EMB];   // source | target parts of the action
  float a_o[16];
  float p_st[32], p_n2[32];   // per staged pair: a_s.z_s + a_t.z_t (from the fp16 snapshots), |z_s|^2 + |z_t|^2 + 1
  int p_r0[32], p_pre[33];    // first candidate row, exclusive prefix of row counts
  uint32_t p_key[32];         // slot << 16 | owned position << 8 | discovered position
  float c_score[CAND_CAP];    // rows still within `margin` of the running maximum, waiting for the float64 re-score
  uint32_t c_key[CAND_CAP];
  int c_row[CAND_CAP];
};

__device__ __forceinline__ float dot8(const uint4 h, const float* a) {
  const __half2* q = reinterpret_cast<const __half2*>(&h);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(q[i]);
    acc = fmaf(f.x, a[2 * i], acc);
    acc = fmaf(f.y, a[2 * i + 1], acc);
  }
  return acc;
}

struct Best {
  double d;
  unsigned long long key;
  int s, t, r;
};

// re-score the pending candidates that are still within `margin` of run_max; empties the list
__device__ __forceinline__ void flush_candidates(const Tables& T, const Params& P, const State& S, const LaneAction& A,
                                                 SelShared& sh, int& ncand, float threshold, int b, double na, int lane,
                                                 const uint8_t* oorder, const uint8_t* dorder, Best& best) {
  for (int i = 0; i < ncand; ++i) {
    if (sh.c_score[i] < threshold) continue;               // NaN scores are kept (comparison false)
    const uint32_t k = sh.c_key[i];
    const int r = sh.c_row[i];
    const int slot = (int)(k >> 16), s = oorder[(k >> 8) & 0xFF], t = dorder[k & 0xFF];
    const unsigned long long key = ((unsigned long long)k << 40) | (unsigned long long)(unsigned)r;
    const double d = exact_distance(T, P, S, A, sh.a_o, b, s, t, slot, r, na, lane);
    // np.argmin: the first NaN wins if any distance is NaN, else the first minimum (insertion order)
    const bool dn = d != d, bn = best.d != best.d;
    const bool better = best.r < 0 || (dn ? (!bn || key < best.key) : (!bn && (d < best.d || (d == best.d && key < best.key))));
    if (better) { best.d = d; best.key = key; best.s = s; best.t = t; best.r = r; }
  }
  ncand = 0;
}

// Table scan.  Pairs (owned source x discovered target, in insertion order) are staged 32 at a time: lane-per-pair
// for the two 64-wide dot products against the half-precision snapshot rows, then the candidate rows of the 32
// staged pairs are flattened over the lanes (binary search in the row-count prefix), so short and long candidate
// lists cost the same per row.  Rows whose float32 score is within `margin` of the running maximum are parked in a
// small list; the list is pruned as the maximum grows and its survivors are re-scored in float64 (every row
// within `margin` of the FINAL maximum is guaranteed to be among them, because the running maximum only grows).
__global__ void __launch_bounds__(SEL_WARPS * 32) decode_select_kernel(Tables T, Params P, State S,
                                                                      const float* __restrict__ actions, int vt_stride,
                                                                      int32_t* __restrict__ sel_out,
                                                                      double* __restrict__ dist_out) {
  __shared__ SelShared sh_all[SEL_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * SEL_WARPS + warp;
  if (b >= P.B) return;
  SelShared& sh = sh_all[warp];
  const int flags = scalar(S, P, S_FLAGS, b);
  if (flags & (FL_DONE | FL_TRUNC | FL_NEEDS_RESET)) {
    if (lane == 0) {
      reinterpret_cast<int4*>(S.sel)[b] = make_int4(0, 0, 0, 0);
      S.dist[b] = 0.0;
      if (sel_out) reinterpret_cast<int4*>(sel_out)[b] = make_int4(0, 0, 0, 0);
      if (dist_out) dist_out[b] = 0.0;
    }
    return;
  }
  const float* act = actions + (size_t)b * P.act_stride;
  LaneAction A;
  double na2 = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) A.av[4 * i + j] = act[2 * NODE_EMB + 128 * i + 4 * lane + j];
  A.as = make_float2(act[2 * lane], act[2 * lane + 1]);
  A.at = make_float2(act[NODE_EMB + 2 * lane], act[NODE_EMB + 2 * lane + 1]);
  const float ao = lane < OUTCOME_DIM ? act[2 * NODE_EMB + VULN_EMB + lane] : 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) na2 = fma((double)A.av[i], (double)A.av[i], na2);
  na2 = fma((double)A.as.x, (double)A.as.x, na2); na2 = fma((double)A.as.y, (double)A.as.y, na2);
  na2 = fma((double)A.at.x, (double)A.at.x, na2); na2 = fma((double)A.at.y, (double)A.at.y, na2);
  na2 = fma((double)ao, (double)ao, na2);
  sh.a_st[2 * lane] = A.as.x; sh.a_st[2 * lane + 1] = A.as.y;
  sh.a_st[NODE_EMB + 2 * lane] = A.at.x; sh.a_st[NODE_EMB + 2 * lane + 1] = A.at.y;
  if (lane < 16) sh.a_o[lane] = ao;
  na2 = warp_sum(na2);
  const double na = sqrt(na2);
  __syncwarp();

  const int sc = scalar(S, P, S_SCENARIO, b);
  const int node_off = T.sc_node_off[sc];
  const int starter = scalar(S, P, S_STARTER, b);
  const int n_disc = scalar(S, P, S_N_DISC, b), n_owned = scalar(S, P, S_N_OWNED, b);
  const uint8_t* dorder = S.disc_order + (size_t)b * P.ncap;
  const uint8_t* oorder = S.owned_order + (size_t)b * P.ncap;
  const uint8_t* ps = S.pair_slot + (size_t)b * P.ncap * P.ncap;
  const float* vt = S.vt + (size_t)b * vt_stride;
  const float margin_s = P.margin * (float)na;

  float run_max = -INFINITY;
  Best best{INFINITY, ~0ull, 0, 0, -1};
  int ncand = 0;

  const int combos = n_owned * n_disc;
  for (int cbase = 0; cbase < combos; cbase += 32) {
    // ---- phase A: one (source, target) combination per lane ----
    const int c = cbase + lane;
    bool live = false;
    float st = 0.f, n2 = 0.f;
    int r0 = 0, cnt = 0;
    uint32_t key = 0;
    if (c < combos) {
      const int op = c / n_disc, dp = c - op * n_disc;
      const int s = oorder[op], t = dorder[dp];
      const int slot = ps[s * P.ncap + t];
      if (slot != 0xFF) {
        const size_t zbase = ((size_t)b * P.slots + slot) * P.ncap;
        const uint4* zs = reinterpret_cast<const uint4*>(S.z16_hist + (zbase + s) * NODE_EMB);
        const uint4* zt = reinterpret_cast<const uint4*>(S.z16_hist + (zbase + t) * NODE_EMB);
        uint4 hs[NODE_EMB / 8], ht[NODE_EMB / 8];
#pragma unroll
        for (int i = 0; i < NODE_EMB / 8; ++i) { hs[i] = zs[i]; ht[i] = zt[i]; }
        n2 = S.zn2_hist[zbase + s] + S.zn2_hist[zbase + t] + 1.f;
        const int g = node_off + t;
        r0 = (s == t) ? T.nd_row_off[2 * g] : T.nd_row_off[2 * g + 1];
        cnt = T.nd_row_off[2 * g + 2] - r0;
#pragma unroll
        for (int i = 0; i < NODE_EMB / 8; ++i) {
          st += dot8(hs[i], sh.a_st + 8 * i);
          st += dot8(ht[i], sh.a_st + NODE_EMB + 8 * i);
        }
        key = ((uint32_t)slot << 16) | ((uint32_t)op << 8) | (uint32_t)dp;
        live = cnt > 0;
      }
    }
    const unsigned lmask = __ballot_sync(0xFFFFFFFFu, live);
    const int npairs = __popc(lmask);
    if (npairs == 0) continue;
    const int idx = __popc(lmask & ((1u << lane) - 1u));
    if (live) { sh.p_st[idx] = st; sh.p_n2[idx] = n2; sh.p_r0[idx] = r0; sh.p_key[idx] = key; sh.p_pre[idx + 1] = cnt; }
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

    // ---- phase B: candidate rows of the staged pairs, flattened over the lanes ----
    for (int j0 = 0; j0 < total; j0 += 32) {
      const int j = j0 + lane;
      float score = -INFINITY;
      bool valid = false;
      int pi = 0, r = 0;
      if (j < total) {
        int lo = 0, hi = npairs;           // largest pi with p_pre[pi] <= j
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (sh.p_pre[mid] <= j) lo = mid; else hi = mid;
        }
        pi = lo;
        r = sh.p_r0[pi] + (j - sh.p_pre[pi]);
        const uint32_t packed = T.row_packed[r];
        const int kind = (packed >> 20) & 15;
        const uint32_t k = sh.p_key[pi];
        const int s = oorder[(k >> 8) & 0xFF], t = dorder[k & 0xFF];
        if (!row_filtered(P, kind, s, t, starter)) {
          const int u = packed & 0xFFFFF, oh = (packed >> 24) & 15;
          score = (sh.p_st[pi] + vt[u] + sh.a_o[oh]) * rsqrtf(sh.p_n2[pi] + (float)T.vnorm2[u]);
          valid = true;
        }
      }
      float cmax = valid ? score : -INFINITY;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cmax = fmaxf(cmax, __shfl_xor_sync(0xFFFFFFFFu, cmax, o));
      run_max = fmaxf(run_max, cmax);                       // fmaxf drops NaN; NaN rows are candidates below
      unsigned cand = __ballot_sync(0xFFFFFFFFu, valid && !(score < run_max - margin_s));
      while (cand) {
        const int src_lane = __ffs(cand) - 1;
        cand &= cand - 1;
        if (ncand == CAND_CAP) flush_candidates(T, P, S, A, sh, ncand, run_max - margin_s, b, na, lane, oorder, dorder, best);
        const float cs = __shfl_sync(0xFFFFFFFFu, score, src_lane);
        const int cpi = __shfl_sync(0xFFFFFFFFu, pi, src_lane);
        const int cr = __shfl_sync(0xFFFFFFFFu, r, src_lane);
        if (lane == 0) { sh.c_score[ncand] = cs; sh.c_key[ncand] = sh.p_key[cpi]; sh.c_row[ncand] = cr; }
        ++ncand;
        __syncwarp();
      }
    }
    __syncwarp();
  }
  flush_candidates(T, P, S, A, sh, ncand, run_max - margin_s, b, na, lane, oorder, dorder, best);
  if (lane == 0) {
    int4 out = make_int4(starter, starter, 0, 0);
    double d = 1.0;
    if (best.r >= 0) {
      const uint32_t packed = T.row_packed[best.r];
      out = make_int4(best.s, best.t, T.vi_ulocal[T.row_inst[best.r]], (int)((packed >> 20) & 15));
      d = best.d;
    } else {
      atomicExch(S.errflag, 3);   // empty action table: outside the reference's domain (cdist would raise)
    }
    reinterpret_cast<int4*>(S.sel)[b] = out;
    S.dist[b] = d;
    if (sel_out) reinterpret_cast<int4*>(sel_out)[b] = out;
    if (dist_out) dist_out[b] = d;
  }
}

cudaError_t launch_decode_select(const Tables& T, const Params& P, const State& S, const float* actions, int vt_stride,
                                 int32_t* sel_out, double* dist_out, cudaStream_t stream) {
  const int grid = (P.B + SEL_WARPS - 1) / SEL_WARPS;
  decode_select_kernel<<<grid, SEL_WARPS * 32, 0, stream>>>(T, P, S, actions, vt_stride, sel_out, dist_out);
  return cudaGetLastError();
}

}  // namespace cbs
