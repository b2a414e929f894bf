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

__global__ void __launch_bounds__(256) decode_gemm_simt_kernel(const float* __restrict__ actions, const float* __restrict__ vemb,
                                                               float* __restrict__ vt, int B, int Ug, int vt_stride) {
  __shared__ float As[GT_K][GT_M + 1];
  __shared__ float Bs[GT_K][GT_N + 1];
  const int m0 = blockIdx.x * GT_M, n0 = blockIdx.y * GT_N;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < VULN_EMB; k0 += GT_K) {
    for (int i = threadIdx.x; i < GT_M * GT_K; i += 256) {
      const int r = i / GT_K, k = i % GT_K;
      const int m = m0 + r;
      As[k][r] = m < B ? actions[(size_t)m * ACTION_DIM + 2 * NODE_EMB + k0 + k] : 0.f;
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

cudaError_t launch_decode_gemm_simt(const float* actions, const float* vemb, float* vt, int B, int Ug, int vt_stride,
                                    cudaStream_t stream) {
  dim3 grid((B + GT_M - 1) / GT_M, (Ug + GT_N - 1) / GT_N);
  decode_gemm_simt_kernel<<<grid, 256, 0, stream>>>(actions, vemb, vt, B, Ug, vt_stride);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// candidate scan + float64 re-score + argmin
// ------------------------------------------------------------------------------------------------
constexpr int SEL_WARPS = 4;

struct RowRef { int s, t, slot, r; unsigned long long key; };

__device__ __forceinline__ bool row_filtered(const Params& P, int kind, int s, int t, int starter) {
  if ((s == t && kind == K_LATERAL) || kind == K_CREDACCESS) return true;                         // compressed:532
  if (P.remove_all && (P.goal == GOAL_CONTROL || P.goal == GOAL_DISCOVERY) && kind == K_DOS) return true;   // :536-538
  if (P.remove_main && kind == K_DOS && t == starter) return true;                                // :541-543
  return false;
}

// float64 cosine distance of one table row, warp-cooperative; mirrors scipy's cdist 'cosine'
// (1 - u.v / (|u| |v|) on float64 inputs; the action is float32 widened to float64, compressed:582)
__device__ double exact_distance(const Tables& T, const Params& P, const State& S, const float* __restrict__ act, int b,
                                 const RowRef& rr, double na, int lane) {
  const uint32_t packed = T.row_packed[rr.r];
  const int u = packed & 0xFFFFF, oh = (packed >> 24) & 15;
  const float* zs = S.z_hist + (((size_t)b * P.slots + rr.slot) * P.ncap + rr.s) * NODE_EMB;
  const float* zt = S.z_hist + (((size_t)b * P.slots + rr.slot) * P.ncap + rr.t) * NODE_EMB;
  double dot = 0.0, ne2 = 0.0;
#pragma unroll
  for (int i = lane; i < NODE_EMB; i += 32) {
    const double a = act[i], z = zs[i], a2 = act[NODE_EMB + i], z2 = zt[i];
    dot = fma(a, z, dot); ne2 = fma(z, z, ne2);
    dot = fma(a2, z2, dot); ne2 = fma(z2, z2, ne2);
  }
  const double* v = T.vemb64 + (size_t)u * VULN_EMB;
  const float* av = act + 2 * NODE_EMB;
  for (int i = lane; i < VULN_EMB; i += 32) dot = fma((double)av[i], v[i], dot);
  dot = warp_sum(dot);
  ne2 = warp_sum(ne2);
  dot += (double)act[2 * NODE_EMB + VULN_EMB + oh];
  ne2 += T.vnorm2[u] + 1.0;
  double c = dot / (na * sqrt(ne2));
  if (fabs(c) > 1.0) c = copysign(1.0, c);
  return 1.0 - c;
}

__global__ void __launch_bounds__(SEL_WARPS * 32) decode_select_kernel(Tables T, Params P, State S,
                                                                      const float* __restrict__ actions, int vt_stride,
                                                                      int32_t* __restrict__ sel_out,
                                                                      double* __restrict__ dist_out) {
  __shared__ float s_ao[SEL_WARPS][16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * SEL_WARPS + warp;
  if (b >= P.B) return;
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
  const float* act = actions + (size_t)b * ACTION_DIM;
  // lane l holds channels 2l, 2l+1 of the source and target parts (matches the half2 snapshot layout)
  const float as0 = act[2 * lane], as1 = act[2 * lane + 1];
  const float at0 = act[NODE_EMB + 2 * lane], at1 = act[NODE_EMB + 2 * lane + 1];
  if (lane < OUTCOME_DIM) s_ao[warp][lane] = act[2 * NODE_EMB + VULN_EMB + lane];
  double na2 = 0.0;
  for (int i = lane; i < ACTION_DIM; i += 32) { const double a = act[i]; na2 = fma(a, a, na2); }
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

  // Single pass.  run_max only grows, so a row within `margin` of the final maximum is also within `margin` of
  // run_max when it is visited: every such row is re-scored in float64 on the spot; rows that were only
  // "records so far" cost a few extra re-scores.
  float run_max = -INFINITY;
  double best_d = INFINITY;
  unsigned long long best_key = ~0ull;
  RowRef best{0, 0, 0, -1, ~0ull};

  for (int op = 0; op < n_owned; ++op) {
    const int s = oorder[op];
    for (int dp = 0; dp < n_disc; ++dp) {
      const int t = dorder[dp];
      const int slot = ps[s * P.ncap + t];
      if (slot == 0xFF) continue;
      const __half2* z16 = reinterpret_cast<const __half2*>(S.z16_hist + ((size_t)b * P.slots + slot) * P.ncap * NODE_EMB);
      const float2 zs = __half22float2(z16[s * (NODE_EMB / 2) + lane]);
      const float2 zt = __half22float2(z16[t * (NODE_EMB / 2) + lane]);
      const float st = warp_sum(as0 * zs.x + as1 * zs.y + at0 * zt.x + at1 * zt.y);
      const float* zn = S.zn2_hist + ((size_t)b * P.slots + slot) * P.ncap;
      const float n2 = zn[s] + zn[t] + 1.f;
      const int g = node_off + t;
      const int r0 = (s == t) ? T.nd_row_off[2 * g] : T.nd_row_off[2 * g + 1];
      const int r1 = T.nd_row_off[2 * g + 2];
      for (int base = r0; base < r1; base += 32) {
        const int r = base + lane;
        float score = -INFINITY;
        bool valid = false;
        if (r < r1) {
          const uint32_t packed = T.row_packed[r];
          const int kind = (packed >> 20) & 15;
          if (!row_filtered(P, kind, s, t, starter)) {
            const int u = packed & 0xFFFFF, oh = (packed >> 24) & 15;
            score = (st + vt[u] + s_ao[warp][oh]) * rsqrtf(n2 + (float)T.vnorm2[u]);
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
          RowRef rr;
          rr.s = s; rr.t = t; rr.slot = slot; rr.r = base + src_lane;
          rr.key = ((unsigned long long)slot << 56) | ((unsigned long long)op << 48) | ((unsigned long long)dp << 40) |
                   (unsigned long long)(unsigned)rr.r;
          const double d = exact_distance(T, P, S, act, b, rr, na, lane);
          // np.argmin: the first NaN wins if any distance is NaN, else the first minimum (insertion order)
          const bool dn = d != d, bn = best_d != best_d;
          const bool better = best.r < 0 || (dn ? (!bn || rr.key < best_key)
                                                : (!bn && (d < best_d || (d == best_d && rr.key < best_key))));
          if (better) { best = rr; best_d = d; best_key = rr.key; }
        }
      }
    }
  }
  if (lane == 0) {
    int4 out = make_int4(starter, starter, 0, 0);
    double d = 1.0;
    if (best.r >= 0) {
      const uint32_t packed = T.row_packed[best.r];
      out = make_int4(best.s, best.t, T.vi_ulocal[T.row_inst[best.r]], (int)((packed >> 20) & 15));
      d = best_d;
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
