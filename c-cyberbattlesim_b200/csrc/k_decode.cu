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
#include <cstdlib>

#include "transition.cuh"
#include "subset.cuh"

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
// row_filtered(): cbs_device.cuh

// ------------------------------------------------------------------------------------------------
// decode_select: one warp per env.  Measured with a per-env cycle trace (tools/select_trace.py): the kernel is a
// chain of dependent memory round trips (~1500-2000 cycles each under load), not bandwidth or issue bound, and its
// duration is set by the envs with the longest action tables (max ~3200 rows, mean ~270).  Hence:
//   * everything that does not depend on other loads is requested in one prologue burst (scalars, the two order
//     lists, the env's VT row -> shared memory, the action -> shared memory + its norm);
//   * the candidate-row templates are prefetched one trip ahead and the VT / norm gathers hit shared memory, so a
//     trip of 128 rows costs no exposed global round trip;
//   * parked candidates carry their row template, so the float64 re-score starts its loads immediately.
// Narrower per-env tiles were tried (4 / 8 / 16 lanes: 428 / 212 / 121 us against 83 us for a full warp): lanes per
// env matter more than envs in flight.
// ------------------------------------------------------------------------------------------------
#ifndef CBS_SEL_MINB
#define CBS_SEL_MINB 5   // 96 registers, 20 warps per SM: measured best (4: 78 us, 5: 76, 6: 84, 7: 94, 8: 104; no bound = 215 registers: 117)
#endif
#ifndef CBS_SEL_WARPS
#define CBS_SEL_WARPS 4
#endif
constexpr int SEL_WARPS = CBS_SEL_WARPS;
constexpr int SEL_THREADS = SEL_WARPS * 32;
constexpr int CAND_CAP = 16;
constexpr int RPL = 4;          // rows per lane per trip
constexpr int SEL_VT_SMEM_MAX = 2048;   // (VT[u], |v_u|^2) pairs cached per warp (larger pools fall back to global gathers for the rest).
                                        // The CTA's share is (2 SEL_WARPS + 1) x 4 bytes per cached entry: five CTAs per SM fit up to
                                        // ~1000 global vulnerabilities (the measured workloads: 200 and 600), four up to ~1300, two at 2048.

struct SelWarp {
  float a_st[2 * NODE_EMB];   // source | target parts of the action
  float a_o[16];
  float2 p_sn[32];            // per staged pair: (a_s.z_s + a_t.z_t from the fp16 snapshots, |z_s|^2 + |z_t|^2 + 1)
  uint32_t p_fm[32];          // per staged pair: bit k set = rows of outcome kind k never enter the table (row_filtered)
  int p_r0[32], p_pre[33];    // first candidate row, exclusive prefix of row counts
  uint32_t p_key[32];         // insertion epoch << 24 | source's insertion position << 16 | discovered position << 8 | index in owned_order
  int p_slot[32];             // snapshot slot the pair's rows carry (== the epoch unless precise_action_space_positions refreshed them)
  float c_score[CAND_CAP];    // rows still within `margin` of the running maximum, waiting for the float64 re-score
  unsigned long long c_key[CAND_CAP];   // place in the table's order (np.argmin takes the first of equal distances)
  uint32_t c_packed[CAND_CAP];
  int c_row[CAND_CAP], c_slot[CAND_CAP], c_st[CAND_CAP];   // c_st: source | target << 8
  uint8_t oorder[MAX_NODES], dorder[MAX_NODES];
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

// float64 cosine distance of one table row, warp-cooperative; mirrors scipy's cdist 'cosine'
// (1 - u.v / (|u| |v|) on float64 inputs; the action is float32 widened to float64, compressed:582)
__device__ __forceinline__ double exact_distance(const Tables& T, const Params& P, const State& S, const float* __restrict__ act,
                                                 int b, int s, int t, int slot, uint32_t packed, double na, int lane) {
  const int u = packed & 0xFFFFF, oh = (packed >> 24) & 15;
  const size_t zbase = ((size_t)b * P.slots + slot) * P.ncap;
  const float2 zs = reinterpret_cast<const float2*>(S.z_hist + (zbase + s) * NODE_EMB)[lane];
  const float2 zt = reinterpret_cast<const float2*>(S.z_hist + (zbase + t) * NODE_EMB)[lane];
  const double2* v = reinterpret_cast<const double2*>(T.vemb64 + (size_t)u * VULN_EMB) + 2 * lane;
  const float* av = act + 2 * NODE_EMB + 4 * lane;
  double2 vv[12];
  float aa[24];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    vv[2 * i] = v[64 * i]; vv[2 * i + 1] = v[64 * i + 1];
#pragma unroll
    for (int j = 0; j < 4; ++j) aa[4 * i + j] = av[128 * i + j];
  }
  const float as0 = act[2 * lane], as1 = act[2 * lane + 1], at0 = act[NODE_EMB + 2 * lane], at1 = act[NODE_EMB + 2 * lane + 1];
  const float ao = act[2 * NODE_EMB + VULN_EMB + oh];
  const double vn2 = T.vnorm2[u];
  double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    d0 = fma((double)aa[4 * i + 0], vv[2 * i].x, d0);
    d1 = fma((double)aa[4 * i + 1], vv[2 * i].y, d1);
    d2 = fma((double)aa[4 * i + 2], vv[2 * i + 1].x, d2);
    d3 = fma((double)aa[4 * i + 3], vv[2 * i + 1].y, d3);
  }
  d0 = fma((double)as0, (double)zs.x, d0);
  d1 = fma((double)as1, (double)zs.y, d1);
  d2 = fma((double)at0, (double)zt.x, d2);
  d3 = fma((double)at1, (double)zt.y, d3);
  double ne2 = fma((double)zs.x, (double)zs.x, (double)zs.y * (double)zs.y) +
               fma((double)zt.x, (double)zt.x, (double)zt.y * (double)zt.y);
  double dot = warp_sum((d0 + d1) + (d2 + d3));
  ne2 = warp_sum(ne2);
  dot += (double)ao;
  ne2 += vn2 + 1.0;
  double c = dot / (na * sqrt(ne2));
  if (fabs(c) > 1.0) c = copysign(1.0, c);
  return 1.0 - c;
}

struct Best {
  double d;
  unsigned long long key;
  int s, t, r;
  uint32_t packed;
  float score32;      // the float32 scan score of the row that won the float64 re-score (margin diagnostics)
};

// re-score the pending candidates that are still within `margin` of run_max; empties the list
__device__ __forceinline__ void flush_candidates(const Tables& T, const Params& P, const State& S, const float* act, SelWarp& sh,
                                                 int& ncand, float threshold, int b, double na, int lane, Best& best) {
  for (int i = 0; i < ncand; ++i) {
    if (sh.c_score[i] < threshold) continue;               // NaN scores are kept (comparison false)
    const unsigned long long key = sh.c_key[i];
    const int r = sh.c_row[i];
    const int slot = sh.c_slot[i], s = sh.c_st[i] & 0xFF, t = sh.c_st[i] >> 8;
    const double d = exact_distance(T, P, S, act, b, s, t, slot, sh.c_packed[i], na, lane);
    // np.argmin: the first NaN wins if any distance is NaN, else the first minimum (insertion order)
    const bool dn = d != d, bn = best.d != best.d;
    const bool better = best.r < 0 || (dn ? (!bn || key < best.key) : (!bn && (d < best.d || (d == best.d && key < best.key))));
    if (better) { best.d = d; best.key = key; best.s = s; best.t = t; best.r = r; best.packed = sh.c_packed[i]; best.score32 = sh.c_score[i]; }
  }
  ncand = 0;
}

// Once per CTA, after the barrier that follows its schedule lookup: the last CTA to get there clears the cost bins for the next
// transition.  Every CTA counts itself only after its own reads of the bins have returned (their values addressed the loads in
// front of the barrier), so the clear needs no fence — the per-warp fence + atomic at the kernel's end was 4 % of its warp time.
__device__ __forceinline__ void sched_done(const State& S, int buf) {
  if (threadIdx.x == 0) {
    int32_t* cnt = S.bin_cnt + buf * (SCHED_BINS + 1);
    if (atomicAdd(&cnt[SCHED_BINS], 1) == (int)gridDim.x - 1) {
#pragma unroll
      for (int k = 0; k <= SCHED_BINS; ++k) cnt[k] = 0;
    }
  }
}

// what the fused kernel needs to run the transition right after the decode (cbs_step)
struct FusedTransition {
  int enabled;
  const float* uniforms;
  float* reward;
  uint8_t* done;
};

// W1: one-word mask planes and no defender -> the fused transition stages the env's records in registers
// SUBSET: sample_subset_samples — the table is the env's explicit per-class row lists (subset.cuh), scanned one row per lane
template <bool FUSE, bool DEF, bool W1, bool SUBSET>
__global__ void __launch_bounds__(SEL_THREADS, CBS_SEL_MINB) decode_select_kernel(Tables T, Params P, State S, const float* __restrict__ actions,
                                                                   int vt_stride, int vt_cached, int sched_buf, FusedTransition ft,
                                                                   int32_t* __restrict__ sel_out, double* __restrict__ dist_out,
                                                                   long long* __restrict__ trace) {
  extern __shared__ __align__(16) unsigned char sel_smem[];
  SelWarp* sh_all = reinterpret_cast<SelWarp*>(sel_smem);
  float* vn2_sh = reinterpret_cast<float*>(sel_smem + SEL_WARPS * sizeof(SelWarp));          // [vt_cached] shared by the CTA
  float2* vtn_all = reinterpret_cast<float2*>(vn2_sh + vt_cached);                            // [SEL_WARPS][vt_cached] (VT[u], |v_u|^2)
  const long long t_begin = clock64();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Programmatic dependent launch (cbs_types.h): this CTA may have been scheduled while the contraction that produces VT is still
  // draining; nothing of the env state or VT is touched before pdl_wait().  The next kernel (observe) stages its constant
  // weights while this grid's last wave runs.
  pdl_trigger();
  for (int i = threadIdx.x; i < vt_cached; i += SEL_THREADS) vn2_sh[i] = i < T.num_global_vulns ? (float)T.vnorm2[i] : 1.f;
  pdl_wait();
  __syncthreads();     // vn2_sh is complete before the prologues below interleave it with the env's VT row
  // Longest-first schedule: the transition kernel of the previous step binned every env by the size of its action
  // table; warp w takes the w-th env counting from the heaviest bin.  The kernel's duration is set by the largest
  // tables, so they must start first.  (No complete binning -> identity order.)
  int b = blockIdx.x * SEL_WARPS + warp;
  {
    int cnt[SCHED_BINS], tot = 0;
#pragma unroll
    for (int k = 0; k < SCHED_BINS; ++k) { cnt[k] = S.bin_cnt[sched_buf * (SCHED_BINS + 1) + k]; tot += cnt[k]; }
    if (tot == P.B && b < P.B) {
      int w = b;
#pragma unroll
      for (int k = SCHED_BINS - 1; k >= 0; --k) {
        if (w >= 0 && w < cnt[k]) { b = S.bin_list[((size_t)sched_buf * SCHED_BINS + k) * P.B + w]; w = -1; }
        else if (w >= 0) w -= cnt[k];
      }
    }
  }
  const bool in_range = b < P.B;
  SelWarp& sh = sh_all[warp];
  float2* vtn = vtn_all + (size_t)warp * vt_cached;

  // ---- prologue: one burst of independent loads ----
  int flags = FL_NEEDS_RESET, node_off = 0, starter = 0, n_disc = 0, n_owned = 0;
  const float* act = actions;
  double na = 0.0;
  if (in_range) {
    flags = scalar(S, P, S_FLAGS, b);
    node_off = scalar(S, P, S_NODE_OFF, b);
    starter = scalar(S, P, S_STARTER, b);
    n_disc = scalar(S, P, S_N_DISC, b);
    n_owned = scalar(S, P, S_N_OWNED, b);
    act = actions + (size_t)b * P.act_stride;
    const uint32_t* dsrc = reinterpret_cast<const uint32_t*>(S.disc_order + (size_t)b * P.ncap);
    const uint32_t* osrc = reinterpret_cast<const uint32_t*>(S.owned_order + (size_t)b * P.ncap);
    if (lane < P.ncap / 4) {
      reinterpret_cast<uint32_t*>(sh.dorder)[lane] = dsrc[lane];
      reinterpret_cast<uint32_t*>(sh.oorder)[lane] = osrc[lane];
    }
    const float* vt_row = S.vt + (size_t)b * vt_stride;
    for (int i = lane; i < vt_cached; i += 32) vtn[i] = make_float2(vt_row[i], vn2_sh[i]);
    float a[29];
#pragma unroll
    for (int i = 0; i < 29; ++i) { const int e = lane + 32 * i; a[i] = e < ACTION_DIM ? act[e] : 0.f; }
    double n0 = 0.0, n1 = 0.0;
#pragma unroll
    for (int i = 0; i < 29; ++i) {
      const int e = lane + 32 * i;
      if (i < 4) sh.a_st[e] = a[i];                                     // e < 128
      if (i >= 28 && e >= 2 * NODE_EMB + VULN_EMB && e < ACTION_DIM) sh.a_o[e - 2 * NODE_EMB - VULN_EMB] = a[i];
      if (i & 1) n1 = fma((double)a[i], (double)a[i], n1); else n0 = fma((double)a[i], (double)a[i], n0);
    }
    na = sqrt(warp_sum(n0 + n1));
  }
  __syncthreads();
  sched_done(S, sched_buf);
  // a finished env (the reference raises there, cyberbattle_env.py:300-302) decodes to zeros
  const bool active = in_range && !(flags & (FL_DONE | FL_TRUNC | FL_NEEDS_RESET));
  int4 out = make_int4(0, 0, 0, 0);      // lane 0 holds the env's result
  double d = 0.0;
  if (active) {
  const uint8_t* ps = S.pair_slot + (size_t)b * P.ncap * P.ncap;
  const float* vt_g = S.vt + (size_t)b * vt_stride;
  const float margin_s = P.margin * (float)na;
  const int interest = is_node_goal(P) ? T.sc_interest[scalar(S, P, S_SCENARIO, b)] : -1;

  float run_max = -INFINITY;
  Best best{INFINITY, ~0ull, 0, 0, -1, 0u, 0.f};
  int ncand = 0, n_rows = 0, n_live = 0, n_exact = 0;

  const int combos = SUBSET ? 0 : n_owned * n_disc;
  if constexpr (SUBSET) {
    // ---- the explicit table: at most SUB_CLASSES x subset_k rows.  Lane-per-row: entry -> snapshot rows of its pair (half
    //      precision) -> two 64-wide dots, the row template, score; candidates are parked for the float64 re-score as below ----
    const int K = P.subset_k;
    const uint32_t* lists = S.sub_rows + (size_t)b * SUB_CLASSES * K;
    const int32_t* meta = S.sub_meta + (size_t)b * SUB_META;
    int cls_end[SUB_CLASSES];
    int total = 0;
#pragma unroll
    for (int c = 0; c < SUB_CLASSES; ++c) { total += meta[c]; cls_end[c] = total; }
    const uint32_t rank_lo = (uint32_t)meta[11], rank_hi = (uint32_t)meta[12];
    // positions of the nodes in the two order lists (tie order); p_sn (256 bytes) doubles as two byte arrays here
    static_assert(sizeof(SelWarp::p_sn) >= 2 * MAX_NODES, "p_sn holds two MAX_NODES byte arrays in subset mode");
    uint8_t* dpos = reinterpret_cast<uint8_t*>(sh.p_sn);
    uint8_t* opos = dpos + MAX_NODES;
    for (int i = lane; i < n_disc; i += 32) dpos[sh.dorder[i]] = (uint8_t)i;
    if (!DEF) for (int i = lane; i < n_owned; i += 32) opos[sh.oorder[i]] = (uint8_t)i;
    __syncwarp();
    n_rows = total;
    for (int j0 = 0; j0 < total; j0 += 32) {
      const int j = j0 + lane;
      float score = -INFINITY;
      bool valid = false;
      int s = 0, t = 0, slot = 0, r = 0;
      uint32_t packed = 0u;
      unsigned long long key = 0ull;
      if (j < total) {
        int c = 0, first = 0;
#pragma unroll
        for (int q = 0; q < SUB_CLASSES - 1; ++q) if (j >= cls_end[q]) { c = q + 1; first = cls_end[q]; }
        const uint32_t e = lists[c * K + (j - first)];
        s = e & 127; t = (e >> 7) & 127;
        const int rip = (e >> 14) & 255, epoch = (int)(e >> 22);
        slot = ps[s * P.ncap + t];
        const int g = node_off + t;
        r = T.nd_row_off[2 * g + (s == t ? 0 : 1)] + rip;
        packed = T.row_packed[r];
        const size_t zbase = ((size_t)b * P.slots + slot) * P.ncap;
        const uint4* zs = reinterpret_cast<const uint4*>(S.z16_hist + (zbase + s) * NODE_EMB);
        const uint4* zt = reinterpret_cast<const uint4*>(S.z16_hist + (zbase + t) * NODE_EMB);
        uint4 hs[NODE_EMB / 8], ht[NODE_EMB / 8];
#pragma unroll
        for (int i = 0; i < NODE_EMB / 8; ++i) { hs[i] = zs[i]; ht[i] = zt[i]; }
        const float n2 = S.zn2_hist[zbase + s] + S.zn2_hist[zbase + t] + 1.f;
        float st = 0.f;
#pragma unroll
        for (int i = 0; i < NODE_EMB / 8; ++i) {
          st += dot8(hs[i], sh.a_st + 8 * i);
          st += dot8(ht[i], sh.a_st + NODE_EMB + 8 * i);
        }
        const int u = packed & 0xFFFFF, oh = (packed >> 24) & 15;
        const float2 vv = u < vt_cached ? vtn[u] : make_float2(vt_g[u], (float)T.vnorm2[u]);
        score = (st + vv.x + sh.a_o[oh]) * rsqrtf(n2 + vv.y);
        valid = true;
        // table order: the classes in first-appearance order, inside a class the insertion order
        const uint32_t rank = ((c < 8 ? rank_lo : rank_hi) >> (4 * (c & 7))) & 15u;
        const int opk = DEF ? (int)S.pair_opos[(size_t)b * P.ncap * P.ncap + s * P.ncap + t] : (int)opos[s];
        key = ((unsigned long long)rank << 32) |
              (((unsigned long long)epoch << 24) | ((unsigned long long)opk << 16) | ((unsigned long long)dpos[t] << 8) | (unsigned long long)rip);
      }
      float cmax = score;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cmax = fmaxf(cmax, __shfl_xor_sync(0xFFFFFFFFu, cmax, o));
      run_max = fmaxf(run_max, cmax);
      unsigned cand = __ballot_sync(0xFFFFFFFFu, valid && !(score < run_max - margin_s));
      while (cand) {
        const int src = __ffs(cand) - 1;
        cand &= cand - 1;
        if (ncand == CAND_CAP) { n_exact += ncand; flush_candidates(T, P, S, act, sh, ncand, run_max - margin_s, b, na, lane, best); }
        const float cs = __shfl_sync(0xFFFFFFFFu, score, src);
        const int cst = __shfl_sync(0xFFFFFFFFu, s | (t << 8), src);
        const int cslot = __shfl_sync(0xFFFFFFFFu, slot, src);
        const int cr = __shfl_sync(0xFFFFFFFFu, r, src);
        const uint32_t cp = __shfl_sync(0xFFFFFFFFu, packed, src);
        const unsigned long long ck = __shfl_sync(0xFFFFFFFFu, key, src);
        if (lane == 0) { sh.c_score[ncand] = cs; sh.c_key[ncand] = ck; sh.c_st[ncand] = cst; sh.c_slot[ncand] = cslot; sh.c_row[ncand] = cr; sh.c_packed[ncand] = cp; }
        ++ncand;
        __syncwarp();
      }
    }
    __syncwarp();
  }
  for (int cbase = 0; cbase < combos; cbase += 32) {
    // ---- phase A: one (source, target) combination per lane ----
    const int c = cbase + lane;
    bool live = false;
    float st = 0.f, n2 = 0.f;
    int r0 = 0, cnt = 0, slot_of_pair = 0;
    uint32_t key = 0, fm = 0;
    if (c < combos) {
      const int op = c / n_disc, dp = c - op * n_disc;
      const int s = sh.oorder[op], t = sh.dorder[dp];
      const int g = node_off + t;
      const int ra = T.nd_row_off[2 * g], rb = T.nd_row_off[2 * g + 1], rc = T.nd_row_off[2 * g + 2];   // independent of the slot
      const int slot = ps[s * P.ncap + t];
      if (slot != 0xFF) {
        const size_t zbase = ((size_t)b * P.slots + slot) * P.ncap;
        const uint4* zs = reinterpret_cast<const uint4*>(S.z16_hist + (zbase + s) * NODE_EMB);
        const uint4* zt = reinterpret_cast<const uint4*>(S.z16_hist + (zbase + t) * NODE_EMB);
        uint4 hs[NODE_EMB / 8], ht[NODE_EMB / 8];
#pragma unroll
        for (int i = 0; i < NODE_EMB / 8; ++i) { hs[i] = zs[i]; ht[i] = zt[i]; }
        n2 = S.zn2_hist[zbase + s] + S.zn2_hist[zbase + t] + 1.f;
        r0 = (s == t) ? ra : rb;
        cnt = rc - r0;
#pragma unroll
        for (int i = 0; i < NODE_EMB / 8; ++i) {
          st += dot8(hs[i], sh.a_st + 8 * i);
          st += dot8(ht[i], sh.a_st + NODE_EMB + 8 * i);
        }
        // exact ties resolve in table-insertion order: slot, then the source's position in owned_nodes WHEN the pair was
        // added (== its current position unless a defender removed nodes from the list), then the target's position
        const int opk = DEF ? (int)S.pair_opos[(size_t)b * P.ncap * P.ncap + s * P.ncap + t] : op;
        const int epoch = P.precise_positions ? (int)S.pair_epoch[(size_t)b * P.ncap * P.ncap + s * P.ncap + t] : slot;
        key = ((uint32_t)epoch << 24) | ((uint32_t)opk << 16) | ((uint32_t)dp << 8) | (uint32_t)op;
        slot_of_pair = slot;
        live = cnt > 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) fm |= row_filtered(P, k, s, t, starter, interest) ? (1u << k) : 0u;   // folds to three tests
      }
    }
    const unsigned lmask = __ballot_sync(0xFFFFFFFFu, live);
    const int npairs = __popc(lmask);
    n_live += npairs;
    if (npairs == 0) continue;
    const int idx = __popc(lmask & ((1u << lane) - 1u));
    if (live) { sh.p_sn[idx] = make_float2(st, n2); sh.p_fm[idx] = fm; sh.p_r0[idx] = r0; sh.p_key[idx] = key; sh.p_slot[idx] = slot_of_pair; sh.p_pre[idx + 1] = cnt; }
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
    n_rows += total;

    // ---- phase B: candidate rows of the staged pairs, flattened over the lanes, RPL rows per lane per trip; the
    //      row templates of the NEXT trip are requested before the current trip is scored ----
    int pi = 0;                                             // rows are visited in increasing order: the pair index only grows
    int pp_n[RPL];                                          // (a row's table index is recomputed from its pair if it gets parked)
    uint32_t packed_n[RPL];
#pragma unroll
    for (int q = 0; q < RPL; ++q) {
      const int j = q * 32 + lane;
      pp_n[q] = 0; packed_n[q] = 0;
      if (j < total) {
        while (sh.p_pre[pi + 1] <= j) ++pi;
        pp_n[q] = pi;
        packed_n[q] = T.row_packed[sh.p_r0[pi] + (j - sh.p_pre[pi])];
      }
    }
    for (int j0 = 0; j0 < total; j0 += RPL * 32) {
      int pp[RPL];
      uint32_t packed[RPL];
#pragma unroll
      for (int q = 0; q < RPL; ++q) { pp[q] = pp_n[q]; packed[q] = packed_n[q]; }
#pragma unroll
      for (int q = 0; q < RPL; ++q) {                       // prefetch the next trip
        const int j = j0 + RPL * 32 + q * 32 + lane;
        if (j < total) {
          while (sh.p_pre[pi + 1] <= j) ++pi;
          pp_n[q] = pi;
          packed_n[q] = T.row_packed[sh.p_r0[pi] + (j - sh.p_pre[pi])];
        }
      }
      float score[RPL];
      bool valid[RPL];
      float cmax = -INFINITY;
#pragma unroll
      for (int q = 0; q < RPL; ++q) {
        const int j = j0 + q * 32 + lane;
        score[q] = -INFINITY; valid[q] = false;
        if (j < total) {
          const int kind = (packed[q] >> 20) & 15;
          if (!((sh.p_fm[pp[q]] >> kind) & 1u)) {
            const int u = packed[q] & 0xFFFFF, oh = (packed[q] >> 24) & 15;
            const float2 vv = u < vt_cached ? vtn[u] : make_float2(vt_g[u], (float)T.vnorm2[u]);
            const float2 sn = sh.p_sn[pp[q]];
            score[q] = (sn.x + vv.x + sh.a_o[oh]) * rsqrtf(sn.y + vv.y);
            valid[q] = true;
            cmax = fmaxf(cmax, score[q]);
          }
        }
      }
      {   // warp maximum in one REDUX on the order-preserving integer image of the float (cmax is never NaN: fmaxf dropped them)
        int ci = __float_as_int(cmax);
        ci ^= (ci >> 31) & 0x7FFFFFFF;
        ci = __reduce_max_sync(0xFFFFFFFFu, ci);
        ci ^= (ci >> 31) & 0x7FFFFFFF;
        cmax = __int_as_float(ci);
      }
      run_max = fmaxf(run_max, cmax);                       // fmaxf drops NaN; NaN rows are candidates below
      bool any_c = false;
#pragma unroll
      for (int q = 0; q < RPL; ++q) any_c |= valid[q] && !(score[q] < run_max - margin_s);
      if (!__any_sync(0xFFFFFFFFu, any_c)) continue;       // the common trip: nothing within the margin of the running maximum
#pragma unroll
      for (int q = 0; q < RPL; ++q) {
        unsigned cand = __ballot_sync(0xFFFFFFFFu, valid[q] && !(score[q] < run_max - margin_s));
        while (cand) {
          const int src = __ffs(cand) - 1;
          cand &= cand - 1;
          if (ncand == CAND_CAP) { n_exact += ncand; flush_candidates(T, P, S, act, sh, ncand, run_max - margin_s, b, na, lane, best); }
          const float cs = __shfl_sync(0xFFFFFFFFu, score[q], src);
          const int cpi = __shfl_sync(0xFFFFFFFFu, pp[q], src);
          const int cr = sh.p_r0[cpi] + (j0 + q * 32 + src - sh.p_pre[cpi]);
          const uint32_t cp = __shfl_sync(0xFFFFFFFFu, packed[q], src);
          if (lane == 0) {
            const uint32_t pk = sh.p_key[cpi];
            sh.c_score[ncand] = cs; sh.c_key[ncand] = ((unsigned long long)pk << 32) | (unsigned long long)(unsigned)cr;
            sh.c_st[ncand] = (int)sh.oorder[pk & 0xFF] | ((int)sh.dorder[(pk >> 8) & 0xFF] << 8);
            sh.c_slot[ncand] = sh.p_slot[cpi]; sh.c_row[ncand] = cr; sh.c_packed[ncand] = cp;
          }
          ++ncand;
          __syncwarp();
        }
      }
    }
    __syncwarp();
  }
  n_exact += ncand;
  flush_candidates(T, P, S, act, sh, ncand, run_max - margin_s, b, na, lane, best);
  if (lane == 0) {
    out = make_int4(starter, starter, 0, 0);
    d = 1.0;
    if (best.r >= 0) {
      out = make_int4(best.s, best.t, T.row_ulocal[best.r], (int)((best.packed >> 20) & 15));
      d = best.d;
      // margin diagnostics: the float64 winner's float32 score sat in the outer half of the re-score window, i.e. the scan's
      // error (TF32 operand rounding of the contraction + half-precision snapshot rows) used more than half of the margin
      if (run_max - best.score32 > 0.5f * margin_s) atomicAdd(S.errflag + 2, 1);
    } else {
      atomicExch(S.errflag, 3);   // empty action table: outside the reference's domain (cdist would raise)
    }
    if (trace) {
      long long* tr = trace + (size_t)b * 6;
      tr[0] = clock64() - t_begin; tr[1] = n_rows; tr[2] = n_live; tr[3] = n_exact; tr[4] = combos; tr[5] = t_begin;
    }
  }
  }   // active
  if (in_range && lane == 0) {
    reinterpret_cast<int4*>(S.sel)[b] = out;
    S.dist[b] = d;
    if (sel_out) reinterpret_cast<int4*>(sel_out)[b] = out;
    if (dist_out) dist_out[b] = d;
  }
  // fused step: the CTA's transitions run together on the first lanes of warp 0 once all its warps have decoded.  One
  // lane per warp right after its decode cost the same ~700 instructions per ENV at 1/32 lane utilisation — a fifth of
  // the kernel's issue slots; the CTA's resources are held until its slowest warp is done either way.
  if (FUSE) {
    __shared__ int4 sh_out[SEL_WARPS];
    __shared__ double sh_d[SEL_WARPS];
    __shared__ int sh_b[SEL_WARPS];
    if (lane == 0) { sh_out[warp] = out; sh_d[warp] = d; sh_b[warp] = in_range ? b : -1; }
    __syncthreads();
    if (warp == 0 && lane < SEL_WARPS && sh_b[lane] >= 0) {
      const int bb = sh_b[lane];
      TransitionIn<W1> tin;
      tin.issue(P, S, bb, ft.uniforms);
      tin.sl = sh_out[lane]; tin.dist = sh_d[lane];
      transition_env<DEF, true, W1>(T, P, S, bb, tin, T.sc_pack, ft.uniforms != nullptr, false, sched_buf ^ 1, ft.reward, ft.done, nullptr, nullptr);
      tin.store_hot(S, bb);
    }
  }
}

long long* g_sel_trace = nullptr;   // debug: per-env {cycles, rows, live pairs, parked candidates, combos, start clock}

cudaError_t launch_decode_select(const Tables& T, const Params& P, const State& S, const float* actions, int vt_stride,
                                 int sched_buf, int fuse_transition, const float* uniforms, float* reward, uint8_t* done,
                                 int32_t* sel_out, double* dist_out, cudaStream_t stream) {
  const FusedTransition ft{fuse_transition, uniforms, reward, done};
  const int vt_cached = vt_stride <= SEL_VT_SMEM_MAX ? vt_stride : SEL_VT_SMEM_MAX;
  const size_t smem = SEL_WARPS * sizeof(SelWarp) + (size_t)(2 * SEL_WARPS + 1) * vt_cached * sizeof(float);
  const int which = ((fuse_transition ? 1 : 0) | (P.defender ? 2 : (P.words == 1 ? 4 : 0))) + (P.subset_k ? 6 : 0);
  using KernelFn = void (*)(Tables, Params, State, const float*, int, int, int, FusedTransition, int32_t*, double*, long long*);
  const KernelFn kernels[12] = {decode_select_kernel<false, false, false, false>, decode_select_kernel<true, false, false, false>,
                                decode_select_kernel<false, true, false, false>,  decode_select_kernel<true, true, false, false>,
                                decode_select_kernel<false, false, true, false>,  decode_select_kernel<true, false, true, false>,
                                decode_select_kernel<false, false, false, true>,  decode_select_kernel<true, false, false, true>,
                                decode_select_kernel<false, true, false, true>,   decode_select_kernel<true, true, false, true>,
                                decode_select_kernel<false, false, true, true>,   decode_select_kernel<true, false, true, true>};
  const KernelFn kern = kernels[which];
  if (smem > 48 * 1024) {
    cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(kern), smem);
    if (e != cudaSuccess) return e;
  }
  const int grid = (P.B + SEL_WARPS - 1) / SEL_WARPS;
  return launch_pdl(kern, dim3(grid), dim3(SEL_THREADS), smem, stream, true, T, P, S, actions, vt_stride, vt_cached, sched_buf, ft, sel_out,
                    dist_out, g_sel_trace);
}

}  // namespace cbs
