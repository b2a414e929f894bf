// k_decode_tc.cu — the decode contraction  VT[B, Ug] = A_v[B, 768] x Vemb[Ug, 768]^T  on tcgen05 (TF32 in, FP32
// accumulate in TMEM), operands staged by TMA.
//
// This is the one GEMM-shaped piece of the step path (find_closest_action_embedding,
// _env/cyberbattle_env_compressed.py:570-590: the 768-wide vulnerability part of every action against every
// vulnerability embedding).  TF32 keeps the float32 inputs as they are (no conversion pass); its ~2^-11 operand
// rounding is covered by decode_select's float64 re-score margin, so the chosen action is unaffected.
//
// Shape: one CTA per 128 envs x (<=256) vulnerabilities; K = 768 in 24 slabs of 32 floats (= one 128-byte
// swizzle row), 4-stage TMA -> smem ring; warp 0 lane 0 issues TMA, warp 1 lane 0 issues 4 UMMAs (K = 8) per
// slab and commits to the ring's empty barriers; all four warps drain the 128 x N accumulator from TMEM with
// tcgen05.ld.32x32b and store float4 rows.  A dense [B,905] action tensor has rows 3620 bytes apart (not a
// multiple of 16), which TMA cannot address, so pack_actions first copies the 768-float slice into an aligned
// [B,768] slab; with a 16-byte-multiple row pitch (cbs_set_action_stride) TMA reads in place.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <type_traits>

#include "cbs_types.h"

namespace cbs {
extern long long* g_sel_trace;   // cbs_debug_select_trace's buffer (debug builds of this file append their own rows)

namespace {

// envs per CTA = UMMA M: 128, or 64 when 128-row tiles would leave more than half of the SMs without a CTA (8192 envs:
// 64 tiles of 128 on 148 SMs).  The kernel is bound by staging the A slab per CTA, so twice as many half-height tiles
// roughly halve its duration.  With M = 64 the accumulator occupies lanes 0-15 of each warp's 32-lane TMEM
// sub-partition: tile row r lives in lane (r / 16) * 32 + r % 16.
constexpr int BK = 32;         // floats per K slab = 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 8;      // tf32: 32 bytes per instruction
constexpr int MAX_STAGES = 8;   // ring depth is chosen per launch: as many stages as fit in shared memory
constexpr int NT_MAX = 256;    // UMMA N limit
constexpr int TMEM_COLS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a protocol bug must end the kernel with an error flag, never hang the GPU
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int32_t* errflag) {
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  atomicExch(errflag, 9);
  return false;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout): start address >> 4 in
// [0,14), stride byte offset (8 rows x 128 B = 1024) >> 4 in [32,46), version 1 in [46,48), layout type 2 in [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

}  // namespace

__global__ void __launch_bounds__(256) pack_actions_kernel(const float* __restrict__ actions, float* __restrict__ packed, int B,
                                                           int act_stride) {
  const int row = blockIdx.x * 2 + (threadIdx.x >> 7);
  if (row >= B) return;
  const float* src = actions + (size_t)row * act_stride + 2 * NODE_EMB;
  float* dst = packed + (size_t)row * VULN_EMB;
  for (int i = threadIdx.x & 127; i < VULN_EMB; i += 128) dst[i] = src[i];
}

// LDGSTS_A = false: both operands arrive by TMA (A from the repacked slab, or in place when the caller's action rows
//            have a 16-byte-multiple pitch).  128 threads.
// LDGSTS_A = true:  the caller's action rows are only 4-byte aligned (dense [B,905]); four extra producer warps stage
//            the A slab with 4-byte cp.async straight into the 128-byte-swizzled layout the UMMA descriptor expects
//            (16-byte chunk c of row r lands at chunk c ^ (r & 7)) and arrive on the stage's full barrier when their
//            copies land (cp.async.mbarrier.arrive.noinc); no repack pass over HBM.  256 threads.
#ifdef CBS_GEMM_TRACE
__device__ long long* g_sel_trace_dev = nullptr;
#endif
template <bool LDGSTS_A, int BM>
__global__ void __launch_bounds__(LDGSTS_A ? 256 : 128, 1) decode_gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                                 const __grid_constant__ CUtensorMap map_b,
                                                                                 const float* __restrict__ actions, int act_stride,
                                                                                 float* __restrict__ vt, int B, int Upad, int nt_box,
                                                                                 int vt_stride, int32_t* errflag, int direct, int STAGES, int wait_early) {
#ifdef CBS_GEMM_TRACE   // debug build (-DCBS_GEMM_TRACE): per-CTA, per-warp phase stamps, read by tools/gemm_trace.py
  long long tt[6] = {0, 0, 0, 0, 0, 0};
#define GT(k) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt[k]))
#else
#define GT(k)
#endif
  GT(0);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_STAGE_BYTES = BM * BK * 4;   // 16 KB / 8 KB
  const int stage_bytes = A_STAGE_BYTES + nt_box * BK * 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), tfull = smem_u32(bars + 2 * STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * NT_MAX;
  const int nt = min(NT_MAX, Upad - n0);             // columns this CTA produces (multiple of 16)
  pdl_trigger();   // the next kernel (decode_select) may be scheduled as SMs free up; it waits for this grid before reading VT
  constexpr int KB = VULN_EMB / BK;                  // 24 slabs

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, LDGSTS_A ? 1 + 128 : 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  GT(1);
  // wait_early (the default): nothing is READ before the previous kernel on the stream has completed - only the barrier / TMEM
  // set-up above overlaps it.  With actions the caller declared pre-staged (cbs_set_actions_prestaged) the whole operand pipeline
  // runs first and only the VT stores wait.
  if (wait_early) pdl_wait();

  if (warp == 0 && lane == 0) {
    // ---- TMA producer ----
    const uint32_t bytes = (uint32_t)(LDGSTS_A ? stage_bytes - A_STAGE_BYTES : stage_bytes);
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      if (!mbar_wait(empty0 + 8 * s, ph ^ 1, errflag)) break;
      mbar_expect_tx(full0 + 8 * s, bytes);
      const uint32_t sa = smem_u32(smem + s * stage_bytes);
      // direct: map_a is the caller's action tensor itself (row pitch a multiple of 16 bytes), the 768-float
      // vulnerability part starts 128 floats into each row; otherwise map_a is the repacked [B,768] slab
      if (!LDGSTS_A) tma_load_2d(sa, &map_a, full0 + 8 * s, (direct ? 2 * NODE_EMB : 0) + kb * BK, m0);
      tma_load_2d(sa + A_STAGE_BYTES, &map_b, full0 + 8 * s, kb * BK, n0);
    }
  } else if (warp == 1 && lane == 0) {
    // ---- MMA issuer ----  instruction descriptor: D=F32 (bit 4), A=B=TF32 (2 at bits 7 and 10), K-major both, N>>3 at 17, M>>4 at 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(nt >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    bool ok = true;
    for (int kb = 0; kb < KB && ok; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      ok = mbar_wait(full0 + 8 * s, ph, errflag);
      if (!ok) break;
      if (LDGSTS_A) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // cp.async wrote through the generic proxy
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa = smem_u32(smem + s * stage_bytes);
      const uint64_t ad = umma_desc(sa), bd = umma_desc(sa + A_STAGE_BYTES);
#pragma unroll
      for (int k = 0; k < BK / UMMA_K; ++k)   // advance 32 bytes along K inside the swizzle row: +2 in the >>4 address field
        umma_tf32(tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) ? 1u : 0u);
      umma_commit(empty0 + 8 * s);            // frees the smem stage when these MMAs retire
    }
    umma_commit(tfull);                       // accumulator complete
  } else if (LDGSTS_A && warp >= 4) {
    // ---- A producers (4 warps): 32 coalesced 4-byte cp.async per thread per slab, rows beyond B are zero-filled ----
    const int t = threadIdx.x - 128;
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      if (!mbar_wait(empty0 + 8 * s, ph ^ 1, errflag)) break;
      const uint32_t sa = smem_u32(smem + s * stage_bytes);
#pragma unroll 8
      for (int i = 0; i < BM * BK / 128; ++i) {
        const int idx = i * 128 + t, r = idx >> 5, k = idx & 31;
        const uint32_t dst = sa + r * 128 + ((((uint32_t)k >> 2) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)k & 3u) * 4;
        const int row = m0 + r;
        const float* src = actions + (size_t)(row < B ? row : 0) * act_stride + 2 * NODE_EMB + kb * BK + k;
        const uint32_t nbytes = row < B ? 4u : 0u;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full0 + 8 * s) : "memory");
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  __syncwarp();

  // ---- epilogue: TMEM -> registers -> global.  Every warp owns TMEM lanes [32w, 32w+32): tile rows 32w + lane
  //      (M = 128) or rows 16w + lane on its lanes 0-15 (M = 64) ----
  GT(2);
  const bool ready = warp < 4 && mbar_wait(tfull, 0, errflag);
  GT(3);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // Everything above reads only the caller's actions and the constant embedding table; VT is still being read by the previous
  // step's decode_select if that step's kernels have not drained (programmatic dependent launch, cbs_types.h).
  pdl_wait();
  const bool lane_has_row = BM == 128 || lane < 16;
  const int row = lane_has_row ? m0 + warp * (BM / 4) + lane : B;
  if (ready) {
    for (int c0 = 0; c0 < nt; c0 += 32) {
      uint32_t r[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < B) {
        float4* dst = reinterpret_cast<float4*>(vt + (size_t)row * vt_stride + n0 + c0);
        const int ncols = min(32, vt_stride - n0 - c0);   // multiple of 4
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (4 * i < ncols)
            dst[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                                 __uint_as_float(r[4 * i + 3]));
      }
    }
  }
  GT(4);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  GT(5);
#ifdef CBS_GEMM_TRACE   // rows behind the select kernel's [B][6] trace block (the buffer is cbs_debug_select_trace's)
  if (g_sel_trace_dev && (threadIdx.x & 31) == 0) {
    long long* tr = g_sel_trace_dev + (size_t)B * 6 + ((size_t)blockIdx.x * 8 + warp) * 6;
    for (int k = 0; k < 6; ++k) tr[k] = tt[k];
  }
#endif
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Half-precision variant (default for half-height tiles).  What bounds the TF32 kernel above is, equally, the operands' way into
// the SM and the rate of its K = 8 UMMAs (~220 cycles per tcgen05.mma whatever is knocked out around it).  FP16
// operands halve both: a UMMA covers K = 16 and the B slab is half the bytes.  FP16 has TF32's 10-bit mantissa, so the operand
// rounding the re-score margin of decode_select absorbs is the same; only the exponent range is narrower (the action space is
// Box(-4, 4); values beyond +-65504 are clamped, embeddings below 6e-5 lose bits that cannot decide a 0.5-wide margin).
//   * B: Vemb converted once to FP16 at table load; TMA boxes of 64 halfs (one 128-byte swizzle row) x N.
//   * A: the dense float32 [B,905] rows (4-byte aligned only).  Eight producer warps read them as coalesced, ALIGNED 16-byte
//     vectors (a half-warp per row and slab), each lane also its next vector, pick their floats at the row's shift, convert,
//     and store into the 128-byte-swizzled FP16 tile of the MMA ring; A_AHEAD slabs are held in registers.
//     Measured and replaced: (1) 4-byte cp.async into a float32 ring + a conversion pass through shared memory - 64
//     load/store-unit instructions per thread and slab, copy and conversion serialise on that unit: 32.6 us (TF32 kernel: 22.8;
//     A path knocked out: 17.6); (2) one (row, 16-float segment) per thread, five aligned vectors each - every warp request
//     touches 32 different sectors in 16 lines: 44 us.
//   * 12 ring stages of K = 64 instead of 24 of K = 32; 4 UMMAs (kind::f16, K = 16) per stage.
constexpr int F16_BK = 64;        // halfs per K slab = 128 bytes = one SWIZZLE_128B row
constexpr int F16_BM = 64;
constexpr int A_AHEAD = 3;        // slabs of A a producer thread holds in registers ahead of the conversion (32 floats each)
constexpr int F16_THREADS = 384;  // warp 0 TMA, warp 1 MMA, warps 0-3 epilogue, warps 4-11 A producers

__device__ __forceinline__ float4 ld_nc_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(256) convert_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t n) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) dst[i] = __float2half_rn(fminf(fmaxf(src[i], -65504.f), 65504.f));
}

__global__ void __launch_bounds__(F16_THREADS, 1) decode_gemm_f16_kernel(const __grid_constant__ CUtensorMap map_b,
                                                                 const float* __restrict__ actions, int act_stride,
                                                                 float* __restrict__ vt, int B, int Upad, int nt_box, int vt_stride,
                                                                 int32_t* errflag, int STAGES, int knock, int wait_early) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int BM = F16_BM;
  constexpr int A_STAGE_BYTES = BM * F16_BK * 2;      // 8 KB
  const int stage_bytes = A_STAGE_BYTES + nt_box * F16_BK * 2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), tfull = smem_u32(bars + 2 * STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * NT_MAX;
  const int nt = min(NT_MAX, Upad - n0);
  constexpr int KB = VULN_EMB / F16_BK;               // 12 slabs
  pdl_trigger();
  if (wait_early) pdl_wait();      // (see the TF32 kernel)

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1 + 8); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ---- TMA producer (B) ----
    const uint32_t bytes = (uint32_t)(stage_bytes - A_STAGE_BYTES);
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      if (!mbar_wait(empty0 + 8 * s, ph ^ 1, errflag)) break;
      if (knock & 8) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full0 + 8 * s) : "memory"); continue; }
      mbar_expect_tx(full0 + 8 * s, bytes);
      tma_load_2d(smem_u32(smem + s * stage_bytes) + A_STAGE_BYTES, &map_b, full0 + 8 * s, kb * F16_BK, n0);
    }
  } else if (warp == 1 && lane == 0) {
    // ---- MMA issuer ----  D = F32 (bit 4), A = B = F16 (format 0), K-major both, N >> 3 at 17, M >> 4 at 24
    const uint32_t idesc = (1u << 4) | ((uint32_t)(nt >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    bool ok = true;
    for (int kb = 0; kb < KB && ok; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      ok = mbar_wait(full0 + 8 * s, ph, errflag);
      if (!ok) break;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa = smem_u32(smem + s * stage_bytes);
      const uint64_t ad = umma_desc(sa), bd = umma_desc(sa + A_STAGE_BYTES);
#pragma unroll
      for (int k = 0; k < F16_BK / 16; ++k)   // K = 16 halfs = 32 bytes per instruction: +2 in the >>4 address field
        if (!(knock & 2) || (kb | k) == 0) umma_f16(tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) ? 1u : 0u);
      umma_commit(empty0 + 8 * s);
    }
    umma_commit(tfull);
  } else if (warp >= 4) {
    // ---- A producers (8 warps; warp pw owns tile rows [8 pw, 8 pw + 8)) ----
    // The rows are only 4-byte aligned.  Per slab a warp makes four passes over a PAIR of rows (r, r + 4: the same alignment
    // class for any row pitch, so the shift below is warp-uniform): a half-warp reads one row's 256 bytes as 16 aligned 16-byte
    // vectors - fully coalesced - plus each lane's NEXT vector (the same lines again, an L1 hit), picks its four floats at the
    // row's shift (0-3 floats), converts and stores 8 bytes into the 128-byte-swizzled FP16 tile.  A_AHEAD slabs are held in
    // registers ahead of the one being converted.
    const int pw = warp - 4, v = lane & 15;
    const float4* base[4];
    int shift[4];
    bool live[4];
    uint32_t toff[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int r = 8 * pw + it + 4 * (lane >> 4), row = m0 + r;
      const float* p0 = actions + (size_t)(row < B ? row : 0) * act_stride + 2 * NODE_EMB;
      shift[it] = (int)((reinterpret_cast<uintptr_t>(p0) >> 2) & 3);
      base[it] = reinterpret_cast<const float4*>(p0 - shift[it]) + v;        // 16-byte aligned
      live[it] = row < B && !(knock & 1);
      toff[it] = (uint32_t)r * 128 + ((((uint32_t)v >> 1) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)v & 1u) * 8;
    }
    float4 q[A_AHEAD][4][2];
    auto load = [&](float4 (&d)[4][2], int kk) {
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const float4* g = base[it] + kk * (F16_BK / 4);
        d[it][0] = live[it] ? ld_nc_f4(g) : make_float4(0.f, 0.f, 0.f, 0.f);
        d[it][1] = (live[it] && shift[it]) ? ld_nc_f4(g + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
#pragma unroll
    for (int d = 0; d < A_AHEAD; ++d) load(q[d], d);
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
      uint2 h[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const float4 a = q[kb % A_AHEAD][it][0], c = q[kb % A_AHEAD][it][1];
        float f0, f1, f2, f3;
        switch (shift[it]) {
          case 0: f0 = a.x; f1 = a.y; f2 = a.z; f3 = a.w; break;
          case 1: f0 = a.y; f1 = a.z; f2 = a.w; f3 = c.x; break;
          case 2: f0 = a.z; f1 = a.w; f2 = c.x; f3 = c.y; break;
          default: f0 = a.w; f1 = c.x; f2 = c.y; f3 = c.z; break;
        }
        const __half2 lo = __floats2half2_rn(fminf(fmaxf(f0, -65504.f), 65504.f), fminf(fmaxf(f1, -65504.f), 65504.f));
        const __half2 hi = __floats2half2_rn(fminf(fmaxf(f2, -65504.f), 65504.f), fminf(fmaxf(f3, -65504.f), 65504.f));
        h[it] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
      }
      if (kb + A_AHEAD < KB) load(q[kb % A_AHEAD], kb + A_AHEAD);
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      if (!mbar_wait(empty0 + 8 * s, ph ^ 1, errflag)) break;
      unsigned char* tile = smem + s * stage_bytes;
      if (!(knock & 4)) {
#pragma unroll
        for (int it = 0; it < 4; ++it) *reinterpret_cast<uint2*>(tile + toff[it]) = h[it];
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy stores -> visible to the MMA's async proxy
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full0 + 8 * s) : "memory");
    }
  }
  __syncwarp();

  // ---- epilogue (M = 64: tile row 16 w + lane on lanes 0-15 of warp w's TMEM sub-partition) ----
  const bool ready = warp < 4 && mbar_wait(tfull, 0, errflag);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  pdl_wait();      // VT may still be read by the previous step's decode_select (see the TF32 kernel)
  const int row = lane < 16 ? m0 + warp * (BM / 4) + lane : B;
  if (ready) {
    for (int c0 = 0; c0 < nt; c0 += 32) {
      uint32_t r[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < B) {
        float4* dst = reinterpret_cast<float4*>(vt + (size_t)row * vt_stride + n0 + c0);
        const int ncols = min(32, vt_stride - n0 - c0);   // multiple of 4
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (4 * i < ncols)
            dst[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                                 __uint_as_float(r[4 * i + 3]));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

static bool load_encode() {
  if (g_encode) return true;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return false;
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return true;
}

bool decode_gemm_tc_available() {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  return major == 10 && load_encode();
}

static bool make_map(CUtensorMap* map, const float* base, uint64_t rows, uint32_t box_rows, uint64_t pitch_floats = VULN_EMB,
                     uint64_t row_floats = VULN_EMB) {
  const cuuint64_t dims[2] = {(cuuint64_t)row_floats, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch_floats * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

cudaError_t convert_vemb_f16(const float* vemb32, __half* vemb16, size_t n, cudaStream_t stream) {
  convert_f16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(vemb32, vemb16, n);
  return cudaGetLastError();
}

// true when launch_decode_gemm_f16 is the kernel for this batch.  Opt-in (CBS_GEMM_F16=1): measured on B200 at 8192 envs it does
// not beat the TF32 kernel yet (24.4 us against 22.8 us) - its MMA loop is half as long, but the A stream
// (25 MB of fresh actions from DRAM per launch) then needs more bytes in flight than a register-staged conversion can hold.
bool decode_gemm_f16_applies(int B, int Ug) {
  static const bool opt_in = getenv("CBS_GEMM_F16") != nullptr;
  if (!opt_in || getenv("CBS_GEMM_M128")) return false;
  int dev = 0, num_sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  const int Upad = ((Ug + 15) / 16) * 16;
  const int ntiles_n = (Upad + NT_MAX - 1) / NT_MAX;
  return 2 * ((B + 127) / 128) * ntiles_n <= num_sms + num_sms / 4;
}

cudaError_t launch_decode_gemm_f16(const float* actions, int act_stride, const __half* vemb16, float* vt, int B, int Ug,
                                   int vt_stride, int32_t* errflag, int wait_early, cudaStream_t stream) {
  if (!load_encode()) return cudaErrorNotSupported;
  const int Upad = ((Ug + 15) / 16) * 16;
  const int nt_box = Upad < NT_MAX ? Upad : NT_MAX;
  CUtensorMap map_b;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)VULN_EMB, (cuuint64_t)Ug};
    const cuuint64_t strides[1] = {(cuuint64_t)VULN_EMB * 2};
    const cuuint32_t box[2] = {(cuuint32_t)F16_BK, (cuuint32_t)nt_box};
    const cuuint32_t estr[2] = {1, 1};
    if (g_encode(&map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(vemb16), dims, strides, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  const int ntiles_n = (Upad + NT_MAX - 1) / NT_MAX;
  const size_t stage_bytes = (size_t)F16_BM * F16_BK * 2 + (size_t)nt_box * F16_BK * 2;
  const size_t fixed = 1024 + 256;
  int stages = (int)((225 * 1024 - fixed) / stage_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return cudaErrorInvalidValue;
  const size_t smem = (size_t)stages * stage_bytes + fixed;
  {
    cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(decode_gemm_f16_kernel), smem);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((B + F16_BM - 1) / F16_BM, ntiles_n);
  static const int knock = getenv("CBS_GEMM_KNOCK") ? atoi(getenv("CBS_GEMM_KNOCK")) : 0;   // debug: knock out A copies (1), MMAs (2), the conversion (4), B copies (8)
  return launch_pdl(decode_gemm_f16_kernel, grid, dim3(F16_THREADS), smem, stream, true, map_b, actions, act_stride, vt, B, Upad, nt_box, vt_stride,
                    errflag, stages, knock, wait_early);
}

cudaError_t launch_decode_gemm_tc(const float* actions, int act_stride, const float* vemb, float* a_packed, float* vt, int B,
                                  int Ug, int vt_stride, int32_t* errflag, int wait_early, cudaStream_t stream) {
  if (!load_encode()) return cudaErrorNotSupported;
  const int Upad = ((Ug + 15) / 16) * 16;
  const int nt_box = Upad < NT_MAX ? Upad : NT_MAX;
  // tensor maps are rebuilt per launch (host-side encode, ~1 us each): the operands never move, but keeping the
  // maps out of the handle keeps this translation unit self-contained
  CUtensorMap map_a, map_b;
  // A operand: (1) row pitch a multiple of 16 bytes (e.g. 908 floats): TMA reads the action tensor in place;
  // (2) dense [B,905] (3620-byte rows: TMA cannot address them, an unaligned box start faults as an illegal
  // instruction): producer warps stage it with 4-byte cp.async; (3) CBS_TMA_PACK=1: repack into an aligned slab first
  // (the first working version, kept for comparison).
  static const bool force_pack = getenv("CBS_TMA_PACK") != nullptr;
  const bool direct = (act_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(actions) & 15) == 0);
  const bool ldgsts = !direct && !force_pack;
  if (!make_map(&map_b, vemb, (uint64_t)Ug, (uint32_t)nt_box)) return cudaErrorInvalidValue;
  const int ntiles_n = (Upad + NT_MAX - 1) / NT_MAX;
  int num_sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  static const bool force128 = getenv("CBS_GEMM_M128") != nullptr;   // A/B switch for the measurements in DESIGN.md
  const bool half = !force128 && 2 * ((B + 127) / 128) * ntiles_n <= num_sms + num_sms / 4;   // half-height tiles still fit (about) one wave
  const int bm = half ? 64 : 128;
  if (direct) {
    if (!make_map(&map_a, actions, (uint64_t)B, bm, (uint64_t)act_stride, (uint64_t)ACTION_DIM)) return cudaErrorInvalidValue;
  } else {
    if (!make_map(&map_a, a_packed, (uint64_t)B, bm)) return cudaErrorInvalidValue;
    if (!ldgsts) pack_actions_kernel<<<(B + 1) / 2, 256, 0, stream>>>(actions, a_packed, B, act_stride);
  }
  const size_t stage_bytes = (size_t)bm * BK * 4 + (size_t)nt_box * BK * 4;
  // Launched with a programmatic dependency (cbs_types.h), the kernel leaves kPdlReserve bytes of the SM's shared memory to the CTAs
  // of the NEXT kernel (decode_select, ~17 KB each): they become resident beside this CTA, finish their constant-only head and
  // start their table scans the moment this grid completes, instead of being launched then.  Measured (graph replay, 8192 envs):
  // ring of 6 stages (209 KB, nothing fits beside it) 0.0970 ms per step, 5 or 4 stages 0.0932 ms, 3 stages 0.1042 ms; the kernel
  // itself is no slower with 4-5 stages (beyond 4 stages the operand pipeline is not what bounds it).
  constexpr size_t kPdlReserve = 48 * 1024;
  const bool pdl = (direct || ldgsts) && pdl_enabled();
  const int fit = (int)((225 * 1024 - 1024 - 256) / stage_bytes);
  int stages = pdl ? (int)((225 * 1024 - 1024 - 256 - kPdlReserve) / stage_bytes) : fit;
  if (stages < 4) stages = fit < 4 ? fit : 4;     // the reserve never costs the ring its fourth stage
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  static const char* stages_env = getenv("CBS_GEMM_STAGES");
  if (stages_env && atoi(stages_env) >= 2 && atoi(stages_env) < stages) stages = atoi(stages_env);
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
  using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const float*, int, float*, int, int, int, int, int32_t*, int, int, int);
  const int which = (ldgsts ? 1 : 0) | (half ? 2 : 0);
  const KernelFn kernels[4] = {decode_gemm_tc_kernel<false, 128>, decode_gemm_tc_kernel<true, 128>,
                               decode_gemm_tc_kernel<false, 64>, decode_gemm_tc_kernel<true, 64>};
  {
    cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(kernels[which]), smem);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((B + bm - 1) / bm, ntiles_n);
#ifdef CBS_GEMM_TRACE
  cudaMemcpyToSymbolAsync(g_sel_trace_dev, &g_sel_trace, sizeof(g_sel_trace), 0, cudaMemcpyHostToDevice, stream);
#endif
  // (the repacked slab is written by pack_actions_kernel right before: that variant keeps the plain stream order)
  return launch_pdl(kernels[which], grid, dim3(ldgsts ? 256 : 128), smem, stream, direct || ldgsts, map_a, map_b, actions, act_stride, vt, B,
                    Upad, nt_box, vt_stride, errflag, direct ? 1 : 0, stages, wait_early);
}

}  // namespace cbs
