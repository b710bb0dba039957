// tcgen05 single-head attention for AttnBlockpp (sgmse/backbones/ncsnpp_utils/layerspp.py:75-91):
//   O[q, :] = softmax_k( Q[q,:].K[k,:] / sqrt(C) ) V[k, :]      (the path's one dense contraction)
// q, k: [B][N][C] fp32 (token-major), vT: [B][C][N] fp32 (channel-major, written that way by the V projection's
// epilogue so that both GEMMs see K-major operands), o: [B][N][C] fp32.  N % 128 == 0, C % 16 == 0, C <= 256.
//
// One CTA = one batch element x 128 queries.  Both GEMMs run on the tensor core with the bf16x3 split
// (fp32-grade): S = Q K^T accumulates in TMEM columns [0,128) / [128,256) (double buffered), O = P V in columns [256, 256+C).
// Two passes over the keys instead of online rescaling of the TMEM accumulator:
//   pass 1: S~ = Qhi Khi^T only (1/3 of the MMAs) -> per-row shift m (softmax is shift invariant, the bf16 error of
//           m is irrelevant); pass 2: exact S, P = exp((S - m)/sqrt(C)) written as bf16 hi/lo K-major tiles to
//           shared memory, O += P V, l += row sums; finally O / l.
// Operands: a pre-pass (attn_pack_kernel) writes Q, K and V ONCE as bf16 hi/lo operand images in the no-swizzle K-major core-matrix
// layout, one contiguous block per ring stage (8 KB per 128 rows x 16 channels of Q / K, C x 64 bytes per 16 keys of V); the
// attention CTAs - eight of them share a sample's K and V - fetch a stage with one cp.async.bulk each.  (Converting inside the
// attention kernel cost every CTA a global round trip + conversion per stage, 320 stages per CTA: 276 us; four converting producer
// groups: 229 us.)
// Warp roles: 0-3 softmax + epilogue (TMEM lane = query row), lane 0 of warps 4-7 = four copy issuers that fill the stage rings
// round-robin (a cp.async.bulk costs its issuing thread ~240 cycles), 8 TMEM alloc + single-thread MMA issue.
#include <cuda_bf16.h>

#include "kernels.cuh"

namespace storm {

namespace {

constexpr int A_PGROUPS = 4;                          // copy issuers (lane 0 of warps 4 .. 7)
constexpr int A_WARP_MMA = 4 + A_PGROUPS;             // warps 0-3 softmax, 4-7 issuers, 8 MMA
constexpr int A_THREADS = (A_WARP_MMA + 1) * 32;      // 288
constexpr int TQ = 128, TK = 128;         // queries per CTA, keys per tile
constexpr int SQ = 4, SV = 4;             // ring depths
// every ring slot belongs to ONE issuer (slot = item % depth, issuer = item % A_PGROUPS): issuers are not synchronised
// with each other, and one that ran a whole ring revolution ahead of a slot's other users would alias the barrier parity
static_assert(SQ % A_PGROUPS == 0 && SV % A_PGROUPS == 0, "ring depths must be multiples of the producer group count");
constexpr int QK_STAGE = 16 * 1024;       // Q chunk hi/lo (2 x 4 KB) + K chunk hi/lo (2 x 4 KB)
constexpr int V_STAGE_MAX = 16 * 1024;    // V chunk: C rows x 16 keys x hi/lo
constexpr int P_BYTES = 64 * 1024;        // P tile 128 x 128, hi + lo

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > 8000000000LL) __trap();     // bounded: a protocol bug traps instead of hanging the GPU
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a),
               "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major, no swizzle: 8 rows x 16 B core matrices; SBO = stride between 8-row groups, LBO = between k-cores
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t make_idesc(int n) {   // D f32, A/B bf16, K-major, M = 128
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b, float& ra, float& rb) {
  // one packed conversion (F2FP.BF16.F32.PACK_AB, full-rate pipe) instead of two scalar F2F on the quarter-rate XU pipe
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);          // .x (low half) = a, .y (high half) = b
  const uint32_t u = *reinterpret_cast<const uint32_t*>(&h);
  ra = a - __uint_as_float(u << 16);
  rb = b - __uint_as_float(u & 0xFFFF0000u);
  return u;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// 16 consecutive fp32 (one row of a 16-wide K chunk) -> bf16 hi/lo, written as the two k-cores of row `row` of an
// operand block with `rows` rows: hi plane at base, lo plane at base + plane; k-core stride = rows * 16 bytes
__device__ __forceinline__ void stage_row16(const float* __restrict__ src, uint32_t base, uint32_t plane, int rows, int row) {
  const float4* gp = reinterpret_cast<const float4*>(src);
  float4 v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = __ldg(gp + j);
#pragma unroll
  for (int kh = 0; kh < 2; ++kh) {
    const float4 a = v[2 * kh], c = v[2 * kh + 1];
    float r[8];
    uint4 hi, lo;
    hi.x = pack_bf16x2(a.x, a.y, r[0], r[1]);
    hi.y = pack_bf16x2(a.z, a.w, r[2], r[3]);
    hi.z = pack_bf16x2(c.x, c.y, r[4], r[5]);
    hi.w = pack_bf16x2(c.z, c.w, r[6], r[7]);
    float d0, d1;
    lo.x = pack_bf16x2(r[0], r[1], d0, d1);
    lo.y = pack_bf16x2(r[2], r[3], d0, d1);
    lo.z = pack_bf16x2(r[4], r[5], d0, d1);
    lo.w = pack_bf16x2(r[6], r[7], d0, d1);
    const uint32_t off = base + (uint32_t)(kh * rows + row) * 16u;
    st_shared_v4(off, hi);
    st_shared_v4(off + plane, lo);
  }
}

// ---- operand pre-pass -------------------------------------------------------------------------------
// q, k [B][N][C] and vT [B][C][N] fp32 -> bf16 hi/lo operand images:
//   Q / K block (b, 128-row tile t, 16-channel chunk c), 8192 B: hi plane [k-core kh = 0,1][row 0..127][16 B = 8 channels], lo plane + 4096
//   V block (b, key tile t, 16-key step ks), C x 64 B: hi plane [kh][channel 0..C-1][16 B = 8 keys], lo plane + C x 32
// one thread per 8 consecutive source floats (32 B read, 16 + 16 B written)
__device__ __forceinline__ void pack8(const float* __restrict__ src, uint8_t* __restrict__ hi_dst, uint8_t* __restrict__ lo_dst) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(src)), c = __ldg(reinterpret_cast<const float4*>(src) + 1);
  float r[8];
  uint4 hi, lo;
  hi.x = pack_bf16x2(a.x, a.y, r[0], r[1]);
  hi.y = pack_bf16x2(a.z, a.w, r[2], r[3]);
  hi.z = pack_bf16x2(c.x, c.y, r[4], r[5]);
  hi.w = pack_bf16x2(c.z, c.w, r[6], r[7]);
  float d0, d1;
  lo.x = pack_bf16x2(r[0], r[1], d0, d1);
  lo.y = pack_bf16x2(r[2], r[3], d0, d1);
  lo.z = pack_bf16x2(r[4], r[5], d0, d1);
  lo.w = pack_bf16x2(r[6], r[7], d0, d1);
  *reinterpret_cast<uint4*>(hi_dst) = hi;
  *reinterpret_cast<uint4*>(lo_dst) = lo;
}
__global__ void __launch_bounds__(256) attn_pack_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ vT,
                                                        uint8_t* __restrict__ qp, uint8_t* __restrict__ kp, uint8_t* __restrict__ vp, int B, int N,
                                                        int C) {
  pdl_enter();
  const int G = C >> 3, NC = C >> 4, NT = N / TK, KG = N >> 3;
  const size_t per = (size_t)B * N * G;                      // items per tensor (B * C * KG is the same number)
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < 3 * per; i += (size_t)gridDim.x * blockDim.x) {
    const int which = (int)(i / per);
    const size_t j = i - (size_t)which * per;
    if (which < 2) {
      const int g = (int)(j % G);
      const size_t bn = j / G;
      const int n = (int)(bn % N), b = (int)(bn / N);
      const int c = g >> 1, kh = g & 1, t = n / TK, row = n % TK;
      uint8_t* dst = (which == 0 ? qp : kp) + (((size_t)b * NT + t) * NC + c) * 8192 + (size_t)kh * 2048 + (size_t)row * 16;
      pack8((which == 0 ? q : k) + bn * C + g * 8, dst, dst + 4096);
    } else {
      const int kg = (int)(j % KG);
      const size_t bc = j / KG;
      const int ch = (int)(bc % C), b = (int)(bc / C);
      const int t = kg >> 4, ks = (kg & 15) >> 1, kh = kg & 1;
      uint8_t* dst = vp + (((size_t)b * NT + t) * (TK / 16) + ks) * ((size_t)C * 64) + (size_t)kh * C * 16 + (size_t)ch * 16;
      pack8(vT + bc * N + (size_t)kg * 8, dst, dst + (size_t)C * 32);
    }
  }
}

__global__ void __launch_bounds__(A_THREADS, 1) attn_tc_kernel(const uint8_t* __restrict__ qp, const uint8_t* __restrict__ kp,
                                                              const uint8_t* __restrict__ vp, float* __restrict__ o, int N, int C,
                                                              float scale) {
  pdl_enter();
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, q0 = blockIdx.x * TQ;
  const int NC = C >> 4, NT = N / TK;
  const int v_plane = C * 32;                 // bytes of one V plane (hi or lo): 2 k-cores x C rows x 16 B
  const uint32_t qk_base = smem_u32(smem);
  const uint32_t v_base = qk_base + SQ * QK_STAGE;
  const uint32_t p_base = v_base + SV * V_STAGE_MAX;
  const uint32_t bar_base = p_base + P_BYTES;
  auto qk_full = [&](int i) { return bar_base + 8u * i; };
  auto qk_empty = [&](int i) { return bar_base + 8u * (SQ + i); };
  auto v_full = [&](int i) { return bar_base + 8u * (2 * SQ + i); };
  auto v_empty = [&](int i) { return bar_base + 8u * (2 * SQ + SV + i); };
  // S is double buffered (TMEM columns [0,128) and [128,256)): the Q K^T MMAs of key tile t+1 run under the softmax of tile t
  const uint32_t s_full0 = bar_base + 8u * (2 * SQ + 2 * SV), s_empty0 = s_full0 + 16, p_full = s_full0 + 32, p_empty = s_full0 + 40,
                 o_full = s_full0 + 48, tmem_slot = s_full0 + 56;
  auto s_full = [&](uint32_t n) { return s_full0 + 8u * (n & 1u); };
  auto s_empty = [&](uint32_t n) { return s_empty0 + 8u * (n & 1u); };
  if (tid == 0) {
    for (int i = 0; i < SQ; ++i) { mbar_init(qk_full(i), 1); mbar_init(qk_empty(i), 1); }
    for (int i = 0; i < SV; ++i) { mbar_init(v_full(i), 1); mbar_init(v_empty(i), 1); }
    for (uint32_t i = 0; i < 2; ++i) { mbar_init(s_full(i), 1); mbar_init(s_empty(i), 128); }
    mbar_init(p_full, 128); mbar_init(p_empty, 1); mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == A_WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot) : "memory");
  const uint32_t tm_s = tmem, tm_o = tmem + 256;

  if (warp < 4) {
    // =================================== softmax / epilogue ============================================
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    const int row = warp * 32 + lane;            // query row inside the tile
    uint32_t ns = 0, np = 0;
    float m = -INFINITY;
    for (int t = 0; t < NT; ++t) {               // pass 1: row shift
      mbar_wait(s_full(ns), (ns >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int g = 0; g < TK / 32; ++g) {
        uint32_t r[32];
        tmem_ld32(tm_s + (ns & 1u) * TK + lane_addr + g * 32, r);
#pragma unroll
        for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(r[j]));
      }
      tc_fence_before();
      mbar_arrive(s_empty(ns));
      ++ns;
    }
    float l = 0.f;
    for (int t = 0; t < NT; ++t) {               // pass 2: P tiles
      mbar_wait(s_full(ns), (ns >> 1) & 1u);
      tc_fence_after();
      mbar_wait(p_empty, (np & 1) ^ 1);          // previous P tile consumed by its MMAs
#pragma unroll 1
      for (int g = 0; g < TK / 32; ++g) {
        uint32_t r[32];
        tmem_ld32(tm_s + (ns & 1u) * TK + lane_addr + g * 32, r);
        float pv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          pv[j] = expf((__uint_as_float(r[j]) - m) * scale);
          l += pv[j];
        }
#pragma unroll
        for (int kg = 0; kg < 4; ++kg) {         // 8 keys = one k-core row of this query
          float rr[8];
          uint4 hi, lo;
          hi.x = pack_bf16x2(pv[kg * 8 + 0], pv[kg * 8 + 1], rr[0], rr[1]);
          hi.y = pack_bf16x2(pv[kg * 8 + 2], pv[kg * 8 + 3], rr[2], rr[3]);
          hi.z = pack_bf16x2(pv[kg * 8 + 4], pv[kg * 8 + 5], rr[4], rr[5]);
          hi.w = pack_bf16x2(pv[kg * 8 + 6], pv[kg * 8 + 7], rr[6], rr[7]);
          float d0, d1;
          lo.x = pack_bf16x2(rr[0], rr[1], d0, d1);
          lo.y = pack_bf16x2(rr[2], rr[3], d0, d1);
          lo.z = pack_bf16x2(rr[4], rr[5], d0, d1);
          lo.w = pack_bf16x2(rr[6], rr[7], d0, d1);
          const uint32_t off = p_base + (uint32_t)((g * 4 + kg) * TQ + row) * 16u;   // [k-core][row][16 B]
          st_shared_v4(off, hi);
          st_shared_v4(off + P_BYTES / 2, lo);
        }
      }
      tc_fence_before();
      mbar_arrive(s_empty(ns));
      fence_async_shared();
      mbar_arrive(p_full);
      ++ns; ++np;
    }
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = 1.0f / l;
    float* orow = o + ((size_t)b * N + q0 + row) * C;
    for (int g = 0; g < C / 32 + ((C % 32) ? 1 : 0); ++g) {
      uint32_t r[32];
      tmem_ld32(tm_o + lane_addr + g * 32, r);
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (g * 32 + j >= C) break;
        *reinterpret_cast<float4*>(orow + g * 32 + j) = make_float4(__uint_as_float(r[j]) * inv, __uint_as_float(r[j + 1]) * inv,
                                                                     __uint_as_float(r[j + 2]) * inv, __uint_as_float(r[j + 3]) * inv);
      }
    }
  } else if (warp < A_WARP_MMA) {
    // =================================== copy issuers ====================================================
    // item n of a ring belongs to issuer n % A_PGROUPS; every issuer walks the whole schedule in the MMA's order and skips
    // foreign items (a ring slot is always filled by the same issuer, see the static_assert above)
    if (lane == 0) {
      const int grp = warp - 4;
      const int NTq = N / TQ;
      uint32_t nq = 0, nv = 0;                   // ring item counters (all issuers count alike)
      const uint8_t* qblk = qp + ((size_t)b * NTq + blockIdx.x) * NC * 8192;
      const uint32_t v_bytes = (uint32_t)C * 64u;
      auto load_qk = [&](int t) {
        const uint8_t* kblk = kp + ((size_t)b * NT + t) * NC * 8192;
        for (int c = 0; c < NC; ++c, ++nq) {
          if ((int)(nq % A_PGROUPS) != grp) continue;
          const int sq = (int)(nq % SQ);
          mbar_wait(qk_empty(sq), ((nq / SQ) & 1u) ^ 1u);
          const uint32_t st = qk_base + sq * QK_STAGE;
          mbar_arrive_expect_tx(qk_full(sq), 16384u);
          bulk_g2s(st, qblk + (size_t)c * 8192, 8192u, qk_full(sq));              // Q: hi at +0, lo at +4096
          bulk_g2s(st + 8192, kblk + (size_t)c * 8192, 8192u, qk_full(sq));       // K: hi at +8192, lo at +12288
        }
      };
      auto load_v = [&](int t) {
        const uint8_t* vblk = vp + ((size_t)b * NT + t) * (TK / 16) * (size_t)v_bytes;
        for (int ks = 0; ks < TK / 16; ++ks, ++nv) {
          if ((int)(nv % A_PGROUPS) != grp) continue;
          const int sv = (int)(nv % SV);
          mbar_wait(v_empty(sv), ((nv / SV) & 1u) ^ 1u);
          mbar_arrive_expect_tx(v_full(sv), v_bytes);
          bulk_g2s(v_base + sv * V_STAGE_MAX, vblk + (size_t)ks * v_bytes, v_bytes, v_full(sv));
        }
      };
      // the same order as the MMA issuer below (a different order can deadlock: both rings are consumed in issue order)
      for (int t = 0; t < NT; ++t) load_qk(t);                 // pass 1
      load_qk(0);                                               // pass 2
      for (int t = 0; t < NT; ++t) {
        if (t + 1 < NT) load_qk(t + 1);
        load_v(t);
      }
    }
  } else if (lane == 0) {
    // =================================== MMA issuer =====================================================
    const uint32_t idesc_s = make_idesc(TK), idesc_o = make_idesc(C);
    int sq = 0, sv = 0;
    uint32_t pq = 0, pvp = 0, ns = 0, np = 0;
    auto issue_qk = [&](bool exact) {            // S(next buffer) = Q K^T of the next key tile in ring order
      mbar_wait(s_empty(ns), ((ns >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d = tm_s + (ns & 1u) * TK;
      for (int c = 0; c < NC; ++c) {
        mbar_wait(qk_full(sq), pq);
        tc_fence_after();
        const uint32_t st = qk_base + sq * QK_STAGE;
        const uint64_t q_hi = make_desc(st, TQ * 16, 128), q_lo = make_desc(st + 4096, TQ * 16, 128);
        const uint64_t k_hi = make_desc(st + 8192, TK * 16, 128), k_lo = make_desc(st + 12288, TK * 16, 128);
        umma_bf16(d, q_hi, k_hi, idesc_s, c > 0);
        if (exact) {
          umma_bf16(d, q_lo, k_hi, idesc_s, 1u);
          umma_bf16(d, q_hi, k_lo, idesc_s, 1u);
        }
        umma_commit(qk_empty(sq));
        if (++sq == SQ) { sq = 0; pq ^= 1; }
      }
      umma_commit(s_full(ns));
      ++ns;
    };
    auto issue_pv = [&](int t) {                 // O += P(t) V(t)
      mbar_wait(p_full, np & 1);
      tc_fence_after();
      for (int ks = 0; ks < TK / 16; ++ks) {
        mbar_wait(v_full(sv), pvp);
        tc_fence_after();
        const uint32_t st = v_base + sv * V_STAGE_MAX;
        const uint32_t pa = p_base + (uint32_t)(2 * ks) * TQ * 16;
        const uint64_t p_hi = make_desc(pa, TQ * 16, 128), p_lo = make_desc(pa + P_BYTES / 2, TQ * 16, 128);
        const uint64_t v_hi = make_desc(st, (uint32_t)C * 16, 128), v_lo = make_desc(st + v_plane, (uint32_t)C * 16, 128);
        umma_bf16(tm_o, p_hi, v_hi, idesc_o, (t > 0 || ks > 0) ? 1u : 0u);
        umma_bf16(tm_o, p_lo, v_hi, idesc_o, 1u);
        umma_bf16(tm_o, p_hi, v_lo, idesc_o, 1u);
        umma_commit(v_empty(sv));
        if (++sv == SV) { sv = 0; pvp ^= 1; }
      }
      umma_commit(p_empty);
      ++np;
    };
    for (int t = 0; t < NT; ++t) issue_qk(false);              // pass 1: hi x hi only, row shift
    issue_qk(true);                                            // pass 2: S(t+1) is issued before P(t) V(t), i.e. under softmax(t)
    for (int t = 0; t < NT; ++t) {
      if (t + 1 < NT) issue_qk(true);
      issue_pv(t);
    }
    umma_commit(o_full);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == A_WARP_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

}  // namespace

bool attention_tc_eligible(int N, int C) { return N >= TK && N % TK == 0 && C % 16 == 0 && C >= 16 && C <= 256; }

size_t attention_tc_scratch_bytes(int B, int N, int C) { return (size_t)3 * B * N * C * 4; }

// scratch: attention_tc_scratch_bytes(B, N, C) of device memory for the packed operands (the caller's workspace)
int launch_attention_tc(const float* q, const float* k, const float* vT, float* o, int B, int N, int C, void* scratch, cudaStream_t st) {
  if (!attention_tc_eligible(N, C)) return fail(-1, "attention_tc: shape not eligible");
  if (!scratch) return fail(-1, "attention_tc: no scratch buffer");
  const size_t per = (size_t)B * N * C * 4;
  uint8_t* qp = reinterpret_cast<uint8_t*>(scratch);
  uint8_t* kp = qp + per;
  uint8_t* vp = kp + per;
  {
    const size_t items = (size_t)3 * B * N * (C / 8);
    const unsigned blocks = (unsigned)((items + 255) / 256 > 148 * 16 ? 148 * 16 : (items + 255) / 256);
    storm_launch(attn_pack_kernel, blocks, 256, 0, st, q, k, vT, qp, kp, vp, B, N, C);
    STORM_LAUNCH_CHECK();
  }
  const int smem = SQ * QK_STAGE + SV * V_STAGE_MAX + P_BYTES + 8 * (2 * SQ + 2 * SV + 8) + 16;
  STORM_SET_SMEM_ONCE((attn_tc_kernel), 200 * 1024);
  storm_launch_m(2, attn_tc_kernel, dim3(N / TQ, B), A_THREADS, smem, st, (const uint8_t*)qp, (const uint8_t*)kp, (const uint8_t*)vp, o, N, C,
                 1.0f / sqrtf((float)C));
  STORM_LAUNCH_CHECK();
  return 0;
}

}  // namespace storm
