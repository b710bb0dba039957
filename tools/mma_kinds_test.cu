// Numerical + rate check of the tcgen05.mma kinds a "1 fp16 pass + 2 fp8 correction passes" operand split would use:
//   kind::f16 with F16 x F16 (K = 16)   and   kind::f8f6f4 with E4M3 x E5M2 / E5M2 x E4M3 (K = 32),
// both with the no-swizzle K-major core-matrix layout of conv_tc.cu (core matrix = 8 rows x 16 bytes).
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/mma_kinds_test tools/mma_kinds_test.cu
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// D = f32; afmt / bfmt: kind::f16 0 = F16, 1 = BF16; kind::f8f6f4 0 = E4M3, 1 = E5M2
__host__ __device__ inline uint32_t make_idesc(int n, int afmt, int bfmt) {
  return (1u << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a),
               "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_f8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a),
               "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// block-scaled fp4: kind::mxf4 (e2m1 x e2m1, one ue8m0 scale per 32 elements of K, K = 64 per instruction); scale factors in TMEM.
// Instruction descriptor of the block-scaled kinds (no accumulator-format field: bits 4-5 are the B scale-factor id):
// a_format [7,10) = b_format [10,13) = 1 (e2m1), K-major, N >> 3 at [17,23), scale_format [23] = 1 (ue8m0), M >> 4 at [24,29), K64 [31] = 0
__host__ __device__ inline uint32_t make_idesc_mxf4(int n) {
  return (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (1u << 23) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_mxf4(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc, uint32_t sfa, uint32_t sfb) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc),
               "r"(acc), "r"(sfa), "r"(sfb) : "memory");
}
__device__ __forceinline__ void commit_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  uint32_t ok = 0;
  while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// operand images are prepared on the host in the core-matrix layout: element (row r, k) of an operand with `eb` bytes per
// element lives at byte (r/8)*SBO + (kb/16)*LBO + (r%8)*16 + kb%16 with kb = k*eb, LBO = 128, SBO = 256 (2 k-cores per row group)
constexpr int N = 64;
__global__ void __launch_bounds__(128, 1) numeric(const uint8_t* img, float* out) {
  // img: [A f16 4096][B f16 N*32][A8a e4m3 4096][B8a e5m2 N*32][A8b e5m2 4096][B8b e4m3 N*32]
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int total = 3 * (4096 + N * 32);
  for (int i = threadIdx.x; i < total; i += blockDim.x) smem[i] = img[i];
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t s0 = smem_u32(smem);
    const uint32_t per = 4096 + N * 32;
    umma_f16(tm + 0, make_desc(s0, 128, 256), make_desc(s0 + 4096, 128, 256), make_idesc(N, 0, 0), 0);
    umma_f8(tm + 64, make_desc(s0 + per, 128, 256), make_desc(s0 + per + 4096, 128, 256), make_idesc(N, 0, 1), 0);
    umma_f8(tm + 128, make_desc(s0 + 2 * per, 128, 256), make_desc(s0 + 2 * per + 4096, 128, 256), make_idesc(N, 1, 0), 0);
    // and all three accumulated into one tile, as the conv would do
    umma_f16(tm + 192, make_desc(s0, 128, 256), make_desc(s0 + 4096, 128, 256), make_idesc(N, 0, 0), 0);
    umma_f8(tm + 192, make_desc(s0 + per, 128, 256), make_desc(s0 + per + 4096, 128, 256), make_idesc(N, 0, 1), 1);
    umma_f8(tm + 192, make_desc(s0 + 2 * per, 128, 256), make_desc(s0 + 2 * per + 4096, 128, 256), make_idesc(N, 1, 0), 1);
    commit_wait(&bar, 0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t = 0; t < 4; ++t)
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t r[32];
      const uint32_t addr = tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t * 64 + c0);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                     "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                     "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                     "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                   : "r"(addr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) out[((size_t)t * 128 + warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256));
}

// ---- numerics of kind::mxf4.block_scale.block32: D[128 x N] = sum over the two 32-element K blocks of sfa[m][blk] * sfb[n][blk] * (A_blk . B_blk)
// img: [A fp4 v0 4096][B fp4 v0 N*32][A fp4 v1 4096][B fp4 v1 N*32][SFA atom 512][SFB atom 512]   (v0: element 2i in the low nibble, v1: high)
__global__ void __launch_bounds__(128, 1) numeric_mxf4(const uint8_t* img, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int per = 4096 + N * 32, total = 2 * per + 1024;
  for (int i = threadIdx.x; i < total; i += blockDim.x) smem[i] = img[i];
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t s0 = smem_u32(smem);
    const uint32_t sfa_s = s0 + 2 * per, sfb_s = sfa_s + 512;
    const uint32_t sfa_t = tm + 128, sfb_t = tm + 132;
    // scale factors: 32 rows x 128 bit per copy (8-row core matrices of 128 contiguous bytes), broadcast to the four lane quadrants
    asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(sfa_t), "l"(make_desc(sfa_s, 128, 128)) : "memory");
    asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(sfb_t), "l"(make_desc(sfb_s, 128, 128)) : "memory");
    umma_mxf4(tm + 0, make_desc(s0, 128, 256), make_desc(s0 + 4096, 128, 256), make_idesc_mxf4(N), 0, sfa_t, sfb_t);
    umma_mxf4(tm + 64, make_desc(s0 + per, 128, 256), make_desc(s0 + per + 4096, 128, 256), make_idesc_mxf4(N), 0, sfa_t, sfb_t);
    commit_wait(&bar, 0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t = 0; t < 2; ++t)
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t r[32];
      const uint32_t addr = tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t * 64 + c0);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                     "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                     "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                     "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                   : "r"(addr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) out[((size_t)t * 128 + warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256));
}

// rate: per "tap" of a 32-channel chunk the new split issues 2 f16 (K=16) + 2 fp8 (K=32) MMAs, the current one 6 bf16 MMAs
__global__ void __launch_bounds__(128, 1) rate(int n, int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x38003800u;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 32 * 1024;
    const uint64_t da = make_desc(a0, 128, 256), db = make_desc(b0, 5440, 160);
    const uint32_t i16 = make_idesc(n, 0, 0), i8a = make_idesc(n, 0, 1), i8b = make_idesc(n, 1, 0), ibf = make_idesc(n, 1, 1);
    long long t0 = clock64();
    uint64_t off = 0;
    for (int i = 0; i < iters; ++i) {
      for (int acc = 0; acc < 2; ++acc) {
        const uint32_t d = tm + acc * n;
        const uint64_t wa = da + 256 * acc;
        if (mode == 0) {            // bf16 x3 on two 16-channel chunks
          umma_f16(d, wa, db + off, ibf, i > 0); umma_f16(d, wa, db + 680 + off, ibf, 1); umma_f16(d, wa + 512, db + off, ibf, 1);
          umma_f16(d, wa + 1024, db + 1360 + off, ibf, 1); umma_f16(d, wa + 1024, db + 2040 + off, ibf, 1); umma_f16(d, wa + 1536, db + 1360 + off, ibf, 1);
        } else if (mode == 2) {     // f16 x2 (K=16 each) + ONE mxf4 (K=64: both correction products of the 32 channels concatenated along K)
          umma_f16(tm, wa, db + off, i16, i > 0); umma_f16(tm, wa + 512, db + 680 + off, i16, 1);
          umma_mxf4(tm, wa + 1024, db + 1360 + off, make_idesc_mxf4(n), 1, tm + 256, tm + 264);
        } else if (mode == 3) {     // mxf4 only: cycles per instruction (timing only - operands and scale factors are whatever is there)
          umma_mxf4(tm, wa + 1024, db + 1360 + off, make_idesc_mxf4(n), i > 0, tm + 256, tm + 264);
        } else if (mode == 4) {     // fp8 only, for comparison
          umma_f8(tm, wa + 1024, db + 1360 + off, i8a, i > 0);
        } else {                    // f16 x2 (K=16 each) + fp8 x2 (K=32 each)
          umma_f16(d, wa, db + off, i16, i > 0); umma_f16(d, wa + 512, db + 680 + off, i16, 1);
          umma_f8(d, wa + 1024, db + 1360 + off, i8a, 1); umma_f8(d, wa + 1536, db + 2040 + off, i8b, 1);
        }
      }
      off = (off + 1) & 7;
    }
    commit_wait(&bar, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

static void put(std::vector<uint8_t>& img, size_t base, int r, int k, int eb, const void* v) {
  const int kb = k * eb;
  memcpy(&img[base + (size_t)(r / 8) * 256 + (kb / 16) * 128 + (r % 8) * 16 + kb % 16], v, eb);
}

int main() {
  // ---- numerics
  const int per = 4096 + N * 32;
  std::vector<uint8_t> img(3 * per, 0);
  auto Af = [](int m, int k) { return (float)((m + 2 * k) % 7 - 3) * 0.5f; };
  auto Bf = [](int n, int k) { return (float)((n * 3 + k) % 5 - 2) * 0.25f; };
  for (int m = 0; m < 128; ++m) for (int k = 0; k < 16; ++k) { __half h = __float2half(Af(m, k)); put(img, 0, m, k, 2, &h); }
  for (int n = 0; n < N; ++n) for (int k = 0; k < 16; ++k) { __half h = __float2half(Bf(n, k)); put(img, 4096, n, k, 2, &h); }
  for (int m = 0; m < 128; ++m) for (int k = 0; k < 32; ++k) {
    __nv_fp8_e4m3 a(Af(m, k)); put(img, per, m, k, 1, &a);
    __nv_fp8_e5m2 b(Af(m, k)); put(img, 2 * per, m, k, 1, &b);
  }
  for (int n = 0; n < N; ++n) for (int k = 0; k < 32; ++k) {
    __nv_fp8_e5m2 a(Bf(n, k)); put(img, per + 4096, n, k, 1, &a);
    __nv_fp8_e4m3 b(Bf(n, k)); put(img, 2 * per + 4096, n, k, 1, &b);
  }
  uint8_t* dimg; float* dout; cudaMalloc(&dimg, img.size()); cudaMalloc(&dout, 4 * 128 * N * 4);
  cudaMemcpy(dimg, img.data(), img.size(), cudaMemcpyHostToDevice);
  numeric<<<1, 128, 3 * per>>>(dimg, dout);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> out(4 * 128 * N);
  cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  const char* names[4] = {"kind::f16 F16xF16 K=16", "kind::f8f6f4 E4M3xE5M2 K=32", "kind::f8f6f4 E5M2xE4M3 K=32", "all three into one accumulator"};
  for (int t = 0; t < 4; ++t) {
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
      double ref16 = 0, ref32 = 0;
      for (int k = 0; k < 16; ++k) ref16 += (double)Af(m, k) * Bf(n, k);
      for (int k = 0; k < 32; ++k) ref32 += (double)Af(m, k) * Bf(n, k);
      const double ref = t == 0 ? ref16 : t == 3 ? ref16 + 2 * ref32 : ref32;
      const double err = fabs(out[((size_t)t * 128 + m) * N + n] - ref);
      if (err > maxerr) maxerr = err;
      bad += err > 1e-6;
    }
    printf("%-34s %s  max abs err %.3g  mismatches %d / %d\n", names[t], cudaGetErrorString(e), maxerr, bad, 128 * N);
  }
  // ---- numerics of the block-scaled fp4 kind
  {
    static const float e2m1[8] = {0.f, 0.5f, 1.f, 1.5f, 2.f, 3.f, 4.f, 6.f};
    auto A4 = [](int m, int k) { return (m * 5 + k * 3) % 15; };            // 4-bit code: bit 3 = sign, bits 0-2 = magnitude index
    auto B4 = [](int n, int k) { return (n * 7 + k * 11 + 3) % 16; };
    auto val = [&](int code) { return (code & 8 ? -1.f : 1.f) * e2m1[code & 7]; };
    auto sfa = [](int m, int blk) { return (m * 3 + blk) % 5 - 2; };          // exponent of the ue8m0 scale
    auto sfb = [](int n, int blk) { return (n + 2 * blk) % 4 - 1; };
    const int per4 = 4096 + N * 32;
    std::vector<uint8_t> im4(2 * per4 + 1024, 0);
    for (int v = 0; v < 2; ++v) {
      auto put4 = [&](size_t base, int r, int k, int code) {
        const int b = k / 2;
        uint8_t& by = im4[base + (size_t)(r / 8) * 256 + (b / 16) * 128 + (r % 8) * 16 + b % 16];
        const bool hi = ((k & 1) != 0) != (v == 1);
        by = hi ? (uint8_t)((by & 0x0F) | (code << 4)) : (uint8_t)((by & 0xF0) | code);
      };
      for (int m = 0; m < 128; ++m) for (int k = 0; k < 64; ++k) put4((size_t)v * per4, m, k, A4(m, k));
      for (int n = 0; n < N; ++n) for (int k = 0; k < 64; ++k) put4((size_t)v * per4 + 4096, n, k, B4(n, k));
    }
    for (int r = 0; r < 128; ++r) for (int j = 0; j < 4; ++j) {
      im4[2 * per4 + (r % 32) * 16 + (r / 32) * 4 + j] = (uint8_t)(127 + (j < 2 ? sfa(r, j) : 0));
      im4[2 * per4 + 512 + (r % 32) * 16 + (r / 32) * 4 + j] = (uint8_t)(127 + ((j < 2 && r < N) ? sfb(r, j) : 0));
    }
    uint8_t* d4; float* o4; cudaMalloc(&d4, im4.size()); cudaMalloc(&o4, 2 * 128 * N * 4);
    cudaMemcpy(d4, im4.data(), im4.size(), cudaMemcpyHostToDevice);
    cudaMemset(o4, 0, 2 * 128 * N * 4);
    numeric_mxf4<<<1, 128, 2 * per4 + 1024>>>(d4, o4);
    e = cudaDeviceSynchronize();
    std::vector<float> o(2 * 128 * N);
    cudaMemcpy(o.data(), o4, o.size() * 4, cudaMemcpyDeviceToHost);
    for (int v = 0; v < 2; ++v) {
      double maxerr = 0; int bad = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int blk = 0; blk < 2; ++blk) {
          double sacc = 0;
          for (int k = 32 * blk; k < 32 * blk + 32; ++k) sacc += (double)val(A4(m, k)) * val(B4(n, k));
          ref += sacc * exp2((double)sfa(m, blk) + sfb(n, blk));
        }
        const double err = fabs(o[((size_t)v * 128 + m) * N + n] - ref);
        if (err > maxerr) maxerr = err;
        bad += err > 1e-4;
        if (m < 2 && n < 3) printf("    v%d D[%d][%d] = %g  (reference %g)\n", v, m, n, o[((size_t)v * 128 + m) * N + n], ref);
      }
      printf("kind::mxf4.block_scale (%s nibble first)  %s  max abs err %.3g  mismatches %d / %d\n", v == 0 ? "low" : "high", cudaGetErrorString(e),
             maxerr, bad, 128 * N);
    }
  }
  // ---- rate
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  for (int n : {256, 240, 128}) for (int mode : {0, 1, 2, 3, 4}) {
    const int iters = 512;
    rate<<<148, 128, 160 * 1024>>>(n, mode, iters, d);
    e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("N=%d %-28s %8.1f cycles per (32 channels x 1 tap x 2 cout blocks)  [%d MMAs]  %s\n", n,
           mode == 0 ? "bf16x3 (6 MMAs per block)" : mode == 1 ? "f16 x2 + fp8 x2 (4 MMAs)" : mode == 2 ? "f16 x2 + mxf4 x1 (3 MMAs)"
           : mode == 3 ? "mxf4 x1 (K=64)" : "fp8 x1 (K=32)", (double)mx / iters, mode == 0 ? 12 : mode == 1 ? 8 : mode == 2 ? 6 : 2, cudaGetErrorString(e));
  }
  return 0;
}
