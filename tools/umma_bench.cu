// Micro-benchmark: raw tcgen05.mma issue rate for the operand layouts conv_tc.cu uses (no data movement).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench umma_bench.cu && ./umma_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a),
               "l"(b), "r"(idesc), "r"(acc) : "memory");
}
struct Cfg { int n; int a_lbo, a_sbo, b_lbo, b_sbo; int nacc; int iters; int a_step; int b_step; int bulk; };

__global__ void __launch_bounds__(128, 1) bench(Cfg c, long long* out, const uint8_t* gsrc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t tm = slot;
  if (threadIdx.x == 32 && c.bulk) {
    // background bulk copies into the upper 64 KB of smem while the MMAs run
    uint32_t dst = smem_u32(smem) + 128 * 1024;
    for (int i = 0; i < c.iters / 6; ++i) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar2)), "r"(16384));
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + (i & 3) * 16384),
                   "l"(gsrc + (size_t)(i & 63) * 16384), "r"(16384), "r"(smem_u32(&bar2)) : "memory");
      uint32_t ok = 0;
      while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(smem_u32(&bar2)), "r"(i & 1) : "memory");
    }
  }
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(c.n);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 32 * 1024;
    long long t0 = clock64();
    if (c.bulk == 2) {
      // lean issue loop: descriptors precomputed, 6 MMAs per iteration, descriptor advance = one 64-bit add
      const uint64_t da0 = make_desc(a0, c.a_lbo, c.a_sbo), da1 = make_desc(a0 + 8192, c.a_lbo, c.a_sbo);
      const uint64_t db0 = make_desc(b0, c.b_lbo, c.b_sbo), db1 = make_desc(b0 + 16384, c.b_lbo, c.b_sbo);
      const uint32_t d0 = tm, d1 = tm + c.n;
      uint64_t off = 0;
      for (int i = 0; i < c.iters; i += 6) {
        umma(d0, da0, db0 + off, idesc, i > 0);
        umma(d0, da0, db1 + off, idesc, 1);
        umma(d0, da1, db0 + off, idesc, 1);
        umma(d1, da0 + 256, db0 + off, idesc, i > 0);
        umma(d1, da0 + 256, db1 + off, idesc, 1);
        umma(d1, da1 + 256, db0 + off, idesc, 1);
        off = (off + 1) & 7;
      }
    } else
    for (int i = 0; i < c.iters; ++i) {
      uint32_t aa = a0 + (i % 8) * c.a_step, bb = b0 + (i % 9) * c.b_step;
      umma(tm + (i % c.nacc) * c.n, make_desc(aa, c.a_lbo, c.a_sbo), make_desc(bb, c.b_lbo, c.b_sbo), idesc, i >= c.nacc);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  uint8_t* g; cudaMalloc(&g, 64 * 16384); cudaMemset(g, 1, 64 * 16384);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct { const char* name; Cfg c; } tests[] = {
    {"modeP N=256 hx=10 (A=w: lbo128 sbo256; B=px: lbo5440 sbo160)", {256, 128, 256, 5440, 160, 2, 4096, 4096, 16, 0}},
    {"modeP N=256 hx=18 (B lbo9792 sbo288)", {256, 128, 256, 9792, 288, 2, 4096, 4096, 16, 0}},
    {"dense N=256 (B lbo4096 sbo128)", {256, 128, 256, 4096, 128, 2, 4096, 4096, 16, 0}},
    {"dense N=256 canonical A/B lbo=128 sbo=256", {256, 128, 256, 128, 256, 2, 4096, 4096, 16, 0}},
    {"modeP N=128 hx=10", {128, 128, 256, 5440, 160, 4, 4096, 4096, 16, 0}},
    {"modeP N=256 1 acc", {256, 128, 256, 5440, 160, 1, 4096, 4096, 16, 0}},
    {"modeP N=256 hx=10 + bulk copies", {256, 128, 256, 5440, 160, 2, 4096, 4096, 16, 1}},
    {"modeC M=px N=256 (A=px lbo5184 sbo288; B=w)", {256, 5184, 288, 128, 256, 2, 4096, 16, 4096, 0}},
    {"LEAN modeP N=256 hx=10", {256, 128, 256, 5440, 160, 2, 4092, 4096, 16, 2}},
    {"LEAN modeP N=128 hx=10", {128, 128, 256, 5440, 160, 2, 4092, 4096, 16, 2}},
    {"LEAN dense N=256 (B lbo4096 sbo128)", {256, 128, 256, 4096, 128, 2, 4092, 4096, 16, 2}},
    {"LEAN N=64", {64, 128, 256, 5440, 160, 2, 4092, 4096, 16, 2}},
  };
  for (auto& t : tests) {
    for (int grid : {1, 148}) {
      bench<<<grid, 128, 200 * 1024>>>(t.c, d, g);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("%-62s grid=%3d  %7.1f cycles/MMA (ideal %d)  %s\n", t.name, grid, (double)mx / t.c.iters, t.c.n / 2, cudaGetErrorString(e));
    }
  }
  return 0;
}
