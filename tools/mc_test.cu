// Stand-alone check of cluster multicast primitives used for weight streaming experiments:
//   (1) cp.async.bulk .multicast::cluster of per-rank slices + per-CTA expect_tx
//   (2) tcgen05.commit .multicast::cluster arriving on a count=cl mbarrier in every CTA
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o mc_test tools/mc_test.cu ; run: ./mc_test
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  for (long i = 0; i < 20000000; ++i) if (mbar_try(bar, parity)) return true;
  return false;
}
__device__ __forceinline__ void bulk_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t crank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

constexpr int STAGE = 8192, NSTAGE = 4, ROUNDS = 6;

// out[cta][round][STAGE/4] words; flags[cta] = bit0 timeout on full, bit1 timeout on empty
__global__ void mc_kernel(const uint8_t* src, uint32_t* out, int* flags, int cl, int delay_rank) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t bars = base + NSTAGE * STAGE;
  auto full = [&](int i) { return bars + 8u * i; };
  auto empty = [&](int i) { return bars + 8u * (NSTAGE + i); };
  const uint32_t r = crank();
  const uint16_t mask = (uint16_t)((1u << cl) - 1u);
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(full(i), 1); mbar_init(empty(i), cl); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync();
  int fl = 0;
  const int cta = blockIdx.x;
  for (int round = 0; round < ROUNDS; ++round) {
    const int s = round % NSTAGE;
    const uint32_t ph = (round / NSTAGE) & 1;
    if (threadIdx.x == 0) {
      if ((int)r == delay_rank) { long long t0 = clock64(); while (clock64() - t0 < 200000) {} }   // peers' data arrives before expect_tx
      if (!mbar_wait(empty(s), ph ^ 1)) fl |= 2;
      mbar_expect(full(s), STAGE);
      const uint32_t bytes = STAGE / cl;
      bulk_mc(base + s * STAGE + r * bytes, src + (size_t)round * STAGE + r * bytes, bytes, full(s), mask);
    }
    if (threadIdx.x == 32) { if (!mbar_wait(full(s), ph)) fl |= 1; }
    __syncthreads();
    // wait done by one thread; everyone reads after the barrier (generic-proxy reads of async-proxy writes are
    // ordered by the mbarrier wait of thread 32 + bar.sync)
    if (threadIdx.x == 32) (void)mbar_wait(full(s), ph);
    __syncthreads();
    const uint32_t* sp = reinterpret_cast<const uint32_t*>(smem + s * STAGE);
    for (int i = threadIdx.x; i < STAGE / 4; i += blockDim.x) out[((size_t)cta * ROUNDS + round) * (STAGE / 4) + i] = sp[i];
    __syncthreads();
    if (threadIdx.x == 0) commit_mc(empty(s), mask);      // no MMAs outstanding: arrives immediately in every CTA
  }
  if (threadIdx.x == 0 || threadIdx.x == 32) atomicOr(&flags[cta], fl);
  __syncthreads();
  cluster_sync();
}

int main() {
  const int ncta = 8;
  std::vector<uint8_t> h((size_t)ROUNDS * STAGE);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)((i * 2654435761u) >> 13);
  uint8_t* d; uint32_t* o; int* f;
  cudaMalloc(&d, h.size()); cudaMalloc(&o, (size_t)ncta * ROUNDS * STAGE); cudaMalloc(&f, ncta * 4);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  const int smem = NSTAGE * STAGE + 8 * 2 * NSTAGE;
  cudaFuncSetAttribute(mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int cl : {1, 2, 4}) for (int delay : {-1, 0, 1}) {
    cudaMemset(o, 0, (size_t)ncta * ROUNDS * STAGE); cudaMemset(f, 0, ncta * 4);
    cudaLaunchConfig_t lc{}; lc.gridDim = dim3(ncta); lc.blockDim = dim3(128); lc.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&lc, mc_kernel, (const uint8_t*)d, o, f, cl, delay);
    cudaError_t e2 = cudaDeviceSynchronize();
    std::vector<uint8_t> ho((size_t)ncta * ROUNDS * STAGE); std::vector<int> hf(ncta);
    cudaMemcpy(ho.data(), o, ho.size(), cudaMemcpyDeviceToHost); cudaMemcpy(hf.data(), f, ncta * 4, cudaMemcpyDeviceToHost);
    printf("cl=%d delay_rank=%d launch=%s sync=%s :", cl, delay, cudaGetErrorString(e), cudaGetErrorString(e2));
    for (int c = 0; c < ncta; ++c) {
      long bad = 0;
      for (int rd = 0; rd < ROUNDS; ++rd) for (int i = 0; i < STAGE; ++i) bad += ho[((size_t)c * ROUNDS + rd) * STAGE + i] != h[(size_t)rd * STAGE + i];
      printf(" cta%d[bad=%ld fl=%d]", c, bad, hf[c]);
    }
    printf("\n");
  }
  return 0;
}
