// Shared declarations for the storm_b200 native engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <mutex>
#include <string>
#include <utility>

namespace storm {

// ---- error plumbing ---------------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define STORM_CUDA_TRY(expr)                                                                       \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::storm::fail(-2, std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " + \
                                   __FILE__ + ":" + std::to_string(__LINE__));                     \
  } while (0)

#define STORM_LAUNCH_CHECK()                                                                        \
  do {                                                                                              \
    cudaError_t _e = cudaGetLastError();                                                            \
    if (_e != cudaSuccess)                                                                          \
      return ::storm::fail(-2, std::string("kernel launch failed: ") + cudaGetErrorString(_e) +     \
                                   " at " + __FILE__ + ":" + std::to_string(__LINE__));             \
  } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per DEVICE: remembered per call site and device, so a second handle on another
// GPU of the same process gets it too
#define STORM_SET_SMEM_ONCE(kernel, bytes)                                                           \
  do {                                                                                               \
    static std::mutex mu_;                                                                           \
    static bool done_[64] = {};                                                                      \
    int dev_ = 0;                                                                                    \
    STORM_CUDA_TRY(cudaGetDevice(&dev_));                                                            \
    std::lock_guard<std::mutex> lk_(mu_);                                                            \
    if (dev_ < 0 || dev_ >= 64 || !done_[dev_]) {                                                    \
      STORM_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (bytes))); \
      if (dev_ >= 0 && dev_ < 64) done_[dev_] = true;                                                \
    }                                                                                                \
  } while (0)

#define STORM_TRY(expr)          \
  do {                           \
    int _rc = (expr);            \
    if (_rc != 0) return _rc;    \
  } while (0)

// ---- activation tensors: NHWC fp32, H = frequency bins (F), W = frames (T) ----------------------------
struct Act {
  float* p = nullptr;
  int B = 0, H = 0, W = 0, C = 0;
  double* part = nullptr;   // GroupNorm partial sums [B][nchunk][C][2] (sum, sum of squares), or null
  int nchunk = 0;
  bool part_valid = false;
  size_t elems() const { return (size_t)B * H * W * C; }
};

// ---- implicit-GEMM convolution description -----------------------------------------------------------
// One K-segment of the GEMM: `taps` (9 = 3x3 pad 1, 1 = 1x1) over the channel concatenation
// [src0 (C0 ch) | src1 (C1 ch)], optionally passed through y = x*scale + shift (per sample, channel)
// and SiLU before it enters the convolution (zero padding is applied AFTER this prologue, as
// F.conv2d(padding=1) does on the activated tensor: layerspp.py:243-260).
struct ConvSeg {
  const float* src0;
  const float* src1;
  int C0, C1;
  const float* ss;    // [B][C0+C1][2] scale, shift; null = identity
  int silu;
  int taps;
  const float* w;     // SIMT path: [taps][C0+C1][Cout]
  const void* wtc;    // tcgen05 path: packed shared-memory operand images (TcWeights::img, see conv_tc.cuh)
  int wprec;          // operand split the image was packed in (TcWeights::prec)
  float wscale;       // power-of-two factor the packed weights carry (TcWeights::scale; 1 for the bf16 x 3 split)
};

struct ConvParams {
  ConvSeg seg[2];
  int nseg;
  int B, H, W, Cout;
  const float* bias;       // [Cout] or null
  const float* temb;       // [B][temb_stride] per-sample bias or null
  int temb_stride;
  const float* res;        // residual NHWC [B][H][W][Cout] or null
  // residual given at HALF resolution, NHWC [B][H/2][W/2][Cout]: the epilogue adds its x2 FIR upsampling (taps [1,3,3,1],
  // up_or_down_sampling.py:195-217: o[2i] = x[i-1]/4 + 3x[i]/4, o[2i+1] = 3x[i]/4 + x[i+1]/4 per axis, zero outside).  Used for
  // the 1x1 skip of up blocks, which commutes with the (linear, per-channel) resampling: Conv_2(FIR_up(x)) = FIR_up(Conv_2(x)).
  const float* res_up;
  float out_scale;
  float* out;              // NHWC [B][H][W][Cout]
  double* stats_out;       // optional GroupNorm partial sums of `out`: [B][stats_chunks][Cout][2] (tcgen05 mode P only)
  int stats_chunks;
  long long* dbg;          // optional [16] cycle counters written by CTA 0 (tools/conv_timeline.py)
  int out_cm;              // 1: write `out` channel-major [B][Cout][H*W] (tcgen05 kernel only; attention V operand)
  // Magnitude of the RAW (un-normalised) operands of this launch: [B][raw_ms_n] mean squares (gn_finalize writes them for
  // the GroupNorm groups of the same tensor).  The f16-based operand split of the tcgen05 kernel scales every operand of
  // sample b by one power of two alpha_b <= 1 chosen from max_g raw_ms[b][g], so that a residual stream beyond fp16's
  // range (65504) does not saturate; the epilogue multiplies by 1 / alpha_b.  null: alpha = 1.
  const float* raw_ms;
  int raw_ms_n;
};

int launch_conv_simt(const ConvParams& p, cudaStream_t st);
bool conv_tc_eligible(const ConvParams& p);
// number of GroupNorm partial-sum chunks the tcgen05 kernel would emit for this launch (0 = cannot emit)
int conv_tc_stats_chunks(const ConvParams& p);
int launch_conv_tc(const ConvParams& p, cudaStream_t st);

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + expf(-v)); }
// fast variant for the tensor-core producers (ex2.approx + approximate division, ~1e-6 relative)
// x * sigmoid(x) with the two MUFU approximations issued directly: __expf() wraps ex2 in a denormal-range guard (a compare and
// two predicated multiplies per element) that an activation never needs - ex2.approx.ftz returns 0 below 2^-126 and 1 + 0
// is exact; error ~2 ulp like __fdividef(v, 1 + __expf(-v)).
__device__ __forceinline__ float silu_fast(float v) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return v * r;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch ----------------------------------------------------------------------
// Every kernel of the engine starts with pdl_enter(): it lets the NEXT kernel of the stream be scheduled as soon as this grid's
// CTAs are all resident (the next kernel's CTAs take the SMs this grid's tail leaves idle and run their prologue there), then
// waits until the PREVIOUS grid has completed and its memory is visible.  Nothing global is read or written before the wait,
// so the stream keeps its serial semantics; without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
  pdl_launch_dependents();
  pdl_wait();
}
int pdl_mode();          // STORM_B200_PDL (default 0 = off) bit 0: small kernels, bit 1: conv kernels carry the launch attribute; engine.cu
// kernel<<<grid, block, smem, st>>>(args...) with the programmatic-stream-serialization attribute when enabled
template <typename... KArgs, typename... Args>
inline cudaError_t storm_launch_m(int mask, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t lc{};
  lc.gridDim = grid;
  lc.blockDim = block;
  lc.dynamicSmemBytes = smem;
  lc.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = at;
  lc.numAttrs = (pdl_mode() & mask) ? 1 : 0;
  return cudaLaunchKernelEx(&lc, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t storm_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  return storm_launch_m(1, kernel, grid, block, smem, st, std::forward<Args>(args)...);
}

}  // namespace storm
