// tcgen05 implicit-GEMM convolution: weight packing entry points (kernel in conv_tc.cu).
#pragma once
#include <vector>

#include "common.cuh"

namespace storm {
// A packed weight tensor: shared-memory operand images in device memory plus what the launcher must know about them.
struct TcWeights {
  void* img = nullptr;   // null when the shape is not eligible for the tcgen05 kernel
  int prec = 0;          // 0: bf16 x 3 (16-channel K chunks); 1: f16 + 2 x fp8 (32-channel K chunks)
  float scale = 1.0f;    // power of two the weights were multiplied by before the split (prec 1 only)
};
// max |w| over n floats (NaN / inf ignored); synchronises the stream (load time only)
int tc_weight_absmax(const float* w, size_t n, float* out_host, cudaStream_t st);
// OIHW fp32 [Cout][Cin][taps] -> operand images.  Two operand splits (see conv_tc.cu): mode 1 "f16 + 2 x fp8" (chosen when
// Cin % 32 == 0, the environment allows it and max_prec >= 1) and mode 0 "bf16 x 3".  Every K segment of one launch must
// use the same mode and, in mode 1, the same scale: `absmax` is the largest |w| over ALL weight tensors of the launch
// (tc_weight_absmax; <= 0: measure this tensor alone).  Pass max_prec = 0 for weights that share a launch with a segment
// whose channel count is not a multiple of 32.
int tc_pack_weights(const float* w_oihw, int Cout, int Cin, int taps, TcWeights* out, std::vector<void*>* owned, cudaStream_t st,
                    int max_prec = 1, float absmax = 0.f);
// [Cin][Cout] fp32 (NIN) -> same layout with taps = 1
int tc_pack_weights_io(const float* w_io, int Cin, int Cout, TcWeights* out, std::vector<void*>* owned, cudaStream_t st,
                       int max_prec = 1, float absmax = 0.f);
// per-device one-time setup (shared-memory carve-out attributes of the kernel instantiations); thread-safe
int conv_tc_init_device();
inline void set_wtc(ConvSeg& s, const TcWeights& w) { s.wtc = w.img; s.wprec = w.prec; s.wscale = w.scale; }
}  // namespace storm
