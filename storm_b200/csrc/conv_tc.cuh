// tcgen05 implicit-GEMM convolution: weight packing entry points (kernel in conv_tc.cu).
#pragma once
#include <vector>

#include "common.cuh"

namespace storm {
// OIHW fp32 [Cout][Cin][taps] -> bf16 hi/lo shared-memory images; *out = nullptr when the shape is not eligible.
int tc_pack_weights(const float* w_oihw, int Cout, int Cin, int taps, void** out, std::vector<void*>* owned, cudaStream_t st);
// [Cin][Cout] fp32 (NIN) -> same layout with taps = 1
int tc_pack_weights_io(const float* w_io, int Cin, int Cout, void** out, std::vector<void*>* owned, cudaStream_t st);
}  // namespace storm
