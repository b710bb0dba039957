// tcgen05 implicit-GEMM convolution: weight packing entry points (kernel in conv_tc.cu).
#pragma once
#include <vector>

#include "common.cuh"

namespace storm {
// OIHW fp32 [Cout][Cin][taps] -> shared-memory operand images; *out = nullptr when the shape is not eligible.
// Two operand splits (see conv_tc.cu): mode 1 "f16 + 2 x fp8" (32-channel K chunks; chosen when Cin % 32 == 0, the
// environment allows it and max_prec >= 1) and mode 0 "bf16 x 3" (16-channel chunks).  The mode is remembered per image;
// every K segment of one launch must use the same mode (pass max_prec = 0 for weights that share a launch with a segment
// whose channel count is not a multiple of 32).
int tc_pack_weights(const float* w_oihw, int Cout, int Cin, int taps, void** out, std::vector<void*>* owned, cudaStream_t st,
                    int max_prec = 1);
// [Cin][Cout] fp32 (NIN) -> same layout with taps = 1
int tc_pack_weights_io(const float* w_io, int Cin, int Cout, void** out, std::vector<void*>* owned, cudaStream_t st,
                       int max_prec = 1);
}  // namespace storm
