// tcgen05 implicit-GEMM convolution - placeholder until the kernel lands.
#include "conv_tc.cuh"
namespace storm {
bool conv_tc_eligible(const ConvParams&) { return false; }
int launch_conv_tc(const ConvParams&, cudaStream_t) { return fail(-3, "conv_tc: not built"); }
int tc_pack_weights(const float*, int, int, int, void** out, std::vector<void*>*, cudaStream_t) { *out = nullptr; return 0; }
int tc_pack_weights_io(const float*, int, int, void** out, std::vector<void*>*, cudaStream_t) { *out = nullptr; return 0; }
}  // namespace storm
