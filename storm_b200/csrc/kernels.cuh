// Launchers of the small (HBM- or latency-bound) kernels of the path.  All tensors NHWC fp32.
#pragma once
#include "common.cuh"

namespace storm {

// ---- weights ---------------------------------------------------------------------------------------
// OIHW [Cout][Cin][taps] -> [taps][Cin][Cout]
int launch_pack_conv_simt(const float* w, float* out, int Cout, int Cin, int taps, cudaStream_t st);
// [rows][cols] -> out[c * out_stride + out_off + r]   (transpose into a wider matrix)
int launch_transpose_into(const float* w, float* out, int rows, int cols, int out_stride, int out_off, cudaStream_t st);
int launch_add_vec(const float* a, const float* b, float* out, int n, float scale_b, cudaStream_t st);

// ---- time embedding (ncsnpp.py:298-317, layerspp.py:39-41, :263) ---------------------------------------
// act[b][4nf] = SiLU(L2(SiLU(L1(fourier(log t)))))
int launch_temb_mlp(const float* t, const float* Wf, const float* W1t, const float* b1, const float* W2t,
                    const float* b2, float* act, int B, int nf, cudaStream_t st);
// out[b][o] = bias[o] + sum_i Wt[i][o] * act[b][i]
int launch_dense_all(const float* act, const float* Wt, const float* bias, float* out, int B, int in_dim, int out_dim,
                     cudaStream_t st);

// ---- GroupNorm statistics (nn.GroupNorm eps=1e-6, groups=min(C/4,32): layerspp.py:219,231,67) ------------
int gn_num_chunks(int HW);
int launch_gn_partial(const Act& a, cudaStream_t st);   // fills a.part
// scale/shift [B][C0+C1][2] from the partials of one or two (concatenated) tensors; ms (optional, [B][G], G = gn_groups(C))
// receives the mean square of every RAW group (what the tensor-core kernel scales un-normalised operands by)
int gn_groups(int C);
int launch_gn_finalize(const Act& a0, const Act* a1, const float* gamma, const float* beta, float* ss, float* ms, cudaStream_t st);

// ---- FIR resampling, taps [1,3,3,1] (up_or_down_sampling.py:195-257) ------------------------------------
// h_out = FIR(SiLU(x*scale+shift)) (if ss) and x_out = FIR(x); mode: 1 = up x2, 2 = down x2.  C % 4 == 0.
int launch_resample(const Act& x, const float* ss, float* h_out, float* x_out, int mode, cudaStream_t st);
// small-channel scalar variant (input pyramid): C input channels -> CO >= C output channels, the extra ones zero
int launch_fir_down_small(const float* x, float* out, int B, int H, int W, int C, int CO, int Creal, cudaStream_t st);

// ---- input / pyramid / head ---------------------------------------------------------------------------
// complex planes -> NHWC real [B][F][T][2nc] with 2x-1 (ncsnpp.py:290-296,321-323); plane ch of sample b
// starts at p[ch] + b * bstride[ch] complex elements
struct UnpackSrc { const float* p[3]; long long bstride[3]; };
// out has CP >= 2nc channels per pixel; channels beyond 2nc are written as zeros (tensor-core K padding)
int launch_unpack_input(UnpackSrc src, float* out, int B, int nc, int CP, int HW, cudaStream_t st);
int launch_fill(float* p, float v, int n, cudaStream_t st);
// out = h + bias + W(pyr): Combine 'sum' with 1x1 conv Cin->C (layerspp.py:52-57); w [C][Cin]
int launch_combine(const float* h, const float* pyr, int pyr_stride, const float* w, const float* bias, float* out, int B, int HW,
                   int C, int Cin, cudaStream_t st);
// [Cout][Cin][taps] -> [Cout][CinPad][taps] with zero channels appended
int launch_pad_cin(const float* w, float* out, int Cout, int Cin, int CinPad, int taps, cudaStream_t st);
// combined 1x1 matrix [Cout][C2 + CP]: columns [0,C2) = w2 ([Cout][C2]), columns [C2, C2+Cc) = scale * wc ([Cout][Cc]), rest zero
int launch_concat_1x1(const float* w2, int C2, const float* wc, int Cc, int CP, float scale, float* out, int Cout, cudaStream_t st);
// progressive output branch: conv3x3_{C->Co}(SiLU(GN(h))) + bias (+ FIR_up(prev)) (ncsnpp.py:389-410);
// w packed [9][C][8] (Co <= 8, zero padded)
int launch_conv_out_small(const Act& h, const float* ss, const float* w, const float* bias, const float* prev,
                          float* out, int Co, cudaStream_t st);
// Tensor-core route of the same branch: Z[b,y,x,tap*Co+co] = sum_c act(h)[b,y,x,c] * W[co][c][tap] is a 1x1 GEMM
// (N = 9*Co padded to a multiple of 16) run by the tcgen05 kernel; this gathers the 9 shifted partial sums:
// out = bias + sum_tap Z[y+dy-1, x+dx-1, tap*Co+co] (+ FIR_up(prev)).
int launch_pack_convout_gemm(const float* w_oihw, float* out_io, int Co, int Cin, int Npad, cudaStream_t st);
int launch_tap_gather(const float* Z, int Npad, const float* bias, const float* prev, float* out, int B, int H, int W, int Co,
                      cudaStream_t st);
// d = conv1x1_{Ci->2}(pyr / t) packed to complex (ncsnpp.py:440-449); w [2][Ci]
int launch_head(const float* pyr, const float* t, const float* w, const float* bias, float* out_c, int B, int HW,
                int Ci, cudaStream_t st);

// ---- attention (layerspp.py:75-91) ------------------------------------------------------------------
// o[b][q][:] = softmax_k(q.k / sqrt(C)) v ; q,k,v,o: [B][N][C]
int launch_attention_simt(const float* q, const float* k, const float* v, float* o, int B, int N, int C, cudaStream_t st);
// tcgen05 version: vT is channel-major [B][C][N]; needs N % 128 == 0, C % 16 == 0, C <= 256
bool attention_tc_eligible(int N, int C);
size_t attention_tc_scratch_bytes(int B, int N, int C);
int launch_attention_tc(const float* q, const float* k, const float* vT, float* o, int B, int N, int C, void* scratch, cudaStream_t st);

// ---- sampler elementwise updates (sdes.py:233-237, predictors.py:41-69, correctors.py:37-93) ----------
struct UpdateCoef { float a, b, c, e; };   // x_mean = a x + b y + c d ; x = x_mean + e z
// x, y, d: complex [B][n]; z from `noise` (complex [B][n]) or Philox(seed, utt_ids[b], draw)
// coef_dev (optional): device UpdateCoef overriding `coef` (langevin); seed_dev (optional): device seed overriding `seed`
// (CUDA-graph replays of a captured trajectory read seed and utterance ids from device memory)
int launch_sampler_update(const float* x, const float* y, const float* d, float* x_mean, float* x_new,
                          UpdateCoef coef, const UpdateCoef* coef_dev, const float* noise, uint64_t seed, const uint64_t* seed_dev,
                          const uint64_t* utt_ids_dev, uint32_t draw, int B, int64_t n, cudaStream_t st);
int launch_philox_fill(uint64_t seed, uint64_t utt_id, uint32_t draw, int64_t n, float* out, cudaStream_t st);
// langevin step size from batch-mean norms (correctors.py:52-58): coef = {1, 0, -step, sqrt(2 step)}
int launch_langevin_coef(const float* d, const float* noise, uint64_t seed, const uint64_t* seed_dev, const uint64_t* utt_ids_dev,
                         uint32_t draw, float snr, int B, int64_t n, double* scratch, UpdateCoef* coef_out, cudaStream_t st);
// drift = theta (y - x) + 0.5 g2 d   (PF-ODE drift with score = -d)
int launch_ode_drift(const float* x, const float* y, const float* d, float* out, float theta, float half_g2,
                     int64_t n_total, cudaStream_t st);

// ---- STFT / iSTFT (data_module.py:182-223) -----------------------------------------------------------
// lens (optional, device int[B]): per-utterance true lengths inside the L-sample rows (ragged batch of one frame bucket)
int launch_wave_norm(const float* wav, int B, int L, const int* lens, float* norm, cudaStream_t st);
int launch_stft(const float* wav, const float* norm, int B, int L, const int* lens, float* spec, int T_pad, int n_fft, int hop,
                int sqrt_window, float factor, float expo, cudaStream_t st);
// frames_buf: [B][T_pad][n_fft] scratch
int launch_istft(const float* spec, const float* norm, int B, int T_pad, float* frames_buf, float* wav, int L, const int* lens,
                 int n_fft, int hop, int sqrt_window, float factor, float expo, cudaStream_t st);

// NHWC -> NCHW (debug taps)
// SI-SDR in dB per utterance (util/other.py:82-94); eps = 0 (numpy form) or 1e-10 (torch form)
int launch_si_sdr(const float* s, const float* s_hat, int B, int L, const int* lens, double eps, float* out, cudaStream_t st);
int launch_nhwc_to_nchw(const float* in, float* out, int B, int HW, int C, cudaStream_t st);

}  // namespace storm
