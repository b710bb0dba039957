/*
 * storm_b200.h - C ABI of the B200-native reverse-diffusion enhancement path.
 *
 * Drop-in boundary for ONE hot path of sp-uhh/storm: enhancement.py -> ScoreModel.enhance() ->
 * predictor-corrector sampler -> NCSN++ score network over complex STFT spectrograms.
 * The reference is pure Python; its own native boundary is a torch JIT extension
 * (sgmse/backbones/ncsnpp_utils/op/upfirdn2d.cpp:4-22).  A maintainer binds this library with ctypes
 * (see INTEGRATION.md); storm_b200/_native.py is that binding.
 *
 * Conventions
 *   - plain C types only; every device pointer is a raw CUDA device address (e.g. torch .data_ptr()).
 *   - caller allocates inputs and outputs; the library owns only the weight arena and a workspace pool
 *     that grows on the first call with a new (B, F, T) shape (storm_reserve() pre-builds it).
 *   - every entry point is stream-ordered on the cudaStream_t passed as `void* stream` (0 = default)
 *     and performs no host synchronisation unless documented.
 *   - return value: 0 = ok, negative = error (STORM_E_*); storm_last_error() gives the message.
 *     The Python shim maps STORM_E_INVALID -> ValueError, everything else -> RuntimeError, mirroring
 *     the reference's TORCH_CHECK / registry ValueError behaviour (op/upfirdn2d.cpp:8-10,
 *     util/registry.py:27-30).
 *   - a handle is bound to the CUDA device current at storm_create() and is not thread-safe;
 *     distinct handles are independent (one process per GPU, one handle per process).
 *   - complex spectrograms are interleaved (re, im) float pairs, laid out [B][F][T] with T fastest,
 *     exactly torch.complex64 [B,1,F,T] contiguous (sgmse/model.py:285-286).
 */
#ifndef STORM_B200_H
#define STORM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STORM_OK 0
#define STORM_E_INVALID (-1)   /* bad argument / unsupported configuration  -> ValueError   */
#define STORM_E_CUDA (-2)      /* CUDA runtime error                         -> RuntimeError */
#define STORM_E_STATE (-3)     /* call sequence error (e.g. weights missing) -> RuntimeError */
#define STORM_E_NOMEM (-4)

#define STORM_MAX_LEVELS 8

typedef struct StormHandle StormHandle;

/* Topology of the NCSN++ backbone = the ctor options of NCSNpp that this path builds
 * (sgmse/backbones/ncsnpp.py:40-65; registered variants :36,460-513). */
typedef struct StormNetConfig {
  int32_t nf;                                  /* base width (128; 96 for ncsnpp12M/6M)                  */
  int32_t num_levels;                          /* len(ch_mult)                                           */
  int32_t ch_mult[STORM_MAX_LEVELS];
  int32_t num_res_blocks;
  int32_t num_attn_resolutions;
  int32_t attn_resolutions[STORM_MAX_LEVELS];  /* compared with the F extent like ncsnpp.py:338,385      */
  int32_t image_size;                          /* 256: only places attention blocks at build (ncsnpp.py:78,194) */
  int32_t input_channels;                      /* real channels: 4 = [x.re,x.im,y.re,y.im]; 6 for StoRM  */
  int32_t discriminative;                      /* 1: 2-channel denoiser, no time conditioning (ncsnpp.py:80-86) */
  int32_t engine;                              /* 0 = auto (tcgen05 where eligible), 1 = SIMT fp32 only  */
} StormNetConfig;

/* Sampler options = kwargs of ScoreModel.enhance / get_pc_sampler
 * (sgmse/model.py:273-277, sgmse/sampling/__init__.py:27-32) + OUVESDE ctor (sgmse/sdes.py:167-188). */
enum { STORM_PRED_REVERSE_DIFFUSION = 0, STORM_PRED_EULER_MARUYAMA = 1, STORM_PRED_NONE = 2 };
enum { STORM_CORR_ALD = 0, STORM_CORR_LANGEVIN = 1, STORM_CORR_NONE = 2 };

typedef struct StormSamplerConfig {
  int32_t predictor;
  int32_t corrector;
  int32_t N;                 /* reverse steps                                              */
  int32_t corrector_steps;
  float snr;
  float t_eps;               /* ScoreModel.t_eps, 0.03                                     */
  float theta, sigma_min, sigma_max;   /* OUVE: 1.5, 0.05, 0.5                             */
  int32_t denoise;           /* 1: return x_mean of the last step (sampling/__init__.py:64) */
  int32_t probability_flow;  /* 1: RSDE.discretize PF branch (sdes.py:153-156), fixed-step ODE */
} StormSamplerConfig;

/* STFT front/back end = SpecsDataModule options (sgmse/data_module.py:141-145,195-204). */
typedef struct StormSpecConfig {
  int32_t n_fft;             /* 510 */
  int32_t hop_length;        /* 128 */
  int32_t sqrt_window;       /* 0 = hann (periodic), 1 = sqrthann                          */
  float spec_factor;         /* 0.15 */
  float spec_abs_exponent;   /* 0.5  */
} StormSpecConfig;

const char* storm_last_error(void);
const char* storm_version(void);

/* ---- lifecycle ---------------------------------------------------------------------------------- */
int storm_create(const StormNetConfig* cfg, StormHandle** out);
int storm_destroy(StormHandle* h);

/* Parameter table: names/shapes/order of the reference state_dict (output_layer.*, all_modules.{i}.*),
 * replaces nn.Module.state_dict()/load_state_dict (sgmse/backbones/ncsnpp.py:108,273). */
int64_t storm_param_floats(const StormHandle* h);        /* total fp32 elements of the blob          */
int32_t storm_param_entries(const StormHandle* h);
int storm_param_info(const StormHandle* h, int32_t index, char* name, int32_t name_cap,
                     int64_t shape[4], int32_t* ndim, int64_t* offset_floats);

/* Upload the flat fp32 blob (tensor i occupies [offset_i, offset_i + numel_i), C-contiguous; offsets are
 * 64-float aligned, gaps are ignored; storm_param_floats() is the padded total) and repack it into
 * the kernels' layouts.  `blob` may be a host or a device pointer (is_device).  Replaces
 * load_from_checkpoint + .cuda() (enhancement.py:56-61).  storm_weights_device() exposes the device
 * copy of the raw blob so that rank 0 can ncclBroadcast it. */
int storm_load_weights(StormHandle* h, const float* blob, int64_t n_floats, int32_t is_device, void* stream);
int storm_weights_device(StormHandle* h, float** dev_blob, int64_t* n_floats);
int storm_repack_weights(StormHandle* h, void* stream);  /* after writing into storm_weights_device() */

int storm_reserve(StormHandle* h, int32_t B, int32_t F, int32_t T);
int64_t storm_workspace_bytes(const StormHandle* h);

/* ---- score network -------------------------------------------------------------------------------
 * NCSNpp.forward (sgmse/backbones/ncsnpp.py:281-450).
 *   in  : complex [B][input_channels/2][F][T]  (x then y [then y_denoised])
 *   t   : float   [B]  device (ignored when discriminative)
 *   out : complex [B][F][T]
 * F, T multiples of 2^(num_levels-1). */
int storm_net_forward(StormHandle* h, const float* in, const float* t, float* out,
                      int32_t B, int32_t F, int32_t T, void* stream);

/* Debug taps for layer-level parity tests: copy of activation `name` ("mod{i}", "conv_in", "pyr{lvl}")
 * of the LAST forward, converted to NCHW fp32.  Enable with storm_set_debug_taps(h, 1) first. */
int storm_set_debug_taps(StormHandle* h, int32_t enable);
int storm_get_tap(StormHandle* h, const char* name, float* out_nchw, int64_t cap_floats,
                  int64_t shape[4], void* stream);

/* ---- sampler --------------------------------------------------------------------------------------
 * pc_sampler (sgmse/sampling/__init__.py:54-66) with score = -dnn(cat[x, cond...], t) (sgmse/model.py:127-136).
 *   Y, X : complex [B][F][T].  Y is the SDE's steady-state mean and the centre of the prior (sdes.py:233-237).
 *   cond0, cond1: the conditioning channels stacked after x in the network input.  cond0 == NULL means Y
 *          (ScoreModel: cat[x, y]).  StoRM (model.py:742-756, condition="both") passes Y = Y_denoised,
 *          cond0 = Y_noisy, cond1 = Y_denoised and a 6-channel score network.
 *   noise: NULL -> counter-based Philox keyed by (seed, utt_ids[b], draw index, element) so results do
 *          not depend on batch composition or sharding; else injected complex draws
 *          [n_draws][B][F][T] in the reference's order (prior, then per step corrector..., predictor).
 *   utt_ids: host array [B] or NULL (-> 0..B-1).   nfe_out: host int or NULL. */
int storm_pc_sample(StormHandle* h, const float* Y, const float* cond0, const float* cond1, float* X, int32_t B, int32_t F,
                    int32_t T, const StormSamplerConfig* sc, uint64_t seed, const uint64_t* utt_ids,
                    const float* noise, int32_t* nfe_out, void* stream);
int32_t storm_pc_num_draws(const StormSamplerConfig* sc);
/* The Philox draw the sampler would use: complex normal, var 1/2 per component (torch.randn_like on
 * complex64, sgmse/sdes.py:237). out: complex [n_complex]. */
int storm_philox_normal(uint64_t seed, uint64_t utt_id, uint32_t draw, int64_t n_complex, float* out, void* stream);
/* Probability-flow drift  theta (y - x) - 1/2 g(t)^2 score  for an external ODE solver
 * (sgmse/sampling/__init__.py:102-104, sgmse/sdes.py:124-141). */
int storm_ode_drift(StormHandle* h, const float* X, const float* Y, float t, float* drift,
                    int32_t B, int32_t F, int32_t T, const StormSamplerConfig* sc, void* stream);

/* ---- STFT front / back end -------------------------------------------------------------------------
 * stft + spec_fwd + pad_spec (sgmse/data_module.py:182-186,217-219, sgmse/util/other.py:102-109) and
 * spec_back + istft (data_module.py:188-193,221-223).
 *   wav  : float [B][L] device.  norm: float [B] device = max|wav| per utterance (sgmse/model.py:283-284);
 *          storm_stft divides by it (pass NULL for no normalisation), storm_istft multiplies.
 *   spec : complex [B][n_fft/2+1][T_pad]; frames >= 1 + L/hop are zero (pad_spec). */
int32_t storm_num_frames(int32_t L, const StormSpecConfig* sp);
int32_t storm_padded_frames(int32_t frames);
int storm_wave_norm(const float* wav, int32_t B, int32_t L, float* norm, void* stream);
int storm_stft(const float* wav, const float* norm, int32_t B, int32_t L, float* spec, int32_t T_pad,
               const StormSpecConfig* sp, void* stream);
int storm_istft(StormHandle* h, const float* spec, const float* norm, int32_t B, int32_t T_pad, float* wav, int32_t L,
                const StormSpecConfig* sp, void* stream);

/* Ragged batches.  The reference enhances one file at a time (enhancement.py:66-72), so every utterance has its own
 * length: its own reflect padding, frame count 1 + L_b/hop (data_module.py:217-219), pad_spec amount and istft
 * length (model.py:301).  Utterances whose padded frame counts agree (same storm_padded_frames(storm_num_frames(L_b)))
 * can share one [B][F][T_pad] batch without changing any result, because GroupNorm, attention and the sampler are
 * per sample.  `lengths`: int32 [B] on the DEVICE, L_b <= L (the row stride), all in the frame bucket of L; NULL =
 * every row is L samples.  Samples of a row beyond L_b are ignored on input and written as zeros on output. */
int storm_wave_norm_ragged(const float* wav, int32_t B, int32_t L, const int32_t* lengths, float* norm, void* stream);
int storm_stft_ragged(const float* wav, const float* norm, int32_t B, int32_t L, const int32_t* lengths, float* spec,
                      int32_t T_pad, const StormSpecConfig* sp, void* stream);
int storm_istft_ragged(StormHandle* h, const float* spec, const float* norm, int32_t B, int32_t T_pad, float* wav, int32_t L,
                       const int32_t* lengths, const StormSpecConfig* sp, void* stream);

/* ---- one-call enhancement ---------------------------------------------------------------------------
 * ScoreModel.enhance for a batch (sgmse/model.py:273-310): normalise, STFT, transform, pad, PC-sample,
 * inverse transform, iSTFT, de-normalise.  wav_in / wav_out: float [B][L] device. */
int storm_enhance(StormHandle* h, const float* wav_in, float* wav_out, int32_t B, int32_t L,
                  const StormSamplerConfig* sc, const StormSpecConfig* sp, uint64_t seed,
                  const uint64_t* utt_ids, int32_t* nfe_out, void* stream);
/* same for a ragged batch of one frame bucket (see above); L = row stride = the longest utterance of the batch */
int storm_enhance_ragged(StormHandle* h, const float* wav_in, float* wav_out, int32_t B, int32_t L, const int32_t* lengths,
                         const StormSamplerConfig* sc, const StormSpecConfig* sp, uint64_t seed, const uint64_t* utt_ids,
                         int32_t* nfe_out, void* stream);

/* ---- evaluation metric on the device ------------------------------------------------------------------
 * si_sdr / si_sdr_torch (sgmse/util/other.py:82-94) for a batch: s (reference) and s_hat (estimate) float [B][L]
 * device, optional int32 lengths [B] device (ragged rows), out_db float [B] device.  eps = 0 gives the numpy form
 * (other.py:82-86), eps = 1e-10 the torch form (other.py:88-94).  fp64 accumulation, deterministic. */
int storm_si_sdr(const float* s, const float* s_hat, int32_t B, int32_t L, const int32_t* lengths, double eps,
                 float* out_db, void* stream);

/* Kernel-launch counter (bench.py's gpu_launches). */
int64_t storm_launch_count(const StormHandle* h);
int storm_reset_launch_count(StormHandle* h);

/* CUDA-graph replay of whole sampler trajectories (default on; STORM_B200_GRAPH=0 in the environment disables it for the process):
 * the first storm_pc_sample / storm_enhance call of a (B, F, T, sampler configuration) runs eagerly and sizes the workspace,
 * the second captures the N-step loop (sgmse/sampling/__init__.py:59) into a graph, later calls replay it - results are
 * bit-identical to the eager path.  Injected noise, debug taps and profiling always run eagerly.  The handle keeps up to 8
 * graphs; they are dropped when the weights are re-packed.  storm_graph_replays: replays so far (tests / bench). */
int storm_set_graph(StormHandle* h, int32_t enable);
int64_t storm_graph_replays(const StormHandle* h);

/* Live roofline measurement of the dominant kernel class (the implicit-GEMM convolutions): while enabled,
 * every conv launch is bracketed by CUDA events on its stream.  storm_get_profile synchronises the device
 * and returns, per engine (0 = tcgen05, 1 = SIMT): launches, summed device milliseconds and summed
 * algorithmic FLOPs (2 * B*H*W * Cout * sum_seg taps*Cin) since the last storm_set_profile(h, 1). */
int storm_set_profile(StormHandle* h, int32_t enable);
int storm_get_profile(StormHandle* h, int64_t launches[2], double ms[2], double flops[2]);

/* Test hook: run ONE fused convolution through the chosen kernel (engine 0 = tcgen05, 1 = SIMT):
 *   out = out_scale * ( conv_taps( act(cat[x0, x1]) ; w ) [+ conv_1x1(x2 ; w2)] + bias [+ res] )
 * x*: NHWC fp32 device tensors [B][H][W][C*]; ss: [B][C0+C1][2] scale/shift or NULL; w: OIHW fp32
 * [Cout][C0+C1][taps] device; w2: [Cout][C2] or NULL; out NHWC.  Returns STORM_E_INVALID when the shape
 * is not eligible for the requested kernel. */
int storm_debug_conv(int32_t engine, const float* x0, const float* x1, int32_t C0, int32_t C1, const float* ss, int32_t silu,
                     int32_t taps, const float* w, const float* x2, int32_t C2, const float* w2, int32_t Cout,
                     const float* bias, const float* res, float out_scale, float* out, int32_t B, int32_t H, int32_t W,
                     void* stream);

/* Same, launched twice; the second launch is timed with CUDA events (ms_out) and, when dbg_dev != NULL
 * (device int64[16]), CTA 0 of the tcgen05 kernel records cycle counters: [0] MMA issue loop, [1] waiting for
 * activation stages, [2] waiting for weight stages, [3] until the last MMA retired, [4] producer idle before
 * the epilogue, [5] epilogue; producer thread 0: [6] issuing the global loads (back-pressure), [8] waiting for a
 * free activation stage, [9] / [10] load latency + conversion + stores of segment 0 / 1. */
int storm_debug_conv_timed(int32_t engine, const float* x0, const float* x1, int32_t C0, int32_t C1, const float* ss, int32_t silu,
                           int32_t taps, const float* w, const float* x2, int32_t C2, const float* w2, int32_t Cout,
                           const float* bias, const float* res, float out_scale, float* out, int32_t B, int32_t H, int32_t W,
                           long long* dbg_dev, float* ms_out, void* stream);

/* General form.  res_up: residual given at HALF resolution ([B][H/2][W/2][Cout], tcgen05 engine only); the epilogue adds
 * its x2 FIR upsampling (taps [1,3,3,1], up_or_down_sampling.py:195-217) - how the engine runs the 1x1 skip of up blocks.
 * max_prec: 1 = let the tcgen05 path use the two-pass "f16 + 2 x fp8" operand split when every segment is a multiple of
 * 32 channels, 0 = force the bf16 x 3 split.  ms_out / dbg_dev as in storm_debug_conv_timed (NULL = single launch). */
int storm_debug_conv_ex(int32_t engine, const float* x0, const float* x1, int32_t C0, int32_t C1, const float* ss, int32_t silu,
                        int32_t taps, const float* w, const float* x2, int32_t C2, const float* w2, int32_t Cout,
                        const float* bias, const float* res, const float* res_up, float out_scale, float* out, int32_t B,
                        int32_t H, int32_t W, int32_t max_prec, long long* dbg_dev, float* ms_out, void* stream);

/* Test hook: the FIR resampling pass of a BigGAN res block (layerspp.py:245-258, up_or_down_sampling.py:195-257, taps
 * [1,3,3,1]): x [B][H][W][C] fp32 NHWC; mode 1 = x2 up, 2 = x2 down; h_out = FIR(SiLU(x * scale + shift)) with
 * ss = [B][C] (scale, shift) pairs (NULL: FIR of x itself), x_out = FIR(x); either output may be NULL. */
int storm_debug_resample(const float* x, const float* ss, float* h_out, float* x_out, int32_t B, int32_t H, int32_t W, int32_t C,
                         int32_t mode, void* stream);

/* Test hook: single-head attention o = softmax(q k^T / sqrt(C)) v on [B][N][C] fp32 device tensors
 * (AttnBlockpp core, layerspp.py:82-86).  engine 0 = tcgen05 kernel (v must be given channel-major [B][C][N]),
 * 1 = SIMT kernel (v token-major [B][N][C]). */
int storm_debug_attention(int32_t engine, const float* q, const float* k, const float* v, float* o, int32_t B, int32_t N, int32_t C,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STORM_B200_H */
