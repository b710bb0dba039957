// Small kernels of the reverse-diffusion path: weight packing, time embedding, GroupNorm statistics,
// FIR resampling, input/pyramid/head, attention (SIMT), sampler updates with Philox noise, STFT/iSTFT.
// Every kernel here is HBM- or latency-bound; the FLOPs live in conv_simt.cu / conv_tc.cu.
#include "kernels.cuh"

#include <math.h>

namespace storm {

namespace {
constexpr float PI_F = 3.14159274101257324f;   // float(np.pi), as torch's fp32 scalar multiply sees it

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
}  // namespace

// =====================================================================================================
// weights
// =====================================================================================================
__global__ void pack_conv_simt_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int taps) {
  pdl_enter();
  size_t n = (size_t)Cout * Cin * taps;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int co = i % Cout;
    int ci = (i / Cout) % Cin;
    int tap = i / ((size_t)Cout * Cin);
    out[i] = w[((size_t)co * Cin + ci) * taps + tap];
  }
}
int launch_pack_conv_simt(const float* w, float* out, int Cout, int Cin, int taps, cudaStream_t st) {
  size_t n = (size_t)Cout * Cin * taps;
  storm_launch(pack_conv_simt_kernel, (unsigned)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256), 256, 0, st, w, out, Cout, Cin, taps);
  STORM_LAUNCH_CHECK();
  return 0;
}

__global__ void transpose_into_kernel(const float* __restrict__ w, float* __restrict__ out, int rows, int cols,
                                      int out_stride, int out_off) {
  pdl_enter();
  size_t n = (size_t)rows * cols;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int r = i % rows;
    int c = i / rows;
    out[(size_t)c * out_stride + out_off + r] = w[(size_t)r * cols + c];
  }
}
int launch_transpose_into(const float* w, float* out, int rows, int cols, int out_stride, int out_off, cudaStream_t st) {
  size_t n = (size_t)rows * cols;
  storm_launch(transpose_into_kernel, (unsigned)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256), 256, 0, st, w, out, rows, cols, out_stride, out_off);
  STORM_LAUNCH_CHECK();
  return 0;
}

__global__ void pad_cin_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int CinPad, int taps) {
  pdl_enter();
  const size_t n = (size_t)Cout * CinPad * taps;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int tap = i % taps;
    const int ci = (i / taps) % CinPad;
    const int co = i / ((size_t)taps * CinPad);
    out[i] = ci < Cin ? w[((size_t)co * Cin + ci) * taps + tap] : 0.f;
  }
}
int launch_pad_cin(const float* w, float* out, int Cout, int Cin, int CinPad, int taps, cudaStream_t st) {
  size_t n = (size_t)Cout * CinPad * taps;
  storm_launch(pad_cin_kernel, (unsigned)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256), 256, 0, st, w, out, Cout, Cin, CinPad, taps);
  STORM_LAUNCH_CHECK();
  return 0;
}
__global__ void concat_1x1_kernel(const float* __restrict__ w2, int C2, const float* __restrict__ wc, int Cc, int CP, float scale,
                                  float* __restrict__ out, int Cout) {
  pdl_enter();
  const int K = C2 + CP;
  const size_t n = (size_t)Cout * K;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int k = i % K, co = i / K;
    float v = 0.f;
    if (k < C2) v = w2[(size_t)co * C2 + k];
    else if (k - C2 < Cc) v = scale * wc[(size_t)co * Cc + (k - C2)];
    out[i] = v;
  }
}
int launch_concat_1x1(const float* w2, int C2, const float* wc, int Cc, int CP, float scale, float* out, int Cout, cudaStream_t st) {
  size_t n = (size_t)Cout * (C2 + CP);
  storm_launch(concat_1x1_kernel, (unsigned)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256), 256, 0, st, w2, C2, wc, Cc, CP, scale, out, Cout);
  STORM_LAUNCH_CHECK();
  return 0;
}

__global__ void add_vec_kernel(const float* a, const float* b, float* out, int n, float sb) {
  pdl_enter();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (a ? a[i] : 0.f) + (b ? sb * b[i] : 0.f);
}
int launch_add_vec(const float* a, const float* b, float* out, int n, float scale_b, cudaStream_t st) {
  storm_launch(add_vec_kernel, ceil_div(n, 256), 256, 0, st, a, b, out, n, scale_b);
  STORM_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================
// time embedding
// =====================================================================================================
__global__ void temb_mlp_kernel(const float* __restrict__ t, const float* __restrict__ Wf, const float* __restrict__ W1t,
                                const float* __restrict__ b1, const float* __restrict__ W2t, const float* __restrict__ b2,
                                float* __restrict__ act, int nf) {
  pdl_enter();
  extern __shared__ float sm[];
  const int E = 2 * nf, D = 4 * nf;
  float* e = sm;        // [E]
  float* h = sm + E;    // [D]
  const int b = blockIdx.x;
  // log in double then rounded: the correctly rounded fp32 log (torch.log on CPU is <= 1 ulp)
  const float lt = (float)log((double)t[b]);
  for (int i = threadIdx.x; i < nf; i += blockDim.x) {
    // x[:, None] * W[None, :] * 2 * np.pi evaluated left to right in fp32 (layerspp.py:40)
    float p = __fmul_rn(__fmul_rn(__fmul_rn(lt, Wf[i]), 2.0f), PI_F);
    e[i] = (float)sin((double)p);
    e[nf + i] = (float)cos((double)p);
  }
  __syncthreads();
  // weights are streamed from HBM once per call: keep 8 independent loads in flight per thread
  for (int o = threadIdx.x; o < D; o += blockDim.x) {
    float s = b1[o];
    for (int i = 0; i < E; i += 8) {
      float wv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) wv[j] = __ldg(W1t + (size_t)(i + j) * D + o);
#pragma unroll
      for (int j = 0; j < 8; ++j) s = fmaf(e[i + j], wv[j], s);
    }
    h[o] = silu_f(s);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < D; o += blockDim.x) {
    float s = b2[o];
    for (int i = 0; i < D; i += 8) {
      float wv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) wv[j] = __ldg(W2t + (size_t)(i + j) * D + o);
#pragma unroll
      for (int j = 0; j < 8; ++j) s = fmaf(h[i + j], wv[j], s);
    }
    act[(size_t)b * D + o] = silu_f(s);   // every consumer applies act(temb) first (layerspp.py:263)
  }
}
int launch_temb_mlp(const float* t, const float* Wf, const float* W1t, const float* b1, const float* W2t, const float* b2,
                    float* act, int B, int nf, cudaStream_t st) {
  if (nf % 4 != 0) return fail(-1, "temb_mlp: nf must be a multiple of 4");
  storm_launch(temb_mlp_kernel, B, 512, 6 * nf * sizeof(float), st, t, Wf, W1t, b1, W2t, b2, act, nf);
  STORM_LAUNCH_CHECK();
  return 0;
}

// 64 outputs x KS k-slices per CTA: the 64 x in_dim weight panel is read once, coalesced along the outputs, with
// KS x 8 independent loads in flight per output instead of one serial chain over in_dim (latency-bound before);
// slice partials are summed in a fixed order through shared memory (deterministic).
constexpr int DENSE_KS = 8;   // at most; the launcher picks the largest power of two that divides in_dim / 8
__global__ void __launch_bounds__(64 * DENSE_KS) dense_all_kernel(const float* __restrict__ act, const float* __restrict__ Wt,
                                                                  const float* __restrict__ bias, float* __restrict__ out, int B,
                                                                  int in_dim, int out_dim) {
  pdl_enter();
  extern __shared__ float s_act[];   // [8][in_dim] then [DENSE_KS][64][8] partials
  float* s_part = s_act + 8 * in_dim;
  const int tx = threadIdx.x & 63, ks = threadIdx.x >> 6;
  const int o = blockIdx.x * 64 + tx;
  const int KS = blockDim.x >> 6;
  const int kper = in_dim / KS;                       // multiple of 8 (chosen by the launcher)
  for (int b0 = 0; b0 < B; b0 += 8) {
    const int nb = min(8, B - b0);
    __syncthreads();
    for (int e = threadIdx.x; e < 8 * in_dim; e += blockDim.x)
      s_act[e] = (e / in_dim < nb) ? act[(size_t)b0 * in_dim + e] : 0.f;
    __syncthreads();
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (o < out_dim) {
      for (int i = ks * kper; i < (ks + 1) * kper; i += 8) {
        float wv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) wv[k] = __ldg(Wt + (size_t)(i + k) * out_dim + o);
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(wv[k], s_act[j * in_dim + i + k], acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s_part[(ks * 64 + tx) * 8 + j] = acc[j];
    __syncthreads();
    if (ks == 0 && o < out_dim) {
      const float bv = bias[o];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = bv;
        for (int q = 0; q < KS; ++q) v += s_part[(q * 64 + tx) * 8 + j];
        if (j < nb) out[(size_t)(b0 + j) * out_dim + o] = v;
      }
    }
  }
}
int launch_dense_all(const float* act, const float* Wt, const float* bias, float* out, int B, int in_dim, int out_dim,
                     cudaStream_t st) {
  if (in_dim % 8 != 0 || in_dim > 1024) return fail(-1, "dense_all: in_dim must be a multiple of 8 and <= 1024");
  int ks = DENSE_KS;
  while (ks > 1 && in_dim % (8 * ks) != 0) ks >>= 1;
  const size_t smem = (size_t)(8 * in_dim + ks * 64 * 8) * sizeof(float);
  storm_launch(dense_all_kernel, ceil_div(out_dim, 64), 64 * ks, smem, st, act, Wt, bias, out, B, in_dim, out_dim);
  STORM_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================
// GroupNorm statistics
// =====================================================================================================
int gn_num_chunks(int HW) {
  int n = HW / 256;
  if (n < 1) n = 1;
  if (n > 64) n = 64;
  return n;
}

__global__ void gn_partial_kernel(const float* __restrict__ x, double* __restrict__ part, int HW, int C, int nchunk) {
  pdl_enter();
  extern __shared__ double smd[];   // [rows][C4][8]
  const int chunk = blockIdx.x, b = blockIdx.y;
  const int C4 = C >> 2;
  const int rows = blockDim.x / C4;
  const int c4 = threadIdx.x % C4, r = threadIdx.x / C4;
  const int per = (HW + nchunk - 1) / nchunk;
  const int p0 = chunk * per, p1 = min(HW, p0 + per);
  double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (r < rows) {
    const float4* xp = reinterpret_cast<const float4*>(x + (size_t)b * HW * C) + c4;
    for (int p = p0 + r; p < p1; p += rows) {
      float4 v = __ldg(xp + (size_t)p * C4);
      s[0] += v.x; q[0] += (double)v.x * v.x;
      s[1] += v.y; q[1] += (double)v.y * v.y;
      s[2] += v.z; q[2] += (double)v.z * v.z;
      s[3] += v.w; q[3] += (double)v.w * v.w;
    }
    double* d = smd + ((size_t)r * C4 + c4) * 8;
#pragma unroll
    for (int j = 0; j < 4; ++j) { d[j] = s[j]; d[4 + j] = q[j]; }
  }
  __syncthreads();
  if (threadIdx.x < C4) {
    double ts[4] = {0, 0, 0, 0}, tq[4] = {0, 0, 0, 0};
    for (int rr = 0; rr < rows; ++rr) {
      const double* d = smd + ((size_t)rr * C4 + threadIdx.x) * 8;
#pragma unroll
      for (int j = 0; j < 4; ++j) { ts[j] += d[j]; tq[j] += d[4 + j]; }
    }
    double* o = part + (((size_t)b * nchunk + chunk) * C + threadIdx.x * 4) * 2;
#pragma unroll
    for (int j = 0; j < 4; ++j) { o[2 * j] = ts[j]; o[2 * j + 1] = tq[j]; }
  }
}
int launch_gn_partial(const Act& a, cudaStream_t st) {
  if (a.C % 4 != 0 || a.C / 4 > 256) return fail(-1, "gn_partial: unsupported channel count");
  const int C4 = a.C / 4;
  int threads = 256;
  int rows = threads / C4;
  size_t smem = (size_t)rows * C4 * 8 * sizeof(double);
  STORM_SET_SMEM_ONCE((gn_partial_kernel), 64 * 1024);
  storm_launch(gn_partial_kernel, dim3(a.nchunk, a.B), threads, smem, st, a.p, a.part, a.H * a.W, a.C, a.nchunk);
  STORM_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(128) gn_finalize_kernel(const double* __restrict__ part0, int C0, int nchunk0,
                                                          const double* __restrict__ part1, int C1, int nchunk1,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float* __restrict__ ss, float* __restrict__ ms, double inv_n, int G) {
  pdl_enter();
  const int g = blockIdx.x, b = blockIdx.y;
  const int C = C0 + C1;
  const int cpg = C / G;
  double s = 0, q = 0;
  // items = (chunk, channel-in-group) pairs; loads are issued in batches of 8 before being consumed so that the
  // ~16 dependent round trips of the naive loop collapse into 2-3
  auto item_ptr = [&](int i) -> const double2* {
    const int cc = g * cpg + i % cpg;
    const int ch = i / cpg;
    if (cc < C0) return ch < nchunk0 ? reinterpret_cast<const double2*>(part0 + (((size_t)b * nchunk0 + ch) * C0 + cc) * 2) : nullptr;
    return ch < nchunk1 ? reinterpret_cast<const double2*>(part1 + (((size_t)b * nchunk1 + ch) * C1 + (cc - C0)) * 2) : nullptr;
  };
  const int maxchunk = max(nchunk0, nchunk1);
  const int nitems = cpg * maxchunk;
  for (int i0 = threadIdx.x; i0 < nitems; i0 += 8 * blockDim.x) {
    double2 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = i0 + j * blockDim.x;
      const double2* pp = i < nitems ? item_ptr(i) : nullptr;
      v[j] = pp ? __ldg(pp) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { s += v[j].x; q += v[j].y; }
  }
  __shared__ double rs[32], rq[32];
  s = warp_sum_d(s); q = warp_sum_d(q);
  if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rq[threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double S = 0, Q = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { S += rs[i]; Q += rq[i]; }
    // fp64 division and square root are software sequences of ~100 instructions each on this part (one thread, on the critical
    // path of a 4-us kernel that runs 45 times per evaluation): multiply by the host-computed 1/n, and refine a single-precision
    // rsqrt with two Newton steps in fp64 (relative error ~1e-7 -> 1e-14 -> below 1 ulp)
    const double mean = S * inv_n;
    const double msq = Q * inv_n;
    double var = msq - mean * mean;
    if (var < 0) var = 0;
    const double ve = var + 1e-6;
    double r = (double)rsqrtf((float)ve);
    r = r * (1.5 - 0.5 * ve * r * r);
    r = r * (1.5 - 0.5 * ve * r * r);
    rs[0] = mean;
    rq[0] = r;
    if (ms) ms[(size_t)b * G + g] = (float)msq;          // mean square of the RAW group: magnitude of the un-normalised tensor
  }
  __syncthreads();
  const double mean = rs[0], rstd = rq[0];
  for (int i = threadIdx.x; i < cpg; i += blockDim.x) {
    int cc = g * cpg + i;
    double sc = (double)gamma[cc] * rstd;
    float* o = ss + ((size_t)b * C + cc) * 2;
    o[0] = (float)sc;
    o[1] = (float)((double)beta[cc] - mean * sc);
  }
}
int gn_groups(int C) { return C / 4 < 32 ? C / 4 : 32; }     // nn.GroupNorm(num_groups=min(C // 4, 32)), layerspp.py:219
int launch_gn_finalize(const Act& a0, const Act* a1, const float* gamma, const float* beta, float* ss, float* ms, cudaStream_t st) {
  const int C = a0.C + (a1 ? a1->C : 0);
  const int G = gn_groups(C);
  if (G < 1 || C % G != 0) return fail(-1, "gn_finalize: channels not divisible into groups");
  storm_launch(gn_finalize_kernel, dim3(G, a0.B), 128, 0, st, a0.part, a0.C, a0.nchunk, a1 ? a1->part : nullptr, a1 ? a1->C : 0,
                                                     a1 ? a1->nchunk : 0, gamma, beta, ss, ms, 1.0 / ((double)(C / G) * a0.H * a0.W), G);
  STORM_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================
// FIR resampling
// =====================================================================================================
__device__ __forceinline__ float4 act4(float4 v, const float* ss4) {
  // ss4 -> 4 consecutive (scale, shift) pairs
  float4 a = __ldg(reinterpret_cast<const float4*>(ss4));
  float4 b = __ldg(reinterpret_cast<const float4*>(ss4) + 1);
  float4 r;
  r.x = silu_f(fmaf(v.x, a.x, a.y));
  r.y = silu_f(fmaf(v.y, a.z, a.w));
  r.z = silu_f(fmaf(v.z, b.x, b.y));
  r.w = silu_f(fmaf(v.w, b.z, b.w));
  return r;
}

__global__ void resample_kernel(const float* __restrict__ x, const float* __restrict__ ss, float* __restrict__ h_out,
                                float* __restrict__ x_out, int B, int H, int W, int C, int mode) {
  pdl_enter();
  const int C4 = C >> 2;
  const int OH = mode == 1 ? 2 * H : H / 2, OW = mode == 1 ? 2 * W : W / 2;
  const size_t total = (size_t)B * OH * OW * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = i % C4;
    size_t r = i / C4;
    const int ox = r % OW; r /= OW;
    const int oy = r % OH;
    const int b = r / OH;
    const float* ssp = ss ? ss + ((size_t)b * C + c4 * 4) * 2 : nullptr;
    float4 ax = make_float4(0, 0, 0, 0), ah = make_float4(0, 0, 0, 0);
    int ys[4], xs[4];
    float wy[4], wx[4];
    int ny, nx;
    if (mode == 1) {   // up: o[2i] = x[i-1]/4 + 3x[i]/4 ; o[2i+1] = 3x[i]/4 + x[i+1]/4
      ny = nx = 2;
      int iy = oy >> 1, ix = ox >> 1;
      if (oy & 1) { ys[0] = iy; wy[0] = 0.75f; ys[1] = iy + 1; wy[1] = 0.25f; }
      else        { ys[0] = iy - 1; wy[0] = 0.25f; ys[1] = iy; wy[1] = 0.75f; }
      if (ox & 1) { xs[0] = ix; wx[0] = 0.75f; xs[1] = ix + 1; wx[1] = 0.25f; }
      else        { xs[0] = ix - 1; wx[0] = 0.25f; xs[1] = ix; wx[1] = 0.75f; }
    } else {           // down: o[i] = (x[2i-1] + 3x[2i] + 3x[2i+1] + x[2i+2]) / 8
      ny = nx = 4;
      const float k[4] = {0.125f, 0.375f, 0.375f, 0.125f};
#pragma unroll
      for (int j = 0; j < 4; ++j) { ys[j] = 2 * oy - 1 + j; wy[j] = k[j]; xs[j] = 2 * ox - 1 + j; wx[j] = k[j]; }
    }
    for (int jy = 0; jy < ny; ++jy) {
      if ((unsigned)ys[jy] >= (unsigned)H) continue;
      for (int jx = 0; jx < nx; ++jx) {
        if ((unsigned)xs[jx] >= (unsigned)W) continue;
        const float wgt = wy[jy] * wx[jx];
        float4 v = __ldg(reinterpret_cast<const float4*>(x + (((size_t)b * H + ys[jy]) * W + xs[jx]) * C) + c4);
        ax.x = fmaf(wgt, v.x, ax.x); ax.y = fmaf(wgt, v.y, ax.y); ax.z = fmaf(wgt, v.z, ax.z); ax.w = fmaf(wgt, v.w, ax.w);
        if (h_out) {
          float4 a = ssp ? act4(v, ssp) : v;
          ah.x = fmaf(wgt, a.x, ah.x); ah.y = fmaf(wgt, a.y, ah.y); ah.z = fmaf(wgt, a.z, ah.z); ah.w = fmaf(wgt, a.w, ah.w);
        }
      }
    }
    const size_t o = (((size_t)b * OH + oy) * OW + ox) * C4 + c4;
    if (x_out) reinterpret_cast<float4*>(x_out)[o] = ax;
    if (h_out) reinterpret_cast<float4*>(h_out)[o] = ah;
  }
}
// Tiled variants: every input element is loaded and activated exactly once per CTA (shared memory), then
// both outputs (FIR of the activated tensor and FIR of the raw tensor) are produced from the staged tile.
// UP:   8x8 input pixels (+1 halo) x 32 channels -> 16x16 output pixels.   smem 2 x 10x10x32 floats
// DOWN: 8x8 output pixels <- 18x18 input pixels x 16 channels.             smem 2 x 18x18x20 floats (padded)
template <int MODE>
__global__ void __launch_bounds__(256) resample_tiled_kernel(const float* __restrict__ x, const float* __restrict__ ss,
                                                             float* __restrict__ h_out, float* __restrict__ x_out, int H, int W, int C) {
  pdl_enter();
  constexpr int CG = MODE == 1 ? 32 : 16;            // channels per CTA
  constexpr int C4 = CG / 4;
  constexpr int TI = MODE == 1 ? 10 : 18;            // staged input tile edge
  constexpr int PS = MODE == 1 ? 32 : 20;            // floats per staged pixel (padded for DOWN)
  extern __shared__ __align__(16) float rs_smem[];
  float* sAct = rs_smem;
  float* sRaw = rs_smem + TI * TI * PS;
  const int OH = MODE == 1 ? 2 * H : H / 2, OW = MODE == 1 ? 2 * W : W / 2;
  const int tiles_x = MODE == 1 ? (W + 7) / 8 : (OW + 7) / 8;
  const int tyi = blockIdx.x / tiles_x, txi = blockIdx.x % tiles_x;
  const int c0 = blockIdx.y * CG;
  const int b = blockIdx.z;
  // first staged input pixel
  const int iy0 = MODE == 1 ? tyi * 8 - 1 : tyi * 16 - 1;
  const int ix0 = MODE == 1 ? txi * 8 - 1 : txi * 16 - 1;
  // all loads of the staged tile are issued before the first is consumed: one DRAM round trip per CTA instead of one per
  // loop iteration (the kernel was latency-bound: 1.4 TB/s on the down variant)
  constexpr int NE = TI * TI * C4, NIT = (NE + 255) / 256;
  float4 rawv[NIT];
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int e = threadIdx.x + it * 256;
    const int c4 = e % C4, pix = e / C4;
    const int py = pix / TI, px = pix - py * TI;
    const int y = iy0 + py, xx = ix0 + px;
    rawv[it] = make_float4(0, 0, 0, 0);
    if (e < NE && (unsigned)y < (unsigned)H && (unsigned)xx < (unsigned)W)
      rawv[it] = __ldg(reinterpret_cast<const float4*>(x + (((size_t)b * H + y) * W + xx) * C + c0) + c4);
  }
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int e = threadIdx.x + it * 256;
    if (e >= NE) break;
    const int c4 = e % C4, pix = e / C4;
    const int py = pix / TI, px = pix - py * TI;
    const int y = iy0 + py, xx = ix0 + px;
    const float4 raw = rawv[it];
    float4 act = raw;
    if (ss && (unsigned)y < (unsigned)H && (unsigned)xx < (unsigned)W) {   // outside the image both tensors are the FIR's zero padding
      const float4* sp = reinterpret_cast<const float4*>(ss + ((size_t)b * C + c0 + c4 * 4) * 2);
      const float4 a = __ldg(sp), d = __ldg(sp + 1);
      act.x = silu_fast(fmaf(raw.x, a.x, a.y));
      act.y = silu_fast(fmaf(raw.y, a.z, a.w));
      act.z = silu_fast(fmaf(raw.z, d.x, d.y));
      act.w = silu_fast(fmaf(raw.w, d.z, d.w));
    }
    *reinterpret_cast<float4*>(sAct + pix * PS + c4 * 4) = act;
    *reinterpret_cast<float4*>(sRaw + pix * PS + c4 * 4) = raw;
  }
  __syncthreads();
  constexpr int TO = MODE == 1 ? 16 : 8;             // output tile edge
  for (int e = threadIdx.x; e < TO * TO * C4; e += 256) {
    const int c4 = e % C4, op = e / C4;
    const int oyl = op / TO, oxl = op - oyl * TO;
    const int oy = tyi * TO + oyl, ox = txi * TO + oxl;
    if (oy >= OH || ox >= OW) continue;
    float4 ah = make_float4(0, 0, 0, 0), ax = make_float4(0, 0, 0, 0);
    if (MODE == 1) {
      // o[2i] = x[i-1]/4 + 3x[i]/4 ; o[2i+1] = 3x[i]/4 + x[i+1]/4   (staged tile starts at input pixel i0 - 1)
      const int ly = (oyl >> 1) + ((oyl & 1) ? 1 : 0), lx = (oxl >> 1) + ((oxl & 1) ? 1 : 0);   // first tap, staged coords
      const float wy0 = (oyl & 1) ? 0.75f : 0.25f, wx0 = (oxl & 1) ? 0.75f : 0.25f;
#pragma unroll
      for (int jy = 0; jy < 2; ++jy)
#pragma unroll
        for (int jx = 0; jx < 2; ++jx) {
          const float wgt = (jy ? 1.0f - wy0 : wy0) * (jx ? 1.0f - wx0 : wx0);
          const int pix = (ly + jy) * TI + lx + jx;
          const float4 va = *reinterpret_cast<const float4*>(sAct + pix * PS + c4 * 4);
          const float4 vr = *reinterpret_cast<const float4*>(sRaw + pix * PS + c4 * 4);
          ah.x = fmaf(wgt, va.x, ah.x); ah.y = fmaf(wgt, va.y, ah.y); ah.z = fmaf(wgt, va.z, ah.z); ah.w = fmaf(wgt, va.w, ah.w);
          ax.x = fmaf(wgt, vr.x, ax.x); ax.y = fmaf(wgt, vr.y, ax.y); ax.z = fmaf(wgt, vr.z, ax.z); ax.w = fmaf(wgt, vr.w, ax.w);
        }
    } else {
      const float k[4] = {0.125f, 0.375f, 0.375f, 0.125f};
#pragma unroll
      for (int jy = 0; jy < 4; ++jy)
#pragma unroll
        for (int jx = 0; jx < 4; ++jx) {
          const float wgt = k[jy] * k[jx];
          const int pix = (2 * oyl + jy) * TI + 2 * oxl + jx;
          const float4 va = *reinterpret_cast<const float4*>(sAct + pix * PS + c4 * 4);
          const float4 vr = *reinterpret_cast<const float4*>(sRaw + pix * PS + c4 * 4);
          ah.x = fmaf(wgt, va.x, ah.x); ah.y = fmaf(wgt, va.y, ah.y); ah.z = fmaf(wgt, va.z, ah.z); ah.w = fmaf(wgt, va.w, ah.w);
          ax.x = fmaf(wgt, vr.x, ax.x); ax.y = fmaf(wgt, vr.y, ax.y); ax.z = fmaf(wgt, vr.z, ax.z); ax.w = fmaf(wgt, vr.w, ax.w);
        }
    }
    const size_t o = ((((size_t)b * OH + oy) * OW + ox) * C + c0) / 4 + c4;
    if (h_out) reinterpret_cast<float4*>(h_out)[o] = ah;
    if (x_out) reinterpret_cast<float4*>(x_out)[o] = ax;
  }
}

// DOWN, separable and streamed through registers: warp = one output column, lane = 4 channels (a warp reads 512 contiguous
// bytes per pixel), every thread walks R output rows = 2R + 2 input rows; per input row 4 loads (columns 2ox-1 .. 2ox+2, shared
// with the neighbouring warp through L1), the horizontal FIR, and two vertical accumulations.  9 loads per output instead of 16
// shared-memory reads, no staging pass, no barrier; GroupNorm scale / shift of the thread's channels live in registers.
template <int R>
__global__ void __launch_bounds__(256) resample_down_strip_kernel(const float* __restrict__ x, const float* __restrict__ ss,
                                                                  float* __restrict__ h_out, float* __restrict__ x_out, int H, int W, int C) {
  pdl_enter();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int OH = H / 2, OW = W / 2;
  const int ncg = C / 128;
  const int cg = blockIdx.z % ncg, b = blockIdx.z / ncg;
  const int c = cg * 128 + lane * 4;
  const int ox = blockIdx.x * 8 + warp;
  if (ox >= OW) return;
  const int oy0 = blockIdx.y * R;
  float4 s0 = make_float4(1.f, 0.f, 1.f, 0.f), s1 = s0;          // (scale, shift) pairs of the 4 channels
  if (ss) {
    const float4* sp = reinterpret_cast<const float4*>(ss + ((size_t)b * C + c) * 2);
    s0 = __ldg(sp);
    s1 = __ldg(sp + 1);
  }
  const float k[4] = {0.125f, 0.375f, 0.375f, 0.125f};
  const float* xb = x + (size_t)b * H * W * C + c;
  float4 aH = make_float4(0, 0, 0, 0), aX = aH, bH = aH, bX = aH;     // a: output row r, b: output row r - 1
#pragma unroll
  for (int r = 0; r <= R; ++r) {
    float4 v[2][4];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int y = 2 * (oy0 + r) - 1 + t;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int xx = 2 * ox - 1 + j;
        v[t][j] = make_float4(0, 0, 0, 0);
        if ((unsigned)y < (unsigned)H && (unsigned)xx < (unsigned)W) v[t][j] = __ldg(reinterpret_cast<const float4*>(xb + ((size_t)y * W + xx) * C));
      }
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int y = 2 * (oy0 + r) - 1 + t;
      float4 hx = make_float4(0, 0, 0, 0), hh = hx;                // horizontal FIR of the raw / activated row
      if ((unsigned)y < (unsigned)H) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int xx = 2 * ox - 1 + j;
          const float4 raw = v[t][j];
          hx.x = fmaf(k[j], raw.x, hx.x); hx.y = fmaf(k[j], raw.y, hx.y); hx.z = fmaf(k[j], raw.z, hx.z); hx.w = fmaf(k[j], raw.w, hx.w);
          if (h_out && (unsigned)xx < (unsigned)W) {                // outside the image the activated tensor is zero padding too
            float4 a = raw;
            if (ss) {
              a.x = silu_fast(fmaf(raw.x, s0.x, s0.y));
              a.y = silu_fast(fmaf(raw.y, s0.z, s0.w));
              a.z = silu_fast(fmaf(raw.z, s1.x, s1.y));
              a.w = silu_fast(fmaf(raw.w, s1.z, s1.w));
            }
            hh.x = fmaf(k[j], a.x, hh.x); hh.y = fmaf(k[j], a.y, hh.y); hh.z = fmaf(k[j], a.z, hh.z); hh.w = fmaf(k[j], a.w, hh.w);
          }
        }
      }
      // input row 2(oy0+r)-1+t is tap t of output row r and tap t + 2 of output row r - 1
      aH.x = fmaf(k[t], hh.x, aH.x); aH.y = fmaf(k[t], hh.y, aH.y); aH.z = fmaf(k[t], hh.z, aH.z); aH.w = fmaf(k[t], hh.w, aH.w);
      aX.x = fmaf(k[t], hx.x, aX.x); aX.y = fmaf(k[t], hx.y, aX.y); aX.z = fmaf(k[t], hx.z, aX.z); aX.w = fmaf(k[t], hx.w, aX.w);
      bH.x = fmaf(k[t + 2], hh.x, bH.x); bH.y = fmaf(k[t + 2], hh.y, bH.y); bH.z = fmaf(k[t + 2], hh.z, bH.z); bH.w = fmaf(k[t + 2], hh.w, bH.w);
      bX.x = fmaf(k[t + 2], hx.x, bX.x); bX.y = fmaf(k[t + 2], hx.y, bX.y); bX.z = fmaf(k[t + 2], hx.z, bX.z); bX.w = fmaf(k[t + 2], hx.w, bX.w);
    }
    if (r > 0 && oy0 + r - 1 < OH) {
      const size_t o = (((size_t)b * OH + oy0 + r - 1) * OW + ox) * C + c;
      if (h_out) *reinterpret_cast<float4*>(h_out + o) = bH;
      if (x_out) *reinterpret_cast<float4*>(x_out + o) = bX;
    }
    bH = aH; bX = aX;
    aH = make_float4(0, 0, 0, 0); aX = aH;
  }
}

// UP, streamed the same way: warp = one input column (two output columns), lane = 4 channels, R input rows (+ one halo row on
// each side) per thread; per input row 3 loads, the horizontal pair H0 = a[i-1]/4 + 3a[i]/4, H1 = 3a[i]/4 + a[i+1]/4, and with the
// previous row's pair the two output rows 2(i-1)+1 = 3H[i-1]/4 + H[i]/4 and 2i = H[i-1]/4 + 3H[i]/4.
template <int R>
__global__ void __launch_bounds__(256) resample_up_strip_kernel(const float* __restrict__ x, const float* __restrict__ ss,
                                                                float* __restrict__ h_out, float* __restrict__ x_out, int H, int W, int C) {
  pdl_enter();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int OW = 2 * W;
  const int ncg = C / 128;
  const int cg = blockIdx.z % ncg, b = blockIdx.z / ncg;
  const int c = cg * 128 + lane * 4;
  const int ix = blockIdx.x * 8 + warp;
  if (ix >= W) return;
  const int iy0 = blockIdx.y * R;
  float4 s0 = make_float4(1.f, 0.f, 1.f, 0.f), s1 = s0;
  if (ss) {
    const float4* sp = reinterpret_cast<const float4*>(ss + ((size_t)b * C + c) * 2);
    s0 = __ldg(sp);
    s1 = __ldg(sp + 1);
  }
  const float* xb = x + (size_t)b * H * W * C + c;
  auto lin = [](float wa, const float4& a, float wb, const float4& bq) {
    return make_float4(fmaf(wb, bq.x, wa * a.x), fmaf(wb, bq.y, wa * a.y), fmaf(wb, bq.z, wa * a.z), fmaf(wb, bq.w, wa * a.w));
  };
  const float4 zero = make_float4(0, 0, 0, 0);
  float4 pH0 = zero, pH1 = zero, pX0 = zero, pX1 = zero;           // horizontal pairs of the previous input row
#pragma unroll
  for (int r = -1; r <= R; ++r) {
    const int y = iy0 + r;
    float4 v[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int xx = ix - 1 + j;
      v[j] = zero;
      if ((unsigned)y < (unsigned)H && (unsigned)xx < (unsigned)W) v[j] = __ldg(reinterpret_cast<const float4*>(xb + ((size_t)y * W + xx) * C));
    }
    float4 cH0 = zero, cH1 = zero, cX0 = zero, cX1 = zero;
    if ((unsigned)y < (unsigned)H) {
      if (x_out) { cX0 = lin(0.25f, v[0], 0.75f, v[1]); cX1 = lin(0.75f, v[1], 0.25f, v[2]); }
      if (h_out) {
        float4 a[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int xx = ix - 1 + j;
          a[j] = v[j];
          if (ss && (unsigned)xx < (unsigned)W) {                  // outside the image the activated tensor is zero padding too
            a[j].x = silu_fast(fmaf(v[j].x, s0.x, s0.y));
            a[j].y = silu_fast(fmaf(v[j].y, s0.z, s0.w));
            a[j].z = silu_fast(fmaf(v[j].z, s1.x, s1.y));
            a[j].w = silu_fast(fmaf(v[j].w, s1.z, s1.w));
          }
        }
        cH0 = lin(0.25f, a[0], 0.75f, a[1]);
        cH1 = lin(0.75f, a[1], 0.25f, a[2]);
      }
    }
    // output rows 2(y-1)+1 (belongs to input row y-1) and 2y (input row y)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int iy = y - 1 + half;                                 // owning input row
      if (iy < iy0 || iy >= iy0 + R || iy >= H) continue;
      const int oy = 2 * iy + (half ? 0 : 1);
      const float wp = half ? 0.25f : 0.75f, wc = half ? 0.75f : 0.25f;
      const size_t o = (((size_t)b * 2 * H + oy) * OW + 2 * ix) * C + c;
      if (h_out) {
        *reinterpret_cast<float4*>(h_out + o) = lin(wp, pH0, wc, cH0);
        *reinterpret_cast<float4*>(h_out + o + C) = lin(wp, pH1, wc, cH1);
      }
      if (x_out) {
        *reinterpret_cast<float4*>(x_out + o) = lin(wp, pX0, wc, cX0);
        *reinterpret_cast<float4*>(x_out + o + C) = lin(wp, pX1, wc, cX1);
      }
    }
    pH0 = cH0; pH1 = cH1; pX0 = cX0; pX1 = cX1;
  }
}

int launch_resample(const Act& x, const float* ss, float* h_out, float* x_out, int mode, cudaStream_t st) {
  if (x.C % 4 != 0) return fail(-1, "resample: C must be a multiple of 4");
  if (mode == 2 && ((x.H | x.W) & 1)) return fail(-1, "resample: downsampling needs even extents");
  static const bool no_strip = getenv("STORM_B200_RESAMPLE_TILED") && atoi(getenv("STORM_B200_RESAMPLE_TILED")) != 0;
  if (mode == 1 && x.C % 128 == 0 && !no_strip) {
    constexpr int R = 8;
    dim3 grid(ceil_div(x.W, 8), ceil_div(x.H, R), x.B * (x.C / 128));
    storm_launch(resample_up_strip_kernel<R>, grid, 256, 0, st, x.p, ss, h_out, x_out, x.H, x.W, x.C);
    STORM_LAUNCH_CHECK();
    return 0;
  }
  if (mode == 2 && x.C % 128 == 0 && !no_strip) {
    constexpr int R = 8;
    dim3 grid(ceil_div(x.W / 2, 8), ceil_div(x.H / 2, R), x.B * (x.C / 128));
    storm_launch(resample_down_strip_kernel<R>, grid, 256, 0, st, x.p, ss, h_out, x_out, x.H, x.W, x.C);
    STORM_LAUNCH_CHECK();
    return 0;
  }
  if (mode == 1 && x.C % 32 == 0) {
    dim3 grid(ceil_div(x.H, 8) * ceil_div(x.W, 8), x.C / 32, x.B);
    storm_launch(resample_tiled_kernel<1>, grid, 256, 2 * 10 * 10 * 32 * sizeof(float), st, x.p, ss, h_out, x_out, x.H, x.W, x.C);
    STORM_LAUNCH_CHECK();
    return 0;
  }
  if (mode == 2 && x.C % 16 == 0) {
    STORM_SET_SMEM_ONCE((resample_tiled_kernel<2>), 64 * 1024);
    dim3 grid(ceil_div(x.H / 2, 8) * ceil_div(x.W / 2, 8), x.C / 16, x.B);
    storm_launch(resample_tiled_kernel<2>, grid, 256, 2 * 18 * 18 * 20 * sizeof(float), st, x.p, ss, h_out, x_out, x.H, x.W, x.C);
    STORM_LAUNCH_CHECK();
    return 0;
  }
  const int OH = mode == 1 ? 2 * x.H : x.H / 2, OW = mode == 1 ? 2 * x.W : x.W / 2;
  size_t total = (size_t)x.B * OH * OW * (x.C / 4);
  unsigned blocks = (unsigned)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  storm_launch(resample_kernel, blocks, 256, 0, st, x.p, ss, h_out, x_out, x.B, x.H, x.W, x.C, mode);
  STORM_LAUNCH_CHECK();
  return 0;
}

// C input channels -> CO >= C output channels (channels beyond C are written as zeros: K padding of the tensor-core kernel).
// Only the first Creal input channels carry data (the rest are that same zero padding): one thread per output pixel and
// group of 4 real channels issues its 16 tap loads together, the remaining channels are stored as zeros.
__global__ void __launch_bounds__(256) fir_down_small_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int H, int W,
                                                             int C, int CO, int Creal) {
  pdl_enter();
  const int OH = H / 2, OW = W / 2;
  const int n4 = (Creal + 3) / 4;                                  // float4 groups with data
  const size_t total = (size_t)B * OH * OW * n4;
  const float k[4] = {0.125f, 0.375f, 0.375f, 0.125f};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = i % n4;
    size_t r = i / n4;
    const int ox = r % OW; r /= OW;
    const int oy = r % OH;
    const int b = r / OH;
    float4 v[16];
#pragma unroll
    for (int jy = 0; jy < 4; ++jy)
#pragma unroll
      for (int jx = 0; jx < 4; ++jx) {
        const int y = 2 * oy - 1 + jy, xx = 2 * ox - 1 + jx;
        v[jy * 4 + jx] = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((unsigned)y < (unsigned)H && (unsigned)xx < (unsigned)W)
          v[jy * 4 + jx] = __ldg(reinterpret_cast<const float4*>(x + (((size_t)b * H + y) * W + xx) * C) + c4);
      }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int jy = 0; jy < 4; ++jy)                                   // same tap order as the scalar loop it replaces
#pragma unroll
      for (int jx = 0; jx < 4; ++jx) {
        const float wgt = k[jy] * k[jx];
        acc.x = fmaf(wgt, v[jy * 4 + jx].x, acc.x);
        acc.y = fmaf(wgt, v[jy * 4 + jx].y, acc.y);
        acc.z = fmaf(wgt, v[jy * 4 + jx].z, acc.z);
        acc.w = fmaf(wgt, v[jy * 4 + jx].w, acc.w);
      }
    float4* o = reinterpret_cast<float4*>(out + (((size_t)b * OH + oy) * OW + ox) * CO);
    o[c4] = acc;
    if (c4 == 0)
      for (int q = n4; q < CO / 4; ++q) o[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
int launch_fir_down_small(const float* x, float* out, int B, int H, int W, int C, int CO, int Creal, cudaStream_t st) {
  if ((C & 3) || (CO & 3) || Creal > C || Creal < 1 || CO < C) return fail(-1, "fir_down_small: channel counts must be multiples of 4, Creal <= C <= CO");
  size_t total = (size_t)B * (H / 2) * (W / 2) * ((Creal + 3) / 4);
  unsigned blocks = (unsigned)((total + 255) / 256 > 148 * 8 ? 148 * 8 : (total + 255) / 256);
  if (blocks == 0) blocks = 1;
  storm_launch(fir_down_small_kernel, blocks, 256, 0, st, x, out, B, H, W, C, CO, Creal);
  STORM_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================
// input / pyramid / head
// =====================================================================================================
__global__ void unpack_input_kernel(UnpackSrc src, float* __restrict__ out, int B, int nc, int CP, int HW) {
  pdl_enter();
  const size_t total = (size_t)B * HW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int b = i / HW;
    const int p = i % HW;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};          // 2x - 1 of the up to 3 complex channels, then zeros
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
      if (ch < nc) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(src.p[ch]) + (size_t)b * src.bstride[ch] + p);
        f[2 * ch] = 2.0f * v.x - 1.0f;
        f[2 * ch + 1] = 2.0f * v.y - 1.0f;
      }
    float4* o = reinterpret_cast<float4*>(out + i * CP);
    o[0] = make_float4(f[0], f[1], f[2], f[3]);
    o[1] = make_float4(f[4], f[5], f[6], f[7]);
    for (int q = 2; q < CP / 4; ++q) o[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
int launch_unpack_input(UnpackSrc src, float* out, int B, int nc, int CP, int HW, cudaStream_t st) {
  if (nc < 1 || nc > 3 || CP < 8 || (CP & 3)) return fail(-1, "unpack_input: 1..3 complex channels, padded width a multiple of 4 and >= 8");
  size_t total = (size_t)B * HW;
  unsigned blocks = (unsigned)((total + 255) / 256 > 148 * 8 ? 148 * 8 : (total + 255) / 256);
  storm_launch(unpack_input_kernel, blocks, 256, 0, st, src, out, B, nc, CP, HW);
  STORM_LAUNCH_CHECK();
  return 0;
}
__global__ void fill_kernel(float* p, float v, int n) {
  pdl_enter();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
int launch_fill(float* p, float v, int n, cudaStream_t st) {
  storm_launch(fill_kernel, ceil_div(n, 256), 256, 0, st, p, v, n);
  STORM_LAUNCH_CHECK();
  return 0;
}

__global__ void combine_kernel(const float* __restrict__ h, const float* __restrict__ pyr, int pstride, const float* __restrict__ w,
                               const float* __restrict__ bias, float* __restrict__ out, size_t npix, int C, int Cin) {
  pdl_enter();
  const int C4 = C >> 2;
  const size_t total = npix * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = i % C4;
    const size_t p = i / C4;
    float4 hv = __ldg(reinterpret_cast<const float4*>(h) + i);
    float acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = bias[c4 * 4 + j];
    for (int k = 0; k < Cin; ++k) {
      float pv = __ldg(pyr + p * pstride + k);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(__ldg(w + (size_t)(c4 * 4 + j) * Cin + k), pv, acc[j]);
    }
    reinterpret_cast<float4*>(out)[i] = make_float4(acc[0] + hv.x, acc[1] + hv.y, acc[2] + hv.z, acc[3] + hv.w);
  }
}
int launch_combine(const float* h, const float* pyr, int pyr_stride, const float* w, const float* bias, float* out, int B, int HW, int C,
                   int Cin, cudaStream_t st) {
  size_t total = (size_t)B * HW * (C / 4);
  unsigned blocks = (unsigned)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  storm_launch(combine_kernel, blocks, 256, 0, st, h, pyr, pyr_stride, w, bias, out, (size_t)B * HW, C, Cin);
  STORM_LAUNCH_CHECK();
  return 0;
}

// ---- progressive output branch: conv3x3 C -> Co<=8 with GN+SiLU prologue ------------------------------
// CTA = 16 x 32 output pixels, 128 threads, each thread a 2x2 pixel block x NCO output channels (register
// blocking: 16 activations + 9 weight vectors feed 36*NCO FMAs per input channel).
namespace {
constexpr int OTH = 16, OTW = 32;       // output tile
constexpr int OKC = 8;                  // channels per staged chunk
constexpr int OHH = OTH + 2, OHW = OTW + 2;
constexpr int OHP = OHW + 2;            // padded row pitch (even -> float2 loads stay aligned)
}
template <int NCO>
__global__ void __launch_bounds__(128) conv_out_small_kernel(const float* __restrict__ x, const float* __restrict__ ss,
                                                             const float* __restrict__ w, const float* __restrict__ bias,
                                                             const float* __restrict__ prev, float* __restrict__ out,
                                                             int H, int W, int C, int Co) {
  pdl_enter();
  __shared__ __align__(16) float sA[OKC][OHH][OHP];
  __shared__ __align__(16) float sW[9][OKC][8];
  const int tiles_x = (W + OTW - 1) / OTW;
  const int ty0 = (blockIdx.x / tiles_x) * OTH, tx0 = (blockIdx.x % tiles_x) * OTW;
  const int b = blockIdx.y;
  const int bx = threadIdx.x & 15, by = threadIdx.x >> 4;    // 2x2 block coordinates inside the tile
  float acc[4][NCO];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NCO; ++j) acc[i][j] = 0.f;
  for (int c0 = 0; c0 < C; c0 += OKC) {
    for (int e = threadIdx.x; e < OKC * OHH * OHW; e += 128) {
      const int ch = e % OKC;
      const int pix = e / OKC;
      const int hy = pix / OHW, hx = pix - hy * OHW;
      const int y = ty0 + hy - 1, xx = tx0 + hx - 1;
      float v = 0.f;
      if ((unsigned)y < (unsigned)H && (unsigned)xx < (unsigned)W && c0 + ch < C) {
        v = __ldg(x + (((size_t)b * H + y) * W + xx) * C + c0 + ch);
        if (ss) {
          const float* s2 = ss + ((size_t)b * C + c0 + ch) * 2;
          v = silu_fast(fmaf(v, __ldg(s2), __ldg(s2 + 1)));
        }
      }
      sA[ch][hy][hx] = v;
    }
    for (int e = threadIdx.x; e < 9 * OKC * 8; e += 128) {
      const int j = e % 8;
      const int ch = (e / 8) % OKC;
      const int tap = e / (8 * OKC);
      sW[tap][ch][j] = (c0 + ch < C) ? __ldg(w + ((size_t)tap * C + c0 + ch) * 8 + j) : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int ch = 0; ch < OKC; ++ch) {
      float a[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float2 v0 = *reinterpret_cast<const float2*>(&sA[ch][2 * by + r][2 * bx]);
        const float2 v1 = *reinterpret_cast<const float2*>(&sA[ch][2 * by + r][2 * bx + 2]);
        a[r][0] = v0.x; a[r][1] = v0.y; a[r][2] = v1.x; a[r][3] = v1.y;
      }
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          float wv[NCO];
          const float4 w0 = *reinterpret_cast<const float4*>(&sW[dy * 3 + dx][ch][0]);
          wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w;
          if (NCO == 8) {
            const float4 w1 = *reinterpret_cast<const float4*>(&sW[dy * 3 + dx][ch][4]);
            wv[NCO - 4] = w1.x; wv[NCO - 3] = w1.y; wv[NCO - 2] = w1.z; wv[NCO - 1] = w1.w;
          }
#pragma unroll
          for (int py = 0; py < 2; ++py)
#pragma unroll
            for (int px = 0; px < 2; ++px) {
              const float av = a[py + dy][px + dx];
#pragma unroll
              for (int j = 0; j < NCO; ++j) acc[py * 2 + px][j] = fmaf(av, wv[j], acc[py * 2 + px][j]);
            }
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int py = 0; py < 2; ++py)
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      const int y = ty0 + 2 * by + py, xx = tx0 + 2 * bx + px;
      if (y >= H || xx >= W) continue;
      for (int j = 0; j < Co; ++j) {
        float v = acc[py * 2 + px][j < NCO ? j : 0] + bias[j];
        if (prev) {   // + FIR_up(prev) at (y, xx): prev is [B][H/2][W/2][Co]
          const int PH = H / 2, PW = W / 2;
          const int iy = y >> 1, ix = xx >> 1;
          int y0, y1, x0, x1; float wy0, wy1, wx0, wx1;
          if (y & 1) { y0 = iy; wy0 = 0.75f; y1 = iy + 1; wy1 = 0.25f; } else { y0 = iy - 1; wy0 = 0.25f; y1 = iy; wy1 = 0.75f; }
          if (xx & 1) { x0 = ix; wx0 = 0.75f; x1 = ix + 1; wx1 = 0.25f; } else { x0 = ix - 1; wx0 = 0.25f; x1 = ix; wx1 = 0.75f; }
          float u = 0.f;
          const float* pb = prev + (size_t)b * PH * PW * Co + j;
          if ((unsigned)y0 < (unsigned)PH) {
            if ((unsigned)x0 < (unsigned)PW) u = fmaf(wy0 * wx0, __ldg(pb + ((size_t)y0 * PW + x0) * Co), u);
            if ((unsigned)x1 < (unsigned)PW) u = fmaf(wy0 * wx1, __ldg(pb + ((size_t)y0 * PW + x1) * Co), u);
          }
          if ((unsigned)y1 < (unsigned)PH) {
            if ((unsigned)x0 < (unsigned)PW) u = fmaf(wy1 * wx0, __ldg(pb + ((size_t)y1 * PW + x0) * Co), u);
            if ((unsigned)x1 < (unsigned)PW) u = fmaf(wy1 * wx1, __ldg(pb + ((size_t)y1 * PW + x1) * Co), u);
          }
          v = u + v;   // pyramid = upsample(pyramid) + pyramid_h (ncsnpp.py:405-410)
        }
        out[(((size_t)b * H + y) * W + xx) * Co + j] = v;
      }
    }
}
int launch_conv_out_small(const Act& h, const float* ss, const float* w, const float* bias, const float* prev, float* out,
                          int Co, cudaStream_t st) {
  if (Co > 8) return fail(-1, "conv_out_small: Co > 8");
  dim3 grid(ceil_div(h.W, OTW) * ceil_div(h.H, OTH), h.B);
  if (Co <= 4) storm_launch(conv_out_small_kernel<4>, grid, 128, 0, st, h.p, ss, w, bias, prev, out, h.H, h.W, h.C, Co);
  else storm_launch(conv_out_small_kernel<8>, grid, 128, 0, st, h.p, ss, w, bias, prev, out, h.H, h.W, h.C, Co);
  STORM_LAUNCH_CHECK();
  return 0;
}

__global__ void pack_convout_gemm_kernel(const float* __restrict__ w, float* __restrict__ out, int Co, int Cin, int Npad) {
  pdl_enter();
  const int total = Cin * Npad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i % Npad, c = i / Npad;
    float v = 0.f;
    if (n < 9 * Co) {
      const int tap = n / Co, co = n - tap * Co;
      v = w[((size_t)co * Cin + c) * 9 + tap];
    }
    out[i] = v;
  }
}
int launch_pack_convout_gemm(const float* w_oihw, float* out_io, int Co, int Cin, int Npad, cudaStream_t st) {
  storm_launch(pack_convout_gemm_kernel, ceil_div(Cin * Npad, 256), 256, 0, st, w_oihw, out_io, Co, Cin, Npad);
  STORM_LAUNCH_CHECK();
  return 0;
}

__global__ void tap_gather_kernel(const float* __restrict__ Z, int Npad, const float* __restrict__ bias, const float* __restrict__ prev,
                                  float* __restrict__ out, int B, int H, int W, int Co) {
  pdl_enter();
  const size_t total = (size_t)B * H * W * Co;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int j = i % Co;
    size_t r = i / Co;
    const int xx = r % W; r /= W;
    const int y = r % H;
    const int b = r / H;
    float v = 0.f;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = y + dy - 1;
      if ((unsigned)yy >= (unsigned)H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int xs = xx + dx - 1;
        if ((unsigned)xs >= (unsigned)W) continue;
        v += __ldg(Z + (((size_t)b * H + yy) * W + xs) * Npad + (dy * 3 + dx) * Co + j);
      }
    }
    v += bias[j];
    if (prev) {
      const int PH = H / 2, PW = W / 2;
      const int iy = y >> 1, ix = xx >> 1;
      int y0, y1, x0, x1; float wy0, wy1, wx0, wx1;
      if (y & 1) { y0 = iy; wy0 = 0.75f; y1 = iy + 1; wy1 = 0.25f; } else { y0 = iy - 1; wy0 = 0.25f; y1 = iy; wy1 = 0.75f; }
      if (xx & 1) { x0 = ix; wx0 = 0.75f; x1 = ix + 1; wx1 = 0.25f; } else { x0 = ix - 1; wx0 = 0.25f; x1 = ix; wx1 = 0.75f; }
      float u = 0.f;
      const float* pb = prev + (size_t)b * PH * PW * Co + j;
      if ((unsigned)y0 < (unsigned)PH) {
        if ((unsigned)x0 < (unsigned)PW) u = fmaf(wy0 * wx0, __ldg(pb + ((size_t)y0 * PW + x0) * Co), u);
        if ((unsigned)x1 < (unsigned)PW) u = fmaf(wy0 * wx1, __ldg(pb + ((size_t)y0 * PW + x1) * Co), u);
      }
      if ((unsigned)y1 < (unsigned)PH) {
        if ((unsigned)x0 < (unsigned)PW) u = fmaf(wy1 * wx0, __ldg(pb + ((size_t)y1 * PW + x0) * Co), u);
        if ((unsigned)x1 < (unsigned)PW) u = fmaf(wy1 * wx1, __ldg(pb + ((size_t)y1 * PW + x1) * Co), u);
      }
      v = u + v;
    }
    out[i] = v;
  }
}
int launch_tap_gather(const float* Z, int Npad, const float* bias, const float* prev, float* out, int B, int H, int W, int Co,
                      cudaStream_t st) {
  size_t total = (size_t)B * H * W * Co;
  unsigned blocks = (unsigned)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  storm_launch(tap_gather_kernel, blocks, 256, 0, st, Z, Npad, bias, prev, out, B, H, W, Co);
  STORM_LAUNCH_CHECK();
  return 0;
}

__global__ void head_kernel(const float* __restrict__ pyr, const float* __restrict__ t, const float* __restrict__ w,
                            const float* __restrict__ bias, float2* __restrict__ out, int B, int HW, int Ci) {
  pdl_enter();
  const size_t total = (size_t)B * HW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int b = i / HW;
    float re = bias[0], im = bias[1];
    const float tb = t ? t[b] : 1.0f;
    for (int k = 0; k < Ci; ++k) {
      float v = __ldg(pyr + i * Ci + k);
      if (t) v = v / tb;                     // h / used_sigmas before the output conv (ncsnpp.py:441-446)
      re = fmaf(w[k], v, re);
      im = fmaf(w[Ci + k], v, im);
    }
    out[i] = make_float2(re, im);
  }
}
int launch_head(const float* pyr, const float* t, const float* w, const float* bias, float* out_c, int B, int HW, int Ci,
                cudaStream_t st) {
  size_t total = (size_t)B * HW;
  unsigned blocks = (unsigned)((total + 255) / 256 > 148 * 8 ? 148 * 8 : (total + 255) / 256);
  storm_launch(head_kernel, blocks, 256, 0, st, pyr, t, w, bias, reinterpret_cast<float2*>(out_c), B, HW, Ci);
  STORM_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================
// attention (SIMT flash-style): one CTA = 16 queries, key tiles of 32
// =====================================================================================================
namespace {
constexpr int AQ = 16, AK = 32;
}
__global__ void __launch_bounds__(256) attention_simt_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                             const float* __restrict__ v, float* __restrict__ o, int N, int C,
                                                             float scale) {
  pdl_enter();
  extern __shared__ float sm[];
  const int CP = C + 4;
  float* Qs = sm;                    // [AQ][CP]
  float* Ks = Qs + AQ * CP;          // [AK][CP]
  float* Vs = Ks + AK * CP;          // [AK][C]
  float* Ss = Vs + AK * C;           // [AQ][AK]
  float* alpha = Ss + AQ * AK;       // [AQ]
  float* lsum = alpha + AQ;          // [AQ]
  const int b = blockIdx.y;
  const int q0 = blockIdx.x * AQ;
  const int tid = threadIdx.x;
  const int C4 = C >> 2;
  const float* qb = q + (size_t)b * N * C;
  const float* kb = k + (size_t)b * N * C;
  const float* vb = v + (size_t)b * N * C;
  for (int e = tid; e < AQ * C4; e += 256) {
    int r = e / C4, c4 = e % C4;
    float4 val = make_float4(0, 0, 0, 0);
    if (q0 + r < N) val = __ldg(reinterpret_cast<const float4*>(qb + (size_t)(q0 + r) * C) + c4);
    *reinterpret_cast<float4*>(Qs + r * CP + c4 * 4) = val;
  }
  float O[AQ];
#pragma unroll
  for (int i = 0; i < AQ; ++i) O[i] = 0.f;
  const int sq = tid >> 4;      // query row handled in the score phase
  const int sk = tid & 15;      // keys sk and sk + 16
  float m_run = -INFINITY, l_run = 0.f;   // replicated across the 16 lanes of a query row
  for (int k0 = 0; k0 < N; k0 += AK) {
    __syncthreads();
    for (int e = tid; e < AK * C4; e += 256) {
      int r = e / C4, c4 = e % C4;
      float4 kv = make_float4(0, 0, 0, 0), vv = make_float4(0, 0, 0, 0);
      if (k0 + r < N) {
        kv = __ldg(reinterpret_cast<const float4*>(kb + (size_t)(k0 + r) * C) + c4);
        vv = __ldg(reinterpret_cast<const float4*>(vb + (size_t)(k0 + r) * C) + c4);
      }
      *reinterpret_cast<float4*>(Ks + r * CP + c4 * 4) = kv;
      *reinterpret_cast<float4*>(Vs + r * C + c4 * 4) = vv;
    }
    __syncthreads();
    // ---- scores
    float s0 = 0.f, s1 = 0.f;
    {
      const float4* qp = reinterpret_cast<const float4*>(Qs + sq * CP);
      const float4* k0p = reinterpret_cast<const float4*>(Ks + sk * CP);
      const float4* k1p = reinterpret_cast<const float4*>(Ks + (sk + 16) * CP);
      for (int c4 = 0; c4 < C4; ++c4) {
        float4 a = qp[c4], b0 = k0p[c4], b1 = k1p[c4];
        s0 = fmaf(a.x, b0.x, s0); s0 = fmaf(a.y, b0.y, s0); s0 = fmaf(a.z, b0.z, s0); s0 = fmaf(a.w, b0.w, s0);
        s1 = fmaf(a.x, b1.x, s1); s1 = fmaf(a.y, b1.y, s1); s1 = fmaf(a.z, b1.z, s1); s1 = fmaf(a.w, b1.w, s1);
      }
    }
    s0 = (k0 + sk < N) ? s0 * scale : -INFINITY;
    s1 = (k0 + sk + 16 < N) ? s1 * scale : -INFINITY;
    float mx = fmaxf(s0, s1);
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    const float m_new = fmaxf(m_run, mx);
    const float p0 = expf(s0 - m_new), p1 = expf(s1 - m_new);
    float ps = p0 + p1;
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
    const float al = expf(m_run - m_new);
    l_run = l_run * al + ps;
    m_run = m_new;
    Ss[sq * AK + sk] = p0;
    Ss[sq * AK + sk + 16] = p1;
    if (sk == 0) alpha[sq] = al;
    __syncthreads();
    // ---- O = O * alpha + P V   (thread = channel)
    if (tid < C) {
#pragma unroll
      for (int i = 0; i < AQ; ++i) O[i] *= alpha[i];
      for (int kk = 0; kk < AK; kk += 4) {
        float v0 = Vs[(kk + 0) * C + tid], v1 = Vs[(kk + 1) * C + tid], v2 = Vs[(kk + 2) * C + tid], v3 = Vs[(kk + 3) * C + tid];
#pragma unroll
        for (int i = 0; i < AQ; ++i) {
          float4 p = *reinterpret_cast<const float4*>(Ss + i * AK + kk);
          O[i] = fmaf(p.x, v0, O[i]); O[i] = fmaf(p.y, v1, O[i]); O[i] = fmaf(p.z, v2, O[i]); O[i] = fmaf(p.w, v3, O[i]);
        }
      }
    }
  }
  if (sk == 0) lsum[sq] = l_run;
  __syncthreads();
  if (tid < C) {
#pragma unroll
    for (int i = 0; i < AQ; ++i)
      if (q0 + i < N) o[((size_t)b * N + q0 + i) * C + tid] = O[i] / lsum[i];
  }
}
int launch_attention_simt(const float* q, const float* k, const float* v, float* o, int B, int N, int C, cudaStream_t st) {
  if (C > 256 || C % 4 != 0) return fail(-1, "attention_simt: C must be <= 256 and a multiple of 4");
  size_t smem = ((size_t)AQ * (C + 4) + (size_t)AK * (C + 4) + (size_t)AK * C + AQ * AK + 2 * AQ) * sizeof(float);
  STORM_SET_SMEM_ONCE((attention_simt_kernel), 100 * 1024);
  storm_launch(attention_simt_kernel, dim3(ceil_div(N, AQ), B), 256, smem, st, q, k, v, o, N, C, 1.0f / sqrtf((float)C));
  STORM_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================
// sampler updates + Philox
// =====================================================================================================
namespace {
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
// complex standard normal (var 1/2 per component), counter = (element, draw, utterance id), key = seed
__device__ __forceinline__ float2 philox_cnormal(uint64_t seed, uint64_t utt, uint32_t draw, uint32_t idx) {
  uint32_t c[4] = {idx, draw, (uint32_t)utt, (uint32_t)(utt >> 32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  const float u1 = ((float)(c[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0,1)
  const float u2 = ((float)(c[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float r = sqrtf(-logf(u1));   // sqrt(-2 ln u1) * sqrt(1/2)
  float s, co;
  sincospif(2.0f * u2, &s, &co);
  return make_float2(r * co, r * s);
}
}  // namespace

__global__ void sampler_update_kernel(const float2* __restrict__ x, const float2* __restrict__ y, const float2* __restrict__ d,
                                      float2* __restrict__ x_mean, float2* __restrict__ x_new, UpdateCoef coef,
                                      const UpdateCoef* __restrict__ coef_dev, const float2* __restrict__ noise, uint64_t seed,
                                      const uint64_t* __restrict__ seed_dev, const uint64_t* __restrict__ utt_ids, uint32_t draw, int64_t n) {
  pdl_enter();
  const int b = blockIdx.y;
  if (coef_dev) coef = *coef_dev;
  if (seed_dev) seed = *seed_dev;             // captured trajectories read the seed from device memory
  const uint64_t utt = utt_ids ? utt_ids[b] : (uint64_t)b;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const size_t g = (size_t)b * n + i;
    float2 m = make_float2(0.f, 0.f);
    if (coef.a != 0.f) { float2 v = x[g]; m.x = coef.a * v.x; m.y = coef.a * v.y; }
    if (coef.b != 0.f) { float2 v = y[g]; m.x = fmaf(coef.b, v.x, m.x); m.y = fmaf(coef.b, v.y, m.y); }
    if (coef.c != 0.f) { float2 v = d[g]; m.x = fmaf(coef.c, v.x, m.x); m.y = fmaf(coef.c, v.y, m.y); }
    if (x_mean) x_mean[g] = m;
    if (x_new) {
      float2 z = make_float2(0.f, 0.f);
      if (coef.e != 0.f || noise) z = noise ? noise[g] : philox_cnormal(seed, utt, draw, (uint32_t)i);
      x_new[g] = make_float2(fmaf(coef.e, z.x, m.x), fmaf(coef.e, z.y, m.y));
    }
  }
}
int launch_sampler_update(const float* x, const float* y, const float* d, float* x_mean, float* x_new, UpdateCoef coef,
                          const UpdateCoef* coef_dev, const float* noise, uint64_t seed, const uint64_t* seed_dev,
                          const uint64_t* utt_ids_dev, uint32_t draw, int B, int64_t n, cudaStream_t st) {
  unsigned bx = (unsigned)((n + 255) / 256 > 148 * 4 ? 148 * 4 : (n + 255) / 256);
  storm_launch(sampler_update_kernel, dim3(bx, B), 256, 0, st, 
      reinterpret_cast<const float2*>(x), reinterpret_cast<const float2*>(y), reinterpret_cast<const float2*>(d),
      reinterpret_cast<float2*>(x_mean), reinterpret_cast<float2*>(x_new), coef, coef_dev,
      reinterpret_cast<const float2*>(noise), seed, seed_dev, utt_ids_dev, draw, n);
  STORM_LAUNCH_CHECK();
  return 0;
}

__global__ void philox_fill_kernel(uint64_t seed, uint64_t utt, uint32_t draw, int64_t n, float2* out) {
  pdl_enter();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = philox_cnormal(seed, utt, draw, (uint32_t)i);
}
int launch_philox_fill(uint64_t seed, uint64_t utt_id, uint32_t draw, int64_t n, float* out, cudaStream_t st) {
  unsigned bx = (unsigned)((n + 255) / 256 > 148 * 4 ? 148 * 4 : (n + 255) / 256);
  storm_launch(philox_fill_kernel, bx, 256, 0, st, seed, utt_id, draw, n, reinterpret_cast<float2*>(out));
  STORM_LAUNCH_CHECK();
  return 0;
}

// scratch: [B][64][2] doubles
__global__ void sqnorm_partial_kernel(const float2* __restrict__ d, const float2* __restrict__ noise, uint64_t seed,
                                      const uint64_t* __restrict__ seed_dev, const uint64_t* __restrict__ utt_ids, uint32_t draw,
                                      int64_t n, double* scratch) {
  pdl_enter();
  const int b = blockIdx.y;
  if (seed_dev) seed = *seed_dev;
  const uint64_t utt = utt_ids ? utt_ids[b] : (uint64_t)b;
  double sd = 0, sz = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float2 v = d[(size_t)b * n + i];
    float2 z = noise ? noise[(size_t)b * n + i] : philox_cnormal(seed, utt, draw, (uint32_t)i);
    sd += (double)v.x * v.x + (double)v.y * v.y;
    sz += (double)z.x * z.x + (double)z.y * z.y;
  }
  __shared__ double r0[8], r1[8];
  sd = warp_sum_d(sd); sz = warp_sum_d(sz);
  if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = sd; r1[threadIdx.x >> 5] = sz; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, c = 0;
    for (int i = 0; i < 8; ++i) { a += r0[i]; c += r1[i]; }
    scratch[((size_t)b * gridDim.x + blockIdx.x) * 2] = a;
    scratch[((size_t)b * gridDim.x + blockIdx.x) * 2 + 1] = c;
  }
}
__global__ void langevin_coef_kernel(const double* scratch, int B, int nblk, float snr, UpdateCoef* out) {
  pdl_enter();
  if (threadIdx.x != 0) return;
  float gn = 0.f, zn = 0.f;   // mean over the batch of per-sample L2 norms (correctors.py:53-54)
  for (int b = 0; b < B; ++b) {
    double a = 0, c = 0;
    for (int i = 0; i < nblk; ++i) { a += scratch[((size_t)b * nblk + i) * 2]; c += scratch[((size_t)b * nblk + i) * 2 + 1]; }
    gn += (float)sqrt(a);
    zn += (float)sqrt(c);
  }
  gn /= (float)B; zn /= (float)B;
  float r = snr * zn / gn;
  float step = r * r * 2.0f;
  out->a = 1.f; out->b = 0.f; out->c = -step; out->e = sqrtf(step * 2.0f);
}
int launch_langevin_coef(const float* d, const float* noise, uint64_t seed, const uint64_t* seed_dev, const uint64_t* utt_ids_dev,
                         uint32_t draw, float snr, int B, int64_t n, double* scratch, UpdateCoef* coef_out, cudaStream_t st) {
  const int nblk = 64;
  storm_launch(sqnorm_partial_kernel, dim3(nblk, B), 256, 0, st, reinterpret_cast<const float2*>(d), reinterpret_cast<const float2*>(noise),
                                                       seed, seed_dev, utt_ids_dev, draw, n, scratch);
  STORM_LAUNCH_CHECK();
  storm_launch(langevin_coef_kernel, 1, 32, 0, st, scratch, B, nblk, snr, coef_out);
  STORM_LAUNCH_CHECK();
  return 0;
}

__global__ void ode_drift_kernel(const float2* __restrict__ x, const float2* __restrict__ y, const float2* __restrict__ d,
                                 float2* __restrict__ out, float theta, float half_g2, int64_t n) {
  pdl_enter();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float2 xv = x[i], yv = y[i], dv = d[i];
    out[i] = make_float2(fmaf(half_g2, dv.x, theta * (yv.x - xv.x)), fmaf(half_g2, dv.y, theta * (yv.y - xv.y)));
  }
}
int launch_ode_drift(const float* x, const float* y, const float* d, float* out, float theta, float half_g2, int64_t n_total,
                     cudaStream_t st) {
  unsigned bx = (unsigned)((n_total + 255) / 256 > 148 * 8 ? 148 * 8 : (n_total + 255) / 256);
  storm_launch(ode_drift_kernel, bx, 256, 0, st, reinterpret_cast<const float2*>(x), reinterpret_cast<const float2*>(y),
                                       reinterpret_cast<const float2*>(d), reinterpret_cast<float2*>(out), theta, half_g2, n_total);
  STORM_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================
// STFT / iSTFT (direct DFT of length n_fft = 510 = 2*3*5*17: not a power of two, data_module.py:143,232)
// =====================================================================================================
// `lens` (optional, device): true length of every utterance inside its L-sample row (ragged batches of one frame bucket)
__global__ void wave_norm_kernel(const float* __restrict__ wav, int L, const int* __restrict__ lens, float* __restrict__ norm) {
  pdl_enter();
  const int b = blockIdx.x;
  const int Lb = lens ? min(lens[b], L) : L;
  float m = 0.f;
  for (int i = threadIdx.x; i < Lb; i += blockDim.x) m = fmaxf(m, fabsf(wav[(size_t)b * L + i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float r[32];
  if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, r[i]);
    norm[b] = m;
  }
}
int launch_wave_norm(const float* wav, int B, int L, const int* lens, float* norm, cudaStream_t st) {
  storm_launch(wave_norm_kernel, B, 1024, 0, st, wav, L, lens, norm);
  STORM_LAUNCH_CHECK();
  return 0;
}

namespace {
__device__ __forceinline__ float window_at(int n, int n_fft, int sqrt_window) {
  // torch.hann_window(n_fft, periodic=True) = 0.5 - 0.5 cos(2 pi n / n_fft)   (data_module.py:19-25)
  double w = 0.5 - 0.5 * cospi(2.0 * (double)n / (double)n_fft);
  return sqrt_window ? (float)sqrt(w) : (float)w;
}
}  // namespace

// smem: frame[n_fft] floats, tw[n_fft] float2
__global__ void __launch_bounds__(256) stft_kernel(const float* __restrict__ wav, const float* __restrict__ norm, int Lrow,
                                                   const int* __restrict__ lens, float2* __restrict__ spec, int T_pad, int n_fft,
                                                   int hop, int sqrt_window, float factor, float expo) {
  pdl_enter();
  extern __shared__ float sm[];
  float* frame = sm;
  float2* tw = reinterpret_cast<float2*>(sm + n_fft + (n_fft & 1));
  const int f = blockIdx.x, b = blockIdx.y;
  const int nb = n_fft / 2 + 1;
  const int L = lens ? min(lens[b], Lrow) : Lrow;      // this utterance's own length: its own reflect padding and frame count
  const int n_frames = 1 + L / hop;
  if (f >= n_frames) {   // pad_spec: zero frames up to the next multiple of 64 (util/other.py:102-109)
    for (int k = threadIdx.x; k < nb; k += blockDim.x) spec[((size_t)b * nb + k) * T_pad + f] = make_float2(0.f, 0.f);
    return;
  }
  const float nrm = norm ? norm[b] : 1.0f;
  const int pad = n_fft / 2;
  for (int n = threadIdx.x; n < n_fft; n += blockDim.x) {
    int j = f * hop + n - pad;                 // center=True, reflect padding (torch.stft)
    if (j < 0) j = -j;
    if (j >= L) j = 2 * (L - 1) - j;
    j = max(0, min(j, L - 1));                 // only reachable for signals shorter than the padding (rejected by the callers)
    float v = wav[(size_t)b * Lrow + j];
    if (norm) v = v / nrm;                     // y = y / norm_factor (model.py:284)
    frame[n] = v * window_at(n, n_fft, sqrt_window);
    double s, c;
    sincospi(2.0 * (double)n / (double)n_fft, &s, &c);
    tw[n] = make_float2((float)c, (float)s);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nb; k += blockDim.x) {
    float re = 0.f, im = 0.f;
    int idx = 0;
    for (int n = 0; n < n_fft; ++n) {
      const float2 w = tw[idx];
      const float v = frame[n];
      re = fmaf(v, w.x, re);
      im = fmaf(-v, w.y, im);
      idx += k;
      if (idx >= n_fft) idx -= n_fft;
    }
    // spec_fwd: |z|^e e^{i angle z} * factor (data_module.py:182-186)
    if (expo != 1.0f) {
      const float mag = hypotf(re, im);
      if (mag > 0.f) {
        const float m2 = (expo == 0.5f) ? sqrtf(mag) : powf(mag, expo);
        re = m2 * (re / mag);
        im = m2 * (im / mag);
      } else {
        re = 0.f; im = 0.f;
      }
    }
    spec[((size_t)b * nb + k) * T_pad + f] = make_float2(re * factor, im * factor);
  }
}
int launch_stft(const float* wav, const float* norm, int B, int L, const int* lens, float* spec, int T_pad, int n_fft, int hop,
                int sqrt_window, float factor, float expo, cudaStream_t st) {
  if (L <= n_fft / 2) return fail(-1, "stft: signal shorter than the reflect padding");
  const int n_frames = 1 + L / hop;
  if (T_pad < n_frames) return fail(-1, "stft: T_pad smaller than the frame count");
  size_t smem = (size_t)(n_fft + (n_fft & 1)) * sizeof(float) + (size_t)n_fft * sizeof(float2);
  storm_launch(stft_kernel, dim3(T_pad, B), 256, smem, st, wav, norm, L, lens, reinterpret_cast<float2*>(spec), T_pad, n_fft, hop,
                                                sqrt_window, factor, expo);
  STORM_LAUNCH_CHECK();
  return 0;
}

// frames_buf[b][f][n] = window[n] * irfft(spec_back(spec[b][:, f]))[n]
__global__ void __launch_bounds__(256) istft_frames_kernel(const float2* __restrict__ spec, int T_pad, float* __restrict__ frames_buf,
                                                           int n_fft, int sqrt_window, float factor, float expo) {
  pdl_enter();
  extern __shared__ float sm[];
  const int nb = n_fft / 2 + 1;
  float2* X = reinterpret_cast<float2*>(sm);          // [nb]
  float2* tw = X + nb;                                // [n_fft]
  const int f = blockIdx.x, b = blockIdx.y;
  for (int k = threadIdx.x; k < nb; k += blockDim.x) {
    float2 z = spec[((size_t)b * nb + k) * T_pad + f];
    float re = z.x / factor, im = z.y / factor;        // spec / spec_factor (data_module.py:189)
    if (expo != 1.0f) {
      const float mag = hypotf(re, im);
      if (mag > 0.f) {
        const float inv = 1.0f / expo;
        const float m2 = (expo == 0.5f) ? mag * mag : powf(mag, inv);
        re = m2 * (re / mag);
        im = m2 * (im / mag);
      } else {
        re = 0.f; im = 0.f;
      }
    }
    X[k] = make_float2(re, im);
  }
  for (int n = threadIdx.x; n < n_fft; n += blockDim.x) {
    double s, c;
    sincospi(2.0 * (double)n / (double)n_fft, &s, &c);
    tw[n] = make_float2((float)c, (float)s);
  }
  __syncthreads();
  const float invn = 1.0f / (float)n_fft;
  for (int n = threadIdx.x; n < n_fft; n += blockDim.x) {
    // c2r inverse of a one-sided spectrum (imaginary parts of DC / Nyquist ignored), n_fft even
    float acc = 0.f;
    int idx = n;   // k * n mod n_fft for k = 1
    for (int k = 1; k < nb - 1; ++k) {
      const float2 w = tw[idx];
      const float2 xk = X[k];
      acc = fmaf(xk.x, w.x, acc);
      acc = fmaf(-xk.y, w.y, acc);
      idx += n;
      if (idx >= n_fft) idx -= n_fft;
    }
    float v = X[0].x + ((n & 1) ? -X[nb - 1].x : X[nb - 1].x) + 2.0f * acc;
    v *= invn;
    frames_buf[((size_t)b * T_pad + f) * n_fft + n] = v * window_at(n, n_fft, sqrt_window);
  }
}
__global__ void istft_ola_kernel(const float* __restrict__ frames_buf, const float* __restrict__ norm, int T_pad,
                                 float* __restrict__ wav, int L, const int* __restrict__ lens, int n_fft, int hop, int sqrt_window) {
  pdl_enter();
  const int b = blockIdx.y;
  const int Lb = lens ? min(lens[b], L) : L;           // torch.istft(length = this utterance's length); the row tail is zeroed
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < L; m += gridDim.x * blockDim.x) {
    if (m >= Lb) { wav[(size_t)b * L + m] = 0.f; continue; }
    const int mp = m + n_fft / 2;
    int f_lo = (mp - n_fft + 1 + hop - 1) / hop;
    if (mp - n_fft + 1 < 0) f_lo = 0;
    int f_hi = mp / hop;
    if (f_hi > T_pad - 1) f_hi = T_pad - 1;
    float y = 0.f, env = 0.f;
    for (int f = f_lo; f <= f_hi; ++f) {
      const int n = mp - f * hop;
      const float w = window_at(n, n_fft, sqrt_window);
      y += frames_buf[((size_t)b * T_pad + f) * n_fft + n];
      env = fmaf(w, w, env);
    }
    float v = (f_hi >= f_lo) ? y / env : 0.f;
    if (norm) v *= norm[b];                    // x_hat * norm_factor (model.py:302)
    wav[(size_t)b * L + m] = v;
  }
}
int launch_istft(const float* spec, const float* norm, int B, int T_pad, float* frames_buf, float* wav, int L, const int* lens,
                 int n_fft, int hop, int sqrt_window, float factor, float expo, cudaStream_t st) {
  if (n_fft & 1) return fail(-1, "istft: n_fft must be even");
  const int nb = n_fft / 2 + 1;
  size_t smem = (size_t)(nb + n_fft) * sizeof(float2);
  storm_launch(istft_frames_kernel, dim3(T_pad, B), 256, smem, st, reinterpret_cast<const float2*>(spec), T_pad, frames_buf, n_fft, sqrt_window,
                                                        factor, expo);
  STORM_LAUNCH_CHECK();
  storm_launch(istft_ola_kernel, dim3(ceil_div(L, 256) > 512 ? 512 : ceil_div(L, 256), B), 256, 0, st, frames_buf, norm, T_pad, wav, L, lens, n_fft,
                                                                                       hop, sqrt_window);
  STORM_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================
// SI-SDR (sgmse/util/other.py:82-94): alpha = <s_hat, s> / |s|^2 ; 10 log10(|alpha s|^2 / |alpha s - s_hat|^2)
// One CTA per utterance, two passes, fp64 accumulation in a fixed order (deterministic).  eps = 0 reproduces si_sdr
// (numpy), eps = 1e-10 si_sdr_torch.
// =====================================================================================================
namespace {
__device__ __forceinline__ double block_sum_1024(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
  return t;
}
}  // namespace
__global__ void __launch_bounds__(1024) si_sdr_kernel(const float* __restrict__ s, const float* __restrict__ sh, int L,
                                                      const int* __restrict__ lens, double eps, float* __restrict__ out) {
  pdl_enter();
  __shared__ double red[32];
  const int b = blockIdx.x;
  const int Lb = lens ? min(lens[b], L) : L;
  const float* sp = s + (size_t)b * L;
  const float* hp = sh + (size_t)b * L;
  double dot = 0.0, ss = 0.0;
  for (int i = threadIdx.x; i < Lb; i += blockDim.x) {
    const double a = sp[i], c = hp[i];
    dot = fma(c, a, dot);
    ss = fma(a, a, ss);
  }
  dot = block_sum_1024(dot, red);
  ss = block_sum_1024(ss, red);
  const double alpha = dot / ss;
  double num = 0.0, den = 0.0;
  for (int i = threadIdx.x; i < Lb; i += blockDim.x) {
    const double a = alpha * (double)sp[i], d = a - (double)hp[i];
    num = fma(a, a, num);
    den = fma(d, d, den);
  }
  num = block_sum_1024(num, red);
  den = block_sum_1024(den, red);
  if (threadIdx.x == 0) out[b] = (float)(10.0 * log10(eps + num / (eps + den)));
}
int launch_si_sdr(const float* s, const float* s_hat, int B, int L, const int* lens, double eps, float* out, cudaStream_t st) {
  storm_launch(si_sdr_kernel, B, 1024, 0, st, s, s_hat, L, lens, eps, out);
  STORM_LAUNCH_CHECK();
  return 0;
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int HW, int C, size_t total) {
  pdl_enter();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = i % C;
    const size_t r = i / C;
    const int p = r % HW;
    const size_t b = r / HW;
    out[(b * C + c) * HW + p] = in[i];
  }
}
int launch_nhwc_to_nchw(const float* in, float* out, int B, int HW, int C, cudaStream_t st) {
  size_t total = (size_t)B * HW * C;
  unsigned blocks = (unsigned)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  storm_launch(nhwc_to_nchw_kernel, blocks, 256, 0, st, in, out, HW, C, total);
  STORM_LAUNCH_CHECK();
  return 0;
}

}  // namespace storm
