// General-purpose fp32 SIMT implicit-GEMM convolution (3x3 pad 1 / 1x1) over NHWC activations with the
// fused prologue (GroupNorm-apply + SiLU) and epilogue (bias, time-embedding bias, residual, 1/sqrt2).
// It handles every channel count (C % 4 == 0) and partial tiles; the tcgen05 kernel (conv_tc.cu) takes
// over for the shapes that dominate the FLOPs.  Replaces the cuDNN conv2d + ATen GroupNorm/SiLU calls of
// ResnetBlockBigGANpp.forward (sgmse/backbones/ncsnpp_utils/layerspp.py:242-274).
#include "common.cuh"

namespace storm {

namespace {

constexpr int TH = 8, TW = 16;         // output tile: 8 rows x 16 cols = 128 pixels
constexpr int BN = 128;                // output channels per CTA
constexpr int KC = 8;                  // input channels per staged chunk
constexpr int HH = TH + 2, HWD = TW + 2;
constexpr int NT = 256;
constexpr int SA_FLOATS = KC * HH * HWD;          // 1440
constexpr int SW_FLOATS = 9 * KC * BN;            // 9216
constexpr int SMEM_BYTES = (SA_FLOATS + SW_FLOATS) * 4;

__device__ __forceinline__ float load_act(const ConvSeg& s, int b, int y, int x, int c, int H, int W) {
  // post-prologue activation at (b, y, x, channel c of the concatenation); zero outside the image
  if ((unsigned)y >= (unsigned)H || (unsigned)x >= (unsigned)W) return 0.f;
  const int C = s.C0 + s.C1;
  if (c >= C) return 0.f;
  size_t pix = ((size_t)b * H + y) * W + x;
  float v = (c < s.C0) ? __ldg(s.src0 + pix * s.C0 + c) : __ldg(s.src1 + pix * s.C1 + (c - s.C0));
  if (s.ss) {
    const float* ss = s.ss + ((size_t)b * C + c) * 2;
    v = fmaf(v, __ldg(ss), __ldg(ss + 1));
  }
  if (s.silu) v = silu_f(v);
  return v;
}

__global__ void __launch_bounds__(NT, 2) conv_simt_kernel(ConvParams p) {
  pdl_enter();
  extern __shared__ float smem[];
  float* sA = smem;               // [KC][HH][HWD]
  float* sW = smem + SA_FLOATS;   // [9][KC][BN]

  const int tid = threadIdx.x;
  const int tiles_x = (p.W + TW - 1) / TW;
  const int ty0 = (blockIdx.x / tiles_x) * TH;
  const int tx0 = (blockIdx.x % tiles_x) * TW;
  const int n0 = blockIdx.y * BN;
  const int b = blockIdx.z;

  const int cg = tid & 15;        // cout group: couts n0 + cg*4 + {0..3} and n0 + 64 + cg*4 + {0..3}
  const int pg = tid >> 4;        // pixel group 0..15
  const int row = pg >> 1;        // 0..7
  const int xh = (pg & 1) * 8;    // x offset of this thread's 8 consecutive pixels

  float acc[8][8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int n = 0; n < 8; ++n) acc[j][n] = 0.f;

  for (int si = 0; si < p.nseg; ++si) {
    const ConvSeg s = p.seg[si];
    const int C = s.C0 + s.C1;
    const int ntap = s.taps;
    for (int c0 = 0; c0 < C; c0 += KC) {
      // ---- stage activations (prologue applied) -------------------------------------------------------
      for (int e = tid; e < SA_FLOATS; e += NT) {
        int ch = e % KC;
        int pix = e / KC;
        int hy = pix / HWD, hx = pix % HWD;
        float v = 0.f;
        if (ntap == 9 || (hy >= 1 && hy <= TH && hx >= 1 && hx <= TW))
          v = load_act(s, b, ty0 + hy - 1, tx0 + hx - 1, c0 + ch, p.H, p.W);
        sA[(ch * HH + hy) * HWD + hx] = v;
      }
      // ---- stage weights [tap][ch][BN] ---------------------------------------------------------------
      for (int e = tid; e < ntap * KC * (BN / 4); e += NT) {
        int n4 = e % (BN / 4);
        int ch = (e / (BN / 4)) % KC;
        int tap = e / (BN / 4 * KC);
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        int n = n0 + n4 * 4;
        if (c0 + ch < C && n < p.Cout)
          w = __ldg(reinterpret_cast<const float4*>(s.w + ((size_t)tap * C + c0 + ch) * p.Cout + n));
        *reinterpret_cast<float4*>(sW + (tap * KC + ch) * BN + n4 * 4) = w;
      }
      __syncthreads();
      // ---- compute -----------------------------------------------------------------------------------
      if (ntap == 9) {
#pragma unroll 2
        for (int ch = 0; ch < KC; ++ch) {
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            float a[10];
            const float* ap = sA + (ch * HH + row + dy) * HWD + xh;
#pragma unroll
            for (int j = 0; j < 10; ++j) a[j] = ap[j];
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              const float* wp = sW + ((dy * 3 + dx) * KC + ch) * BN + cg * 4;
              float4 w0 = *reinterpret_cast<const float4*>(wp);
              float4 w1 = *reinterpret_cast<const float4*>(wp + 64);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float av = a[j + dx];
                acc[j][0] = fmaf(av, w0.x, acc[j][0]);
                acc[j][1] = fmaf(av, w0.y, acc[j][1]);
                acc[j][2] = fmaf(av, w0.z, acc[j][2]);
                acc[j][3] = fmaf(av, w0.w, acc[j][3]);
                acc[j][4] = fmaf(av, w1.x, acc[j][4]);
                acc[j][5] = fmaf(av, w1.y, acc[j][5]);
                acc[j][6] = fmaf(av, w1.z, acc[j][6]);
                acc[j][7] = fmaf(av, w1.w, acc[j][7]);
              }
            }
          }
        }
      } else {
#pragma unroll 2
        for (int ch = 0; ch < KC; ++ch) {
          const float* ap = sA + (ch * HH + row + 1) * HWD + xh + 1;
          const float* wp = sW + ch * BN + cg * 4;
          float4 w0 = *reinterpret_cast<const float4*>(wp);
          float4 w1 = *reinterpret_cast<const float4*>(wp + 64);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float av = ap[j];
            acc[j][0] = fmaf(av, w0.x, acc[j][0]);
            acc[j][1] = fmaf(av, w0.y, acc[j][1]);
            acc[j][2] = fmaf(av, w0.z, acc[j][2]);
            acc[j][3] = fmaf(av, w0.w, acc[j][3]);
            acc[j][4] = fmaf(av, w1.x, acc[j][4]);
            acc[j][5] = fmaf(av, w1.y, acc[j][5]);
            acc[j][6] = fmaf(av, w1.z, acc[j][6]);
            acc[j][7] = fmaf(av, w1.w, acc[j][7]);
          }
        }
      }
      __syncthreads();
    }
  }

  // ---- epilogue ----------------------------------------------------------------------------------------
  const int y = ty0 + row;
  if (y >= p.H) return;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int n = n0 + half * 64 + cg * 4;
    if (n >= p.Cout) continue;
    float4 add = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias) add = __ldg(reinterpret_cast<const float4*>(p.bias + n));
    if (p.temb) {
      float4 t = __ldg(reinterpret_cast<const float4*>(p.temb + (size_t)b * p.temb_stride + n));
      add.x += t.x; add.y += t.y; add.z += t.z; add.w += t.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int x = tx0 + xh + j;
      if (x >= p.W) continue;
      size_t o = (((size_t)b * p.H + y) * p.W + x) * p.Cout + n;
      float4 v = make_float4(acc[j][half * 4 + 0] + add.x, acc[j][half * 4 + 1] + add.y,
                             acc[j][half * 4 + 2] + add.z, acc[j][half * 4 + 3] + add.w);
      if (p.res) {
        float4 r = __ldg(reinterpret_cast<const float4*>(p.res + o));
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
      }
      v.x *= p.out_scale; v.y *= p.out_scale; v.z *= p.out_scale; v.w *= p.out_scale;
      *reinterpret_cast<float4*>(p.out + o) = v;
    }
  }
}

}  // namespace

int launch_conv_simt(const ConvParams& p, cudaStream_t st) {
  if (p.Cout % 4 != 0) return fail(-1, "conv_simt: Cout must be a multiple of 4");
  for (int i = 0; i < p.nseg; ++i) {
    if (p.seg[i].taps != 9 && p.seg[i].taps != 1) return fail(-1, "conv_simt: taps must be 9 or 1");
    if (!p.seg[i].w) return fail(-3, "conv_simt: missing packed weights");
  }
  STORM_SET_SMEM_ONCE((conv_simt_kernel), SMEM_BYTES);
  dim3 grid(ceil_div(p.W, TW) * ceil_div(p.H, TH), ceil_div(p.Cout, BN), p.B);
  storm_launch(conv_simt_kernel, grid, NT, SMEM_BYTES, st, p);
  STORM_LAUNCH_CHECK();
  return 0;
}

}  // namespace storm
