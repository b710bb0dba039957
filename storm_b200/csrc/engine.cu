// Host side of the native engine: network plan (module list + parameter table mirroring
// sgmse/backbones/ncsnpp.py:110-273), weight arena + repacking, workspace pool, NCSN++ forward
// (ncsnpp.py:281-450), PC sampler loop (sgmse/sampling/__init__.py:54-66) and the C ABI of
// include/storm_b200.h.
#include <nvtx3/nvToolsExt.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/storm_b200.h"
#include "conv_tc.cuh"
#include "kernels.cuh"

namespace storm {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

// ---------------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------------
// Programmatic dependent launch is wired through every kernel (pdl_enter / storm_launch, common.cuh) and was measured inside
// whole benchmark runs: 0 (off) 13.34 utt/s, conv kernels only 13.27-13.34, small kernels only 13.20, everything 13.04 - with the
// trajectory replayed as one CUDA graph the launch gaps are already hidden, and dependents that become resident early only
// take issue slots from the running kernel.  Off by default; STORM_B200_PDL = 1 | 2 | 3 switches it on for the small kernels /
// the conv kernels / both (profiles/r2_bench_history.md).
// NVTX ranges on the host side of the path (enhance > stft / sampler trajectory > network evaluation / istft): header-only
// NVTX 3, a pointer check per call when no tool is attached.  A replayed CUDA graph shows as one "trajectory (graph replay)" range.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

int pdl_mode() {
  static const int m = [] { const char* e = getenv("STORM_B200_PDL"); return e ? atoi(e) : 0; }();
  return m;
}

constexpr int PYR_CH = 16;   // channels of the (zero-padded) network input = level 0 of the input pyramid (one 16-channel K chunk)
constexpr int PYR_DOWN_CH = 32;   // channels of the downsampled pyramid levels: a whole 32-channel K chunk, so that the down blocks'
                                  // Conv_1 + Conv_2 + Combine launches stay on the two-pass operand split

enum ModKind { M_FOURIER, M_LINEAR, M_CONV_IN, M_RES, M_ATTN, M_COMBINE, M_GN, M_CONV_OUT };

struct Mod {
  int index;
  ModKind kind;
  int in_ch = 0, out_ch = 0;
  bool up = false, down = false;
};

struct ParamEntry {
  std::string name;
  int64_t shape[4] = {1, 1, 1, 1};
  int ndim = 0;
  int64_t offset = 0, numel = 0;
};

struct Pool {
  struct Buf { void* p; size_t bytes; bool free; };
  std::vector<Buf> bufs;
  size_t total = 0;
  int get(size_t bytes, void** out) {
    if (bytes == 0) bytes = 256;
    bytes = (bytes + 255) & ~(size_t)255;
    int best = -1;
    for (size_t i = 0; i < bufs.size(); ++i)
      if (bufs[i].free && bufs[i].bytes >= bytes && (best < 0 || bufs[i].bytes < bufs[best].bytes)) best = (int)i;
    if (best >= 0 && bufs[best].bytes <= 2 * bytes + (1 << 20)) {
      bufs[best].free = false;
      *out = bufs[best].p;
      return 0;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(STORM_E_NOMEM, std::string("workspace cudaMalloc failed: ") + cudaGetErrorString(e));
    bufs.push_back({p, bytes, false});
    total += bytes;
    *out = p;
    return 0;
  }
  void put(void* p) {
    if (!p) return;
    for (auto& b : bufs)
      if (b.p == p) { b.free = true; return; }
  }
  void destroy() {
    for (auto& b : bufs) cudaFree(b.p);
    bufs.clear();
    total = 0;
  }
};

// returns its buffers to the pool when the scope ends, early error returns included
struct PoolLease {
  Pool* pool;
  std::vector<void*> held;
  explicit PoolLease(Pool* p) : pool(p) {}
  ~PoolLease() { for (void* q : held) pool->put(q); }
  int get(size_t bytes, void** out) {
    const int rc = pool->get(bytes, out);
    if (rc == 0) held.push_back(*out);
    return rc;
  }
};

struct ResPack {
  const float *gn0_g = nullptr, *gn0_b = nullptr, *gn1_g = nullptr, *gn1_b = nullptr;
  const float *b0 = nullptr;        // Conv_0.bias
  float* b12 = nullptr;             // Conv_1.bias (+ Conv_2.bias)
  float *w0 = nullptr, *w1 = nullptr, *w2 = nullptr;   // SIMT layouts
  TcWeights tc0, tc1, tc2;                              // tcgen05 layouts (tc1 and tc2 share one weight scale: one launch)
  int temb_off = 0;
  bool has_skip_conv = false;
  // down blocks followed by a Combine: Conv_2 and sqrt(2)*Combine.Conv_0 concatenated along K (x' | padded pyramid)
  TcWeights tc2c, tc1c;     // tc1c: Conv_1 packed with the scale it shares with tc2c
  float* b12c = nullptr;
};
struct AttnPack {
  const float *gn_g = nullptr, *gn_b = nullptr;
  const float* w[4] = {nullptr, nullptr, nullptr, nullptr};   // NIN W is already [in][out]
  const float* b[4] = {nullptr, nullptr, nullptr, nullptr};
  TcWeights tc[4];
};
struct SmallConvPack { float* w = nullptr; const float* b = nullptr; TcWeights tc; int npad = 0; };

}  // namespace storm

using namespace storm;

struct StormHandle {
  StormNetConfig cfg;
  std::vector<Mod> mods;
  std::vector<ParamEntry> table;
  std::map<std::string, int> index;
  int64_t n_floats = 0;
  float* raw = nullptr;
  bool weights_loaded = false;
  std::vector<void*> packed;            // every cudaMalloc'd packed buffer
  std::map<int, ResPack> res;
  std::map<int, AttnPack> attn;
  std::map<int, SmallConvPack> conv_in, conv_out, combine;
  float *W1t = nullptr, *W2t = nullptr, *denseWt = nullptr, *denseB = nullptr;
  int dense_total = 0;
  Pool pool;
  int64_t launches = 0;
  bool taps_on = false;
  std::map<std::string, Act> taps;
  // per-step cache of the time-embedding biases: corrector and predictor of one PC step share t
  float* temb_cache = nullptr;
  int temb_cache_B = 0;       // rows allocated
  int temb_cache_rows = 0;    // rows computed at temb_cache_t
  float temb_cache_t = -1.f;
  bool temb_cache_valid = false;
  bool profile_on = false;
  // captured trajectories (see pc_sample)
  struct TrajKey { int B, F, T, ncond; StormSamplerConfig sc; };
  struct TrajGraph {
    TrajKey key{};
    cudaGraphExec_t exec = nullptr;
    float *Y = nullptr, *c0 = nullptr, *c1 = nullptr, *X = nullptr;
    uint64_t *ids = nullptr, *seed = nullptr;
    int nfe = 0;
    int64_t launches_per_replay = 0;
    bool warmed = false, failed = false;
    Pool pool;      // workspace the captured nodes point into: owned by the graph, never handed to anybody else
  };
  std::vector<TrajGraph> graphs;
  cudaStream_t cap_stream = nullptr;
  bool graph_on = true;
  int64_t graph_replays = 0;
  struct ProfRec { cudaEvent_t a, b; int eng; double flops; };
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_free;
  int device = 0;

  bool conditional() const { return !cfg.discriminative; }
  int in_ch() const { return cfg.discriminative ? 2 : cfg.input_channels; }
  const float* P(const std::string& name) const {
    auto it = index.find(name);
    return it == index.end() ? nullptr : raw + table[it->second].offset;
  }
};

namespace storm {

// Every entry point that takes a handle runs on the device the handle was created on (and restores the caller's
// current device afterwards): `model.cuda(1)` from a process whose current device is 0 must not launch device-0 kernels
// on device-1 pointers.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(const StormHandle* h) {
    if (!h) return;
    if (cudaGetDevice(&prev) == cudaSuccess && prev != h->device) switched = cudaSetDevice(h->device) == cudaSuccess;
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

static void add_param(StormHandle* h, const std::string& name, std::initializer_list<int64_t> shape) {
  ParamEntry e;
  e.name = name;
  e.ndim = (int)shape.size();
  int i = 0;
  e.numel = 1;
  for (auto s : shape) { e.shape[i++] = s; e.numel *= s; }
  e.offset = h->n_floats;                         // every tensor starts 256-byte aligned inside the blob
  h->n_floats += (e.numel + 63) / 64 * 64;
  h->index[name] = (int)h->table.size();
  h->table.push_back(e);
}

static int build_plan(StormHandle* h) {
  const StormNetConfig& c = h->cfg;
  if (c.nf <= 0 || c.nf % 4 != 0) return fail(STORM_E_INVALID, "nf must be a positive multiple of 4");
  if (c.num_levels < 1 || c.num_levels > STORM_MAX_LEVELS) return fail(STORM_E_INVALID, "num_levels out of range");
  if (c.num_res_blocks < 1) return fail(STORM_E_INVALID, "num_res_blocks must be >= 1");
  if (!c.discriminative && c.input_channels != 4 && c.input_channels != 6 && c.input_channels != 2)
    return fail(STORM_E_INVALID, "input_channels must be 2, 4 or 6");
  const int nf = c.nf, nres = c.num_levels, inch = h->in_ch();
  auto in_attn = [&](int r) {
    for (int i = 0; i < c.num_attn_resolutions; ++i)
      if (c.attn_resolutions[i] == r) return true;
    return false;
  };
  std::vector<Mod>& mods = h->mods;
  auto add = [&](ModKind k, int ic, int oc, bool up = false, bool down = false) {
    Mod m; m.index = (int)mods.size(); m.kind = k; m.in_ch = ic; m.out_ch = oc; m.up = up; m.down = down;
    mods.push_back(m);
  };
  add(M_FOURIER, 0, nf);
  if (h->conditional()) { add(M_LINEAR, 2 * nf, 4 * nf); add(M_LINEAR, 4 * nf, 4 * nf); }
  add(M_CONV_IN, inch, nf);
  std::vector<int> hs_c{nf};
  int in_c = nf;
  for (int lvl = 0; lvl < nres; ++lvl) {
    const int r = c.image_size >> lvl;
    for (int b = 0; b < c.num_res_blocks; ++b) {
      int out_c = nf * c.ch_mult[lvl];
      add(M_RES, in_c, out_c);
      in_c = out_c;
      if (in_attn(r)) add(M_ATTN, in_c, in_c);
      hs_c.push_back(in_c);
    }
    if (lvl != nres - 1) {
      add(M_RES, in_c, in_c, false, true);
      add(M_COMBINE, inch, in_c);
      hs_c.push_back(in_c);
    }
  }
  in_c = hs_c.back();
  add(M_RES, in_c, in_c);
  add(M_ATTN, in_c, in_c);
  add(M_RES, in_c, in_c);
  for (int lvl = nres - 1; lvl >= 0; --lvl) {
    const int r = c.image_size >> lvl;
    for (int b = 0; b < c.num_res_blocks + 1; ++b) {
      int out_c = nf * c.ch_mult[lvl];
      add(M_RES, in_c + hs_c.back(), out_c);
      hs_c.pop_back();
      in_c = out_c;
    }
    if (in_attn(r)) add(M_ATTN, in_c, in_c);
    add(M_GN, in_c, in_c);
    add(M_CONV_OUT, in_c, inch);
    if (lvl != 0) add(M_RES, in_c, in_c, true, false);
  }
  // parameter table in reference state_dict order
  add_param(h, "output_layer.weight", {2, inch, 1, 1});
  add_param(h, "output_layer.bias", {2});
  const int64_t temb_dim = 4 * nf;
  for (const Mod& m : mods) {
    std::string p = "all_modules." + std::to_string(m.index);
    switch (m.kind) {
      case M_FOURIER: add_param(h, p + ".W", {m.out_ch}); break;
      case M_LINEAR: add_param(h, p + ".weight", {m.out_ch, m.in_ch}); add_param(h, p + ".bias", {m.out_ch}); break;
      case M_CONV_IN:
      case M_CONV_OUT: add_param(h, p + ".weight", {m.out_ch, m.in_ch, 3, 3}); add_param(h, p + ".bias", {m.out_ch}); break;
      case M_GN: add_param(h, p + ".weight", {m.in_ch}); add_param(h, p + ".bias", {m.in_ch}); break;
      case M_COMBINE: add_param(h, p + ".Conv_0.weight", {m.out_ch, m.in_ch, 1, 1}); add_param(h, p + ".Conv_0.bias", {m.out_ch}); break;
      case M_ATTN:
        add_param(h, p + ".GroupNorm_0.weight", {m.in_ch}); add_param(h, p + ".GroupNorm_0.bias", {m.in_ch});
        for (int i = 0; i < 4; ++i) {
          add_param(h, p + ".NIN_" + std::to_string(i) + ".W", {m.in_ch, m.in_ch});
          add_param(h, p + ".NIN_" + std::to_string(i) + ".b", {m.in_ch});
        }
        break;
      case M_RES:
        add_param(h, p + ".GroupNorm_0.weight", {m.in_ch}); add_param(h, p + ".GroupNorm_0.bias", {m.in_ch});
        add_param(h, p + ".Conv_0.weight", {m.out_ch, m.in_ch, 3, 3}); add_param(h, p + ".Conv_0.bias", {m.out_ch});
        add_param(h, p + ".Dense_0.weight", {m.out_ch, temb_dim}); add_param(h, p + ".Dense_0.bias", {m.out_ch});
        add_param(h, p + ".GroupNorm_1.weight", {m.out_ch}); add_param(h, p + ".GroupNorm_1.bias", {m.out_ch});
        add_param(h, p + ".Conv_1.weight", {m.out_ch, m.out_ch, 3, 3}); add_param(h, p + ".Conv_1.bias", {m.out_ch});
        if (m.in_ch != m.out_ch || m.up || m.down) {
          add_param(h, p + ".Conv_2.weight", {m.out_ch, m.in_ch, 1, 1}); add_param(h, p + ".Conv_2.bias", {m.out_ch});
        }
        break;
    }
  }
  return 0;
}

#define LAUNCH(h, call)   \
  do {                    \
    STORM_TRY(call);      \
    ++(h)->launches;      \
  } while (0)

static int dev_alloc(StormHandle* h, size_t floats, float** out) {
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, std::max<size_t>(floats, 64) * sizeof(float));
  if (e != cudaSuccess) return fail(STORM_E_NOMEM, std::string("weight cudaMalloc failed: ") + cudaGetErrorString(e));
  h->packed.push_back(p);
  *out = (float*)p;
  return 0;
}

// captured trajectories bake device pointers into their nodes: drop them whenever something they point at is re-allocated
static void invalidate_graphs(StormHandle* h) {
  for (auto& g : h->graphs) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    for (void* q : {(void*)g.Y, (void*)g.c0, (void*)g.c1, (void*)g.X, (void*)g.ids, (void*)g.seed}) if (q) cudaFree(q);
    g.pool.destroy();
  }
  h->graphs.clear();
}

static void free_packed(StormHandle* h) {
  invalidate_graphs(h);
  for (void* p : h->packed) cudaFree(p);
  h->packed.clear();
  h->res.clear(); h->attn.clear(); h->conv_in.clear(); h->conv_out.clear(); h->combine.clear();
  h->W1t = h->W2t = h->denseWt = h->denseB = nullptr;
}

// Build the kernel-native copies of the weights from the raw blob.
static int repack(StormHandle* h, cudaStream_t st) {
  free_packed(h);
  const int nf = h->cfg.nf;
  const int D = 4 * nf;
  const bool use_tc = h->cfg.engine == 0;
  // Dense_0 of every ResnetBlock -> one [D][dense_total] matrix
  h->dense_total = 0;
  for (const Mod& m : h->mods)
    if (m.kind == M_RES) h->dense_total += m.out_ch;
  if (h->conditional()) {
    STORM_TRY(dev_alloc(h, (size_t)2 * nf * D, &h->W1t));
    STORM_TRY(dev_alloc(h, (size_t)D * D, &h->W2t));
    STORM_TRY(dev_alloc(h, (size_t)D * h->dense_total, &h->denseWt));
    STORM_TRY(dev_alloc(h, (size_t)h->dense_total, &h->denseB));
    LAUNCH(h, launch_transpose_into(h->P("all_modules.1.weight"), h->W1t, D, 2 * nf, D, 0, st));
    LAUNCH(h, launch_transpose_into(h->P("all_modules.2.weight"), h->W2t, D, D, D, 0, st));
  }
  int off = 0;
  for (const Mod& m : h->mods) {
    const std::string p = "all_modules." + std::to_string(m.index);
    if (m.kind == M_RES) {
      ResPack r;
      r.gn0_g = h->P(p + ".GroupNorm_0.weight"); r.gn0_b = h->P(p + ".GroupNorm_0.bias");
      r.gn1_g = h->P(p + ".GroupNorm_1.weight"); r.gn1_b = h->P(p + ".GroupNorm_1.bias");
      r.b0 = h->P(p + ".Conv_0.bias");
      r.has_skip_conv = h->P(p + ".Conv_2.weight") != nullptr;
      STORM_TRY(dev_alloc(h, (size_t)9 * m.in_ch * m.out_ch, &r.w0));
      LAUNCH(h, launch_pack_conv_simt(h->P(p + ".Conv_0.weight"), r.w0, m.out_ch, m.in_ch, 9, st));
      STORM_TRY(dev_alloc(h, (size_t)9 * m.out_ch * m.out_ch, &r.w1));
      LAUNCH(h, launch_pack_conv_simt(h->P(p + ".Conv_1.weight"), r.w1, m.out_ch, m.out_ch, 9, st));
      STORM_TRY(dev_alloc(h, (size_t)m.out_ch, &r.b12));
      if (r.has_skip_conv) {
        STORM_TRY(dev_alloc(h, (size_t)m.in_ch * m.out_ch, &r.w2));
        LAUNCH(h, launch_pack_conv_simt(h->P(p + ".Conv_2.weight"), r.w2, m.out_ch, m.in_ch, 1, st));
        LAUNCH(h, launch_add_vec(h->P(p + ".Conv_1.bias"), h->P(p + ".Conv_2.bias"), r.b12, m.out_ch, 1.0f, st));
      } else {
        LAUNCH(h, launch_add_vec(h->P(p + ".Conv_1.bias"), nullptr, r.b12, m.out_ch, 1.0f, st));
      }
      if (use_tc) {
        STORM_TRY(tc_pack_weights(h->P(p + ".Conv_0.weight"), m.out_ch, m.in_ch, 9, &r.tc0, &h->packed, st));
        // Conv_1 and Conv_2 accumulate into one TMEM tile: one weight scale for both (largest |w| of the two tensors)
        float a1 = 0.f, a2 = 0.f;
        STORM_TRY(tc_weight_absmax(h->P(p + ".Conv_1.weight"), (size_t)9 * m.out_ch * m.out_ch, &a1, st));
        if (r.has_skip_conv) STORM_TRY(tc_weight_absmax(h->P(p + ".Conv_2.weight"), (size_t)m.out_ch * m.in_ch, &a2, st));
        const float a12 = std::max(a1, a2);
        STORM_TRY(tc_pack_weights(h->P(p + ".Conv_1.weight"), m.out_ch, m.out_ch, 9, &r.tc1, &h->packed, st, 1, a12));
        if (r.has_skip_conv) STORM_TRY(tc_pack_weights(h->P(p + ".Conv_2.weight"), m.out_ch, m.in_ch, 1, &r.tc2, &h->packed, st, 1, a12));
      }
      r.temb_off = off;
      if (h->conditional()) {
        LAUNCH(h, launch_transpose_into(h->P(p + ".Dense_0.weight"), h->denseWt, m.out_ch, D, h->dense_total, off, st));
        STORM_CUDA_TRY(cudaMemcpyAsync(h->denseB + off, h->P(p + ".Dense_0.bias"), m.out_ch * sizeof(float),
                                       cudaMemcpyDeviceToDevice, st));
      }
      off += m.out_ch;
      h->res[m.index] = r;
    } else if (m.kind == M_ATTN) {
      AttnPack a;
      a.gn_g = h->P(p + ".GroupNorm_0.weight"); a.gn_b = h->P(p + ".GroupNorm_0.bias");
      for (int i = 0; i < 4; ++i) {
        a.w[i] = h->P(p + ".NIN_" + std::to_string(i) + ".W");
        a.b[i] = h->P(p + ".NIN_" + std::to_string(i) + ".b");
        // the 1x1 projections and the output-branch GEMMs are latency / HBM bound: the 16-channel split fills its pipeline
        // sooner (measured 31 vs 34 us and 34 vs 43 us per launch), the tensor work is negligible either way
        if (use_tc) STORM_TRY(tc_pack_weights_io(a.w[i], m.in_ch, m.in_ch, &a.tc[i], &h->packed, st, 0));
      }
      h->attn[m.index] = a;
    } else if (m.kind == M_CONV_IN) {
      // input channels are zero-padded to PYR_CH = 16 so that the 4 -> nf conv is a tensor-core K chunk
      SmallConvPack s;
      s.b = h->P(p + ".bias");
      float* wpad = nullptr;
      STORM_TRY(dev_alloc(h, (size_t)9 * PYR_CH * m.out_ch, &wpad));
      LAUNCH(h, launch_pad_cin(h->P(p + ".weight"), wpad, m.out_ch, m.in_ch, PYR_CH, 9, st));
      STORM_TRY(dev_alloc(h, (size_t)9 * PYR_CH * m.out_ch, &s.w));
      LAUNCH(h, launch_pack_conv_simt(wpad, s.w, m.out_ch, PYR_CH, 9, st));
      if (use_tc) STORM_TRY(tc_pack_weights(wpad, m.out_ch, PYR_CH, 9, &s.tc, &h->packed, st));
      h->conv_in[m.index] = s;
    } else if (m.kind == M_CONV_OUT) {
      // [9][C][8] zero padded on the output-channel axis
      SmallConvPack s;
      s.b = h->P(p + ".bias");
      STORM_TRY(dev_alloc(h, (size_t)9 * m.in_ch * 8, &s.w));
      STORM_CUDA_TRY(cudaMemsetAsync(s.w, 0, (size_t)9 * m.in_ch * 8 * sizeof(float), st));
      float* tmp = nullptr;
      STORM_TRY(dev_alloc(h, (size_t)9 * m.in_ch * m.out_ch, &tmp));
      LAUNCH(h, launch_pack_conv_simt(h->P(p + ".weight"), tmp, m.out_ch, m.in_ch, 9, st));   // [9][C][Co]
      STORM_CUDA_TRY(cudaMemcpy2DAsync(s.w, 8 * sizeof(float), tmp, m.out_ch * sizeof(float), m.out_ch * sizeof(float),
                                       (size_t)9 * m.in_ch, cudaMemcpyDeviceToDevice, st));
      if (use_tc) {   // tensor-core route: 1x1 GEMM onto 9*Co tap-partials (see launch_tap_gather)
        s.npad = 9 * m.out_ch;     // the tcgen05 kernel takes any Cout <= 256 (weight rows are zero-padded to 128)
        float* wz = nullptr;
        STORM_TRY(dev_alloc(h, (size_t)m.in_ch * s.npad, &wz));
        LAUNCH(h, launch_pack_convout_gemm(h->P(p + ".weight"), wz, m.out_ch, m.in_ch, s.npad, st));
        STORM_TRY(tc_pack_weights_io(wz, m.in_ch, s.npad, &s.tc, &h->packed, st, 0));
      }
      h->conv_out[m.index] = s;
    } else if (m.kind == M_COMBINE) {
      SmallConvPack s;
      s.b = h->P(p + ".Conv_0.bias");
      s.w = const_cast<float*>(h->P(p + ".Conv_0.weight"));   // [C][Cin] as stored
      h->combine[m.index] = s;
      // fold "h + Conv_0(pyramid)" into the preceding down block's Conv_1 launch:
      //   ((W2 x' + sqrt2 Wc pyr) + Conv_1(h1) + (b1 + b2 + sqrt2 bc)) / sqrt2  =  resblock + Wc pyr + bc
      const Mod& pm = h->mods[m.index - 1];
      auto it = h->res.find(pm.index);
      if (use_tc && pm.kind == M_RES && pm.down && it != h->res.end() && it->second.has_skip_conv) {
        ResPack& r = it->second;
        const std::string pp = "all_modules." + std::to_string(pm.index);
        float* wcat = nullptr;
        STORM_TRY(dev_alloc(h, (size_t)pm.out_ch * (pm.in_ch + PYR_DOWN_CH), &wcat));
        LAUNCH(h, launch_concat_1x1(h->P(pp + ".Conv_2.weight"), pm.in_ch, s.w, m.in_ch, PYR_DOWN_CH, 1.41421356237309505f, wcat, pm.out_ch, st));
        // Conv_1 and the concatenated 1x1 share a launch: same operand split, same weight scale -> repack Conv_1 with it
        float a1 = 0.f, ac = 0.f;
        STORM_TRY(tc_weight_absmax(h->P(pp + ".Conv_1.weight"), (size_t)9 * pm.out_ch * pm.out_ch, &a1, st));
        STORM_TRY(tc_weight_absmax(wcat, (size_t)pm.out_ch * (pm.in_ch + PYR_DOWN_CH), &ac, st));
        const float a1c = std::max(a1, ac);
        STORM_TRY(tc_pack_weights(wcat, pm.out_ch, pm.in_ch + PYR_DOWN_CH, 1, &r.tc2c, &h->packed, st, 1, a1c));
        STORM_TRY(tc_pack_weights(h->P(pp + ".Conv_1.weight"), pm.out_ch, pm.out_ch, 9, &r.tc1c, &h->packed, st, r.tc2c.prec, a1c));
        STORM_TRY(dev_alloc(h, (size_t)pm.out_ch, &r.b12c));
        LAUNCH(h, launch_add_vec(r.b12, s.b, r.b12c, pm.out_ch, 1.41421356237309505f, st));
      }
    }
  }
  h->weights_loaded = true;
  h->temb_cache_valid = false;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------
struct Fwd {
  StormHandle* h;
  cudaStream_t st;
  int B;

  int alloc(Act& a, int H, int W, int C) {
    a = Act();
    a.B = B; a.H = H; a.W = W; a.C = C;
    void* p;
    STORM_TRY(h->pool.get(a.elems() * sizeof(float), &p));
    a.p = (float*)p;
    a.nchunk = gn_num_chunks(H * W);
    return 0;
  }
  void release(Act& a) {
    h->pool.put(a.p);
    h->pool.put(a.part);
    a.p = nullptr; a.part = nullptr; a.part_valid = false;
  }
  int stats(Act& a) {
    if (a.part_valid) return 0;
    void* p;
    STORM_TRY(h->pool.get((size_t)a.B * a.nchunk * a.C * 2 * sizeof(double), &p));
    a.part = (double*)p;
    LAUNCH(h, launch_gn_partial(a, st));
    a.part_valid = true;
    return 0;
  }
  // ss: [B][C][2] scale/shift, followed by [B][gn_groups(C)] mean squares of the raw groups (*ms, optional)
  int gn_ss(Act& a0, Act* a1, const float* g, const float* b, float** ss, const float** ms = nullptr) {
    STORM_TRY(stats(a0));
    if (a1) STORM_TRY(stats(*a1));
    void* p;
    const int C = a0.C + (a1 ? a1->C : 0);
    STORM_TRY(h->pool.get(((size_t)B * C * 2 + (size_t)B * 32) * sizeof(float), &p));
    *ss = (float*)p;
    float* m = *ss + (size_t)B * C * 2;
    if (ms) *ms = m;
    LAUNCH(h, launch_gn_finalize(a0, a1, g, b, *ss, m, st));
    return 0;
  }
  int conv(ConvParams& p, Act* out_act = nullptr) {
    const bool tc = h->cfg.engine == 0 && conv_tc_eligible(p);
    if (tc && out_act && !h->taps_on) {
      // the tcgen05 epilogue emits the GroupNorm partial sums of its output for free
      const int nch = conv_tc_stats_chunks(p);
      if (nch > 0) {
        void* q;
        STORM_TRY(h->pool.get((size_t)B * nch * p.Cout * 2 * sizeof(double), &q));
        out_act->part = (double*)q;
        out_act->nchunk = nch;
        p.stats_out = out_act->part;
        p.stats_chunks = nch;
      }
    }
    StormHandle::ProfRec rec{};
    if (h->profile_on) {
      for (cudaEvent_t* e : {&rec.a, &rec.b}) {
        if (!h->ev_free.empty()) { *e = h->ev_free.back(); h->ev_free.pop_back(); }
        else STORM_CUDA_TRY(cudaEventCreate(e));
      }
      double k = 0;
      for (int i = 0; i < p.nseg; ++i) k += (double)p.seg[i].taps * (p.seg[i].C0 + p.seg[i].C1);
      rec.flops = 2.0 * (double)p.B * p.H * p.W * p.Cout * k;
      rec.eng = tc ? 0 : 1;
      STORM_CUDA_TRY(cudaEventRecord(rec.a, st));
    }
    if (tc) { LAUNCH(h, launch_conv_tc(p, st)); }
    else { LAUNCH(h, launch_conv_simt(p, st)); }
    if (h->profile_on) {
      STORM_CUDA_TRY(cudaEventRecord(rec.b, st));
      h->prof.push_back(rec);
    }
    if (p.stats_out) out_act->part_valid = true;
    return 0;
  }
  int tap(const std::string& name, const Act& a) {
    if (!h->taps_on) return 0;
    Act c = a;
    c.part = nullptr; c.part_valid = false;
    void* p = nullptr;
    STORM_CUDA_TRY(cudaMalloc(&p, a.elems() * sizeof(float)));
    STORM_CUDA_TRY(cudaMemcpyAsync(p, a.p, a.elems() * sizeof(float), cudaMemcpyDeviceToDevice, st));
    c.p = (float*)p;
    auto it = h->taps.find(name);
    if (it != h->taps.end()) cudaFree(it->second.p);
    h->taps[name] = c;
    return 0;
  }

  // ResnetBlockBigGANpp.forward (layerspp.py:242-274).  x1 != null: input is cat[x0, x1] (ncsnpp.py:381).
  int res_block(const Mod& m, Act& x0, Act* x1, const float* tembias, Act& out, const Act* pyr_fuse = nullptr) {
    const ResPack& r = h->res.at(m.index);
    const int Cin = x0.C + (x1 ? x1->C : 0);
    if (Cin != m.in_ch) return fail(STORM_E_STATE, "res_block: channel mismatch");
    float* ss0 = nullptr;
    const float* ms0 = nullptr;       // magnitude of the raw block input (for the launches that consume it un-normalised)
    STORM_TRY(gn_ss(x0, x1, r.gn0_g, r.gn0_b, &ss0, &ms0));
    const int ms0_n = gn_groups(Cin);
    Act hr, xr;       // FIR-resampled activated / raw inputs (up/down blocks)
    const bool resamp = m.up || m.down;
    int H = x0.H, W = x0.W;
    if (resamp) {
      if (x1) return fail(STORM_E_STATE, "res_block: resampling block with concatenated input");
      H = m.up ? 2 * x0.H : x0.H / 2;
      W = m.up ? 2 * x0.W : x0.W / 2;
    }
    // ---- Conv_1 (+ Conv_2 / identity skip) launch, described first: whether the up block's skip can run at LOW
    // resolution depends on the tensor-core kernel taking both launches
    ConvParams p1{};
    Act skipL;        // up blocks: S = Conv_2.weight * x at the INPUT resolution (bias folded into b12)
    bool lowres_skip = false;
    ConvParams ps{};
    {
      ConvSeg& s = p1.seg[0];
      s.src0 = nullptr; s.C0 = m.out_ch; s.src1 = nullptr; s.C1 = 0; s.ss = nullptr; s.silu = 1; s.taps = 9; s.w = r.w1;
      set_wtc(s, r.tc1);
      p1.nseg = 1;
      p1.bias = r.b12;
      p1.B = B; p1.H = H; p1.W = W; p1.Cout = m.out_ch;
      p1.out_scale = 0.70710678118654752f;   // (x + h) / sqrt(2), layerspp.py:274
      static const bool no_lowres = getenv("STORM_B200_NO_LOWRES_SKIP") && atoi(getenv("STORM_B200_NO_LOWRES_SKIP")) != 0;
      if (m.up && r.has_skip_conv && h->cfg.engine == 0 && r.tc1.img && r.tc2.img && !no_lowres) {
        // Conv_2 is 1x1 and the FIR resampling is linear and per channel, so they commute:
        //   Conv_2(FIR_up(x)) = FIR_up(Conv_2.weight * x) + b2   (layerspp.py:250-252,268-269)
        // The 1x1 runs on a quarter of the pixels, FIR_up(x) is never materialised, and Conv_1's epilogue adds the
        // upsampled S (4 x 6 low-resolution loads per 4 x 8 output unit instead of 32 residual loads).
        ps.nseg = 1;
        ConvSeg& k = ps.seg[0];
        k.src0 = x0.p; k.C0 = Cin; k.src1 = nullptr; k.C1 = 0; k.ss = nullptr; k.silu = 0; k.taps = 1; k.w = r.w2;
        set_wtc(k, r.tc2);
        ps.B = B; ps.H = x0.H; ps.W = x0.W; ps.Cout = m.out_ch; ps.out_scale = 1.0f;
        ps.raw_ms = ms0; ps.raw_ms_n = ms0_n;
        ConvParams probe = p1;
        probe.seg[0].src0 = x0.p;                 // any non-null pointer: eligibility does not look at the data
        probe.res_up = x0.p;
        lowres_skip = conv_tc_eligible(ps) && conv_tc_eligible(probe);
      }
    }
    if (resamp) {
      STORM_TRY(alloc(hr, H, W, Cin));
      if (!lowres_skip) STORM_TRY(alloc(xr, H, W, Cin));
      LAUNCH(h, launch_resample(x0, ss0, hr.p, lowres_skip ? nullptr : xr.p, m.up ? 1 : 2, st));
    }
    Act h1;
    STORM_TRY(alloc(h1, H, W, m.out_ch));
    {
      ConvParams p{};
      p.nseg = 1;
      ConvSeg& s = p.seg[0];
      if (resamp) { s.src0 = hr.p; s.C0 = Cin; s.src1 = nullptr; s.C1 = 0; s.ss = nullptr; s.silu = 0; }
      else { s.src0 = x0.p; s.C0 = x0.C; s.src1 = x1 ? x1->p : nullptr; s.C1 = x1 ? x1->C : 0; s.ss = ss0; s.silu = 1; }
      s.taps = 9; s.w = r.w0;
      set_wtc(s, r.tc0);
      p.B = B; p.H = H; p.W = W; p.Cout = m.out_ch;
      p.bias = r.b0;
      p.temb = tembias ? tembias + r.temb_off : nullptr;
      p.temb_stride = h->dense_total;
      p.res = nullptr; p.out_scale = 1.0f; p.out = h1.p;
      STORM_TRY(conv(p, &h1));
    }
    if (resamp) release(hr);
    if (lowres_skip) {
      STORM_TRY(alloc(skipL, x0.H, x0.W, m.out_ch));
      ps.out = skipL.p;
      STORM_TRY(conv(ps));
    }
    float* ss1 = nullptr;
    STORM_TRY(gn_ss(h1, nullptr, r.gn1_g, r.gn1_b, &ss1));
    STORM_TRY(alloc(out, H, W, m.out_ch));
    {
      ConvParams& p = p1;
      p.seg[0].src0 = h1.p;
      p.seg[0].ss = ss1;
      if (lowres_skip) {
        p.res_up = skipL.p;
      } else if (r.has_skip_conv) {
        ConvSeg& k = p.seg[1];
        if (resamp) { k.src0 = xr.p; k.C0 = Cin; k.src1 = nullptr; k.C1 = 0; }
        else { k.src0 = x0.p; k.C0 = x0.C; k.src1 = x1 ? x1->p : nullptr; k.C1 = x1 ? x1->C : 0; }
        k.ss = nullptr; k.silu = 0; k.taps = 1; k.w = r.w2;
        set_wtc(k, r.tc2);
        p.raw_ms = ms0; p.raw_ms_n = ms0_n;      // FIR_down(x) (down blocks) is bounded by x
        if (pyr_fuse) {      // Combine folded in: the padded pyramid rides along as 32 more K channels of the skip segment
          k.src1 = pyr_fuse->p; k.C1 = pyr_fuse->C; k.w = nullptr;
          set_wtc(k, r.tc2c);
          set_wtc(p.seg[0], r.tc1c);
          p.bias = r.b12c;
        }
        p.nseg = 2;
      } else {
        p.res = x0.p;     // identity skip: in_ch == out_ch, no resampling, single source
      }
      p.out = out.p;
      STORM_TRY(conv(p, &out));
    }
    h->pool.put(ss1);
    h->pool.put(ss0);
    release(h1);
    if (lowres_skip) release(skipL);
    if (resamp && !lowres_skip) release(xr);
    return 0;
  }

  // AttnBlockpp.forward (layerspp.py:75-91)
  int attn_block(const Mod& m, Act& x, Act& out) {
    const AttnPack& a = h->attn.at(m.index);
    const int C = x.C, N = x.H * x.W;
    float* ss = nullptr;
    STORM_TRY(gn_ss(x, nullptr, a.gn_g, a.gn_b, &ss));
    Act q, k, v, o;
    STORM_TRY(alloc(q, x.H, x.W, C)); STORM_TRY(alloc(k, x.H, x.W, C)); STORM_TRY(alloc(v, x.H, x.W, C));
    Act* qkv[3] = {&q, &k, &v};
    bool use_tc_attn = h->cfg.engine == 0 && attention_tc_eligible(N, C);
    for (int i = 0; i < 3; ++i) {
      ConvParams p{};
      p.nseg = 1;
      ConvSeg& s = p.seg[0];
      s.src0 = x.p; s.C0 = C; s.src1 = nullptr; s.C1 = 0; s.ss = ss; s.silu = 0; s.taps = 1; s.w = a.w[i]; set_wtc(s, a.tc[i]);
      p.B = B; p.H = x.H; p.W = x.W; p.Cout = C; p.bias = a.b[i]; p.out_scale = 1.0f; p.out = qkv[i]->p;
      if (i == 2 && use_tc_attn) {
        if (conv_tc_eligible(p)) p.out_cm = 1;     // V channel-major: K-major operand of the P.V tensor-core GEMM
        else use_tc_attn = false;
      }
      STORM_TRY(conv(p));
    }
    h->pool.put(ss);
    STORM_TRY(alloc(o, x.H, x.W, C));
    if (use_tc_attn) {
      void* scratch = nullptr;                   // packed bf16 hi/lo operand images of q, k, v
      STORM_TRY(h->pool.get(attention_tc_scratch_bytes(B, N, C), &scratch));
      const int rc_attn = launch_attention_tc(q.p, k.p, v.p, o.p, B, N, C, scratch, st);
      h->pool.put(scratch);
      if (rc_attn != 0) return rc_attn;
      h->launches += 2;                          // operand pre-pass + attention kernel
    }
    else { LAUNCH(h, launch_attention_simt(q.p, k.p, v.p, o.p, B, N, C, st)); }
    release(q); release(k); release(v);
    STORM_TRY(alloc(out, x.H, x.W, C));
    {
      ConvParams p{};
      p.nseg = 1;
      ConvSeg& s = p.seg[0];
      s.src0 = o.p; s.C0 = C; s.src1 = nullptr; s.C1 = 0; s.ss = nullptr; s.silu = 0; s.taps = 1; s.w = a.w[3]; set_wtc(s, a.tc[3]);
      p.B = B; p.H = x.H; p.W = x.W; p.Cout = C; p.bias = a.b[3]; p.res = x.p; p.out_scale = 0.70710678118654752f; p.out = out.p;
      STORM_TRY(conv(p, &out));
    }
    release(o);
    return 0;
  }
};

// per-step cache of the time-embedding biases; re-allocation drops the captured trajectories that point at the old buffer
static int ensure_temb_cache(StormHandle* h, int B) {
  if (h->temb_cache_B >= B) return 0;
  invalidate_graphs(h);
  if (h->temb_cache) cudaFree(h->temb_cache);
  h->temb_cache = nullptr;
  STORM_CUDA_TRY(cudaMalloc((void**)&h->temb_cache, (size_t)B * h->dense_total * sizeof(float)));
  h->temb_cache_B = B;
  h->temb_cache_valid = false;
  return 0;
}

// t_uniform: when non-null, every batch element has this time value (the PC sampler's vec_t) and the
// time-embedding biases are reused across calls with the same value
static int net_forward(StormHandle* h, UnpackSrc src, const float* t, float* out, int B, int F, int T, cudaStream_t st,
                       const float* t_uniform = nullptr) {
  NvtxRange nvtx_range("storm_b200: NCSN++ evaluation");
  if (!h->weights_loaded) return fail(STORM_E_STATE, "weights not loaded: call storm_load_weights first");
  const StormNetConfig& c = h->cfg;
  const int nres = c.num_levels;
  const int div = 1 << (nres - 1);
  if (B < 1 || F < div || T < div || F % div || T % div)
    return fail(STORM_E_INVALID, "F and T must be positive multiples of 2^(num_levels-1) = " + std::to_string(div));
  if (h->conditional() && !t) return fail(STORM_E_INVALID, "time conditioning tensor t is required");
  Fwd f{h, st, B};
  const int inch = h->in_ch();
  const int nf = c.nf;
  auto in_attn = [&](int r) {
    for (int i = 0; i < c.num_attn_resolutions; ++i)
      if (c.attn_resolutions[i] == r) return true;
    return false;
  };
  if (h->taps_on) {
    for (auto& kv : h->taps) cudaFree(kv.second.p);
    h->taps.clear();
  }

  // ---- time embedding -> per-block biases --------------------------------------------------------------
  float* tembias = nullptr;
  float* act = nullptr;
  bool tembias_pooled = false;
  if (h->conditional()) {
    void* p;
    bool hit = false;
    if (t_uniform) {
      STORM_TRY(ensure_temb_cache(h, B));
      tembias = h->temb_cache;
      hit = h->temb_cache_valid && h->temb_cache_t == *t_uniform && h->temb_cache_rows >= B;
    } else {
      STORM_TRY(h->pool.get((size_t)B * h->dense_total * sizeof(float), &p)); tembias = (float*)p;
      tembias_pooled = true;
    }
    if (!hit) {
      STORM_TRY(h->pool.get((size_t)B * 4 * nf * sizeof(float), &p)); act = (float*)p;
      LAUNCH(h, launch_temb_mlp(t, h->P("all_modules.0.W"), h->W1t, h->P("all_modules.1.bias"), h->W2t,
                                h->P("all_modules.2.bias"), act, B, nf, st));
      LAUNCH(h, launch_dense_all(act, h->denseWt, h->denseB, tembias, B, 4 * nf, h->dense_total, st));
      if (t_uniform) { h->temb_cache_valid = true; h->temb_cache_t = *t_uniform; h->temb_cache_rows = B; }
    }
  }

  // ---- input ---------------------------------------------------------------------------------------------
  size_t mi = h->conditional() ? 3 : 1;
  Act x0;      // [B][F][T][PYR_CH]: 2x-1 of the real/imag channels, zero padded (also level 0 of the input pyramid)
  STORM_TRY(f.alloc(x0, F, T, PYR_CH));
  LAUNCH(h, launch_unpack_input(src, x0.p, B, inch / 2, PYR_CH, F * T, st));
  std::vector<Act> hs;
  {
    const Mod& m = h->mods[mi];
    const SmallConvPack& s = h->conv_in.at(m.index);
    Act a;
    STORM_TRY(f.alloc(a, F, T, nf));
    ConvParams p{};
    p.nseg = 1;
    p.seg[0].src0 = x0.p; p.seg[0].C0 = PYR_CH; p.seg[0].taps = 9; p.seg[0].w = s.w; set_wtc(p.seg[0], s.tc);
    p.B = B; p.H = F; p.W = T; p.Cout = nf; p.bias = s.b; p.out_scale = 1.0f; p.out = a.p;
    static const bool no_tc_in = getenv("STORM_B200_NO_FUSE_INPUT") && atoi(getenv("STORM_B200_NO_FUSE_INPUT")) != 0;
    if (no_tc_in) { LAUNCH(h, launch_conv_simt(p, st)); } else { STORM_TRY(f.conv(p, &a)); }
    STORM_TRY(f.tap("conv_in", a));
    hs.push_back(a);
    ++mi;
  }
  Act pyr_in = x0;     // input pyramid (ncsnpp.py:326-328)
  auto tapmod = [&](const Mod& m, const Act& a) { return f.tap("mod" + std::to_string(m.index), a); };

  // ---- down path -----------------------------------------------------------------------------------------
  for (int lvl = 0; lvl < nres; ++lvl) {
    for (int b = 0; b < c.num_res_blocks; ++b) {
      Act o;
      STORM_TRY(f.res_block(h->mods[mi], hs.back(), nullptr, tembias, o));
      STORM_TRY(tapmod(h->mods[mi], o));
      ++mi;
      if (in_attn(o.H)) {
        Act o2;
        STORM_TRY(f.attn_block(h->mods[mi], o, o2));
        STORM_TRY(tapmod(h->mods[mi], o2));
        ++mi;
        f.release(o);
        o = o2;
      }
      hs.push_back(o);
    }
    if (lvl != nres - 1) {
      Act pd;      // input pyramid one level down (ncsnpp.py:353)
      STORM_TRY(f.alloc(pd, pyr_in.H / 2, pyr_in.W / 2, PYR_DOWN_CH));
      LAUNCH(h, launch_fir_down_small(pyr_in.p, pd.p, B, pyr_in.H, pyr_in.W, pyr_in.C, PYR_DOWN_CH, inch, st));
      f.release(pyr_in);
      pyr_in = pd;
      const Mod& dm = h->mods[mi];
      static const bool no_fuse_in = getenv("STORM_B200_NO_FUSE_INPUT") && atoi(getenv("STORM_B200_NO_FUSE_INPUT")) != 0;
      const bool fuse_combine = !no_fuse_in && !h->taps_on && h->cfg.engine == 0 && h->res.at(dm.index).tc2c.img != nullptr;
      Act o;
      STORM_TRY(f.res_block(dm, hs.back(), nullptr, tembias, o, fuse_combine ? &pyr_in : nullptr));
      ++mi;
      if (fuse_combine) {     // `o` already is Combine(pyramid, resblock(h)) (layerspp.py:52-57)
        ++mi;
        hs.push_back(o);
      } else {
        STORM_TRY(tapmod(dm, o));
        const SmallConvPack& s = h->combine.at(h->mods[mi].index);
        Act o2;
        STORM_TRY(f.alloc(o2, o.H, o.W, o.C));
        LAUNCH(h, launch_combine(o.p, pyr_in.p, pyr_in.C, s.w, s.b, o2.p, B, o.H * o.W, o.C, inch, st));
        STORM_TRY(tapmod(h->mods[mi], o2));
        ++mi;
        f.release(o);
        hs.push_back(o2);
      }
    }
  }
  f.release(pyr_in);

  // ---- middle ----------------------------------------------------------------------------------------------
  Act hcur;
  {
    Act a, b2, c2;
    STORM_TRY(f.res_block(h->mods[mi], hs.back(), nullptr, tembias, a)); STORM_TRY(tapmod(h->mods[mi], a)); ++mi;
    STORM_TRY(f.attn_block(h->mods[mi], a, b2)); STORM_TRY(tapmod(h->mods[mi], b2)); ++mi;
    f.release(a);
    STORM_TRY(f.res_block(h->mods[mi], b2, nullptr, tembias, c2)); STORM_TRY(tapmod(h->mods[mi], c2)); ++mi;
    f.release(b2);
    hcur = c2;
  }

  // ---- up path ---------------------------------------------------------------------------------------------
  Act pyr;   // progressive output pyramid, [B][H][W][inch]
  bool have_pyr = false;
  for (int lvl = nres - 1; lvl >= 0; --lvl) {
    for (int b = 0; b < c.num_res_blocks + 1; ++b) {
      Act skip = hs.back();
      hs.pop_back();
      Act o;
      STORM_TRY(f.res_block(h->mods[mi], hcur, &skip, tembias, o));
      STORM_TRY(tapmod(h->mods[mi], o));
      ++mi;
      f.release(hcur);
      f.release(skip);
      hcur = o;
    }
    if (in_attn(hcur.H)) {
      Act o2;
      STORM_TRY(f.attn_block(h->mods[mi], hcur, o2));
      STORM_TRY(tapmod(h->mods[mi], o2));
      ++mi;
      f.release(hcur);
      hcur = o2;
    }
    {
      const Mod& g = h->mods[mi];
      const Mod& co = h->mods[mi + 1];
      const std::string gp = "all_modules." + std::to_string(g.index);
      float* ss = nullptr;
      STORM_TRY(f.gn_ss(hcur, nullptr, h->P(gp + ".weight"), h->P(gp + ".bias"), &ss));
      const SmallConvPack& s = h->conv_out.at(co.index);
      Act np;
      STORM_TRY(f.alloc(np, hcur.H, hcur.W, inch));
      bool done = false;
      if (h->cfg.engine == 0 && s.tc.img) {
        ConvParams p{};
        p.nseg = 1;
        p.seg[0].src0 = hcur.p; p.seg[0].C0 = hcur.C; p.seg[0].ss = ss; p.seg[0].silu = 1; p.seg[0].taps = 1; set_wtc(p.seg[0], s.tc);
        p.B = B; p.H = hcur.H; p.W = hcur.W; p.Cout = s.npad; p.out_scale = 1.0f;
        if (conv_tc_eligible(p)) {
          void* zq;
          STORM_TRY(h->pool.get((size_t)B * hcur.H * hcur.W * s.npad * sizeof(float), &zq));
          p.out = (float*)zq;
          STORM_TRY(f.conv(p));
          LAUNCH(h, launch_tap_gather(p.out, s.npad, s.b, have_pyr ? pyr.p : nullptr, np.p, B, hcur.H, hcur.W, inch, st));
          h->pool.put(zq);
          done = true;
        }
      }
      if (!done) LAUNCH(h, launch_conv_out_small(hcur, ss, s.w, s.b, have_pyr ? pyr.p : nullptr, np.p, inch, st));
      h->pool.put(ss);
      if (have_pyr) f.release(pyr);
      pyr = np;
      have_pyr = true;
      STORM_TRY(f.tap("pyr" + std::to_string(lvl), pyr));
      mi += 2;
    }
    if (lvl != 0) {
      Act o;
      STORM_TRY(f.res_block(h->mods[mi], hcur, nullptr, tembias, o));
      STORM_TRY(tapmod(h->mods[mi], o));
      ++mi;
      f.release(hcur);
      hcur = o;
    }
  }
  f.release(hcur);
  if (!hs.empty() || mi != h->mods.size()) return fail(STORM_E_STATE, "forward: module walk out of sync");

  // ---- head ------------------------------------------------------------------------------------------------
  LAUNCH(h, launch_head(pyr.p, h->conditional() ? t : nullptr, h->P("output_layer.weight"), h->P("output_layer.bias"), out, B,
                        F * T, inch, st));
  f.release(pyr);
  if (act) h->pool.put(act);
  if (tembias_pooled) h->pool.put(tembias);
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// OUVE SDE scalars (sdes.py:200-228) and the PC sampler
// ---------------------------------------------------------------------------------------------------
static double ouve_std(const StormSamplerConfig& s, double t) {
  const double ls = log((double)s.sigma_max / (double)s.sigma_min), th = s.theta, sm = s.sigma_min;
  return sqrt(sm * sm * exp(-2 * th * t) * (exp(2 * (th + ls) * t) - 1) * ls / (th + ls));
}
static double ouve_diffusion(const StormSamplerConfig& s, double t) {
  const double ls = log((double)s.sigma_max / (double)s.sigma_min);
  return (double)s.sigma_min * pow((double)s.sigma_max / (double)s.sigma_min, t) * sqrt(2 * ls);
}
// torch.linspace(1, eps, N) in fp32 (sampling/__init__.py:58)
static float linspace_at(float start, float end, int steps, int i) {
  if (steps == 1) return start;
  const float step = (end - start) / (float)(steps - 1);
  return i < steps / 2 ? start + step * (float)i : end - step * (float)(steps - i - 1);
}

static int check_sampler(const StormSamplerConfig* sc) {
  if (!sc) return fail(STORM_E_INVALID, "sampler config is null");
  if (sc->predictor < 0 || sc->predictor > 2) return fail(STORM_E_INVALID, "Predictor with this id unknown.");
  if (sc->corrector < 0 || sc->corrector > 2) return fail(STORM_E_INVALID, "Corrector with this id unknown.");
  if (sc->N < 1 || sc->corrector_steps < 0) return fail(STORM_E_INVALID, "N must be >= 1 and corrector_steps >= 0");
  if (!(sc->sigma_min > 0) || !(sc->sigma_max > sc->sigma_min)) return fail(STORM_E_INVALID, "need 0 < sigma_min < sigma_max");
  return 0;
}

// The N-step loop.  ids_dev / seed_dev: device copies of the utterance ids and the seed (null: row index / `seed`); everything
// this function enqueues is stream work on device-resident operands, so it can run eagerly or under stream capture.
static int pc_sample_body(StormHandle* h, const float* Y, const float* cond0, const float* cond1, float* X, int B, int F, int T,
                          const StormSamplerConfig& sc, uint64_t seed, const uint64_t* seed_dev, const uint64_t* ids_dev,
                          const float* noise, int* nfe_out, cudaStream_t st) {
  NvtxRange nvtx_range("storm_b200: sampler trajectory");
  const int64_t n = (int64_t)F * T;
  const int ncplx = h->in_ch() / 2;
  if (ncplx == 3 && !cond1) return fail(STORM_E_INVALID, "6-channel score net needs two conditioning spectrograms");
  if (ncplx == 2 && cond1) return fail(STORM_E_INVALID, "4-channel score net takes one conditioning spectrogram");
  if (ncplx < 2) return fail(STORM_E_INVALID, "sampler needs a score network with (x, y) input");
  if (!cond0) cond0 = Y;
  void* p;
  float *xt, *xm, *d, *tdev;
  double* scratch = nullptr;
  UpdateCoef* coef_dev = nullptr;
  Pool* const pool0 = &h->pool;
  PoolLease lease(pool0);                                // every early return below gives the buffers back
  STORM_TRY(lease.get((size_t)B * n * 8, &p)); xt = (float*)p;
  STORM_TRY(lease.get((size_t)B * n * 8, &p)); xm = (float*)p;
  STORM_TRY(lease.get((size_t)B * n * 8, &p)); d = (float*)p;
  STORM_TRY(lease.get((size_t)B * sizeof(float), &p)); tdev = (float*)p;
  if (sc.corrector == STORM_CORR_LANGEVIN) {
    STORM_TRY(lease.get((size_t)B * 64 * 2 * sizeof(double), &p)); scratch = (double*)p;
    STORM_TRY(lease.get(sizeof(UpdateCoef), &p)); coef_dev = (UpdateCoef*)p;
  }
  uint32_t draw = 0;
  auto noise_at = [&](uint32_t dr) -> const float* { return noise ? noise + (size_t)dr * B * n * 2 : nullptr; };
  auto score = [&](float tval) -> int {   // d = dnn(cat[x, y(, cond)], t); score = -d (model.py:127-132)
    LAUNCH(h, launch_fill(tdev, tval, B, st));
    UnpackSrc src{};
    src.p[0] = xt; src.bstride[0] = n;
    src.p[1] = cond0; src.bstride[1] = n;
    if (ncplx == 3) { src.p[2] = cond1; src.bstride[2] = n; }
    return net_forward(h, src, tdev, d, B, F, T, st, &tval);
  };
  // prior: x_T = y + randn * std(T=1) (sdes.py:233-237); StoRM uses the same form around its mean
  {
    UpdateCoef c{0.f, 1.f, 0.f, (float)ouve_std(sc, 1.0)};
    LAUNCH(h, launch_sampler_update(nullptr, Y, nullptr, nullptr, xt, c, nullptr, noise_at(draw), seed, seed_dev, ids_dev, draw, B, n, st));
    ++draw;
  }
  const int N = sc.N;
  const int n_corr = sc.corrector == STORM_CORR_NONE ? 0 : sc.corrector_steps;
  bool have_mean = false;
  for (int i = 0; i < N; ++i) {
    const float t = linspace_at(1.0f, sc.t_eps, N, i);
    // ---- corrector (correctors.py:45-93)
    for (int k = 0; k < n_corr; ++k) {
      STORM_TRY(score(t));
      if (sc.corrector == STORM_CORR_ALD) {
        const double std = ouve_std(sc, t);
        const double step = (sc.snr * std) * (sc.snr * std) * 2.0;
        UpdateCoef c{1.f, 0.f, (float)(-step), (float)sqrt(step * 2.0)};
        LAUNCH(h, launch_sampler_update(xt, nullptr, d, xm, xt, c, nullptr, noise_at(draw), seed, seed_dev, ids_dev, draw, B, n, st));
      } else {
        STORM_TRY(launch_langevin_coef(d, noise_at(draw), seed, seed_dev, ids_dev, draw, sc.snr, B, n, scratch, coef_dev, st));
        h->launches += 2;
        UpdateCoef c{1.f, 0.f, 0.f, 1.f};
        LAUNCH(h, launch_sampler_update(xt, nullptr, d, xm, xt, c, coef_dev, noise_at(draw), seed, seed_dev, ids_dev, draw, B, n, st));
      }
      ++draw;
      have_mean = true;
    }
    // ---- predictor (predictors.py:41-69; sdes.py:73-90,147-157)
    if (sc.predictor == STORM_PRED_REVERSE_DIFFUSION) {
      STORM_TRY(score(t));
      const double dt = 1.0 / N;
      const double G = ouve_diffusion(sc, t) * sqrt(dt);
      const double pf = sc.probability_flow ? 0.5 : 1.0;
      UpdateCoef c{(float)(1.0 + sc.theta * dt), (float)(-sc.theta * dt), (float)(-G * G * pf), sc.probability_flow ? 0.f : (float)G};
      LAUNCH(h, launch_sampler_update(xt, Y, d, xm, xt, c, nullptr, noise_at(draw), seed, seed_dev, ids_dev, draw, B, n, st));
      ++draw;
      have_mean = true;
    } else if (sc.predictor == STORM_PRED_EULER_MARUYAMA) {
      STORM_TRY(score(t));
      const double g = ouve_diffusion(sc, t);
      UpdateCoef c{(float)(1.0 + sc.theta / N), (float)(-sc.theta / N), (float)(-g * g / N), (float)(g * sqrt(1.0 / N))};
      LAUNCH(h, launch_sampler_update(xt, Y, d, xm, xt, c, nullptr, noise_at(draw), seed, seed_dev, ids_dev, draw, B, n, st));
      ++draw;
      have_mean = true;
    } else {
      // NonePredictor returns (x, x): the mean of this step is the current state (predictors.py:72-80)
      STORM_CUDA_TRY(cudaMemcpyAsync(xm, xt, (size_t)B * n * 8, cudaMemcpyDeviceToDevice, st));
      have_mean = true;
    }
  }
  const float* result = (sc.denoise && have_mean) ? xm : xt;   // sampling/__init__.py:64
  STORM_CUDA_TRY(cudaMemcpyAsync(X, result, (size_t)B * n * 8, cudaMemcpyDeviceToDevice, st));
  if (nfe_out) *nfe_out = N * (n_corr + 1);
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// CUDA-graph replay of whole trajectories (removes the host loop of sampling/__init__.py:59)
// ---------------------------------------------------------------------------------------------------
// One graph per (B, F, T, sampler configuration, number of conditioning spectrograms): the ~6 800 launches of an N = 30
// trajectory are captured ONCE from the eager code above and replayed with cudaGraphLaunch.  Everything that changes between
// calls lives in graph-owned device buffers: mean / conditioning / result spectrograms, utterance ids and the seed (the noise
// kernels read both from device memory).  Workspace pointers are baked into the nodes, so every captured trajectory owns its
// workspace pool (the handle's pool hands buffers to whoever asks next - e.g. the `norm` vector storm_enhance keeps alive
// ACROSS the sampler call - and a replay would scribble over them).
static bool graphs_enabled(const StormHandle* h) {
  static const int env = [] { const char* e = getenv("STORM_B200_GRAPH"); return e ? atoi(e) : 1; }();
  return env != 0 && h->graph_on && !h->taps_on && !h->profile_on;
}

static void destroy_graph(StormHandle::TrajGraph& g) {
  if (g.exec) cudaGraphExecDestroy(g.exec);
  for (void* q : {(void*)g.Y, (void*)g.c0, (void*)g.c1, (void*)g.X, (void*)g.ids, (void*)g.seed}) if (q) cudaFree(q);
  g.pool.destroy();
  g = StormHandle::TrajGraph();
}

static int pc_sample(StormHandle* h, const float* Y, const float* cond0, const float* cond1, float* X, int B, int F, int T,
                     const StormSamplerConfig& sc, uint64_t seed, const uint64_t* utt_ids, const float* noise, int* nfe_out,
                     cudaStream_t st) {
  const size_t spec_bytes = (size_t)B * F * T * 8;
  if (noise || !graphs_enabled(h)) {
    // ---- eager: injected noise (parity tests), debug taps, profiling, STORM_B200_GRAPH=0
    uint64_t* ids_dev = nullptr;
    if (utt_ids) {
      void* p;
      STORM_TRY(h->pool.get((size_t)B * sizeof(uint64_t), &p)); ids_dev = (uint64_t*)p;
      STORM_CUDA_TRY(cudaMemcpyAsync(ids_dev, utt_ids, (size_t)B * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    }
    const int rc = pc_sample_body(h, Y, cond0, cond1, X, B, F, T, sc, seed, nullptr, ids_dev, noise, nfe_out, st);
    h->pool.put(ids_dev);
    return rc;
  }
  if (h->conditional()) STORM_TRY(ensure_temb_cache(h, B));      // before any graph is looked up: growing it invalidates them
  StormHandle::TrajKey key{};
  key.B = B; key.F = F; key.T = T; key.sc = sc; key.ncond = (cond0 ? 1 : 0) + (cond1 ? 1 : 0);
  StormHandle::TrajGraph* g = nullptr;
  for (auto& e : h->graphs)
    if (memcmp(&e.key, &key, sizeof(key)) == 0) { g = &e; break; }
  if (!g) {
    if (h->graphs.size() >= 8) { destroy_graph(h->graphs.front()); h->graphs.erase(h->graphs.begin()); }
    h->graphs.emplace_back();
    g = &h->graphs.back();
    g->key = key;
  }
  if (!g->exec && !g->failed) {
    // ---- first call of this shape: one eager trajectory sizes the workspace pool (no allocation may happen under capture),
    // the second call captures
    if (!g->warmed) {
      g->warmed = true;
      std::swap(h->pool, g->pool);                       // this run allocates the graph's own workspace
      uint64_t* ids_dev = nullptr;
      int rc = 0;
      if (utt_ids) {
        void* p;
        rc = h->pool.get((size_t)B * sizeof(uint64_t), &p);
        if (rc == 0) {
          ids_dev = (uint64_t*)p;
          if (cudaMemcpyAsync(ids_dev, utt_ids, (size_t)B * sizeof(uint64_t), cudaMemcpyHostToDevice, st) != cudaSuccess)
            rc = fail(STORM_E_CUDA, "utterance id upload failed");
        }
      }
      if (rc == 0) rc = pc_sample_body(h, Y, cond0, cond1, X, B, F, T, sc, seed, nullptr, ids_dev, nullptr, nfe_out, st);
      h->pool.put(ids_dev);
      std::swap(h->pool, g->pool);
      return rc;
    }
    bool ok = cudaMalloc((void**)&g->Y, spec_bytes) == cudaSuccess && cudaMalloc((void**)&g->X, spec_bytes) == cudaSuccess &&
              cudaMalloc((void**)&g->ids, (size_t)B * sizeof(uint64_t)) == cudaSuccess && cudaMalloc((void**)&g->seed, sizeof(uint64_t)) == cudaSuccess;
    if (ok && cond0) ok = cudaMalloc((void**)&g->c0, spec_bytes) == cudaSuccess;
    if (ok && cond1) ok = cudaMalloc((void**)&g->c1, spec_bytes) == cudaSuccess;
    cudaGraph_t graph = nullptr;
    if (ok) {
      h->temb_cache_valid = false;                       // the captured sequence must not depend on what ran before it
      std::swap(h->pool, g->pool);                       // nodes point into the graph's own workspace (sized by the warm run)
      const size_t pool_before = h->pool.total;
      // capture on a private stream: the caller's stream may be the legacy default stream (torch's default), which cannot
      // be captured; nothing executes during capture, the instantiated graph is launched on the caller's stream
      if (!h->cap_stream) ok = cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking) == cudaSuccess;
      cudaError_t e = cudaSuccess;
      if (ok) { e = cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeRelaxed); ok = e == cudaSuccess; }
      if (ok) {
        int nfe = 0;
        const int64_t l0 = h->launches;
        const int rc = pc_sample_body(h, g->Y, g->c0, g->c1, g->X, B, F, T, sc, 0, g->seed, g->ids, nullptr, &nfe, h->cap_stream);
        g->nfe = nfe;
        g->launches_per_replay = h->launches - l0;
        h->launches = l0;                                // captured, not executed
        e = cudaStreamEndCapture(h->cap_stream, &graph);
        ok = rc == 0 && e == cudaSuccess && graph != nullptr && h->pool.total == pool_before;
        if (!ok && getenv("STORM_B200_GRAPH_DEBUG"))
          fprintf(stderr, "storm_b200: trajectory capture failed: rc=%d (%s) cuda=%s pool %zu -> %zu\n", rc, rc ? g_last_error.c_str() : "ok",
                  cudaGetErrorString(e), pool_before, h->pool.total);
        h->temb_cache_valid = false;                     // ... and nothing after it may trust what the capture pretended to compute
      }
      std::swap(h->pool, g->pool);
      if (ok) {
        e = cudaGraphInstantiate(&g->exec, graph, 0);
        ok = e == cudaSuccess;
        if (!ok && getenv("STORM_B200_GRAPH_DEBUG")) fprintf(stderr, "storm_b200: graph instantiation failed: %s\n", cudaGetErrorString(e));
      }
      if (graph) cudaGraphDestroy(graph);
    }
    if (!ok) {
      cudaGetLastError();
      StormHandle::TrajKey k2 = g->key;
      destroy_graph(*g);
      g->key = k2; g->failed = true; g->warmed = true;
    }
  }
  if (g->failed || !g->exec) {
    uint64_t* ids_dev = nullptr;
    if (utt_ids) {
      void* p;
      STORM_TRY(h->pool.get((size_t)B * sizeof(uint64_t), &p)); ids_dev = (uint64_t*)p;
      STORM_CUDA_TRY(cudaMemcpyAsync(ids_dev, utt_ids, (size_t)B * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    }
    const int rc = pc_sample_body(h, Y, cond0, cond1, X, B, F, T, sc, seed, nullptr, ids_dev, nullptr, nfe_out, st);
    h->pool.put(ids_dev);
    return rc;
  }
  // ---- replay
  std::vector<uint64_t> ids(B);
  for (int b = 0; b < B; ++b) ids[b] = utt_ids ? utt_ids[b] : (uint64_t)b;
  STORM_CUDA_TRY(cudaMemcpyAsync(g->ids, ids.data(), (size_t)B * sizeof(uint64_t), cudaMemcpyHostToDevice, st));   // pageable source: staged
  STORM_CUDA_TRY(cudaMemcpyAsync(g->seed, &seed, sizeof(uint64_t), cudaMemcpyHostToDevice, st));                    // before the call returns
  STORM_CUDA_TRY(cudaMemcpyAsync(g->Y, Y, spec_bytes, cudaMemcpyDeviceToDevice, st));
  if (cond0) STORM_CUDA_TRY(cudaMemcpyAsync(g->c0, cond0, spec_bytes, cudaMemcpyDeviceToDevice, st));
  if (cond1) STORM_CUDA_TRY(cudaMemcpyAsync(g->c1, cond1, spec_bytes, cudaMemcpyDeviceToDevice, st));
  {
    NvtxRange r("storm_b200: trajectory (graph replay)");
    STORM_CUDA_TRY(cudaGraphLaunch(g->exec, st));
  }
  STORM_CUDA_TRY(cudaMemcpyAsync(X, g->X, spec_bytes, cudaMemcpyDeviceToDevice, st));
  h->temb_cache_valid = false;
  h->launches += g->launches_per_replay;
  ++h->graph_replays;
  if (nfe_out) *nfe_out = g->nfe;
  return 0;
}

}  // namespace storm

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

const char* storm_last_error(void) { return g_last_error.c_str(); }
const char* storm_version(void) { return "storm_b200 0.1 (sm_100a)"; }

int storm_create(const StormNetConfig* cfg, StormHandle** out) {
  if (!cfg || !out) return fail(STORM_E_INVALID, "storm_create: null argument");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(STORM_E_CUDA, "storm_create: no CUDA device available (this library has no CPU path)");
  StormHandle* h = new StormHandle();
  h->cfg = *cfg;
  cudaGetDevice(&h->device);
  int rc = conv_tc_init_device();
  if (rc == 0) rc = build_plan(h);
  if (rc != 0) { delete h; return rc; }
  void* p = nullptr;
  e = cudaMalloc(&p, (size_t)h->n_floats * sizeof(float));
  if (e != cudaSuccess) { delete h; return fail(STORM_E_NOMEM, "storm_create: weight arena allocation failed"); }
  h->raw = (float*)p;
  *out = h;
  return 0;
}

int storm_destroy(StormHandle* h) {
  if (!h) return 0;
  DeviceGuard dg(h);
  cudaDeviceSynchronize();
  free_packed(h);          // also drops the captured trajectories
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  for (auto& kv : h->taps) cudaFree(kv.second.p);
  for (auto& r : h->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto e : h->ev_free) cudaEventDestroy(e);
  if (h->temb_cache) cudaFree(h->temb_cache);
  h->pool.destroy();
  cudaFree(h->raw);
  delete h;
  return 0;
}

int64_t storm_param_floats(const StormHandle* h) { return h ? h->n_floats : 0; }
int32_t storm_param_entries(const StormHandle* h) { return h ? (int32_t)h->table.size() : 0; }
int storm_param_info(const StormHandle* h, int32_t i, char* name, int32_t cap, int64_t shape[4], int32_t* ndim, int64_t* off) {
  if (!h || i < 0 || i >= (int)h->table.size()) return fail(STORM_E_INVALID, "storm_param_info: index out of range");
  const ParamEntry& e = h->table[i];
  if (name && cap > 0) { strncpy(name, e.name.c_str(), cap - 1); name[cap - 1] = 0; }
  if (shape) for (int k = 0; k < 4; ++k) shape[k] = k < e.ndim ? e.shape[k] : 1;
  if (ndim) *ndim = e.ndim;
  if (off) *off = e.offset;
  return 0;
}

int storm_load_weights(StormHandle* h, const float* blob, int64_t n, int32_t is_device, void* stream) {
  if (!h || !blob) return fail(STORM_E_INVALID, "storm_load_weights: null argument");
  DeviceGuard dg(h);
  if (n != h->n_floats)
    return fail(STORM_E_INVALID, "storm_load_weights: blob has " + std::to_string(n) + " floats, expected " + std::to_string(h->n_floats));
  cudaStream_t st = (cudaStream_t)stream;
  STORM_CUDA_TRY(cudaMemcpyAsync(h->raw, blob, (size_t)n * sizeof(float), is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  return repack(h, st);
}
int storm_weights_device(StormHandle* h, float** dev, int64_t* n) {
  if (!h) return fail(STORM_E_INVALID, "null handle");
  if (dev) *dev = h->raw;
  if (n) *n = h->n_floats;
  return 0;
}
int storm_repack_weights(StormHandle* h, void* stream) {
  if (!h) return fail(STORM_E_INVALID, "null handle");
  DeviceGuard dg(h);
  return repack(h, (cudaStream_t)stream);
}

int64_t storm_workspace_bytes(const StormHandle* h) { return h ? (int64_t)h->pool.total : 0; }
int64_t storm_launch_count(const StormHandle* h) { return h ? h->launches : 0; }
int storm_reset_launch_count(StormHandle* h) { if (h) h->launches = 0; return 0; }

int storm_net_forward(StormHandle* h, const float* in, const float* t, float* out, int32_t B, int32_t F, int32_t T, void* stream) {
  if (!h || !in || !out) return fail(STORM_E_INVALID, "storm_net_forward: null argument");
  DeviceGuard dg(h);
  const int nc = h->in_ch() / 2;
  UnpackSrc src{};
  for (int ch = 0; ch < nc; ++ch) { src.p[ch] = in + (size_t)ch * F * T * 2; src.bstride[ch] = (long long)nc * F * T; }
  return net_forward(h, src, t, out, B, F, T, (cudaStream_t)stream);
}

int storm_reserve(StormHandle* h, int32_t B, int32_t F, int32_t T) {
  if (!h) return fail(STORM_E_INVALID, "null handle");
  DeviceGuard dg(h);
  const int nc = h->in_ch() / 2;
  void *a, *b, *c;
  STORM_TRY(h->pool.get((size_t)B * nc * F * T * 8, &a));
  STORM_TRY(h->pool.get((size_t)B * F * T * 8, &b));
  STORM_TRY(h->pool.get((size_t)B * 4, &c));
  STORM_CUDA_TRY(cudaMemset(a, 0, (size_t)B * nc * F * T * 8));
  int rc = launch_fill((float*)c, 0.5f, B, 0);
  if (rc == 0) rc = storm_net_forward(h, (const float*)a, (const float*)c, (float*)b, B, F, T, nullptr);
  cudaDeviceSynchronize();
  h->pool.put(a); h->pool.put(b); h->pool.put(c);
  return rc;
}

int storm_set_debug_taps(StormHandle* h, int32_t enable) {
  if (!h) return fail(STORM_E_INVALID, "null handle");
  h->taps_on = enable != 0;
  if (!h->taps_on) { for (auto& kv : h->taps) cudaFree(kv.second.p); h->taps.clear(); }
  return 0;
}
int storm_get_tap(StormHandle* h, const char* name, float* out, int64_t cap, int64_t shape[4], void* stream) {
  if (!h || !name) return fail(STORM_E_INVALID, "null argument");
  DeviceGuard dg(h);
  auto it = h->taps.find(name);
  if (it == h->taps.end()) return fail(STORM_E_INVALID, std::string("no tap named ") + name);
  const Act& a = it->second;
  if (shape) { shape[0] = a.B; shape[1] = a.C; shape[2] = a.H; shape[3] = a.W; }
  if (!out) return 0;
  if (cap < (int64_t)a.elems()) return fail(STORM_E_INVALID, "tap buffer too small");
  return launch_nhwc_to_nchw(a.p, out, a.B, a.H * a.W, a.C, (cudaStream_t)stream);
}

int32_t storm_pc_num_draws(const StormSamplerConfig* sc) {
  if (!sc) return 0;
  const int n_corr = sc->corrector == STORM_CORR_NONE ? 0 : sc->corrector_steps;
  return 1 + sc->N * (n_corr + (sc->predictor == STORM_PRED_NONE ? 0 : 1));
}

int storm_pc_sample(StormHandle* h, const float* Y, const float* cond0, const float* cond1, float* X, int32_t B, int32_t F, int32_t T,
                    const StormSamplerConfig* sc, uint64_t seed, const uint64_t* utt_ids, const float* noise, int32_t* nfe_out,
                    void* stream) {
  if (!h || !Y || !X) return fail(STORM_E_INVALID, "storm_pc_sample: null argument");
  DeviceGuard dg(h);
  STORM_TRY(check_sampler(sc));
  return pc_sample(h, Y, cond0, cond1, X, B, F, T, *sc, seed, utt_ids, noise, nfe_out, (cudaStream_t)stream);
}

int storm_philox_normal(uint64_t seed, uint64_t utt_id, uint32_t draw, int64_t n, float* out, void* stream) {
  if (!out || n < 0) return fail(STORM_E_INVALID, "storm_philox_normal: bad argument");
  return launch_philox_fill(seed, utt_id, draw, n, out, (cudaStream_t)stream);
}

int storm_ode_drift(StormHandle* h, const float* X, const float* Y, float t, float* drift, int32_t B, int32_t F, int32_t T,
                    const StormSamplerConfig* sc, void* stream) {
  if (!h || !X || !Y || !drift) return fail(STORM_E_INVALID, "storm_ode_drift: null argument");
  DeviceGuard dg(h);
  STORM_TRY(check_sampler(sc));
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = (int64_t)F * T;
  void* p;
  STORM_TRY(h->pool.get((size_t)B * n * 8, &p)); float* d = (float*)p;
  STORM_TRY(h->pool.get((size_t)B * 4, &p)); float* tdev = (float*)p;
  LAUNCH(h, launch_fill(tdev, t, B, st));
  UnpackSrc src{};
  src.p[0] = X; src.bstride[0] = n; src.p[1] = Y; src.bstride[1] = n;
  STORM_TRY(net_forward(h, src, tdev, d, B, F, T, st));
  const double g = ouve_diffusion(*sc, t);
  LAUNCH(h, launch_ode_drift(X, Y, d, drift, sc->theta, (float)(0.5 * g * g), (int64_t)B * n, st));
  h->pool.put(d); h->pool.put(tdev);
  return 0;
}

int storm_set_graph(StormHandle* h, int32_t enable) {
  if (!h) return fail(STORM_E_INVALID, "null handle");
  h->graph_on = enable != 0;
  return 0;
}
int64_t storm_graph_replays(const StormHandle* h) { return h ? h->graph_replays : 0; }

int storm_set_profile(StormHandle* h, int32_t enable) {
  if (!h) return fail(STORM_E_INVALID, "null handle");
  for (auto& r : h->prof) { h->ev_free.push_back(r.a); h->ev_free.push_back(r.b); }
  h->prof.clear();
  h->profile_on = enable != 0;
  return 0;
}
int storm_get_profile(StormHandle* h, int64_t launches[2], double ms[2], double flops[2]) {
  if (!h) return fail(STORM_E_INVALID, "null handle");
  DeviceGuard dg(h);
  STORM_CUDA_TRY(cudaDeviceSynchronize());
  for (int i = 0; i < 2; ++i) { launches[i] = 0; ms[i] = 0; flops[i] = 0; }
  FILE* dump = nullptr;
  if (const char* path = getenv("STORM_B200_PROFILE_DUMP")) dump = fopen(path, "w");   // per-launch list (tools/layer_sustained.py)
  for (auto& r : h->prof) {
    float t = 0.f;
    STORM_CUDA_TRY(cudaEventElapsedTime(&t, r.a, r.b));
    launches[r.eng] += 1; ms[r.eng] += t; flops[r.eng] += r.flops;
    if (dump) fprintf(dump, "%d,%.6f,%.0f\n", r.eng, t, r.flops);
  }
  if (dump) fclose(dump);
  return 0;
}

int storm_debug_conv(int32_t engine, const float* x0, const float* x1, int32_t C0, int32_t C1, const float* ss, int32_t silu,
                     int32_t taps, const float* w, const float* x2, int32_t C2, const float* w2, int32_t Cout, const float* bias,
                     const float* res, float out_scale, float* out, int32_t B, int32_t H, int32_t W, void* stream) {
  return storm_debug_conv_ex(engine, x0, x1, C0, C1, ss, silu, taps, w, x2, C2, w2, Cout, bias, res, nullptr, out_scale, out, B, H, W, 1,
                             nullptr, nullptr, stream);
}

int storm_debug_conv_timed(int32_t engine, const float* x0, const float* x1, int32_t C0, int32_t C1, const float* ss, int32_t silu,
                           int32_t taps, const float* w, const float* x2, int32_t C2, const float* w2, int32_t Cout, const float* bias,
                           const float* res, float out_scale, float* out, int32_t B, int32_t H, int32_t W, long long* dbg_dev,
                           float* ms_out, void* stream) {
  return storm_debug_conv_ex(engine, x0, x1, C0, C1, ss, silu, taps, w, x2, C2, w2, Cout, bias, res, nullptr, out_scale, out, B, H, W, 1,
                             dbg_dev, ms_out, stream);
}

int storm_debug_conv_ex(int32_t engine, const float* x0, const float* x1, int32_t C0, int32_t C1, const float* ss, int32_t silu,
                        int32_t taps, const float* w, const float* x2, int32_t C2, const float* w2, int32_t Cout, const float* bias,
                        const float* res, const float* res_up, float out_scale, float* out, int32_t B, int32_t H, int32_t W,
                        int32_t max_prec, long long* dbg_dev, float* ms_out, void* stream) {
  if (!x0 || !w || !out) return fail(STORM_E_INVALID, "storm_debug_conv: null argument");
  if (res_up && engine != 0) return fail(STORM_E_INVALID, "storm_debug_conv: the half-resolution residual needs the tcgen05 engine");
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<void*> owned;
  ConvParams p{};
  p.nseg = x2 ? 2 : 1;
  p.seg[0].src0 = x0; p.seg[0].src1 = x1; p.seg[0].C0 = C0; p.seg[0].C1 = x1 ? C1 : 0; p.seg[0].ss = ss; p.seg[0].silu = silu;
  p.seg[0].taps = taps;
  if (x2) { p.seg[1].src0 = x2; p.seg[1].C0 = C2; p.seg[1].taps = 1; }
  p.B = B; p.H = H; p.W = W; p.Cout = Cout; p.bias = bias; p.res = res; p.res_up = res_up; p.out_scale = out_scale; p.out = out;
  int rc = 0;
  const int Cin = C0 + (x1 ? C1 : 0);
  if (engine == 0) {
    TcWeights t0, t1;
    // one operand split and one weight scale for the whole launch: the 32-channel mode only if every segment allows it
    const int mp = (max_prec >= 1 && C0 % 32 == 0 && (!x1 || C1 % 32 == 0) && (!w2 || C2 % 32 == 0)) ? 1 : 0;
    float a0 = 0.f, a1 = 0.f;
    rc = tc_weight_absmax(w, (size_t)Cout * Cin * taps, &a0, st);
    if (rc == 0 && w2) rc = tc_weight_absmax(w2, (size_t)Cout * C2, &a1, st);
    const float amax = std::max(a0, a1);
    if (rc == 0) rc = tc_pack_weights(w, Cout, Cin, taps, &t0, &owned, st, mp, amax);
    if (rc == 0 && w2) rc = tc_pack_weights(w2, Cout, C2, 1, &t1, &owned, st, mp, amax);
    set_wtc(p.seg[0], t0);
    if (x2) set_wtc(p.seg[1], t1);
    if (rc == 0 && !conv_tc_eligible(p)) rc = fail(STORM_E_INVALID, "storm_debug_conv: shape not eligible for the tcgen05 kernel");
    if (rc == 0) rc = launch_conv_tc(p, st);
    if (rc == 0 && ms_out) {     // timed repeat (warm weights), optionally with the CTA-0 cycle counters
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      p.dbg = dbg_dev;
      cudaEventRecord(e0, st);
      rc = launch_conv_tc(p, st);
      cudaEventRecord(e1, st);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(ms_out, e0, e1);
      cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
  } else {
    float *s0 = nullptr, *s1 = nullptr;
    if (cudaMalloc((void**)&s0, (size_t)taps * Cin * Cout * 4) != cudaSuccess) return fail(STORM_E_NOMEM, "debug conv alloc");
    owned.push_back(s0);
    rc = launch_pack_conv_simt(w, s0, Cout, Cin, taps, st);
    if (rc == 0 && x2) {
      if (cudaMalloc((void**)&s1, (size_t)C2 * Cout * 4) != cudaSuccess) return fail(STORM_E_NOMEM, "debug conv alloc");
      owned.push_back(s1);
      rc = launch_pack_conv_simt(w2, s1, Cout, C2, 1, st);
    }
    p.seg[0].w = s0; p.seg[1].w = s1;
    if (rc == 0) rc = launch_conv_simt(p, st);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  for (void* q : owned) cudaFree(q);
  if (rc == 0 && e != cudaSuccess) return fail(STORM_E_CUDA, std::string("storm_debug_conv: ") + cudaGetErrorString(e));
  return rc;
}

int storm_debug_resample(const float* x, const float* ss, float* h_out, float* x_out, int32_t B, int32_t H, int32_t W, int32_t C,
                         int32_t mode, void* stream) {
  if (!x || (!h_out && !x_out) || (mode != 1 && mode != 2)) return fail(STORM_E_INVALID, "storm_debug_resample: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  Act a{};
  a.p = const_cast<float*>(x); a.B = B; a.H = H; a.W = W; a.C = C;
  int rc = launch_resample(a, ss, h_out, x_out, mode, st);
  cudaError_t e = cudaStreamSynchronize(st);
  if (rc == 0 && e != cudaSuccess) return fail(STORM_E_CUDA, std::string("storm_debug_resample: ") + cudaGetErrorString(e));
  return rc;
}

int storm_debug_attention(int32_t engine, const float* q, const float* k, const float* v, float* o, int32_t B, int32_t N, int32_t C,
                          void* stream) {
  if (!q || !k || !v || !o) return fail(STORM_E_INVALID, "storm_debug_attention: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (engine == 0) {
    if (!attention_tc_eligible(N, C)) return fail(STORM_E_INVALID, "storm_debug_attention: shape not eligible for the tcgen05 kernel");
    void* scratch = nullptr;
    if (cudaMalloc(&scratch, attention_tc_scratch_bytes(B, N, C)) != cudaSuccess) return fail(STORM_E_NOMEM, "storm_debug_attention: scratch");
    rc = launch_attention_tc(q, k, v, o, B, N, C, scratch, st);
    cudaStreamSynchronize(st);
    cudaFree(scratch);
  } else {
    rc = launch_attention_simt(q, k, v, o, B, N, C, st);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  if (rc == 0 && e != cudaSuccess) return fail(STORM_E_CUDA, std::string("storm_debug_attention: ") + cudaGetErrorString(e));
  return rc;
}

int32_t storm_num_frames(int32_t L, const StormSpecConfig* sp) { return sp && sp->hop_length > 0 ? 1 + L / sp->hop_length : 0; }
int32_t storm_padded_frames(int32_t frames) { return (frames + 63) / 64 * 64; }

int storm_wave_norm_ragged(const float* wav, int32_t B, int32_t L, const int32_t* lengths, float* norm, void* stream) {
  if (!wav || !norm) return fail(STORM_E_INVALID, "storm_wave_norm: null argument");
  return launch_wave_norm(wav, B, L, lengths, norm, (cudaStream_t)stream);
}
int storm_wave_norm(const float* wav, int32_t B, int32_t L, float* norm, void* stream) {
  return storm_wave_norm_ragged(wav, B, L, nullptr, norm, stream);
}
int storm_stft_ragged(const float* wav, const float* norm, int32_t B, int32_t L, const int32_t* lengths, float* spec, int32_t T_pad,
                      const StormSpecConfig* sp, void* stream) {
  if (!wav || !spec || !sp) return fail(STORM_E_INVALID, "storm_stft: null argument");
  return launch_stft(wav, norm, B, L, lengths, spec, T_pad, sp->n_fft, sp->hop_length, sp->sqrt_window, sp->spec_factor,
                     sp->spec_abs_exponent, (cudaStream_t)stream);
}
int storm_stft(const float* wav, const float* norm, int32_t B, int32_t L, float* spec, int32_t T_pad, const StormSpecConfig* sp,
               void* stream) {
  return storm_stft_ragged(wav, norm, B, L, nullptr, spec, T_pad, sp, stream);
}
int storm_istft_ragged(StormHandle* h, const float* spec, const float* norm, int32_t B, int32_t T_pad, float* wav, int32_t L,
                       const int32_t* lengths, const StormSpecConfig* sp, void* stream) {
  if (!h || !spec || !wav || !sp) return fail(STORM_E_INVALID, "storm_istft: null argument");
  DeviceGuard dg(h);
  void* p;
  STORM_TRY(h->pool.get((size_t)B * T_pad * sp->n_fft * sizeof(float), &p));
  int rc = launch_istft(spec, norm, B, T_pad, (float*)p, wav, L, lengths, sp->n_fft, sp->hop_length, sp->sqrt_window, sp->spec_factor,
                        sp->spec_abs_exponent, (cudaStream_t)stream);
  h->launches += 2;
  h->pool.put(p);
  return rc;
}
int storm_istft(StormHandle* h, const float* spec, const float* norm, int32_t B, int32_t T_pad, float* wav, int32_t L,
                const StormSpecConfig* sp, void* stream) {
  return storm_istft_ragged(h, spec, norm, B, T_pad, wav, L, nullptr, sp, stream);
}

int storm_si_sdr(const float* s, const float* s_hat, int32_t B, int32_t L, const int32_t* lengths, double eps, float* out_db,
                 void* stream) {
  if (!s || !s_hat || !out_db || B < 1 || L < 1) return fail(STORM_E_INVALID, "storm_si_sdr: bad argument");
  return launch_si_sdr(s, s_hat, B, L, lengths, eps, out_db, (cudaStream_t)stream);
}

int storm_enhance_ragged(StormHandle* h, const float* wav_in, float* wav_out, int32_t B, int32_t L, const int32_t* lengths,
                         const StormSamplerConfig* sc, const StormSpecConfig* sp, uint64_t seed, const uint64_t* utt_ids,
                         int32_t* nfe_out, void* stream) {
  if (!h || !wav_in || !wav_out || !sp) return fail(STORM_E_INVALID, "storm_enhance: null argument");
  DeviceGuard dg(h);
  STORM_TRY(check_sampler(sc));
  cudaStream_t st = (cudaStream_t)stream;
  const int F = sp->n_fft / 2 + 1;
  const int T_pad = storm_padded_frames(storm_num_frames(L, sp));
  void* p;
  PoolLease lease(&h->pool);
  STORM_TRY(lease.get((size_t)B * 4, &p)); float* norm = (float*)p;
  STORM_TRY(lease.get((size_t)B * F * T_pad * 8, &p)); float* Y = (float*)p;
  STORM_TRY(lease.get((size_t)B * F * T_pad * 8, &p)); float* X = (float*)p;
  NvtxRange nvtx_range("storm_b200: enhance");
  {
    NvtxRange r("storm_b200: normalise + STFT");
    LAUNCH(h, launch_wave_norm(wav_in, B, L, lengths, norm, st));
    LAUNCH(h, launch_stft(wav_in, norm, B, L, lengths, Y, T_pad, sp->n_fft, sp->hop_length, sp->sqrt_window, sp->spec_factor,
                          sp->spec_abs_exponent, st));
  }
  STORM_TRY(pc_sample(h, Y, nullptr, nullptr, X, B, F, T_pad, *sc, seed, utt_ids, nullptr, nfe_out, st));
  {
    NvtxRange r("storm_b200: iSTFT");
    STORM_TRY(storm_istft_ragged(h, X, norm, B, T_pad, wav_out, L, lengths, sp, stream));
  }
  return 0;
}
int storm_enhance(StormHandle* h, const float* wav_in, float* wav_out, int32_t B, int32_t L, const StormSamplerConfig* sc,
                  const StormSpecConfig* sp, uint64_t seed, const uint64_t* utt_ids, int32_t* nfe_out, void* stream) {
  return storm_enhance_ragged(h, wav_in, wav_out, B, L, nullptr, sc, sp, seed, utt_ids, nfe_out, stream);
}

}  // extern "C"
