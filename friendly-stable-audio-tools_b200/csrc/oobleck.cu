// Oobleck VAE encoder / decoder (reference models/autoencoders.py:45-194) on the tcgen05
// implicit-GEMM convolution of gemm.cuh.
//
// Data layout: activations are channels-last [B, L, C].  Every tensor-core convolution
// reads a 16-bit, already Snake-activated copy of its input (written by the producer's
// epilogue with the consumer's alpha/beta) and, where a ResidualUnit skip needs it, an fp32
// copy of the raw value.  A dilated k=7 convolution is 7 shifted GEMMs accumulated in TMEM
// (TMA zero-fills the padding); a transposed convolution (k = 2s, stride s) is a 2-tap GEMM
// over N = s*Cout columns; a strided convolution (k = 2s, stride s) is a 2s-tap GEMM whose
// taps address the input as (phase, row) through a 4-D tensor map.  The encoder's first
// convolution (2 -> 128 channels) is bandwidth-bound and stays on CUDA cores; the decoder's last
// one (128 -> 2) runs through the same GEMM with a mostly empty N tile.
// Weight-norm (w = g * v / ||v||, torch.nn.utils.weight_norm via dac.nn.layers) is folded
// once at load time.
#include <algorithm>
#include <cmath>
#include <tuple>
#include <map>
#include <string>
#include <vector>

#include "../../include/satb200.h"
#include "common.cuh"
#include "resunit.cuh"
#include "conv_halo.cuh"
#include "kernels.h"

namespace satb {

// ---------------------------------------------------------------- small kernels
namespace {

// scale[i] = g[i] / || v[i, :, :] ||   (one block per dim-0 slice)
__global__ void wn_scale_kernel(const float* __restrict__ g, const float* __restrict__ v, float* __restrict__ scale,
                                int slice) {
  __shared__ float red[32];
  const int i = blockIdx.x;
  const float* vs = v + static_cast<size_t>(i) * slice;
  float s = 0.f;
  for (int j = threadIdx.x; j < slice; j += blockDim.x) s += vs[j] * vs[j];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) scale[i] = g[i] / sqrtf(s);
  }
}

// GEMM weight layout [tap][n][k] (16-bit) from a weight-normed conv weight.
//   mode 0: Conv1d  v [cout, cin, kk]      -> dst[t][co][ci]            = v[co, ci, t] * scale[co]
//   mode 1: ConvT1d v [cin, cout, 2*up]    -> dst[tap][ph*cout+co][ci]  = v[ci, co, ph + tap*up] * scale[ci]
template <bool BF16>
__global__ void conv_w_prep_kernel(const float* __restrict__ v, const float* __restrict__ scale,
                                   uint16_t* __restrict__ dst, uint16_t* __restrict__ dst_lo, int mode, int cin, int cout,
                                   int kk, int up, size_t total) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(i % cin);
    const size_t rest = i / cin;
    float w;
    if (mode == 0) {
      const int co = static_cast<int>(rest % cout);
      const int t = static_cast<int>(rest / cout);
      w = v[(static_cast<size_t>(co) * cin + ci) * kk + t] * scale[co];
    } else {
      const int n = static_cast<int>(rest % (static_cast<size_t>(up) * cout));
      const int tap = static_cast<int>(rest / (static_cast<size_t>(up) * cout));
      const int ph = n / cout, co = n - ph * cout;
      w = v[(static_cast<size_t>(ci) * cout + co) * kk + ph + tap * up] * scale[ci];
    }
    typename Op16<BF16>::T h = Op16<BF16>::from_float(w);
    dst[i] = *reinterpret_cast<uint16_t*>(&h);
    if (dst_lo) {   // split-operand mode: the part of w the 16-bit value lost
      typename Op16<BF16>::T l = Op16<BF16>::from_float(w - Op16<BF16>::to_float(h));
      dst_lo[i] = *reinterpret_cast<uint16_t*>(&l);
    }
  }
}

__global__ void fold_small_kernel(const float* __restrict__ v, const float* __restrict__ scale, float* __restrict__ dst,
                                  int slice, size_t total) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < total) dst[i] = v[i] * scale[i / slice];
}

__global__ void snake_prep_kernel(const float* __restrict__ alpha, const float* __restrict__ beta,
                                  float* __restrict__ a, float* __restrict__ ib, int c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c) {
    a[i] = expf(alpha[i]);
    ib[i] = 1.0f / (expf(beta[i]) + 0.000000001f);
  }
}

// NCL fp32 -> channels-last 16-bit (no activation): the decoder's latent input.
template <bool BF16>
__global__ void __launch_bounds__(256) ncl_to_nlc16_kernel(const float* __restrict__ x, uint16_t* __restrict__ y,
                                                           uint16_t* __restrict__ y_lo, int C, int L) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* xs = x + static_cast<size_t>(b) * C * L;
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, l = l0 + tx;
    tile[j][tx] = (c < C && l < L) ? xs[static_cast<size_t>(c) * L + l] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int l = l0 + j, c = c0 + tx;
    if (l < L && c < C) {
      typename Op16<BF16>::T h = Op16<BF16>::from_float(tile[tx][j]);
      y[(static_cast<size_t>(b) * L + l) * C + c] = *reinterpret_cast<uint16_t*>(&h);
      if (y_lo) {
        typename Op16<BF16>::T lo = Op16<BF16>::from_float(tile[tx][j] - Op16<BF16>::to_float(h));
        y_lo[(static_cast<size_t>(b) * L + l) * C + c] = *reinterpret_cast<uint16_t*>(&lo);
      }
    }
  }
}

// Encoder input convolution (audio NCL fp32, Cin = 1 or 2 -> C channels, k taps, pad k/2):
// bandwidth-bound; one thread per output channel, a tile of positions per block.
// Writes raw fp32 and the Snake-activated 16-bit copy, channels-last.
template <bool BF16>
__global__ void __launch_bounds__(256) conv_in_kernel(const float* __restrict__ audio, const float* __restrict__ w,
                                                      const float* __restrict__ bias, const float* __restrict__ sn_a,
                                                      const float* __restrict__ sn_ib, void* __restrict__ raw,
                                                      uint16_t* __restrict__ s16, uint16_t* __restrict__ s16_lo, int Cin,
                                                      int C, int64_t T, int kk, int raw16) {
  constexpr int kTile = 64;
  extern __shared__ float sm_in[];  // [Cin][kTile + kk - 1]
  const int b = blockIdx.y;
  const int64_t l0 = static_cast<int64_t>(blockIdx.x) * kTile;
  const int halo = kk / 2, span = kTile + kk - 1;
  for (int i = threadIdx.x; i < Cin * span; i += blockDim.x) {
    const int ci = i / span, j = i - ci * span;
    const int64_t l = l0 + j - halo;
    sm_in[i] = (l >= 0 && l < T) ? audio[(static_cast<size_t>(b) * Cin + ci) * T + l] : 0.f;
  }
  __syncthreads();
  for (int co = threadIdx.x; co < C; co += blockDim.x) {
    float wr[16];  // Cin * kk <= 16 (stereo k7 = 14)
    for (int i = 0; i < Cin * kk; ++i) wr[i] = w[static_cast<size_t>(co) * Cin * kk + i];
    const float bb = bias ? bias[co] : 0.f;
    const float a = sn_a[co], ib = sn_ib[co];
    for (int j = 0; j < kTile; ++j) {
      const int64_t l = l0 + j;
      if (l >= T) break;
      float acc = bb;
      for (int ci = 0; ci < Cin; ++ci)
        for (int t = 0; t < kk; ++t) acc = fmaf(wr[ci * kk + t], sm_in[ci * span + j + t], acc);
      const size_t o = (static_cast<size_t>(b) * T + l) * C + co;
      if (raw16) {
        typename Op16<BF16>::T r = Op16<BF16>::from_float(acc);
        static_cast<uint16_t*>(raw)[o] = *reinterpret_cast<uint16_t*>(&r);
      } else {
        static_cast<float*>(raw)[o] = acc;
      }
      const float act = snake_fast(acc, a, ib);
      typename Op16<BF16>::T h = Op16<BF16>::from_float(act);
      s16[o] = *reinterpret_cast<uint16_t*>(&h);
      if (s16_lo) {
        typename Op16<BF16>::T lo = Op16<BF16>::from_float(act - Op16<BF16>::to_float(h));
        s16_lo[o] = *reinterpret_cast<uint16_t*>(&lo);
      }
    }
  }
}

}  // namespace

struct ConvW {
  int cin = 0, cout = 0, k = 0;
  bool transposed = false, small = false, has_bias = true;
  std::string pfx;
  uint16_t* w16 = nullptr;  // [taps][n][cin]
  float* w32 = nullptr;     // small convs: folded [cout][cin][k]
  float* bias = nullptr;
};
struct SnakeW {
  int c = 0;
  std::string pfx;
  float *a = nullptr, *ib = nullptr;
};

}  // namespace satb

using namespace satb;

struct SatbOobleck {
  SatbOobleckConfig cfg;
  bool bf16 = false;
  int raw16 = 0;                   // 1: the raw skip stream is carried in the 16-bit operand type (fp16 mode), 0: fp32
  bool split3 = false;             // operand_dtype 2 ("fp16x3"): every product as (hi, hi) + (lo, hi) + (hi, lo), see GemmShape
  size_t lo_off = 0;               // bytes from a 16-bit activation buffer to its "lo" half (split3)
  std::vector<int> chans;          // c_mults[i] * channels, i = 0..n (c_mults prepended with 1)
  std::map<std::string, std::pair<float*, long long>> raw;   // state-dict entries (device fp32)
  std::vector<void*> owned;
  std::map<std::string, ConvW> convs;
  std::map<std::string, SnakeW> snakes;
  bool finalized = false;
  // workspace
  void *buf_raw = nullptr, *buf_a = nullptr, *buf_b = nullptr;
  size_t cap_raw = 0, cap_a = 0, cap_b = 0;
  std::map<std::tuple<const void*, int, int, int, int64_t, int64_t, int>, CUtensorMap> tmaps;

  int alloc_bytes(void** p, size_t bytes) {
    cudaError_t e = cudaMalloc(p, bytes < 256 ? 256 : bytes);
    if (e != cudaSuccess) {
      set_last_error(std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
      return -2;
    }
    owned.push_back(*p);
    return 0;
  }
  int ensure(void** buf, size_t* cap, size_t need) {
    if (need <= *cap) return 0;
    if (*buf) cudaFree(*buf);
    *buf = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc(buf, need);
    if (e != cudaSuccess) {
      set_last_error(std::string("cudaMalloc failed: ") + cudaGetErrorString(e) + " (" + std::to_string(need) + " B)");
      return -2;
    }
    *cap = need;
    tmaps.clear();
    return 0;
  }
};

namespace {

int get_raw(SatbOobleck* h, const std::string& name, long long expect, float** out) {
  auto it = h->raw.find(name);
  if (it == h->raw.end()) {
    set_last_error("oobleck: missing weight " + name);
    return -4;
  }
  if (expect >= 0 && it->second.second != expect) {
    set_last_error("oobleck: bad size for " + name + ": got " + std::to_string(it->second.second) + ", expected " +
                   std::to_string(expect));
    return -4;
  }
  *out = it->second.first;
  return 0;
}

int prep_conv(SatbOobleck* h, const std::string& pfx, int cin, int cout, int k, bool transposed, bool small,
              bool has_bias, int up, cudaStream_t st) {
  ConvW c;
  c.cin = cin; c.cout = cout; c.k = k; c.transposed = transposed; c.small = small; c.has_bias = has_bias; c.pfx = pfx;
  float *g, *v;
  const int d0 = transposed ? cin : cout;
  const int slice = (transposed ? cout : cin) * k;
  SATB_PROPAGATE(get_raw(h, pfx + "weight_g", d0, &g));
  SATB_PROPAGATE(get_raw(h, pfx + "weight_v", static_cast<long long>(d0) * slice, &v));
  if (has_bias) SATB_PROPAGATE(get_raw(h, pfx + "bias", cout, &c.bias));
  float* scale;
  SATB_PROPAGATE(h->alloc_bytes(reinterpret_cast<void**>(&scale), static_cast<size_t>(d0) * 4));
  wn_scale_kernel<<<d0, 256, 0, st>>>(g, v, scale, slice);
  count_launch();
  const size_t total = static_cast<size_t>(cin) * cout * k;
  if (small) {
    SATB_PROPAGATE(h->alloc_bytes(reinterpret_cast<void**>(&c.w32), total * 4));
    fold_small_kernel<<<static_cast<int>(ceil_div64(total, 256)), 256, 0, st>>>(v, scale, c.w32, slice, total);
  } else {
    const size_t parts = h->split3 ? 2 : 1;      // [hi block | lo block]
    SATB_PROPAGATE(h->alloc_bytes(reinterpret_cast<void**>(&c.w16), parts * total * 2 + 256 * 128));  // slack for box overreach
    SATB_CHECK_CUDA(cudaMemsetAsync(c.w16, 0, parts * total * 2 + 256 * 128, st));
    uint16_t* w_lo = h->split3 ? c.w16 + total : nullptr;
    int grid = static_cast<int>(ceil_div64(total, 256));
    if (grid > 8192) grid = 8192;
    if (h->bf16)
      conv_w_prep_kernel<true><<<grid, 256, 0, st>>>(v, scale, c.w16, w_lo, transposed ? 1 : 0, cin, cout, k, up, total);
    else
      conv_w_prep_kernel<false><<<grid, 256, 0, st>>>(v, scale, c.w16, w_lo, transposed ? 1 : 0, cin, cout, k, up, total);
  }
  count_launch();
  SATB_CHECK_CUDA(cudaGetLastError());
  h->convs[pfx] = c;
  return 0;
}

int prep_snake(SatbOobleck* h, const std::string& pfx, int c, cudaStream_t st) {
  SnakeW s;
  s.c = c; s.pfx = pfx;
  float *al, *be;
  SATB_PROPAGATE(get_raw(h, pfx + "alpha", c, &al));
  SATB_PROPAGATE(get_raw(h, pfx + "beta", c, &be));
  SATB_PROPAGATE(h->alloc_bytes(reinterpret_cast<void**>(&s.a), static_cast<size_t>(c) * 4));
  SATB_PROPAGATE(h->alloc_bytes(reinterpret_cast<void**>(&s.ib), static_cast<size_t>(c) * 4));
  snake_prep_kernel<<<ceil_div(c, 256), 256, 0, st>>>(al, be, s.a, s.ib, c);
  count_launch();
  SATB_CHECK_CUDA(cudaGetLastError());
  h->snakes[pfx] = s;
  return 0;
}

int get_tmap_a(SatbOobleck* h, const void* in16, int cin, int a_rows, int B, int L_in, int a_stride, const CUtensorMap** out,
               int box_rows = kBlockM) {
  auto key = std::make_tuple(in16, cin, a_rows, B, static_cast<int64_t>(cin), static_cast<int64_t>(L_in) * cin,
                             a_stride | (box_rows << 8));
  auto it = h->tmaps.find(key);
  if (it == h->tmaps.end()) {
    CUtensorMap m;
    SATB_PROPAGATE(make_tmap_a(&m, in16, cin, a_rows, B, cin, static_cast<int64_t>(L_in) * cin, a_stride, box_rows));
    it = h->tmaps.emplace(key, m).first;
  }
  *out = &it->second;
  return 0;
}

int get_tmap_b(SatbOobleck* h, const ConvW& cw, int b_rows, int box, const CUtensorMap** out) {
  auto key = std::make_tuple(static_cast<const void*>(cw.w16), cw.cin, b_rows, -1, static_cast<int64_t>(cw.cin), int64_t(0), box);
  auto it = h->tmaps.find(key);
  if (it == h->tmaps.end()) {
    CUtensorMap m;
    SATB_PROPAGATE(make_tmap_b(&m, cw.w16, cw.cin, b_rows, cw.cin, box));
    it = h->tmaps.emplace(key, m).first;
  }
  *out = &it->second;
  return 0;
}

// One tensor-core convolution.  in16: [B, L_in, cin] 16-bit.  Output positions per item L_out.
//   kind 0: conv k taps, dilation dil, "same" padding           (L_out = L_in)
//   kind 1: transposed conv k = 2*up, stride up, pad ceil(up/2)  (L_out = L_in * up)
//   kind 2: strided conv k = 2*st, stride st, pad ceil(st/2)     (L_out = L_in / st)
template <class Epi, bool BF16>
int run_conv_gemm(SatbOobleck* h, const ConvW& cw, const void* in16, int B, int L_in, int kind, int dil, int factor,
                  const typename Epi::Params& ep, cudaStream_t st) {
  // the default 16-bit decode / encode runs the lean-epilogue instantiation of the GEMM kernels (EpiConv<.., MASKED>)
  if constexpr (std::is_same<Epi, EpiConv<BF16, false>>::value) {
    if (Epi::fast_flags(ep) && conv_epi_masked())
      return run_conv_gemm<EpiConv<BF16, true>, BF16>(h, cw, in16, B, L_in, kind, dil, factor, ep, st);
  }
  GemmShape s;
  s.batches = B;
  s.b_static = 1;   // folded weight-norm weights, written at finalize time
  s.K = cw.cin;
  int a_stride = 1, a_rows = L_in;
  if (kind == 0) {
    s.L = L_in; s.N = cw.cout; s.n_taps = cw.k; s.tap_base = -(cw.k / 2) * dil; s.tap_step = dil; s.b_tap_rows = cw.cout;
    s.stride = 1;
  } else if (kind == 1) {
    s.L = L_in + 1; s.N = factor * cw.cout; s.n_taps = 2; s.tap_base = 0; s.tap_step = -1; s.b_tap_rows = factor * cw.cout;
    s.stride = 1;
  } else {
    SATB_REQUIRE(L_in % factor == 0, "strided conv: length must be a multiple of the stride");
    s.L = L_in / factor; s.N = cw.cout; s.n_taps = cw.k; s.tap_base = -((factor + 1) / 2); s.tap_step = 1;
    s.b_tap_rows = cw.cout; s.stride = factor;
    a_stride = factor; a_rows = L_in / factor;
  }
  const CUtensorMap* tap;
  SATB_PROPAGATE(get_tmap_a(h, in16, cw.cin, a_rows, B, L_in, a_stride, &tap));
  const CUtensorMap& ta = *tap;
  const CUtensorMap* ta2 = nullptr;
  int b_rows = s.n_taps * s.b_tap_rows;
  if (h->split3) {
    SATB_PROPAGATE(get_tmap_a(h, static_cast<const char*>(in16) + h->lo_off, cw.cin, a_rows, B, L_in, a_stride, &ta2));
    s.n_parts = 3;
    s.b_part_rows = b_rows;       // the lo weight block follows the hi block
    b_rows *= 2;
  }
  auto get_b = [&](int box, const CUtensorMap** out) -> int { return get_tmap_b(h, cw, b_rows, box, out); };
  const CUtensorMap* tb;
  // CTA pairs pay off where the mainloop is long (k7 / strided convolutions: >= 3 taps).  The 1x1 convolutions and
  // the transposed convolutions (2 taps) are epilogue-bound, and there the per-tile hand-offs between the two CTAs
  // cost more than the halved B traffic saves: single CTAs measured 495 vs 578 us (stride 2, 128 channels), 159 vs
  // 185 us (stride 4, 512 -> 256), 33 vs 40 and 88 vs 107 us (1x1 + skip at 1024 / 512 channels).
  const bool pair = gemm_use_2cta() && s.L >= 512 && s.n_taps >= 3;
  if (s.N >= 256 && pair) {
    SATB_PROPAGATE(get_b(128, &tb));   // CTA pair: each CTA loads half of the 256-wide B tile
    return launch_gemm_2cta<Epi, 256, BF16>(ta, *tb, s, ep, st, ta2);
  } else if (s.N >= 256) {
    SATB_PROPAGATE(get_b(256, &tb));
    return launch_gemm<Epi, 256, BF16>(ta, *tb, s, ep, st, ta2);
  } else if (s.N == 128 && pair) {
    SATB_PROPAGATE(get_b(64, &tb));    // CTA pair on 256 x 128 tiles: halves the B traffic of the 128-channel layers
    return launch_gemm_2cta<Epi, 128, BF16>(ta, *tb, s, ep, st, ta2);
  } else if (s.N > 64) {
    SATB_PROPAGATE(get_b(128, &tb));
    return launch_gemm<Epi, 128, BF16>(ta, *tb, s, ep, st, ta2);
  }
  SATB_PROPAGATE(get_b(64, &tb));
  return launch_gemm<Epi, 64, BF16>(ta, *tb, s, ep, st, ta2);
}

// ResidualUnit (models/autoencoders.py:45-68).  In: snake1(x) as 16-bit in sA, x as fp32 in raw.
// Out: y = x + conv1(snake2(conv7(.))) as fp32 in raw (if keep_raw) and snake_next(y) as 16-bit in sA;
// sT is scratch (the two pointers are swapped when the fused kernel wrote its output there).
template <bool BF16>
int residual_unit(SatbOobleck* h, const std::string& pfx, int C, int B, int L, int dil, void* raw, void*& sA, void*& sT,
                  const SnakeW* next_snake, bool keep_raw, cudaStream_t st) {
  const ConvW& c7 = h->convs.at(pfx + "layers.1.");
  const ConvW& c1 = h->convs.at(pfx + "layers.3.");
  const SnakeW& s2 = h->snakes.at(pfx + "layers.2.");
  typedef EpiConv<BF16> E;
  typename E::Params e1{c1.bias, raw, keep_raw ? raw : nullptr, sA, next_snake ? next_snake->a : nullptr,
                        next_snake ? next_snake->ib : nullptr, C, L, 1, 0, h->split3 ? static_cast<char*>(sA) + h->lo_off : nullptr, h->raw16};
  if (C == ResUnitCfg::kC && c7.k == ResUnitCfg::kTaps && dil <= ResUnitCfg::kMaxDil && L >= 512 && gemm_use_2cta() &&
      resunit_use_fused() && !h->split3) {
    // one kernel: conv7 -> snake2 -> conv1 -> + skip; reads sA (with a halo), so it must write elsewhere
    const CUtensorMap *ta, *tb7, *tb1;
    SATB_PROPAGATE(get_tmap_a(h, sA, C, L, B, L, 1, &ta, ResUnitCfg::halo_rows(dil)));   // one halo box per k-block
    SATB_PROPAGATE(get_tmap_b(h, c7, c7.k * C, 64, &tb7));
    SATB_PROPAGATE(get_tmap_b(h, c1, C, 64, &tb1));
    e1.s16_out = sT;
    ResUnitParams<BF16> rp{c7.bias, s2.a, s2.ib, e1};
    ResUnitShape rs{L, B, dil};
    SATB_PROPAGATE(launch_resunit<BF16>(*ta, *tb7, *tb1, rs, rp, st));
    std::swap(sA, sT);
    return 0;
  }
  if (C == ResUnit256Cfg::kC && c7.k == ResUnit256Cfg::kTaps && dil <= ResUnit256Cfg::kMaxDil && L >= 512 &&
      gemm_use_2cta() && resunit_use_fused() && !h->split3) {
    const CUtensorMap *ta, *tb7, *tb1;
    SATB_PROPAGATE(get_tmap_a(h, sA, C, L, B, L, 1, &ta, ResUnit256Cfg::halo_rows(dil)));
    SATB_PROPAGATE(get_tmap_b(h, c7, c7.k * C, 128, &tb7));
    SATB_PROPAGATE(get_tmap_b(h, c1, C, 128, &tb1));
    e1.s16_out = sT;
    ResUnitParams<BF16> rp{c7.bias, s2.a, s2.ib, e1};
    ResUnitShape rs{L, B, dil};
    SATB_PROPAGATE(launch_resunit256<BF16>(*ta, *tb7, *tb1, rs, rp, st));
    std::swap(sA, sT);
    return 0;
  }
  // conv7(dil) on sA -> snake2 -> sT ; conv1 on sT -> + x -> raw, snake_next -> sA
  typename E::Params e7{c7.bias, nullptr, nullptr, sT, s2.a, s2.ib, C, L, 1, 0, h->split3 ? static_cast<char*>(sT) + h->lo_off : nullptr, h->raw16};
  SATB_PROPAGATE((run_conv_gemm<E, BF16>(h, c7, sA, B, L, 0, dil, 1, e7, st)));
  SATB_PROPAGATE((run_conv_gemm<E, BF16>(h, c1, sT, B, L, 0, 1, 1, e1, st)));
  return 0;
}

template <bool BF16>
int decode_impl(SatbOobleck* h, const float* z, float* audio, int B, int L, cudaStream_t st) {
  const SatbOobleckConfig& c = h->cfg;
  const int n = c.n_stages;
  // sizes
  size_t max_elems = static_cast<size_t>(B) * L * std::max(c.latent_dim, h->chans[n]);
  {
    int64_t l = L;
    for (int b = 1; b <= n; ++b) {
      const int s = c.strides[n - b];
      l *= s;
      max_elems = std::max(max_elems, static_cast<size_t>(B) * l * h->chans[n - b]);
    }
  }
  SATB_PROPAGATE(h->ensure(&h->buf_raw, &h->cap_raw, max_elems * 4));
  h->lo_off = h->split3 ? ((max_elems * 2 + 255) & ~static_cast<size_t>(255)) : 0;   // [hi | lo] halves of sA / sB
  SATB_PROPAGATE(h->ensure(&h->buf_a, &h->cap_a, max_elems * 2 + h->lo_off));
  SATB_PROPAGATE(h->ensure(&h->buf_b, &h->cap_b, max_elems * 2 + h->lo_off));
  void* raw = h->buf_raw;
  void* sA = h->buf_a;
  void* sB = h->buf_b;
  typedef EpiConv<BF16> E;
  // latent NCL fp32 -> channels-last 16-bit
  {
    dim3 grid(ceil_div(L, 32), ceil_div(c.latent_dim, 32), B);
    ncl_to_nlc16_kernel<BF16><<<grid, 256, 0, st>>>(z, static_cast<uint16_t*>(sB),
                                                    h->split3 ? reinterpret_cast<uint16_t*>(static_cast<char*>(sB) + h->lo_off) : nullptr,
                                                    c.latent_dim, L);
    count_launch();
  }
  // layers.0: conv k7 latent -> chans[n]; epilogue applies block 1's leading Snake
  {
    const ConvW& c0 = h->convs.at("layers.0.");
    const SnakeW& sn = h->snakes.at("layers.1.layers.0.");
    typename E::Params ep{c0.bias, nullptr, nullptr, sA, sn.a, sn.ib, c0.cout, L, 1, 0, h->split3 ? static_cast<char*>(sA) + h->lo_off : nullptr, h->raw16};
    SATB_PROPAGATE((run_conv_gemm<E, BF16>(h, c0, sB, B, L, 0, 1, 1, ep, st)));
  }
  int64_t Lc = L;
  for (int b = 1; b <= n; ++b) {
    const int cin = h->chans[n - b + 1], cout = h->chans[n - b], s = c.strides[n - b];
    const std::string bp = "layers." + std::to_string(b) + ".";
    const ConvW& ct = h->convs.at(bp + "layers.1.");
    const SnakeW& s_ru0 = h->snakes.at(bp + "layers.2.layers.0.");
    const int64_t Lo = Lc * s;
    SATB_REQUIRE(Lo < (int64_t(1) << 31) && static_cast<int64_t>(B) * Lo * cout < (int64_t(1) << 40), "decoder: sequence too long");
    // transposed conv reads sA [B, Lc, cin], writes raw + snake(ru0) into sB
    typename E::Params et{ct.bias, nullptr, raw, sB, s_ru0.a, s_ru0.ib, cout, static_cast<int>(Lo), s, (s + 1) / 2, h->split3 ? static_cast<char*>(sB) + h->lo_off : nullptr, h->raw16};
    SATB_PROPAGATE((run_conv_gemm<E, BF16>(h, ct, sA, B, static_cast<int>(Lc), 1, 1, s, et, st)));
    std::swap(sA, sB);  // sA now holds the residual units' input
    for (int j = 0; j < 3; ++j) {
      static const int dils[3] = {1, 3, 9};
      const SnakeW* next;
      if (j < 2)
        next = &h->snakes.at(bp + "layers." + std::to_string(3 + j) + ".layers.0.");
      else if (b < n)
        next = &h->snakes.at("layers." + std::to_string(b + 1) + ".layers.0.");
      else
        next = &h->snakes.at("layers." + std::to_string(n + 1) + ".");
      SATB_PROPAGATE((residual_unit<BF16>(h, bp + "layers." + std::to_string(2 + j) + ".", cout, B, static_cast<int>(Lo),
                                          dils[j], raw, sA, sB, next, j < 2, st)));
    }
    Lc = Lo;
    (void)cin;
  }
  // final conv k7 chans[0] -> audio channels (no bias, optional tanh): the 128 -> 2 contraction runs as a
  // 7-tap GEMM with the N tile mostly empty (8x wasted MMA work is still ~10x faster than the
  // shared-memory-bound CUDA-core version it replaces: 1.21 ms -> bandwidth-bound)
  {
    const ConvW& cf = h->convs.at("layers." + std::to_string(n + 2) + ".");
    EpiStoreNCL::Params ep{audio, nullptr, cf.cout, static_cast<int>(Lc), c.final_tanh};
    ConvHaloShape hs{static_cast<int>(Lc), B, cf.cin, cf.k, 1, cf.cout};
    if (conv_halo_enabled() && !h->split3 && cf.cout <= ConvHaloCfg::kBN && cf.cin % kBlockK == 0 && cf.cin <= 256 &&
        ConvHaloCfg::halo_rows(hs) <= 256) {
      // every activation row is fetched once per tile instead of once per tap (see conv_halo.cuh)
      const CUtensorMap *ta, *tb;
      SATB_PROPAGATE(get_tmap_a(h, sA, cf.cin, static_cast<int>(Lc), B, static_cast<int>(Lc), 1, &ta, ConvHaloCfg::halo_rows(hs)));
      SATB_PROPAGATE(get_tmap_b(h, cf, cf.k * cf.cout, ConvHaloCfg::kBN, &tb));
      SATB_PROPAGATE((launch_conv_halo<EpiStoreNCL, BF16>(*ta, *tb, hs, ep, st)));
    } else {
      SATB_PROPAGATE((run_conv_gemm<EpiStoreNCL, BF16>(h, cf, sA, B, static_cast<int>(Lc), 0, 1, 1, ep, st)));
    }
  }
  SATB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <bool BF16>
int encode_impl(SatbOobleck* h, const float* audio, float* latents, int B, int64_t T, cudaStream_t st) {
  const SatbOobleckConfig& c = h->cfg;
  const int n = c.n_stages;
  int64_t ratio = 1;
  for (int i = 0; i < n; ++i) ratio *= c.strides[i];
  SATB_REQUIRE(T % ratio == 0, "encoder: audio length must be a multiple of the downsampling ratio");
  SATB_REQUIRE(T < (int64_t(1) << 31), "encoder: sequence too long");
  size_t max_elems = 0;
  {
    int64_t l = T;
    for (int i = 0; i <= n; ++i) {
      max_elems = std::max(max_elems, static_cast<size_t>(B) * l * h->chans[i]);
      if (i < n) l /= c.strides[i];
    }
  }
  SATB_PROPAGATE(h->ensure(&h->buf_raw, &h->cap_raw, max_elems * 4));
  h->lo_off = h->split3 ? ((max_elems * 2 + 255) & ~static_cast<size_t>(255)) : 0;   // [hi | lo] halves of sA / sB
  SATB_PROPAGATE(h->ensure(&h->buf_a, &h->cap_a, max_elems * 2 + h->lo_off));
  SATB_PROPAGATE(h->ensure(&h->buf_b, &h->cap_b, max_elems * 2 + h->lo_off));
  void* raw = h->buf_raw;
  void* sA = h->buf_a;
  void* sB = h->buf_b;
  typedef EpiConv<BF16> E;
  // layers.0: conv k7 audio -> channels (CUDA cores), epilogue = block 1 / res unit 0 Snake
  {
    const ConvW& c0 = h->convs.at("layers.0.");
    const SnakeW& sn = h->snakes.at("layers.1.layers.0.layers.0.");
    SATB_REQUIRE(c0.cin * c0.k <= 16, "encoder input conv: in_channels * kernel must be <= 16");
    const size_t smem = static_cast<size_t>(c0.cin) * (64 + c0.k - 1) * 4;
    dim3 grid(static_cast<unsigned>(ceil_div64(T, 64)), B);
    conv_in_kernel<BF16><<<grid, 256, smem, st>>>(audio, c0.w32, c0.bias, sn.a, sn.ib, raw, static_cast<uint16_t*>(sA),
                                                   h->split3 ? reinterpret_cast<uint16_t*>(static_cast<char*>(sA) + h->lo_off) : nullptr,
                                                   c0.cin, c0.cout, T, c0.k, h->raw16);
    count_launch();
  }
  int64_t Lc = T;
  for (int b = 1; b <= n; ++b) {
    const int cin = h->chans[b - 1], s = c.strides[b - 1];
    const std::string bp = "layers." + std::to_string(b) + ".";
    for (int j = 0; j < 3; ++j) {
      static const int dils[3] = {1, 3, 9};
      const SnakeW* next = j < 2 ? &h->snakes.at(bp + "layers." + std::to_string(j + 1) + ".layers.0.")
                                 : &h->snakes.at(bp + "layers.3.");
      SATB_PROPAGATE((residual_unit<BF16>(h, bp + "layers." + std::to_string(j) + ".", cin, B, static_cast<int>(Lc), dils[j],
                                          raw, sA, sB, next, j < 2, st)));
    }
    // strided conv reads sA [B, Lc, cin] (already Snake-activated) -> [B, Lc/s, cout]
    const ConvW& cs = h->convs.at(bp + "layers.4.");
    const int64_t Lo = Lc / s;
    const SnakeW* nx = b < n ? &h->snakes.at("layers." + std::to_string(b + 1) + ".layers.0.layers.0.")
                             : &h->snakes.at("layers." + std::to_string(n + 1) + ".");
    typename E::Params ep{cs.bias, nullptr, b < n ? raw : nullptr, sB, nx->a, nx->ib, cs.cout, static_cast<int>(Lo), 1, 0, h->split3 ? static_cast<char*>(sB) + h->lo_off : nullptr, h->raw16};
    SATB_PROPAGATE((run_conv_gemm<E, BF16>(h, cs, sA, B, static_cast<int>(Lc), 2, 1, s, ep, st)));
    std::swap(sA, sB);
    Lc = Lo;
  }
  // final conv k3 chans[n] -> latent_dim, NCL fp32 output
  {
    const ConvW& cf = h->convs.at("layers." + std::to_string(n + 2) + ".");
    EpiStoreNCL::Params ep{latents, cf.bias, cf.cout, static_cast<int>(Lc), 0};
    SATB_PROPAGATE((run_conv_gemm<EpiStoreNCL, BF16>(h, cf, sA, B, static_cast<int>(Lc), 0, 1, 1, ep, st)));
  }
  return 0;
}

}  // namespace

extern "C" {

int satb_oobleck_create(const SatbOobleckConfig* cfg, SatbOobleck** out) {
  SATB_REQUIRE(cfg && out, "null argument");
  SATB_REQUIRE(cfg->n_stages >= 1 && cfg->n_stages <= SATB_MAX_STAGES, "bad number of stages");
  SATB_REQUIRE(cfg->channels % 32 == 0, "channels must be a multiple of 32");
  SATB_REQUIRE(cfg->latent_dim % 8 == 0, "latent_dim must be a multiple of 8");
  SATB_REQUIRE(cfg->in_channels >= 1 && cfg->in_channels <= 2, "audio channels must be 1 or 2");
  SatbOobleck* h = new SatbOobleck();
  h->cfg = *cfg;
  h->bf16 = cfg->operand_dtype == 1;
  h->split3 = cfg->operand_dtype == 2;
  // fp16 operands: the un-activated skip stream is carried in fp16 as well (8 instead of 12 bytes per element and
  // channel through a fused ResidualUnit; measured +23 % on the fp16-operand error floor, tests/test_gpu_baseline_size).
  // bf16 (8 mantissa bits) keeps the fp32 stream.  SATB_RAW=fp32 restores fp32 for A/B measurements.
  h->raw16 = (!h->bf16 && !h->split3 && raw_stream_16bit()) ? 1 : 0;
  h->chans.push_back(cfg->channels);
  for (int i = 0; i < cfg->n_stages; ++i) h->chans.push_back(cfg->c_mults[i] * cfg->channels);
  *out = h;
  return 0;
}

void satb_oobleck_destroy(SatbOobleck* h) {
  if (!h) return;
  for (void* p : h->owned) cudaFree(p);
  for (auto& kv : h->raw) cudaFree(kv.second.first);
  if (h->buf_raw) cudaFree(h->buf_raw);
  if (h->buf_a) cudaFree(h->buf_a);
  if (h->buf_b) cudaFree(h->buf_b);
  delete h;
}

int satb_oobleck_load_weight(SatbOobleck* h, const char* name, const float* src, long long numel, void* stream) {
  SATB_REQUIRE(h && name && src && numel > 0, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto it = h->raw.find(name);
  if (it != h->raw.end() && it->second.second != numel) {
    cudaFree(it->second.first);
    h->raw.erase(it);
    it = h->raw.end();
  }
  float* dst;
  if (it == h->raw.end()) {
    SATB_CHECK_CUDA(cudaMalloc(&dst, numel * sizeof(float)));
    h->raw[name] = std::make_pair(dst, numel);
  } else {
    dst = it->second.first;
  }
  SATB_CHECK_CUDA(cudaMemcpyAsync(dst, src, numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
  h->finalized = false;
  return 0;
}

int satb_oobleck_finalize(SatbOobleck* h, void* stream) {
  SATB_REQUIRE(h, "null handle");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const SatbOobleckConfig& c = h->cfg;
  const int n = c.n_stages;
  for (void* p : h->owned) cudaFree(p);
  h->owned.clear();
  h->convs.clear();
  h->snakes.clear();
  h->tmaps.clear();
  auto res_unit = [&](const std::string& pfx, int C) -> int {
    SATB_PROPAGATE(prep_snake(h, pfx + "layers.0.", C, st));
    SATB_PROPAGATE(prep_conv(h, pfx + "layers.1.", C, C, 7, false, false, true, 1, st));
    SATB_PROPAGATE(prep_snake(h, pfx + "layers.2.", C, st));
    SATB_PROPAGATE(prep_conv(h, pfx + "layers.3.", C, C, 1, false, false, true, 1, st));
    return 0;
  };
  if (c.is_decoder) {
    SATB_PROPAGATE(prep_conv(h, "layers.0.", c.latent_dim, h->chans[n], 7, false, false, true, 1, st));
    for (int b = 1; b <= n; ++b) {
      const int cin = h->chans[n - b + 1], cout = h->chans[n - b], s = c.strides[n - b];
      const std::string bp = "layers." + std::to_string(b) + ".";
      SATB_PROPAGATE(prep_snake(h, bp + "layers.0.", cin, st));
      SATB_PROPAGATE(prep_conv(h, bp + "layers.1.", cin, cout, 2 * s, true, false, true, s, st));
      for (int j = 0; j < 3; ++j) SATB_PROPAGATE(res_unit(bp + "layers." + std::to_string(2 + j) + ".", cout));
    }
    SATB_PROPAGATE(prep_snake(h, "layers." + std::to_string(n + 1) + ".", h->chans[0], st));
    SATB_PROPAGATE(prep_conv(h, "layers." + std::to_string(n + 2) + ".", h->chans[0], c.in_channels, 7, false, false, false, 1, st));
    SATB_REQUIRE(h->chans[0] % 2 == 0, "decoder: channels must be even");
  } else {
    SATB_PROPAGATE(prep_conv(h, "layers.0.", c.in_channels, h->chans[0], 7, false, true, true, 1, st));
    for (int b = 1; b <= n; ++b) {
      const int cin = h->chans[b - 1], cout = h->chans[b], s = c.strides[b - 1];
      const std::string bp = "layers." + std::to_string(b) + ".";
      for (int j = 0; j < 3; ++j) SATB_PROPAGATE(res_unit(bp + "layers." + std::to_string(j) + ".", cin));
      SATB_PROPAGATE(prep_snake(h, bp + "layers.3.", cin, st));
      SATB_PROPAGATE(prep_conv(h, bp + "layers.4.", cin, cout, 2 * s, false, false, true, 1, st));
    }
    SATB_PROPAGATE(prep_snake(h, "layers." + std::to_string(n + 1) + ".", h->chans[n], st));
    SATB_PROPAGATE(prep_conv(h, "layers." + std::to_string(n + 2) + ".", h->chans[n], c.latent_dim, 3, false, false, true, 1, st));
  }
  SATB_CHECK_CUDA(cudaStreamSynchronize(st));
  h->finalized = true;
  return 0;
}

int satb_oobleck_decode(SatbOobleck* h, const float* z, float* audio, int B, int L, void* stream) {
  SATB_REQUIRE(h && h->finalized, "oobleck: weights not finalized");
  SATB_REQUIRE(h->cfg.is_decoder, "oobleck: handle is an encoder");
  SATB_REQUIRE(z && audio && B >= 1 && L >= 1, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return h->bf16 ? decode_impl<true>(h, z, audio, B, L, st) : decode_impl<false>(h, z, audio, B, L, st);
}

int satb_oobleck_encode(SatbOobleck* h, const float* audio, float* latents, int B, long long T, void* stream) {
  SATB_REQUIRE(h && h->finalized, "oobleck: weights not finalized");
  SATB_REQUIRE(!h->cfg.is_decoder, "oobleck: handle is a decoder");
  SATB_REQUIRE(audio && latents && B >= 1 && T >= 1, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return h->bf16 ? encode_impl<true>(h, audio, latents, B, T, st) : encode_impl<false>(h, audio, latents, B, T, st);
}

}  // extern "C"
