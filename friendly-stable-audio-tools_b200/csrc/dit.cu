// Host orchestration of the DiffusionTransformer forward (reference models/dit.py:135-364,
// models/transformer.py:656-809) on the kernels of this library.  One handle per
// (device, model); all work is enqueued on the caller's stream; no allocation and no
// synchronisation inside satb_dit_forward once the workspace has been reserved, so a
// whole denoise step can be captured in a CUDA graph.
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/satb200.h"
#include "common.cuh"
#include "gemm.cuh"
#include "kernels.h"

namespace satb {

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&p, need);
    if (e != cudaSuccess) {
      set_last_error(std::string("cudaMalloc failed: ") + cudaGetErrorString(e) + " (" + std::to_string(need) + " B)");
      return -2;
    }
    bytes = need;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <class T>
  T* as() const {
    return static_cast<T*>(p);
  }
};

struct TmapCache {
  typedef std::tuple<const void*, int, int, int, int64_t, int64_t, int> Key;
  std::map<Key, CUtensorMap> maps;
  int get_a(const void* ptr, int K, int L, int batches, int64_t rs, int64_t bs, const CUtensorMap** out) {
    Key k(ptr, K, L, batches, rs, bs, -1);
    auto it = maps.find(k);
    if (it == maps.end()) {
      CUtensorMap m;
      SATB_PROPAGATE(make_tmap_a(&m, ptr, K, L, batches, rs, bs));
      it = maps.emplace(k, m).first;
    }
    *out = &it->second;
    return 0;
  }
  int get_b(const void* ptr, int K, int rows, int64_t rs, int box_rows, const CUtensorMap** out) {
    Key k(ptr, K, rows, 0, rs, 0, box_rows);
    auto it = maps.find(k);
    if (it == maps.end()) {
      CUtensorMap m;
      SATB_PROPAGATE(make_tmap_b(&m, ptr, K, rows, rs, box_rows));
      it = maps.emplace(k, m).first;
    }
    *out = &it->second;
    return 0;
  }
};

// Flat Linear: C[M, N] = A[M, K] * W[N, K]^T with a fused epilogue.
template <class Epi, int BN, bool BF16>
static int linear(TmapCache& tc, const void* A, int64_t lda, int M, int K, const void* W, int N,
                  const typename Epi::Params& ep, cudaStream_t stream) {
  const CUtensorMap *ta, *tb;
  SATB_PROPAGATE(tc.get_a(A, K, M, 1, lda, static_cast<int64_t>(M) * lda, &ta));
  GemmShape s;
  s.L = M; s.batches = 1; s.N = N; s.K = K; s.n_taps = 1; s.tap_base = 0; s.tap_step = 0; s.b_tap_rows = N; s.stride = 1;
  s.b_static = 1;   // W is a weight matrix prepared at finalize time
  if constexpr (BN == 256 || BN == 128) {
    // CTA-pair kernel (256 x BN tiles, half of B per CTA).  Also for BN = 128: a single CTA streaming
    // 128 x 128 x 16 MMAs needs 128 B/clk of shared-memory operand reads, the SM's limit; the pair needs 96.
    // (Choosing single CTAs where they save a round of the persistent grid -- QKV: 594 pair tiles = 8.03 -> 9
    // rounds vs 1170 single tiles = 7.9 -> 8 -- was measured ~3 % slower: the pair's lower smem traffic wins.
    // 192-wide pair tiles -- 792 tiles = 10.7 -> 11 rounds of 3/4 the work = 8.25 -- measured 6 % slower as well,
    // 2.66 vs 2.50 ms per step: A is re-read a third more often.)
    if (gemm_use_2cta() && M >= 1024) {
      SATB_PROPAGATE(tc.get_b(W, K, N, K, BN / 2, &tb));
      return launch_gemm_2cta<Epi, BN, BF16>(*ta, *tb, s, ep, stream);
    }
  }
  SATB_PROPAGATE(tc.get_b(W, K, N, K, BN, &tb));
  return launch_gemm<Epi, BN, BF16>(*ta, *tb, s, ep, stream);
}

// Picks the N tile (256 or 128) that wastes less of the last wave of the persistent grid; the
// 128-wide tile streams as many smem bytes per MMA cycle as the tensor pipe can take, so it is
// only preferred when it clearly wins on wave quantisation.
template <class Epi, bool BF16>
static int linear_auto(TmapCache& tc, const void* A, int64_t lda, int M, int K, const void* W, int N,
                       const typename Epi::Params& ep, cudaStream_t stream) {
  const double sms = device_sm_count();
  auto eff = [&](int bn) {
    const double waves = static_cast<double>(ceil_div(M, kBlockM)) * ceil_div(N, bn) / sms;
    return waves / std::ceil(waves);
  };
  if (N % 128 == 0 && eff(128) * 0.9 > eff(256)) return linear<Epi, 128, BF16>(tc, A, lda, M, K, W, N, ep, stream);
  return linear<Epi, 256, BF16>(tc, A, lda, M, K, W, N, ep, stream);
}

struct LayerW {
  float *pre_g = nullptr, *pre_b = nullptr, *ca_g = nullptr, *ca_b = nullptr, *ff_g = nullptr, *ff_b = nullptr;
  uint16_t *w_qkv = nullptr, *w_o = nullptr, *w_q = nullptr, *w_kv = nullptr, *w_co = nullptr, *w_ff1 = nullptr,
           *w_ff2 = nullptr;
  float *b_ff1 = nullptr, *b_ff2 = nullptr;
  // LayerNorm folded into the following GEMM (LnFold): c = W gamma, d = W beta per output column
  float *c_qkv = nullptr, *d_qkv = nullptr, *c_q = nullptr, *d_q = nullptr, *c_ff1 = nullptr, *d_ff1 = nullptr;
};

}  // namespace satb

using namespace satb;

struct SatbDit {
  SatbDitConfig cfg;
  int D, H, dh, C, Cin, ct, ce, gd, ge, ffi, depth, F, nf;   // C = output channels, Cin = C + input_concat_dim
  int pdim = 0;               // prepend_cond_dim
  float *pe0_w = nullptr, *pe2_w = nullptr;   // to_prepend_embed (fp32, bias-free)
  int Pp = 0;                 // prepend-conditioning tokens of the current conditioning
  DevBuf ws_prep;             // their embeddings [B, Pp, D] + scratch
  bool bf16, adaln, qk_norm = false;
  bool ln_fused = false;   // prepend-mode blocks without qk_norm: no LayerNorm kernels (see dit_forward_impl)
  int P;  // prepended tokens: 1 (the global-conditioning token; 0 in adaLN mode) + Pp
  std::vector<LayerW> layers;
  std::vector<void*> owned;   // every cudaMalloc of weight storage
  // globals
  float *ts_w = nullptr, *te0_w = nullptr, *te0_b = nullptr, *te2_w = nullptr, *te2_b = nullptr;
  uint16_t *ce0_w = nullptr, *ce2_w = nullptr;
  float *ge0_w = nullptr, *ge2_w = nullptr;
  float *pre_w = nullptr, *post_w = nullptr, *pin_w = nullptr, *pout_w = nullptr, *inv_freq = nullptr;
  uint16_t *w_in16 = nullptr, *w_out16 = nullptr;
  float* w_ssg = nullptr;   // [depth*6D, D] fp32 (adaLN)
  int* ff_perm = nullptr;   // SwiGLU row interleave
  std::map<std::string, int> loaded;
  bool finalized = false;
  // conditioning state
  int B = 0, Mctx = 0, Rc = 0;  // Rc = rows that run cross-attention
  bool cfg_on = false, has_cross = false, has_global = false;
  // workspace
  TmapCache tmaps;
  DevBuf ws_h, ws_a16, ws_qkv, ws_attn, ws_q16, ws_ff, ws_ain, ws_y, ws_small, ws_cond, ws_kv, ws_rope, ws_stats;
  int rope_len = 0;
  int res_R = 0, res_L = 0, res_P = -1;
  // optional per-category CUDA-event timing (bench.py roofline)
  bool prof_on = false;
  struct ProfRec { int cat; cudaEvent_t a, b; };
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;
  cudaEvent_t prof_event() {
    if (!prof_pool.empty()) { cudaEvent_t e = prof_pool.back(); prof_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }

  template <class T>
  int alloc(T** p, size_t n) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T) < 256 ? 256 : n * sizeof(T));
    if (e != cudaSuccess) {
      set_last_error(std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
      return -2;
    }
    owned.push_back(q);
    *p = static_cast<T*>(q);
    return 0;
  }
};

static bool ends_with(const std::string& s, const std::string& suf) {
  return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}

extern "C" {

const char* satb_last_error(void) { return get_last_error(); }
unsigned long long satb_launch_count(void) { return g_launch_count; }
void satb_reset_launch_count(void) { g_launch_count = 0; }
void satb_add_launch_count(unsigned long long n) { g_launch_count += n; }
int satb_abi_version(void) { return SATB_ABI_VERSION; }

int satb_dit_create(const SatbDitConfig* cfg, SatbDit** out) {
  SATB_REQUIRE(cfg && out, "null argument");
  SATB_REQUIRE(cfg->embed_dim % 128 == 0, "embed_dim must be a multiple of 128");
  SATB_REQUIRE(cfg->num_heads > 0 && cfg->embed_dim / cfg->num_heads == 64 && cfg->embed_dim % cfg->num_heads == 0,
               "head dim must be 64");
  SATB_REQUIRE(cfg->io_channels % 8 == 0 && cfg->io_channels % 32 == 0, "io_channels must be a multiple of 32");
  SATB_REQUIRE(cfg->patch_size == 1, "patch_size 1 only");
  SATB_REQUIRE(cfg->input_concat_dim >= 0 && cfg->input_concat_dim % 8 == 0, "input_concat_dim must be a multiple of 8");
  SATB_REQUIRE(cfg->prepend_cond_dim >= 0 && cfg->prepend_cond_dim % 4 == 0, "prepend_cond_dim must be a multiple of 4");
  SATB_REQUIRE(!(cfg->prepend_cond_dim > 0 && cfg->global_cond_type == 1),
               "prepend conditioning is supported with global_cond_type \"prepend\" only");
  SatbDit* d = new SatbDit();
  d->cfg = *cfg;
  d->D = cfg->embed_dim;
  d->H = cfg->num_heads;
  d->dh = d->D / d->H;
  d->C = cfg->io_channels;
  d->Cin = cfg->io_channels + (cfg->input_concat_dim > 0 ? cfg->input_concat_dim : 0);
  d->pdim = cfg->prepend_cond_dim > 0 ? cfg->prepend_cond_dim : 0;
  d->ct = cfg->cond_token_dim;
  d->ce = cfg->project_cond_tokens ? d->D : d->ct;
  d->gd = cfg->global_cond_dim;
  d->ge = cfg->project_global_cond ? d->D : d->gd;
  d->ffi = 4 * d->D;
  d->depth = cfg->depth;
  d->F = 128;  // timestep_features_dim 256 = cos | sin of 128 frequencies (models/dit.py:41-43)
  const int rot = d->dh / 2 > 32 ? d->dh / 2 : 32;  // models/transformer.py:737
  d->nf = rot / 2;
  d->bf16 = cfg->operand_dtype == 1;
  d->adaln = cfg->global_cond_type == 1;
  d->qk_norm = cfg->qk_norm != 0;
  d->P = d->adaln ? 0 : 1;
  if (d->ct > 0) {
    if (d->ce % 64 != 0 || d->H % (d->ce / 64) != 0) {
      delete d;
      set_last_error("cond embed dim must be a multiple of 64 with kv heads dividing num_heads");
      return -1;
    }
  }
  if (d->gd > 0 && d->ge != d->D) {
    delete d;
    set_last_error("global embed dim must equal embed_dim");
    return -1;
  }
  SATB_REQUIRE(d->nf == 16, "rotary dim must be 32 (head dim 64)");
  d->layers.resize(d->depth);
  *out = d;
  return 0;
}

void satb_dit_destroy(SatbDit* d) {
  if (!d) return;
  for (void* p : d->owned) cudaFree(p);
  d->ws_h.release(); d->ws_a16.release(); d->ws_qkv.release(); d->ws_attn.release(); d->ws_q16.release();
  d->ws_ff.release(); d->ws_ain.release(); d->ws_y.release(); d->ws_small.release(); d->ws_cond.release();
  d->ws_kv.release(); d->ws_rope.release(); d->ws_stats.release(); d->ws_prep.release();
  delete d;
}

// Upload one state-dict entry (fp32, device pointer, reference key relative to
// DiffusionTransformer).  Big matrices are cast to the 16-bit operand
// type here, once; small tensors stay fp32.
int satb_dit_load_weight(SatbDit* d, const char* name_c, const float* src, long long numel, void* stream_v) {
  SATB_REQUIRE(d && name_c && src, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  const std::string name(name_c);
  const int D = d->D, C = d->C, Cin = d->Cin;
  auto copy_f32 = [&](float** dst, long long expect) -> int {
    SATB_REQUIRE(numel == expect, ("bad size for " + name).c_str());
    if (!*dst) SATB_PROPAGATE(d->alloc(dst, expect));
    SATB_CHECK_CUDA(cudaMemcpyAsync(*dst, src, expect * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
  };
  auto cast16 = [&](uint16_t** dst, int rows, int cols, const int* perm) -> int {
    SATB_REQUIRE(numel == static_cast<long long>(rows) * cols, ("bad size for " + name).c_str());
    if (!*dst) SATB_PROPAGATE(d->alloc(dst, static_cast<size_t>(rows) * cols));
    return launch_cast_rows(src, *dst, perm, rows, cols, cols, cols, d->bf16, st);
  };
  d->finalized = false;
  d->loaded[name] = 1;
  if (name == "timestep_features.weight") return copy_f32(&d->ts_w, d->F);
  if (name == "to_timestep_embed.0.weight") return copy_f32(&d->te0_w, static_cast<long long>(D) * 2 * d->F);
  if (name == "to_timestep_embed.0.bias") return copy_f32(&d->te0_b, D);
  if (name == "to_timestep_embed.2.weight") return copy_f32(&d->te2_w, static_cast<long long>(D) * D);
  if (name == "to_timestep_embed.2.bias") return copy_f32(&d->te2_b, D);
  if (name == "to_cond_embed.0.weight") return cast16(&d->ce0_w, d->ce, d->ct, nullptr);
  if (name == "to_cond_embed.2.weight") return cast16(&d->ce2_w, d->ce, d->ce, nullptr);
  if (name == "to_global_embed.0.weight") return copy_f32(&d->ge0_w, static_cast<long long>(d->ge) * d->gd);
  if (name == "to_global_embed.2.weight") return copy_f32(&d->ge2_w, static_cast<long long>(d->ge) * d->ge);
  if (name == "preprocess_conv.weight") return copy_f32(&d->pre_w, static_cast<long long>(Cin) * Cin);
  if (name == "to_prepend_embed.0.weight" && d->pdim > 0) return copy_f32(&d->pe0_w, static_cast<long long>(D) * d->pdim);
  if (name == "to_prepend_embed.2.weight" && d->pdim > 0) return copy_f32(&d->pe2_w, static_cast<long long>(D) * D);
  if (name == "postprocess_conv.weight") return copy_f32(&d->post_w, static_cast<long long>(C) * C);
  if (name == "transformer.project_in.weight") return copy_f32(&d->pin_w, static_cast<long long>(D) * Cin);
  if (name == "transformer.project_out.weight") return copy_f32(&d->pout_w, static_cast<long long>(C) * D);
  if (name == "transformer.rotary_pos_emb.inv_freq") return copy_f32(&d->inv_freq, d->nf);
  const std::string lp = "transformer.layers.";
  if (name.compare(0, lp.size(), lp) == 0) {
    const size_t dot = name.find('.', lp.size());
    SATB_REQUIRE(dot != std::string::npos, ("bad key " + name).c_str());
    const int li = atoi(name.substr(lp.size(), dot - lp.size()).c_str());
    SATB_REQUIRE(li >= 0 && li < d->depth, ("layer index out of range in " + name).c_str());
    LayerW& L = d->layers[li];
    const std::string k = name.substr(dot + 1);
    if (k == "pre_norm.gamma") return copy_f32(&L.pre_g, D);
    if (k == "pre_norm.beta") return copy_f32(&L.pre_b, D);
    if (k == "cross_attend_norm.gamma") return copy_f32(&L.ca_g, D);
    if (k == "cross_attend_norm.beta") return copy_f32(&L.ca_b, D);
    if (k == "ff_norm.gamma") return copy_f32(&L.ff_g, D);
    if (k == "ff_norm.beta") return copy_f32(&L.ff_b, D);
    if (k == "self_attn.to_qkv.weight") return cast16(&L.w_qkv, 3 * D, D, nullptr);
    if (k == "self_attn.to_out.weight") return cast16(&L.w_o, D, D, nullptr);
    if (k == "cross_attn.to_q.weight") return cast16(&L.w_q, D, D, nullptr);
    if (k == "cross_attn.to_kv.weight") return cast16(&L.w_kv, 2 * d->ce, d->ce, nullptr);
    if (k == "cross_attn.to_out.weight") return cast16(&L.w_co, D, D, nullptr);
    if (k == "ff.ff.0.proj.weight" || k == "ff.ff.0.proj.bias") {
      if (!d->ff_perm) {
        // interleave so that every 64-row group = 32 value rows then their 32 gate rows
        std::vector<int> perm(2 * d->ffi);
        for (int n = 0; n < 2 * d->ffi; ++n) {
          const int g = n / 64, w = n % 64;
          perm[n] = w < 32 ? g * 32 + w : d->ffi + g * 32 + (w - 32);
        }
        SATB_PROPAGATE(d->alloc(&d->ff_perm, perm.size()));
        SATB_CHECK_CUDA(cudaMemcpy(d->ff_perm, perm.data(), perm.size() * sizeof(int), cudaMemcpyHostToDevice));
      }
      if (ends_with(k, "weight")) return cast16(&L.w_ff1, 2 * d->ffi, D, d->ff_perm);
      SATB_REQUIRE(numel == 2 * d->ffi, ("bad size for " + name).c_str());
      if (!L.b_ff1) SATB_PROPAGATE(d->alloc(&L.b_ff1, 2 * d->ffi));
      return launch_gather_f32(src, L.b_ff1, d->ff_perm, 2 * d->ffi, st);
    }
    if (k == "ff.ff.2.weight") return cast16(&L.w_ff2, D, d->ffi, nullptr);
    if (k == "ff.ff.2.bias") return copy_f32(&L.b_ff2, D);
    if (k == "to_scale_shift_gate.1.weight") {
      SATB_REQUIRE(numel == 6LL * D * D, ("bad size for " + name).c_str());
      if (!d->w_ssg) SATB_PROPAGATE(d->alloc(&d->w_ssg, static_cast<size_t>(d->depth) * 6 * D * D));
      SATB_CHECK_CUDA(cudaMemcpyAsync(d->w_ssg + static_cast<size_t>(li) * 6 * D * D, src, numel * sizeof(float),
                                      cudaMemcpyDeviceToDevice, st));
      return 0;
    }
  }
  d->loaded.erase(name);
  set_last_error("unknown DiT weight key: " + name);
  return -4;
}

// Folds the 1x1 pre/post convolutions into project_in / project_out
// (x + Wpre x then Win: Win (I + Wpre); Wout then y + Wpost y: (I + Wpost) Wout;
// models/dit.py:197,224 with transformer.py:778,807) and checks completeness.
int satb_dit_finalize(SatbDit* d, void* stream_v) {
  SATB_REQUIRE(d, "null handle");
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  const int D = d->D, C = d->C;
  SATB_REQUIRE(d->ts_w && d->te0_w && d->te0_b && d->te2_w && d->te2_b, "timestep embedding weights missing");
  SATB_REQUIRE(d->pin_w && d->pout_w && d->pre_w && d->post_w, "project_in/out or pre/post conv weights missing");
  SATB_REQUIRE(d->inv_freq, "rotary inv_freq missing");
  if (d->ct > 0) SATB_REQUIRE(d->ce0_w && d->ce2_w, "to_cond_embed weights missing");
  if (d->gd > 0) SATB_REQUIRE(d->ge0_w && d->ge2_w, "to_global_embed weights missing");
  for (int i = 0; i < d->depth; ++i) {
    const LayerW& L = d->layers[i];
    SATB_REQUIRE(L.pre_g && L.ff_g && L.w_qkv && L.w_o && L.w_ff1 && L.w_ff2, "transformer layer weights missing");
    if (d->ct > 0) SATB_REQUIRE(L.ca_g && L.w_q && L.w_kv && L.w_co, "cross-attention weights missing");
  }
  if (d->adaln) SATB_REQUIRE(d->w_ssg, "adaLN to_scale_shift_gate weights missing");
  SATB_CHECK_CUDA(cudaStreamSynchronize(st));
  const int Cin = d->Cin;
  if (d->pdim > 0) SATB_REQUIRE(d->pe0_w && d->pe2_w, "to_prepend_embed weights missing");
  std::vector<float> pin(static_cast<size_t>(D) * Cin), pout(static_cast<size_t>(C) * D), pre(Cin * Cin), post(C * C);
  SATB_CHECK_CUDA(cudaMemcpy(pin.data(), d->pin_w, pin.size() * 4, cudaMemcpyDeviceToHost));
  SATB_CHECK_CUDA(cudaMemcpy(pout.data(), d->pout_w, pout.size() * 4, cudaMemcpyDeviceToHost));
  SATB_CHECK_CUDA(cudaMemcpy(pre.data(), d->pre_w, pre.size() * 4, cudaMemcpyDeviceToHost));
  SATB_CHECK_CUDA(cudaMemcpy(post.data(), d->post_w, post.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<float> fin(static_cast<size_t>(D) * Cin), fout(static_cast<size_t>(C) * D);
  for (int n = 0; n < D; ++n)
    for (int c = 0; c < Cin; ++c) {
      double acc = pin[static_cast<size_t>(n) * Cin + c];
      for (int j = 0; j < Cin; ++j) acc += static_cast<double>(pin[static_cast<size_t>(n) * Cin + j]) * pre[j * Cin + c];
      fin[static_cast<size_t>(n) * Cin + c] = static_cast<float>(acc);
    }
  for (int c = 0; c < C; ++c)
    for (int k = 0; k < D; ++k) {
      double acc = pout[static_cast<size_t>(c) * D + k];
      for (int j = 0; j < C; ++j) acc += static_cast<double>(post[c * C + j]) * pout[static_cast<size_t>(j) * D + k];
      fout[static_cast<size_t>(c) * D + k] = static_cast<float>(acc);
    }
  float *tmp_in = nullptr, *tmp_out = nullptr;
  SATB_CHECK_CUDA(cudaMalloc(&tmp_in, fin.size() * 4));
  SATB_CHECK_CUDA(cudaMalloc(&tmp_out, fout.size() * 4));
  SATB_CHECK_CUDA(cudaMemcpy(tmp_in, fin.data(), fin.size() * 4, cudaMemcpyHostToDevice));
  SATB_CHECK_CUDA(cudaMemcpy(tmp_out, fout.data(), fout.size() * 4, cudaMemcpyHostToDevice));
  if (!d->w_in16) SATB_PROPAGATE(d->alloc(&d->w_in16, fin.size()));
  if (!d->w_out16) SATB_PROPAGATE(d->alloc(&d->w_out16, fout.size()));
  int rc = launch_cast_rows(tmp_in, d->w_in16, nullptr, D, Cin, Cin, Cin, d->bf16, st);
  if (rc == 0) rc = launch_cast_rows(tmp_out, d->w_out16, nullptr, C, D, D, D, d->bf16, st);
  cudaStreamSynchronize(st);
  cudaFree(tmp_in);
  cudaFree(tmp_out);
  SATB_PROPAGATE(rc);
  // LayerNorm folded into the consumer GEMMs (opt-in with SATB_LN=fold, plain prepend-mode blocks only; measured slower
  // than the LayerNorm kernels, see runtime.cu): per-column vectors c = W gamma, d = W beta of every LayerNorm -> Linear pair
  d->ln_fused = !d->adaln && !d->qk_norm && ln_fold_enabled() && d->D == 6 * 256;
  if (d->ln_fused) {
    for (int i = 0; i < d->depth; ++i) {
      LayerW& L = d->layers[i];
      auto prep = [&](float** c, float** dd, const uint16_t* w, const float* g, const float* b, int rows) -> int {
        // beta is a zero buffer in the reference's LayerNorm (transformer.py:200-206): then d = W beta = 0 and the
        // epilogues skip it (d pointer null)
        bool has_beta = false;
        if (b) {
          std::vector<float> hb(D);
          SATB_CHECK_CUDA(cudaMemcpy(hb.data(), b, D * sizeof(float), cudaMemcpyDeviceToHost));
          for (float v : hb) has_beta = has_beta || v != 0.f;
        }
        if (!*c) SATB_PROPAGATE(d->alloc(c, rows));
        float* scratch = nullptr;
        if (has_beta && !*dd) SATB_PROPAGATE(d->alloc(dd, rows));
        if (!has_beta) {
          *dd = nullptr;
          SATB_PROPAGATE(d->alloc(&scratch, rows));
        }
        return launch_ln_fold_vectors(w, g, has_beta ? b : nullptr, *c, has_beta ? *dd : scratch, rows, D, d->bf16, st);
      };
      SATB_PROPAGATE(prep(&L.c_qkv, &L.d_qkv, L.w_qkv, L.pre_g, L.pre_b, 3 * D));
      if (d->ct > 0) SATB_PROPAGATE(prep(&L.c_q, &L.d_q, L.w_q, L.ca_g, L.ca_b, D));
      SATB_PROPAGATE(prep(&L.c_ff1, &L.d_ff1, L.w_ff1, L.ff_g, L.ff_b, 2 * d->ffi));
    }
    SATB_CHECK_CUDA(cudaStreamSynchronize(st));
  }
  d->tmaps.maps.clear();
  d->finalized = true;
  return 0;
}

}  // extern "C"

static int ensure_rope(SatbDit* d, int N_seq) {
  if (d->rope_len == N_seq) return 0;
  // models/transformer.py:130-155: freqs[p, j] = float(p) * inv_freq[j] in fp32; cos/sin in fp32.
  std::vector<float> inv(d->nf), tab(static_cast<size_t>(2) * N_seq * d->nf);
  SATB_CHECK_CUDA(cudaMemcpy(inv.data(), d->inv_freq, d->nf * 4, cudaMemcpyDeviceToHost));
  for (int p = 0; p < N_seq; ++p)
    for (int j = 0; j < d->nf; ++j) {
      const float f = static_cast<float>(p) * inv[j];
      tab[static_cast<size_t>(p) * d->nf + j] = cosf(f);
      tab[static_cast<size_t>(N_seq) * d->nf + static_cast<size_t>(p) * d->nf + j] = sinf(f);
    }
  SATB_PROPAGATE(d->ws_rope.ensure(tab.size() * 4));
  SATB_CHECK_CUDA(cudaMemcpy(d->ws_rope.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
  d->rope_len = N_seq;
  return 0;
}

extern "C" {

// Reserve every activation buffer for R rows of L latent tokens (synchronous; call
// before capturing a CUDA graph).  Grow-only.
int satb_dit_reserve(SatbDit* d, int R, int L) {
  SATB_REQUIRE(d && d->finalized, "weights not finalized");
  SATB_REQUIRE(R >= 1 && L >= 1, "bad shape");
  const int N_seq = L + d->P;
  const size_t M = static_cast<size_t>(R) * N_seq;
  const int D = d->D;
  SATB_PROPAGATE(d->ws_h.ensure(M * D * 4));
  SATB_PROPAGATE(d->ws_a16.ensure(M * D * 2));
  SATB_PROPAGATE(d->ws_qkv.ensure(M * 3 * D * 2));
  SATB_PROPAGATE(d->ws_attn.ensure(M * D * 2));
  SATB_PROPAGATE(d->ws_q16.ensure(M * D * 2));
  SATB_PROPAGATE(d->ws_ff.ensure(M * d->ffi * 2));
  SATB_PROPAGATE(d->ws_ain.ensure(M * d->Cin * 2));
  SATB_PROPAGATE(d->ws_y.ensure(M * d->C * 4));
  SATB_PROPAGATE(d->ws_stats.ensure(3 * M * kLnSlots * sizeof(float2)));
  SATB_PROPAGATE(ensure_rope(d, N_seq));
  d->tmaps.maps.clear();
  d->res_R = R;
  d->res_L = L;
  d->res_P = d->P;
  return 0;
}

}  // extern "C"

struct SmallWs {
  float *fourier, *te_h, *tok, *ge_h, *ge, *ssg;
};
static SmallWs small_ws(SatbDit* d, int R) {
  SmallWs s;
  float* p = d->ws_small.as<float>();
  s.fourier = p; p += static_cast<size_t>(R) * 2 * d->F;
  s.te_h = p;    p += static_cast<size_t>(R) * d->D;
  s.tok = p;     p += static_cast<size_t>(R) * d->D;
  s.ge_h = p;    p += static_cast<size_t>(R) * d->D;
  s.ge = p;      p += static_cast<size_t>(R) * d->D;
  s.ssg = p;
  return s;
}

extern "C" {

// Prepend conditioning of the next prepare_cond: embeds = W2 silu(W0 prepend) (dit.py:75-81,157-161), kept as fp32
// tokens [B, Pp, D]; the unconditional CFG rows use zeros (to_prepend_embed is bias-free, so MLP(0) = 0, dit.py:309-311).
int satb_dit_set_prepend_cond(SatbDit* d, const float* prepend, int B, int n_tokens, void* stream_v) {
  SATB_REQUIRE(d && d->finalized, "weights not finalized");
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  if (!prepend || n_tokens <= 0) {
    d->Pp = 0;
    d->P = d->adaln ? 0 : 1;
    return 0;
  }
  SATB_REQUIRE(d->pdim > 0 && !d->adaln, "this model has no prepend conditioning");
  SATB_REQUIRE(B >= 1, "bad batch");
  const size_t rows = static_cast<size_t>(B) * n_tokens;
  SATB_REQUIRE(rows <= 4096, "too many prepend tokens");
  SATB_PROPAGATE(d->ws_prep.ensure(2 * rows * d->D * sizeof(float)));
  float* emb = d->ws_prep.as<float>();
  float* hid = emb + rows * d->D;
  SATB_PROPAGATE(launch_skinny_linear(prepend, d->pe0_w, nullptr, nullptr, hid, static_cast<int>(rows), d->pdim, d->D, 1, st));
  SATB_PROPAGATE(launch_skinny_linear(hid, d->pe2_w, nullptr, nullptr, emb, static_cast<int>(rows), d->D, d->D, 0, st));
  d->Pp = n_tokens;
  d->P = 1 + n_tokens;
  return 0;
}

// Step-invariant conditioning work hoisted out of the sampler loop:
// to_cond_embed, to_global_embed and every layer's cross-attention k/v projection.
//   cross [B, Mctx, ct] fp32 or null; neg_cross [B, Mctx, ct] fp32 or null (already masked);
//   global [B, gd] fp32 or null; use_cfg: rows are doubled (cond rows first, uncond rows second).
int satb_dit_prepare_cond(SatbDit* d, const float* cross, const float* neg_cross, const float* global, int B, int Mctx,
                          int use_cfg, void* stream_v) {
  SATB_REQUIRE(d && d->finalized, "weights not finalized");
  SATB_REQUIRE(B >= 1, "bad batch");
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  const int D = d->D;
  d->B = B;
  d->cfg_on = use_cfg != 0;
  d->has_cross = cross != nullptr && d->ct > 0;
  d->has_global = global != nullptr && d->gd > 0;
  d->Mctx = d->has_cross ? Mctx : 0;
  // rows with a non-null context: cond rows, plus the uncond rows iff a negative prompt is given
  // (a null (zero) context makes the bias-free cross-attention branch exactly 0: SURVEY.md H5)
  d->Rc = d->has_cross ? ((d->cfg_on && neg_cross) ? 2 * B : B) : 0;
  SATB_REQUIRE(!(neg_cross && !d->cfg_on), "negative conditioning requires CFG");
  SATB_PROPAGATE(d->ws_small.ensure(static_cast<size_t>(2 * B) * (2 * d->F + 4 * D + 6 * D * d->depth) * 4 + 4096));
  SmallWs sw = small_ws(d, 2 * B);
  if (d->has_global) {
    SATB_PROPAGATE(launch_skinny_linear(global, d->ge0_w, nullptr, nullptr, sw.ge_h, B, d->gd, d->ge, 1, st));
    SATB_PROPAGATE(launch_skinny_linear(sw.ge_h, d->ge2_w, nullptr, nullptr, sw.ge, B, d->ge, d->ge, 0, st));
  }
  if (d->has_cross) {
    SATB_REQUIRE(Mctx >= 1, "empty cross-attention context");
    const size_t rows = static_cast<size_t>(d->Rc) * Mctx;
    const size_t in_b = rows * d->ct * 2, mid_b = rows * d->ce * 2;
    SATB_PROPAGATE(d->ws_cond.ensure(in_b + 2 * mid_b + 1024));
    uint16_t* in16 = d->ws_cond.as<uint16_t>();
    uint16_t* mid16 = in16 + rows * d->ct;
    uint16_t* ce16 = mid16 + rows * d->ce;
    SATB_PROPAGATE(launch_cast_rows(cross, in16, nullptr, B * Mctx, d->ct, d->ct, d->ct, d->bf16, st));
    if (d->Rc == 2 * B)
      SATB_PROPAGATE(launch_cast_rows(neg_cross, in16 + static_cast<size_t>(B) * Mctx * d->ct, nullptr, B * Mctx,
                                      d->ct, d->ct, d->ct, d->bf16, st));
    SATB_PROPAGATE(d->ws_kv.ensure(static_cast<size_t>(d->depth) * rows * 2 * d->ce * 2));
    d->tmaps.maps.clear();
    const int Mr = static_cast<int>(rows);
    if (d->bf16) {
      typedef EpiStore16<true> E;
      SATB_PROPAGATE((linear<E, 128, true>(d->tmaps, in16, d->ct, Mr, d->ct, d->ce0_w, d->ce, E::Params{mid16, d->ce, nullptr, 1}, st)));
      SATB_PROPAGATE((linear<E, 128, true>(d->tmaps, mid16, d->ce, Mr, d->ce, d->ce2_w, d->ce, E::Params{ce16, d->ce, nullptr, 0}, st)));
      for (int i = 0; i < d->depth; ++i) {
        uint16_t* kv = d->ws_kv.as<uint16_t>() + static_cast<size_t>(i) * rows * 2 * d->ce;
        if (d->qk_norm) {
          typedef EpiHeadNorm16<true> EN;   // k heads normalised (transformer.py:433-436), v as is
          SATB_PROPAGATE((linear<EN, 128, true>(d->tmaps, ce16, d->ce, Mr, d->ce, d->layers[i].w_kv, 2 * d->ce,
                                                 EN::Params{kv, 2 * d->ce, d->ce, 0, 1, nullptr, nullptr}, st)));
        } else {
          SATB_PROPAGATE((linear<E, 128, true>(d->tmaps, ce16, d->ce, Mr, d->ce, d->layers[i].w_kv, 2 * d->ce, E::Params{kv, 2 * d->ce, nullptr, 0}, st)));
        }
      }
    } else {
      typedef EpiStore16<false> E;
      SATB_PROPAGATE((linear<E, 128, false>(d->tmaps, in16, d->ct, Mr, d->ct, d->ce0_w, d->ce, E::Params{mid16, d->ce, nullptr, 1}, st)));
      SATB_PROPAGATE((linear<E, 128, false>(d->tmaps, mid16, d->ce, Mr, d->ce, d->ce2_w, d->ce, E::Params{ce16, d->ce, nullptr, 0}, st)));
      for (int i = 0; i < d->depth; ++i) {
        uint16_t* kv = d->ws_kv.as<uint16_t>() + static_cast<size_t>(i) * rows * 2 * d->ce;
        if (d->qk_norm) {
          typedef EpiHeadNorm16<false> EN;   // k heads normalised (transformer.py:433-436), v as is
          SATB_PROPAGATE((linear<EN, 128, false>(d->tmaps, ce16, d->ce, Mr, d->ce, d->layers[i].w_kv, 2 * d->ce,
                                                 EN::Params{kv, 2 * d->ce, d->ce, 0, 1, nullptr, nullptr}, st)));
        } else {
          SATB_PROPAGATE((linear<E, 128, false>(d->tmaps, ce16, d->ce, Mr, d->ce, d->layers[i].w_kv, 2 * d->ce, E::Params{kv, 2 * d->ce, nullptr, 0}, st)));
        }
      }
    }
  }
  return 0;
}

}  // extern "C"

// RAII helper: records a start/stop event pair around a kernel sequence when profiling is on
struct ProfScope {
  SatbDit* d; int cat; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
  ProfScope(SatbDit* d_, int cat_, cudaStream_t st_) : d(d_), cat(cat_), st(st_) {
    if (d->prof_on) { a = d->prof_event(); b = d->prof_event(); cudaEventRecord(a, st); }
  }
  ~ProfScope() {
    if (a) { cudaEventRecord(b, st); d->prof_recs.push_back({cat, a, b}); }
  }
};
enum { PROF_FF_IN = 0, PROF_FF_OUT, PROF_QKV, PROF_ATTN_SELF, PROF_ATTN_OUT, PROF_CROSS, PROF_LN, PROF_OTHER, PROF_NCAT };

template <bool BF16>
static int dit_forward_impl(SatbDit* d, const float* x, const float* t, float* out, int B, int L, float cfg_scale,
                            float scale_phi, cudaStream_t st, float* hidden_out) {
  const int D = d->D, C = d->C, H = d->H, P = d->P;
  const int R = d->cfg_on ? 2 * B : B;
  const int N_seq = L + P;
  const int M = R * N_seq;
  const int Mc = d->Rc * N_seq;  // rows running cross-attention (a prefix of the row space)
  SmallWs sw = small_ws(d, 2 * d->B);
  float* h = d->ws_h.as<float>();
  uint16_t* a16 = d->ws_a16.as<uint16_t>();
  uint16_t* qkv = d->ws_qkv.as<uint16_t>();
  uint16_t* att = d->ws_attn.as<uint16_t>();
  uint16_t* q16 = d->ws_q16.as<uint16_t>();
  uint16_t* ff = d->ws_ff.as<uint16_t>();
  uint16_t* ain = d->ws_ain.as<uint16_t>();
  float* y = d->ws_y.as<float>();
  const float* cos_tab = d->ws_rope.as<float>();
  const float* sin_tab = cos_tab + static_cast<size_t>(N_seq) * d->nf;

  // timestep embedding (+ global embedding) -> conditioning token / adaLN vector  (dit.py:176-195)
  SATB_PROPAGATE(launch_fourier(t, d->ts_w, sw.fourier, B, d->F, st));
  // te_h = silu(W0 f + b0); tok = W2 te_h + b2 (+ global embed); in adaLN mode only silu(tok) is consumed
  SATB_PROPAGATE(launch_skinny_linear(sw.fourier, d->te0_w, d->te0_b, nullptr, sw.te_h, B, 2 * d->F, D, 1, st));
  SATB_PROPAGATE(launch_skinny_linear(sw.te_h, d->te2_w, d->te2_b, d->has_global ? sw.ge : nullptr, sw.tok, B, D, D,
                                      d->adaln ? 1 : 0, st));
  // latent -> token rows, project_in (with the 1x1 pre-conv folded), prepend token
  SATB_PROPAGATE(launch_dit_pre(x, ain, R, B, d->Cin, L, P, BF16, st));
  SATB_PROPAGATE((linear<EpiStore32, 256, BF16>(d->tmaps, ain, d->Cin, M, d->Cin, d->w_in16, D, EpiStore32::Params{h, D, nullptr}, st)));
  const int64_t ssg_ld = static_cast<int64_t>(d->depth) * 6 * D;
  if (P > 0) {
    SATB_PROPAGATE(launch_write_prepend(sw.tok, d->Pp > 0 ? d->ws_prep.as<float>() : nullptr, h, R, B, N_seq, D, d->Pp, st));
  } else {
    // adaLN: all layers' scale/shift/gate in one skinny GEMM (transformer.py:648-651,667)
    SATB_PROPAGATE(launch_skinny_linear(sw.tok, d->w_ssg, nullptr, nullptr, sw.ssg, B, D, d->depth * 6 * D, 0, st));
    SATB_PROPAGATE(launch_gate_sigmoid(sw.ssg, B, d->depth, D, st));
  }

  // Plain prepend-mode blocks run WITHOUT LayerNorm kernels (d->ln_fused): every residual GEMM (self out-proj, cross
  // out-proj, FF-out) writes h, the 16-bit x16 = h * gamma of the LayerNorm that follows and that row's partial
  // (sum, sum of squares); the GEMM after the LayerNorm reads x16 and applies mean / rstd in its epilogue (LnFold,
  // gemm.cuh).  Only the very first LayerNorm of the forward (block 0, whose input comes from project_in) is a kernel.
  // 8 launches per block instead of 11.  adaLN / qk_norm models keep the LayerNorm kernels below.
  const bool fused = d->ln_fused;
  float2* S1 = d->ws_stats.as<float2>();                 // partial row sums in front of: the QKV projection
  float2* S2 = S1 + static_cast<size_t>(M) * kLnSlots;   //   the cross-attention q projection (conditional rows)
  float2* S3 = S2 + static_cast<size_t>(M) * kLnSlots;   //   the feed-forward input projection
  auto fold = [&](const float2* st_, const float* c, const float* dd) {
    return LnFold{st_, c, dd, 1.0f / static_cast<float>(D), 1e-5f, kLnSlots};
  };
  const LnFold no_ln{nullptr, nullptr, nullptr, 0.f, 0.f, 0};
  for (int i = 0; i < d->depth; ++i) {
    const LayerW& W = d->layers[i];
    const float* ssg_l = d->adaln ? sw.ssg + static_cast<size_t>(i) * 6 * D : nullptr;
    // ---- self-attention: LN -> QKV GEMM (+RoPE) -> attention -> out-proj (+residual)
    if (!fused || i == 0) {
      ProfScope ps(d, PROF_LN, st);
      SATB_PROPAGATE(launch_layernorm(h, W.pre_g, W.pre_b, a16, M, D, ssg_l, ssg_l ? ssg_l + D : nullptr, ssg_ld, N_seq, B, BF16, st));
    }
    {
      ProfScope ps(d, PROF_QKV, st);
      if (d->qk_norm) {
        typedef EpiHeadNorm16<BF16> E;   // q, k heads L2-normalised, then rotary
        typename E::Params ep{qkv, 3 * D, 2 * D, 2 * D, N_seq, cos_tab, sin_tab};
        SATB_PROPAGATE((linear<E, 256, BF16>(d->tmaps, a16, D, M, D, W.w_qkv, 3 * D, ep, st)));
      } else {
        if (fused && i > 0) {
          typedef EpiQkvRope<BF16, true> E;
          typename E::Params ep{qkv, 3 * D, 2 * D, N_seq, cos_tab, sin_tab, fold(S1, W.c_qkv, W.d_qkv)};
          SATB_PROPAGATE((linear<E, 256, BF16>(d->tmaps, a16, D, M, D, W.w_qkv, 3 * D, ep, st)));
        } else {
          typedef EpiQkvRope<BF16> E;
          typename E::Params ep{qkv, 3 * D, 2 * D, N_seq, cos_tab, sin_tab, no_ln};
          SATB_PROPAGATE((linear<E, 256, BF16>(d->tmaps, a16, D, M, D, W.w_qkv, 3 * D, ep, st)));
        }
      }
    }
    {
      ProfScope ps(d, PROF_ATTN_SELF, st);
      const int64_t qs = static_cast<int64_t>(N_seq) * 3 * D;
      SATB_PROPAGATE(launch_attention_tc(qkv, qkv, qkv, att, 3 * D, 3 * D, 3 * D, D, qs, qs, qs,
                                         static_cast<int64_t>(N_seq) * D, 3 * D, 3 * D, 3 * D, 0, D, 2 * D, R, H, H,
                                         N_seq, N_seq, BF16, st));
    }
    {
      ProfScope ps(d, PROF_ATTN_OUT, st);
      if (fused) {
        // rows < Mc go on to the cross-attention LayerNorm, the others straight to the feed-forward LayerNorm
        typedef EpiResidualLN<BF16> E;
        typename E::Params ep{h, D, nullptr, a16, W.ca_g, W.ff_g, S2, S3, Mc};
        SATB_PROPAGATE((linear<E, 256, BF16>(d->tmaps, att, D, M, D, W.w_o, D, ep, st)));
      } else {
        EpiResidual::Params ep{h, D, nullptr, ssg_l ? ssg_l + 2 * D : nullptr, N_seq, static_cast<int>(ssg_ld), B};
        SATB_PROPAGATE((linear_auto<EpiResidual, BF16>(d->tmaps, att, D, M, D, W.w_o, D, ep, st)));
      }
    }
    // ---- cross-attention on the rows that have a non-null context
    if (Mc > 0) {
      ProfScope ps(d, PROF_CROSS, st);
      const int Hkv = d->ce / 64;
      if (!fused) SATB_PROPAGATE(launch_layernorm(h, W.ca_g, W.ca_b, a16, Mc, D, nullptr, nullptr, 0, N_seq, 1, BF16, st));
      if (d->qk_norm) {
        typedef EpiHeadNorm16<BF16> E;
        typename E::Params ep{q16, D, D, 0, N_seq, nullptr, nullptr};
        SATB_PROPAGATE((linear_auto<E, BF16>(d->tmaps, a16, D, Mc, D, W.w_q, D, ep, st)));
      } else {
        if (fused) {
          typedef EpiStore16<BF16, true> E;
          typename E::Params ep{q16, D, nullptr, 0, fold(S2, W.c_q, W.d_q)};
          SATB_PROPAGATE((linear_auto<E, BF16>(d->tmaps, a16, D, Mc, D, W.w_q, D, ep, st)));
        } else {
          typedef EpiStore16<BF16> E;
          typename E::Params ep{q16, D, nullptr, 0, no_ln};
          SATB_PROPAGATE((linear_auto<E, BF16>(d->tmaps, a16, D, Mc, D, W.w_q, D, ep, st)));
        }
      }
      const uint16_t* kv = d->ws_kv.as<uint16_t>() + static_cast<size_t>(i) * d->Rc * d->Mctx * 2 * d->ce;
      const int64_t kvs = static_cast<int64_t>(d->Mctx) * 2 * d->ce;
      SATB_PROPAGATE(launch_attention_tc(q16, kv, kv, att, D, 2 * d->ce, 2 * d->ce, D, static_cast<int64_t>(N_seq) * D,
                                         kvs, kvs, static_cast<int64_t>(N_seq) * D, D, 2 * d->ce, 2 * d->ce, 0, 0,
                                         d->ce, d->Rc, H, Hkv, N_seq, d->Mctx, BF16, st));
      if (fused) {
        typedef EpiResidualLN<BF16> E;
        typename E::Params ep{h, D, nullptr, a16, W.ff_g, W.ff_g, S3, S3, Mc};
        SATB_PROPAGATE((linear<E, 256, BF16>(d->tmaps, att, D, Mc, D, W.w_co, D, ep, st)));
      } else {
        EpiResidual::Params ep{h, D, nullptr, nullptr, N_seq, 0, 1};
        SATB_PROPAGATE((linear_auto<EpiResidual, BF16>(d->tmaps, att, D, Mc, D, W.w_co, D, ep, st)));
      }
    }
    // ---- feed-forward: LN -> GEMM (+bias, SwiGLU) -> GEMM (+bias, +residual)
    if (!fused) {
      ProfScope ps(d, PROF_LN, st);
      SATB_PROPAGATE(launch_layernorm(h, W.ff_g, W.ff_b, a16, M, D, ssg_l ? ssg_l + 3 * D : nullptr,
                                      ssg_l ? ssg_l + 4 * D : nullptr, ssg_ld, N_seq, B, BF16, st));
    }
    {
      ProfScope ps(d, PROF_FF_IN, st);
      if (fused) {
        typedef EpiSwiglu<BF16, true> E;
        typename E::Params ep{ff, d->ffi, W.b_ff1, fold(S3, W.c_ff1, W.d_ff1)};
        SATB_PROPAGATE((linear<E, 256, BF16>(d->tmaps, a16, D, M, D, W.w_ff1, 2 * d->ffi, ep, st)));
      } else {
        typedef EpiSwiglu<BF16> E;
        typename E::Params ep{ff, d->ffi, W.b_ff1, no_ln};
        SATB_PROPAGATE((linear<E, 256, BF16>(d->tmaps, a16, D, M, D, W.w_ff1, 2 * d->ffi, ep, st)));
      }
    }
    {
      ProfScope ps(d, PROF_FF_OUT, st);
      if (fused) {
        // prepares the next block's first LayerNorm; after the last block x16 is the plain 16-bit copy project_out reads
        const bool last = i + 1 == d->depth;
        typedef EpiResidualLN<BF16> E;
        typename E::Params ep{h, D, W.b_ff2, a16, last ? nullptr : d->layers[i + 1].pre_g, nullptr, last ? nullptr : S1, nullptr, M};
        SATB_PROPAGATE((linear<E, 256, BF16>(d->tmaps, ff, d->ffi, M, d->ffi, W.w_ff2, D, ep, st)));
      } else {
        EpiResidual::Params ep{h, D, W.b_ff2, ssg_l ? ssg_l + 5 * D : nullptr, N_seq, static_cast<int>(ssg_ld), B};
        SATB_PROPAGATE((linear_auto<EpiResidual, BF16>(d->tmaps, ff, d->ffi, M, d->ffi, W.w_ff2, D, ep, st)));
      }
    }
  }
  if (hidden_out)
    SATB_CHECK_CUDA(cudaMemcpyAsync(hidden_out, h, static_cast<size_t>(M) * D * 4, cudaMemcpyDeviceToDevice, st));
  // project_out (with the 1x1 post-conv folded) needs 16-bit input: the last FF-out epilogue already wrote it, or cast
  if (!fused || d->depth == 0) SATB_PROPAGATE(launch_cast_rows(h, a16, nullptr, M, D, D, D, BF16, st));
  SATB_PROPAGATE((linear<EpiStore32, 64, BF16>(d->tmaps, a16, D, M, D, d->w_out16, C, EpiStore32::Params{y, C, nullptr}, st)));
  SATB_PROPAGATE(launch_dit_post(y, out, B, C, L, N_seq, P, d->cfg_on ? 1 : 0, cfg_scale, scale_phi, st));
  return 0;
}

extern "C" {

// One denoiser call: x [B, C, L] fp32, t [B] fp32 -> out [B, C, L] fp32 (all device
// pointers, caller-owned).  Mirrors DiffusionTransformer.forward (models/dit.py:228-364)
// for the conditioning registered by satb_dit_prepare_cond.
int satb_dit_forward(SatbDit* d, const float* x, const float* t, float* out, int B, int L, float cfg_scale,
                     float scale_phi, void* stream_v) {
  SATB_REQUIRE(d && d->finalized, "weights not finalized");
  SATB_REQUIRE(B == d->B, "batch size differs from satb_dit_prepare_cond");
  const int R = d->cfg_on ? 2 * B : B;
  if (R > d->res_R || L != d->res_L || d->P != d->res_P) SATB_PROPAGATE(satb_dit_reserve(d, R, L));
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  return d->bf16 ? dit_forward_impl<true>(d, x, t, out, B, L, cfg_scale, scale_phi, st, nullptr)
                 : dit_forward_impl<false>(d, x, t, out, B, L, cfg_scale, scale_phi, st, nullptr);
}

// Per-category kernel timing with CUDA events on the launching stream (bench.py roofline):
// categories 0 ff_in GEMM, 1 ff_out GEMM, 2 qkv GEMM, 3 self-attention core, 4 attn out GEMM,
// 5 cross-attention (LN + q GEMM + core + out GEMM), 6 LayerNorm.
int satb_dit_profile(SatbDit* d, int enable) {
  SATB_REQUIRE(d, "null handle");
  d->prof_on = enable != 0;
  return 0;
}
// Synchronises, sums the recorded intervals per category into ms[8] / count[8], and clears them.
int satb_dit_profile_read(SatbDit* d, float* ms, int* count) {
  SATB_REQUIRE(d && ms && count, "null argument");
  for (int i = 0; i < PROF_NCAT; ++i) { ms[i] = 0.f; count[i] = 0; }
  for (auto& r : d->prof_recs) {
    SATB_CHECK_CUDA(cudaEventSynchronize(r.b));
    float t = 0.f;
    SATB_CHECK_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    ms[r.cat] += t;
    count[r.cat] += 1;
    d->prof_pool.push_back(r.a);
    d->prof_pool.push_back(r.b);
  }
  d->prof_recs.clear();
  return 0;
}

// Debug/test variant that also returns the residual stream after the last block
// ([R * (L + P), D] fp32) for comparison with the reference's hidden_states.
int satb_dit_forward_debug(SatbDit* d, const float* x, const float* t, float* out, float* hidden, int B, int L,
                           float cfg_scale, float scale_phi, void* stream_v) {
  SATB_REQUIRE(d && d->finalized, "weights not finalized");
  SATB_REQUIRE(B == d->B, "batch size differs from satb_dit_prepare_cond");
  const int R = d->cfg_on ? 2 * B : B;
  if (R > d->res_R || L != d->res_L || d->P != d->res_P) SATB_PROPAGATE(satb_dit_reserve(d, R, L));
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  return d->bf16 ? dit_forward_impl<true>(d, x, t, out, B, L, cfg_scale, scale_phi, st, hidden)
                 : dit_forward_impl<false>(d, x, t, out, B, L, cfg_scale, scale_phi, st, hidden);
}

}  // extern "C"
