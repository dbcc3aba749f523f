// C-ABI entry points for the stand-alone primitives (see include/satb200.h).
#include "../../include/satb200.h"
#include "common.cuh"
#include "gemm.cuh"
#include "kernels.h"

using namespace satb;

extern "C" {

int satb_snake_beta(const float* x, const float* alpha, const float* beta, float* y, int B, int C, long long T,
                    int logscale, void* stream) {
  SATB_REQUIRE(x && alpha && beta && y, "null argument");
  return launch_snake_beta(x, alpha, beta, y, B, C, T, logscale, static_cast<cudaStream_t>(stream));
}

int satb_layernorm(const float* x, const float* gamma, const float* beta, void* out16, int rows, int D, int bf16,
                   void* stream) {
  SATB_REQUIRE(x && gamma && out16, "null argument");
  return launch_layernorm(x, gamma, beta, out16, rows, D, nullptr, nullptr, 0, 1, 1, bf16 != 0,
                          static_cast<cudaStream_t>(stream));
}

int satb_linear_f32out(const void* a16, const void* w16, float* c, int M, int N, int K, int bf16, void* stream) {
  SATB_REQUIRE(a16 && w16 && c, "null argument");
  SATB_REQUIRE(M >= 1 && N >= 32 && N % 32 == 0 && K >= 8 && K % 8 == 0, "linear: need N % 32 == 0 and K % 8 == 0");
  CUtensorMap ta, tb;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GemmShape s;
  s.L = M; s.batches = 1; s.N = N; s.K = K; s.n_taps = 1; s.tap_base = 0; s.tap_step = 0; s.b_tap_rows = N; s.stride = 1;
  EpiStore32::Params ep{c, N, nullptr};
  SATB_PROPAGATE(make_tmap_a(&ta, a16, K, M, 1, K, static_cast<int64_t>(M) * K));
  if ((N % 256 == 0 || N > 256) && gemm_use_2cta() && M >= 1024) {
    SATB_PROPAGATE(make_tmap_b(&tb, w16, K, N, K, 128));
    return bf16 ? launch_gemm_2cta<EpiStore32, 256, true>(ta, tb, s, ep, st) : launch_gemm_2cta<EpiStore32, 256, false>(ta, tb, s, ep, st);
  }
  if (N % 256 == 0 || N > 256) {
    SATB_PROPAGATE(make_tmap_b(&tb, w16, K, N, K, 256));
    return bf16 ? launch_gemm<EpiStore32, 256, true>(ta, tb, s, ep, st) : launch_gemm<EpiStore32, 256, false>(ta, tb, s, ep, st);
  } else if (N > 64) {
    SATB_PROPAGATE(make_tmap_b(&tb, w16, K, N, K, 128));
    return bf16 ? launch_gemm<EpiStore32, 128, true>(ta, tb, s, ep, st) : launch_gemm<EpiStore32, 128, false>(ta, tb, s, ep, st);
  }
  SATB_PROPAGATE(make_tmap_b(&tb, w16, K, N, K, 64));
  return bf16 ? launch_gemm<EpiStore32, 64, true>(ta, tb, s, ep, st) : launch_gemm<EpiStore32, 64, false>(ta, tb, s, ep, st);
}

int satb_sampler_update(const float* x, const float* v, const float* den_1, const float* den_2, const float* noise,
                        float* den, float* x_next, float* x_in_next, long long n, float c_skip, float c_out, float a,
                        float b, float c, float d, float s, float c_in_next, void* stream) {
  SATB_REQUIRE(x && v && den && x_next, "null argument");
  return launch_sampler_update(x, v, den_1, den_2, noise, den, x_next, x_in_next, n, c_skip, c_out, a, b, c, d, s,
                               c_in_next, static_cast<cudaStream_t>(stream));
}

int satb_attention(const void* q16, const void* k16, const void* v16, void* o16, int B, int H, int Hkv, int Nq, int Nk,
                   int bf16, void* stream) {
  SATB_REQUIRE(q16 && k16 && v16 && o16, "null argument");
  const int64_t dq = static_cast<int64_t>(H) * 64, dk = static_cast<int64_t>(Hkv) * 64;
  return launch_attention_tc(q16, k16, v16, o16, dq, dk, dk, dq, Nq * dq, Nk * dk, Nk * dk, Nq * dq, static_cast<int>(dq),
                             static_cast<int>(dk), static_cast<int>(dk), 0, 0, 0, B, H, Hkv, Nq, Nk, bf16 != 0,
                             static_cast<cudaStream_t>(stream));
}

int satb_debug_attention_occupancy(int dyn_smem, int carveout_pct) { return debug_attention_occupancy(dyn_smem, carveout_pct); }

// Debug: same as satb_attention, plus a clock64 trace of one CTA into dbg[tiles * 12] (device memory).
int satb_attention_trace(const void* q16, const void* k16, const void* v16, void* o16, int B, int H, int Hkv, int Nq,
                         int Nk, int bf16, unsigned long long* dbg, void* stream) {
  const int64_t dq = static_cast<int64_t>(H) * 64, dk = static_cast<int64_t>(Hkv) * 64;
  return launch_attention_tc(q16, k16, v16, o16, dq, dk, dk, dq, Nq * dq, Nk * dk, Nk * dk, Nq * dq, static_cast<int>(dq),
                             static_cast<int>(dk), static_cast<int>(dk), 0, 0, 0, B, H, Hkv, Nq, Nk, bf16 != 0,
                             static_cast<cudaStream_t>(stream), dbg);
}

}  // extern "C"
