// Internal launch API between translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace satb {

// ---- attention_tc.cu (tcgen05)
int launch_attention_tc(const void* q, const void* k, const void* v, void* o, int64_t ldq, int64_t ldk, int64_t ldv,
                        int64_t ldo, int64_t q_bs, int64_t k_bs, int64_t v_bs, int64_t o_bs, int q_cols, int k_cols,
                        int v_cols, int q_col, int k_col, int v_col, int batch, int H, int H_kv, int Nq, int Nk,
                        bool bf16, cudaStream_t stream, unsigned long long* dbg = nullptr);
int debug_attention_occupancy(int dyn_smem, int carveout_pct);
// debugging switches (environment variables; the defaults are the production path)
bool conv_halo_enabled();     // SATB_CONV_HALO=off: generic 7-tap loads for the final conv (A/B debugging)
bool conv_epi_masked();        // SATB_CONV_EPI=general keeps the combined-epilogue GEMM kernels for the 16-bit convolutions (A/B debugging)
bool resunit_use_fused();      // SATB_RESUNIT=unfused runs the 128-channel ResidualUnits as two GEMM launches (A/B debugging)
bool gemm_use_2cta();          // SATB_GEMM=1cta disables the CTA-pair GEMM (A/B debugging)
bool ln_fold_enabled();        // SATB_LN=fold: LayerNorm folded into the GEMM epilogues (A/B; measured slower, off by default)
bool raw_stream_16bit();       // SATB_RAW=fp32 keeps the Oobleck skip stream in fp32 (A/B debugging)

// ---- elementwise.cu
// LayerNorm over the last dim (eps 1e-5), optional adaLN modulation y*(1+scale)+shift, 16-bit output.
int launch_layernorm(const float* x, const float* gamma, const float* beta, void* out16, int rows, int D,
                     const float* scale, const float* shift, int64_t mod_stride, int rows_per_item, int n_items,
                     bool bf16, cudaStream_t stream);
// c[n] = sum_k W16[n,k] gamma[k], d[n] = sum_k W16[n,k] beta[k] (beta may be null -> d = 0): the per-column vectors of a
// LayerNorm folded into the GEMM that follows it (gemm.cuh: LnFold)
int launch_ln_fold_vectors(const void* w16, const float* gamma, const float* beta, float* c, float* d, int rows, int K,
                           bool bf16, cudaStream_t stream);
// Fused VDenoiser scaling + multistep sampler update + noise (see elementwise.cu).
int launch_sampler_update(const float* x, const float* v, const float* d1, const float* d2, const float* nz, float* den,
                          float* x_next, float* x_in, long long n, float c_skip, float c_out, float A, float B, float C,
                          float D, float S, float c_in_next, cudaStream_t stream);
// SnakeBeta on [B, C, T] fp32 (log-scale alpha/beta per channel).
int launch_snake_beta(const float* x, const float* alpha, const float* beta, float* y, int B, int C, int64_t T,
                      int logscale, cudaStream_t stream);
// x[B_src, C, L] fp32 -> a16[(r*N_seq + P + l), c] (rows r*N_seq .. +P-1 zero), r < R, source row r % B_src.
int launch_dit_pre(const float* x, void* a16, int R, int B_src, int C, int L, int P, bool bf16, cudaStream_t stream);
// Fourier timestep features [B, 2*F]: cat(cos(2*pi*t*w), sin(2*pi*t*w)).
int launch_fourier(const float* t, const float* w, float* out, int B, int F, cudaStream_t stream);
// out[r, n] = act_out(sum_k in[r, k] * W[n, k] + bias[n] (+ add[r, n])); fp32 weights; act_out = SiLU if silu_out.
int launch_skinny_linear(const float* in, const float* W, const float* bias, const float* add, float* out, int R,
                         int K, int N, int silu_out, cudaStream_t stream);
// h[r*N_seq + j, :] = pre[r, j, :] (j < Pp; zeros for rows r >= B or pre == null), h[r*N_seq + Pp, :] = tok[r % B, :]
int launch_write_prepend(const float* tok, const float* pre, float* h, int R, int B, int N_seq, int D, int Pp,
                         cudaStream_t stream);
// in place: x = sigmoid(1 - x) on column ranges [c0, c0+D) and [c1, c1+D) of every 6D-wide layer block
int launch_gate_sigmoid(float* ssg, int rows, int depth, int D, cudaStream_t stream);
// y[R*N_seq, C] fp32 -> out[B, C, L] with CFG combine / rescale (models/dit.py:338-347)
int launch_dit_post(const float* y, float* out, int B, int C, int L, int N_seq, int P, int cfg, float cfg_scale,
                    float scale_phi, cudaStream_t stream);
// generic fp32 -> 16-bit cast with row gather: dst[r, :] = src[perm ? perm[r] : r, :] * row_scale
int launch_cast_rows(const float* src, void* dst, const int* perm, int rows, int cols, int64_t src_ld, int64_t dst_ld,
                     bool bf16, cudaStream_t stream);
int launch_gather_f32(const float* src, float* dst, const int* perm, int n, cudaStream_t stream);
int launch_zero(void* p, size_t bytes, cudaStream_t stream);

}  // namespace satb
