/*
 * satb200 - C ABI of the B200-native Stable Audio denoising hot path.
 *
 * The reference (yukara-ikemiya/friendly-stable-audio-tools) is pure Python/PyTorch and has
 * no FFI of its own; the boundary it offers for this path is a Python object contract
 * (SURVEY.md 8b).  These entry points are what the drop-in Python modules bind through
 * ctypes (INTEGRATION.md shows the stub) and each one names the reference interface it
 * replaces.  Conventions:
 *   - every pointer argument documented as "device" is a CUDA device pointer in
 *     caller-owned storage (e.g. torch.Tensor.data_ptr()); fp32 unless stated;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); all work is
 *     enqueued asynchronously on it;
 *   - return value: 0 = ok, negative = error (message: satb_last_error());
 *   - handles are not re-entrant; use one handle per (device, model).
 */
#ifndef SATB200_H_
#define SATB200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define SATB_ABI_VERSION 3

typedef struct SatbDit SatbDit;
typedef struct SatbOobleck SatbOobleck;

/* Mirrors the constructor kwargs of DiffusionTransformer
 * (reference stable_audio_tools/models/dit.py:15-30) for the continuous_transformer backbone. */
typedef struct SatbDitConfig {
  int io_channels;
  int embed_dim;
  int depth;
  int num_heads;
  int cond_token_dim;       /* 0 = no cross-attention */
  int global_cond_dim;      /* 0 = no global conditioning */
  int project_cond_tokens;  /* dit.py:54 */
  int project_global_cond;  /* dit.py:65 */
  int global_cond_type;     /* 0 = "prepend", 1 = "adaLN" (dit.py:29,185-204) */
  int patch_size;           /* must be 1 */
  int operand_dtype;        /* 0 = fp16 (the reference's autocast dtype), 1 = bf16 */
  int qk_norm;              /* 1 = L2-normalise q and k per head before RoPE / attention
                               (attn_kwargs.qk_norm, models/transformer.py:298,433-436) */
  int input_concat_dim;     /* extra input channels concatenated to x before the 1x1 pre-conv (dit.py:38,163-168); the
                               caller passes x with io_channels + input_concat_dim channels */
  int prepend_cond_dim;     /* width of the prepend conditioning tokens (dit.py:75-81), 0 = none */
} SatbDitConfig;

/* Mirrors OobleckEncoder/OobleckDecoder kwargs (models/autoencoders.py:119-194). */
#define SATB_MAX_STAGES 8
typedef struct SatbOobleckConfig {
  int in_channels;          /* audio channels (encoder input / decoder output) */
  int channels;
  int latent_dim;           /* decoder input channels / encoder output channels */
  int n_stages;             /* len(c_mults) == len(strides) */
  int c_mults[SATB_MAX_STAGES];
  int strides[SATB_MAX_STAGES];
  int final_tanh;           /* decoder only */
  int is_decoder;           /* 1 = OobleckDecoder, 0 = OobleckEncoder */
  int operand_dtype;        /* 0 = fp16, 1 = bf16 (conv operands; accumulation is fp32), 2 = fp16 split hi + lo:
                             * three MMAs per product, ~fp32 accuracy (the reference's own precision for this path) */
} SatbOobleckConfig;

/* ---- library ---------------------------------------------------------------------- */
const char* satb_last_error(void);
int satb_abi_version(void);
unsigned long long satb_launch_count(void);   /* kernels launched by this library so far */
void satb_reset_launch_count(void);
/* Replaying a captured CUDA graph launches kernels this library cannot count: the caller adds the number recorded
 * while the graph was captured. */
void satb_add_launch_count(unsigned long long n);

/* ---- DiT: replaces DiffusionTransformer (models/dit.py:14-364) + ContinuousTransformer
 *      (models/transformer.py:705-809) behind DiTWrapper.forward (models/diffusion.py:491-529) */
int satb_dit_create(const SatbDitConfig* cfg, SatbDit** out);
void satb_dit_destroy(SatbDit* h);
/* One state-dict entry (key relative to DiffusionTransformer, e.g.
 * "transformer.layers.0.self_attn.to_qkv.weight"); src: device fp32, contiguous.
 * Replaces nn.Module.load_state_dict for this module (models/pretrained.py:24). */
int satb_dit_load_weight(SatbDit* h, const char* name, const float* src, long long numel, void* stream);
int satb_dit_finalize(SatbDit* h, void* stream);
/* Pre-allocate activations for `rows` transformer rows (2*B under CFG) of L latent tokens. */
int satb_dit_reserve(SatbDit* h, int rows, int L);
/* Step-invariant conditioning (dit.py:149-154 to_cond_embed / to_global_embed, and every
 * layer's cross-attention to_kv, transformer.py:425): cross [B, Mctx, cond_token_dim],
 * neg_cross (same shape, or NULL), global [B, global_cond_dim] (or NULL); device fp32. */
/* Prepend conditioning for the NEXT satb_dit_prepare_cond (dit.py:157-161,185-195,309-311): prepend [B, n_tokens,
 * prepend_cond_dim] device fp32 (NULL / 0 tokens = none).  Its to_prepend_embed tokens go in front of the
 * global-conditioning token; the unconditional CFG rows get zeros.  "prepend" global_cond_type only. */
int satb_dit_set_prepend_cond(SatbDit* h, const float* prepend, int B, int n_tokens, void* stream);
int satb_dit_prepare_cond(SatbDit* h, const float* cross, const float* neg_cross, const float* global, int B,
                          int Mctx, int use_cfg, void* stream);
/* One denoiser call = DiffusionTransformer.forward (dit.py:228-364): x [B, C, L], t [B] ->
 * out [B, C, L]; CFG combine and rescale (dit.py:338-347) included when use_cfg was set. */
int satb_dit_forward(SatbDit* h, const float* x, const float* t, float* out, int B, int L, float cfg_scale,
                     float scale_phi, void* stream);
/* Same, additionally copying the residual stream after the last block
 * ([rows * (L + prepend), embed_dim]; = info["hidden_states"][-1], transformer.py:804-805). */
int satb_dit_forward_debug(SatbDit* h, const float* x, const float* t, float* out, float* hidden, int B, int L,
                           float cfg_scale, float scale_phi, void* stream);

/* Per-kernel-class CUDA-event timing used by bench.py's roofline line: enable, run forwards,
 * then read ms[8]/count[8] (0 ff_in GEMM, 1 ff_out GEMM, 2 qkv GEMM, 3 self-attention core,
 * 4 attention out GEMM, 5 cross-attention, 6 LayerNorm, 7 unused). */
int satb_dit_profile(SatbDit* h, int enable);
int satb_dit_profile_read(SatbDit* h, float* ms, int* count);

/* ---- primitives exposed for the drop-in modules and the parity tests ---------------- */
/* SnakeBeta.forward (models/blocks.py:330-358): x, y [B, C, T]; alpha, beta [C]. */
int satb_snake_beta(const float* x, const float* alpha, const float* beta, float* y, int B, int C, long long T,
                    int logscale, void* stream);
/* LayerNorm.forward (models/transformer.py:188-206) with 16-bit output (fp16/bf16 bits). */
int satb_layernorm(const float* x, const float* gamma, const float* beta, void* out16, int rows, int D, int bf16,
                   void* stream);
/* C[M, N] (fp32) = A[M, K] * W[N, K]^T, A and W 16-bit (fp16/bf16 bits) row-major: nn.Linear
 * without bias (F.linear call sites transformer.py:422-430,548). */
int satb_linear_f32out(const void* a16, const void* w16, float* c, int M, int N, int K, int bf16, void* stream);
/* One step of the v-objective k-diffusion samplers in a single pass over the latents (replaces the
 * ~20 elementwise torch kernels of K.external.VDenoiser.forward + sample_dpmpp_{2m,3m}_sde's update,
 * reference call sites inference/sampling.py:159,225-228): with v = model(x * c_in, t),
 *   den = c_out v + c_skip x;  x_next = a x + b den + c den_1 + d den_2 + s noise;  x_in_next = x_next * c_in_next.
 * den_1, den_2, noise, x_in_next may be NULL; all tensors fp32 with n elements (n % 4 == 0). */
int satb_sampler_update(const float* x, const float* v, const float* den_1, const float* den_2, const float* noise,
                        float* den, float* x_next, float* x_in_next, long long n, float c_skip, float c_out, float a,
                        float b, float c, float d, float s, float c_in_next, void* stream);
/* softmax(q k^T / sqrt(64)) v (transformer.py:496-536): q [B, Nq, H*64], k/v [B, Nk, Hkv*64],
 * out [B, Nq, H*64]; 16-bit, contiguous. */
int satb_attention(const void* q16, const void* k16, const void* v16, void* o16, int B, int H, int Hkv, int Nq, int Nk,
                   int bf16, void* stream);

/* Debug / profiling: resident CTAs per SM reported by the runtime for the attention kernel with the given dynamic
 * shared-memory size and carveout preference (percent, -1 = unchanged). */
int satb_debug_attention_occupancy(int dyn_smem, int carveout_pct);
/* satb_attention plus a clock64 trace (16 key tiles x 12 slots) of CTA 0's first softmax warp followed by one
 * (SM id, slot, start ns, end ns) record per CTA: dbg must hold 16 * 12 + 4 * gridDim entries (<= 1376). */
int satb_attention_trace(const void* q16, const void* k16, const void* v16, void* o16, int B, int H, int Hkv, int Nq,
                         int Nk, int bf16, unsigned long long* dbg, void* stream);

/* ---- Oobleck VAE: replaces OobleckDecoder / OobleckEncoder.forward
 *      (models/autoencoders.py:119-194) behind AudioAutoencoder.encode/decode (:268-343) */
int satb_oobleck_create(const SatbOobleckConfig* cfg, SatbOobleck** out);
void satb_oobleck_destroy(SatbOobleck* h);
/* State-dict entry relative to the encoder/decoder module ("layers.1.layers.1.weight_v" ...). */
int satb_oobleck_load_weight(SatbOobleck* h, const char* name, const float* src, long long numel, void* stream);
int satb_oobleck_finalize(SatbOobleck* h, void* stream);
/* Decoder: z [B, latent_dim, L] -> audio [B, in_channels, L * prod(strides)]. */
int satb_oobleck_decode(SatbOobleck* h, const float* z, float* audio, int B, int L, void* stream);
/* Encoder: audio [B, in_channels, T] -> pre-bottleneck [B, latent_dim, T / prod(strides)]
 * (the deterministic mean|scale tensor; the VAE sampling stays in PyTorch, bottleneck.py:46-62). */
int satb_oobleck_encode(SatbOobleck* h, const float* audio, float* latents, int B, long long T, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SATB200_H_ */
