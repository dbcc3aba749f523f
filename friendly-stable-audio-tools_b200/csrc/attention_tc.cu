// tcgen05 attention forward, head dim 64, no mask, non-causal, optional GQA:
//   O = softmax(Q K^T / sqrt(64)) V      (reference models/transformer.py:496-536)
//
// One CTA = 128 query rows of one (batch item, head); two CTAs are resident per SM so one
// CTA's softmax overlaps the other's MMAs.  Per 128-key tile:
//   warp 1 (one thread): S = Q K^T  -> TMEM cols [0,128)        tcgen05.mma, A/B from smem
//   warps 4-7 (one query row per thread): row max, then P = exp2(S*c - m) -> TMEM cols
//                        [128,192) as packed 16-bit pairs; row sums in registers (fp32)
//   warp 1: O += P V -> TMEM cols [192,256)                     tcgen05.mma, A from TMEM,
//                        B = V tile in smem addressed MN-major (V is [key][d], d contiguous)
// Q/K/V tiles arrive by TMA (128B swizzle, out-of-range rows zero-filled).  O stays in TMEM
// for the whole pass: the running max only moves when the new row max exceeds it by more
// than 2^8 (lazy rescale: exponentials stay <= 256, exact in fp16/bf16 range, sums in fp32),
// so the O rescale (TMEM load-scale-store) is rare.
#include "common.cuh"
#include "gemm.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace satb {

namespace {

constexpr int kQ = 128;        // query rows per CTA
constexpr int kK = 128;        // keys per tile
constexpr int kD = 64;         // head dim
constexpr int kStagesKV = 2;
constexpr int kTileBytes = kQ * kD * 2;                      // 16 KB
constexpr int kAttnSmem = kTileBytes * (1 + 2 * kStagesKV) + 1024 + 256;
constexpr int kTmemColsAttn = 256;
constexpr uint32_t kColS = 0, kColP = 128, kColO = 192;
constexpr float kRescaleThreshold = 8.0f;                    // log2 units

struct AttnTcArgs {
  uint16_t* o;
  int64_t ldo, o_bs;
  int Nq, Nk, group;
  int q_col, k_col, v_col;   // column offsets (elements) of head 0 inside the q / kv tensor maps
  float scale_log2;
};

// V tile as the MN-major B operand: rows = keys (K dim), 64 contiguous 16-bit d values (128 B,
// one swizzle atom) per row; 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_desc_mnmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <bool BF16>
__global__ void __launch_bounds__(256, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttnTcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + kTileBytes;  // [stage][K | V]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTileBytes * (1 + 2 * kStagesKV));
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + kStagesKV;
  uint64_t* s_full = kv_empty + kStagesKV;
  uint64_t* p_ready = s_full + 1;
  uint64_t* o_full = p_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kQ;
  const int h = blockIdx.y, b = blockIdx.z;
  const int hk = h / p.group;
  const int n_tiles = (p.Nk + kK - 1) / kK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kStagesKV; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 128);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemColsAttn);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      mbar_expect_tx(q_full, kTileBytes);
      tma_load_4d(sQ, &tmQ, q_full, p.q_col + h * kD, 0, q0, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % kStagesKV;
        const uint32_t ph = (j / kStagesKV) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1);
        uint8_t* sk = sKV + st * 2 * kTileBytes;
        mbar_expect_tx(&kv_full[st], 2 * kTileBytes);
        tma_load_4d(sk, &tmK, &kv_full[st], p.k_col + hk * kD, 0, j * kK, b);
        tma_load_4d(sk + kTileBytes, &tmV, &kv_full[st], p.v_col + hk * kD, 0, j * kK, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------------ MMA issuer
      const uint32_t q_addr = smem_u32(sQ);
      auto issue_qk = [&](int j) {
        const int st = j % kStagesKV;
        mbar_wait(&kv_full[st], (j / kStagesKV) & 1);
        tc_fence_after();
        const int nk = min(kK, p.Nk - j * kK);
        const int n_mma = (nk + 15) & ~15;
        const uint32_t idesc = make_idesc_f16(kQ, n_mma, BF16);
        const uint32_t k_addr = smem_u32(sKV + st * 2 * kTileBytes);
#pragma unroll
        for (int ks = 0; ks < kD / 16; ++ks)
          umma_f16_ss(tmem_base + kColS, make_desc_kmajor_sw128(q_addr + ks * 32),
                      make_desc_kmajor_sw128(k_addr + ks * 32), idesc, ks != 0);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_qk(0);
      constexpr uint32_t idesc_pv = make_idesc_f16(kQ, kD, BF16, /*b_mn_major=*/true);
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(p_ready, j & 1);   // S_j consumed, P_j written
        tc_fence_after();
        if (j + 1 < n_tiles) issue_qk(j + 1);
        const int st = j % kStagesKV;
        const int nk = min(kK, p.Nk - j * kK);
        const int ksteps = (nk + 15) >> 4;
        const uint32_t v_addr = smem_u32(sKV + st * 2 * kTileBytes + kTileBytes);
        for (int ks = 0; ks < ksteps; ++ks)
          umma_f16_ts(tmem_base + kColO, tmem_base + kColP + ks * 8, make_desc_mnmajor_sw128(v_addr + ks * 2048),
                      idesc_pv, (j | ks) != 0);
        umma_commit(o_full);
        umma_commit(&kv_empty[st]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------- softmax + epilogue (1 row / thread)
    const int q = warp - 4;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float m_used = -INFINITY, l = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const int nk = min(kK, p.Nk - j * kK);
      const int chunks = (nk + 31) >> 5;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // pass 1: row max of the valid keys
      float mx = -INFINITY;
      for (int c = 0; c < chunks; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_lane + kColS + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c * 32 + i < nk) mx = fmaxf(mx, __uint_as_float(r[i]));
      }
      mx *= p.scale_log2;
      if (j == 0) {
        m_used = mx;
      } else {
        mbar_wait(o_full, (j - 1) & 1);   // P V of the previous tile retired: P and O are free
        tc_fence_after();
        const bool need = mx > m_used + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          const float f = need ? exp2f(m_used - mx) : 1.0f;
          if (need) m_used = mx;
          l *= f;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(t_lane + kColO + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t w[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) w[i] = __float_as_uint(__uint_as_float(r[half * 16 + i]) * f);
              tmem_st_32x16(t_lane + kColO + c * 32 + half * 16, w);
            }
          }
          tmem_st_wait();
        }
      }
      // pass 2: P = exp2(S*c - m), packed 16-bit pairs -> TMEM; row sum in fp32
      for (int c = 0; c < chunks; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_lane + kColS + c * 32, r);
        tmem_ld_wait();
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int k0 = c * 32 + 2 * i;
          const float p0 = k0 < nk ? exp2f(fmaf(__uint_as_float(r[2 * i]), p.scale_log2, -m_used)) : 0.f;
          const float p1 = k0 + 1 < nk ? exp2f(fmaf(__uint_as_float(r[2 * i + 1]), p.scale_log2, -m_used)) : 0.f;
          l += p0 + p1;
          w[i] = Op16<BF16>::pack(p0, p1);
        }
        tmem_st_32x16(t_lane + kColP + c * 16, w);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_ready);
    }
    // epilogue: O / l -> global (128 B per row)
    mbar_wait(o_full, (n_tiles - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l;
    const bool valid = (q0 + row) < p.Nq;
    uint16_t* og = p.o + b * p.o_bs + static_cast<int64_t>(q0 + row) * p.ldo + static_cast<int64_t>(h) * kD;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(t_lane + kColO + c * 32, r);
      tmem_ld_wait();
      if (valid) {
        uint4* dst = reinterpret_cast<uint4*>(og + c * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          dst[i] = make_uint4(Op16<BF16>::pack(__uint_as_float(r[8 * i]) * inv, __uint_as_float(r[8 * i + 1]) * inv),
                              Op16<BF16>::pack(__uint_as_float(r[8 * i + 2]) * inv, __uint_as_float(r[8 * i + 3]) * inv),
                              Op16<BF16>::pack(__uint_as_float(r[8 * i + 4]) * inv, __uint_as_float(r[8 * i + 5]) * inv),
                              Op16<BF16>::pack(__uint_as_float(r[8 * i + 6]) * inv, __uint_as_float(r[8 * i + 7]) * inv));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemColsAttn);
  }
}

int make_tmap_rows(CUtensorMap* m, const void* ptr, int cols, int rows, int batches, int64_t ld, int64_t bs) {
  // (cols, rows, batches) 16-bit, box (64, 128, 1): reuse the A-operand encoder (phase dim = 1)
  return make_tmap_a(m, ptr, cols, rows, batches, ld, bs, 1);
}

}  // namespace

// q / k / v are 16-bit row-major buffers [batch, rows, cols] with row strides ld* and batch strides
// *_bs (elements); head h of q lives at columns q_col + h*64 (k, v likewise with the kv head).
// For the fused QKV buffer pass the same pointer three times with different column offsets.
int launch_attention_tc(const void* q, const void* k, const void* v, void* o, int64_t ldq, int64_t ldk, int64_t ldv,
                        int64_t ldo, int64_t q_bs, int64_t k_bs, int64_t v_bs, int64_t o_bs, int q_cols, int k_cols,
                        int v_cols, int q_col, int k_col, int v_col, int batch, int H, int H_kv, int Nq, int Nk,
                        bool bf16, cudaStream_t stream) {
  SATB_REQUIRE(H % H_kv == 0, "num_heads must be a multiple of kv heads");
  SATB_REQUIRE(Nk >= 1 && Nq >= 1, "empty attention problem");
  SATB_REQUIRE(ldo % 8 == 0, "attention output stride must be 16B aligned");
  CUtensorMap tq, tk, tv;
  SATB_PROPAGATE(make_tmap_rows(&tq, q, q_cols, Nq, batch, ldq, q_bs));
  SATB_PROPAGATE(make_tmap_rows(&tk, k, k_cols, Nk, batch, ldk, k_bs));
  SATB_PROPAGATE(make_tmap_rows(&tv, v, v_cols, Nk, batch, ldv, v_bs));
  AttnTcArgs a;
  a.o = static_cast<uint16_t*>(o);
  a.ldo = ldo; a.o_bs = o_bs;
  a.Nq = Nq; a.Nk = Nk; a.group = H / H_kv;
  a.q_col = q_col; a.k_col = k_col; a.v_col = v_col;
  a.scale_log2 = (1.0f / sqrtf(64.0f)) * 1.4426950408889634f;
  dim3 grid(ceil_div(Nq, kQ), H, batch);
  if (bf16) {
    static bool set = false;
    if (!set) { SATB_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem)); set = true; }
    attn_tc_kernel<true><<<grid, 256, kAttnSmem, stream>>>(tq, tk, tv, a);
  } else {
    static bool set = false;
    if (!set) { SATB_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem)); set = true; }
    attn_tc_kernel<false><<<grid, 256, kAttnSmem, stream>>>(tq, tk, tv, a);
  }
  count_launch();
  SATB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace satb
