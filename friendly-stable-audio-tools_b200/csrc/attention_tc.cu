// tcgen05 attention forward, head dim 64, no mask, non-causal, optional GQA:
//   O = softmax(Q K^T / sqrt(64)) V      (reference models/transformer.py:496-536)
//
// One CTA = 128 query rows of one (batch item, head); two CTAs are resident per SM.  Keys are
// processed in tiles of 64 with double-buffered S and P in TMEM, so the tensor pipe runs ahead of
// the softmax warps:
//   warp 1 (one thread): S[b] = Q K_j^T -> TMEM (tcgen05.mma, smem operands), two tiles ahead
//   warp 3 (one thread): O += P[b] V_j (A = P from TMEM, B = V tile addressed MN-major); the two
//                        products are issued by different threads because a tcgen05.mma issue
//                        costs ~100 cycles of the issuing thread (measured, profiles/)
//   warps 4-7 (one query row per thread): one pass over S[b]: P = exp2(S*c - m_ref) -> TMEM as
//                        packed 16-bit pairs, row sum (fp32) and raw row max in registers
// TMEM columns: S0 [0,64) S1 [64,128) P0 [128,160) P1 [160,192) O [192,256).
// Q/K/V tiles arrive by TMA (128B swizzle, out-of-range rows zero-filled); K tiles are released
// when Q K^T retires, V tiles when P V retires (independent rings).  O stays in TMEM for the whole
// pass: the reference max m_ref only moves when a tile's row max exceeds it by more than 2^8
// (lazy rescale: exponentials stay <= 256, sums in fp32), so the O rescale (TMEM load-scale-store)
// and the recomputation of that tile's P are rare.
#include "common.cuh"
#include "gemm.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace satb {

namespace {

constexpr int kQ = 128;        // query rows per CTA
constexpr int kK = 64;         // keys per tile
constexpr int kD = 64;         // head dim
constexpr int kStagesK = 4;
constexpr int kStagesV = 3;
constexpr int kQBytes = kQ * kD * 2;                         // 16 KB
constexpr int kKVBytes = kK * kD * 2;                        // 8 KB
constexpr int kAttnSmem = kQBytes + (kStagesK + kStagesV) * kKVBytes + 1024 + 256 + 2048;   // + row exchange [2][2][128] fp32
constexpr int kTmemColsAttn = 256;
constexpr uint32_t kColS = 0, kColP = 128, kColO = 192;      // S[b] at kColS + 64 b, P[b] at kColP + 32 b
constexpr float kRescaleThreshold = 8.0f;                    // log2 units

struct AttnTcArgs {
  uint16_t* o;
  int64_t ldo, o_bs;
  int Nq, Nk, group;
  int q_col, k_col, v_col;   // column offsets (elements) of head 0 inside the q / k / v tensor maps
  float scale_log2;
  unsigned long long* dbg;   // optional per-phase clock64 trace of one CTA (profiles/, tests only)
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

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Packed fp32x2 arithmetic (FFMA2 / FADD2 on sm_100): the softmax warps are issue-bound, so the
// scale-and-shift and the row-sum accumulation process two elements per instruction.
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(__float_as_uint(lo)), "r"(__float_as_uint(hi)));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  uint32_t a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v));
  lo = __uint_as_float(a);
  hi = __uint_as_float(b);
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

template <bool BF16>
__global__ void __launch_bounds__(256, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttnTcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kQBytes;
  uint8_t* sV = sK + kStagesK * kKVBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kStagesV * kKVBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + kStagesK;
  uint64_t* v_full = k_empty + kStagesK;
  uint64_t* v_empty = v_full + kStagesV;
  uint64_t* s_full = v_empty + kStagesV;   // [2] MMA -> softmax: S[b] holds Q K_j^T
  uint64_t* p_ready = s_full + 2;          // [2] softmax -> MMA: S[b] consumed, P[b] written
  uint64_t* pv_done = p_ready + 2;         // [2] MMA -> softmax: P[b] V retired (P[b] free, O updated)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

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
    for (int i = 0; i < kStagesK; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < kStagesV; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 128);
      mbar_init(&pv_done[i], 1);
    }
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
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      // ---------------------------------------------------------------- TMA producer
      mbar_expect_tx(q_full, kQBytes);
      tma_load_4d(sQ, &tmQ, q_full, p.q_col + h * kD, 0, q0, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int sk = j % kStagesK, sv = j % kStagesV;
        mbar_wait(&k_empty[sk], ((j / kStagesK) & 1) ^ 1);
        mbar_expect_tx(&k_full[sk], kKVBytes);
        tma_load_4d(sK + sk * kKVBytes, &tmK, &k_full[sk], p.k_col + hk * kD, 0, j * kK, b);
        mbar_wait(&v_empty[sv], ((j / kStagesV) & 1) ^ 1);
        mbar_expect_tx(&v_full[sv], kKVBytes);
        tma_load_4d(sV + sv * kKVBytes, &tmV, &v_full[sv], p.v_col + hk * kD, 0, j * kK, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ------------------------------------------------- MMA issuer 1: S[b] = Q K_j^T, two tiles ahead
      const uint32_t q_addr = smem_u32(sQ);
      mbar_wait(q_full, 0);
      tc_fence_after();
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % kStagesK, sb = j & 1;
        if (j >= 2) {
          mbar_wait(&p_ready[sb], ((j - 2) >> 1) & 1);   // S[sb] of tile j-2 has been consumed
          tc_fence_after();
        }
        mbar_wait(&k_full[st], (j / kStagesK) & 1);
        tc_fence_after();
        const int nk = min(kK, p.Nk - j * kK);
        const int n_mma = (nk + 15) & ~15;
        const uint32_t idesc = make_idesc_f16(kQ, n_mma, BF16);
        const uint32_t k_addr = smem_u32(sK + st * kKVBytes);
#pragma unroll
        for (int ks = 0; ks < kD / 16; ++ks)
          umma_f16_ss(tmem_base + kColS + sb * kK, make_desc_kmajor_sw128(q_addr + ks * 32),
                      make_desc_kmajor_sw128(k_addr + ks * 32), idesc, ks != 0);
        umma_commit(&k_empty[st]);   // the K tile is free as soon as these MMAs retire
        umma_commit(&s_full[sb]);
      }
    }
  } else if (warp == 3) {
    if (elect_one()) {
      // ------------------------------------------------- MMA issuer 2: O += P[b] V_j
      constexpr uint32_t idesc_pv = make_idesc_f16(kQ, kD, BF16, /*b_mn_major=*/true);
      const bool trace = p.dbg != nullptr && blockIdx.x == 3 && blockIdx.y == 11 && blockIdx.z == (gridDim.z / 2);
      for (int j = 0; j < n_tiles; ++j) {
        const int sb = j & 1;
        if (trace) p.dbg[j * 12 + 6] = clock64();
        mbar_wait(&p_ready[sb], (j >> 1) & 1);   // P[sb] written
        tc_fence_after();
        if (trace) p.dbg[j * 12 + 7] = clock64();
        const int st = j % kStagesV;
        mbar_wait(&v_full[st], (j / kStagesV) & 1);
        tc_fence_after();
        if (trace) p.dbg[j * 12 + 8] = clock64();
        const int nk = min(kK, p.Nk - j * kK);
        const int ksteps = (nk + 15) >> 4;
        const uint32_t v_addr = smem_u32(sV + st * kKVBytes);
        for (int ks = 0; ks < ksteps; ++ks)
          umma_f16_ts(tmem_base + kColO, tmem_base + kColP + sb * (kK / 2) + ks * 8,
                      make_desc_mnmajor_sw128(v_addr + ks * 2048), idesc_pv, (j | ks) != 0);
        umma_commit(&pv_done[sb]);
        umma_commit(&v_empty[st]);
        if (trace) p.dbg[j * 12 + 9] = clock64();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------- softmax + epilogue (1 row / thread)
    const int q = warp - 4;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float m_ref = -INFINITY, l = 0.f;
    const float sc = p.scale_log2;
    // One pass over S[sb]: P = exp2(S*c - m_ref) -> P[sb]; returns the row sum, tracks the raw max.
    auto pass_p = [&](int sb, int nk, float& mx_raw) -> float {
      float sum = 0.f;
      uint32_t r0[32], r1[32];
      const uint32_t s_addr = t_lane + kColS + sb * kK, p_addr = t_lane + kColP + sb * (kK / 2);
      auto do_chunk = [&](int c, const uint32_t (&r)[32]) {
        uint32_t w[16];
        if ((c + 1) * 32 <= nk) {
          const uint64_t sc2 = pack2(sc, sc), nm2 = pack2(-m_ref, -m_ref);
          uint64_t sum2 = pack2(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float s0 = __uint_as_float(r[2 * i]), s1 = __uint_as_float(r[2 * i + 1]);
            mx_raw = fmaxf(fmaxf(mx_raw, s0), s1);
            float t0, t1;
            unpack2(ffma2(pack2(s0, s1), sc2, nm2), t0, t1);
            const float p0 = ex2_approx(t0), p1 = ex2_approx(t1);
            sum2 = fadd2(sum2, pack2(p0, p1));
            w[i] = Op16<BF16>::pack(p0, p1);
          }
          float a0, a1;
          unpack2(sum2, a0, a1);
          sum += a0 + a1;
        } else {
          const int lim = nk - c * 32;   // valid columns in this (last) chunk
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float s0 = 2 * i < lim ? __uint_as_float(r[2 * i]) : -INFINITY;
            const float s1 = 2 * i + 1 < lim ? __uint_as_float(r[2 * i + 1]) : -INFINITY;
            mx_raw = fmaxf(fmaxf(mx_raw, s0), s1);
            const float p0 = ex2_approx(fmaf(s0, sc, -m_ref));
            const float p1 = ex2_approx(fmaf(s1, sc, -m_ref));
            sum += p0 + p1;
            w[i] = Op16<BF16>::pack(p0, p1);
          }
        }
        tmem_st_32x16(p_addr + c * 16, w);
      };
      tmem_ld_32x32(s_addr, r0);
      if (nk > 32) tmem_ld_32x32(s_addr + 32, r1);
      tmem_ld_wait();
      do_chunk(0, r0);
      if (nk > 32) do_chunk(1, r1);
      return sum;
    };
    const bool trace = p.dbg != nullptr && blockIdx.x == 3 && blockIdx.y == 11 && blockIdx.z == (gridDim.z / 2) &&
                       warp == 4 && lane == 0;
#define SATB_TRACE(idx) do { if (trace) p.dbg[j * 12 + (idx)] = clock64(); } while (0)
    // a warp whose 32 rows are all beyond Nq (ragged last query tile: 1025 = 8*128 + 1) only keeps
    // the barrier protocol going; its P rows are never read back through O
    const bool warp_active = (q0 + q * 32) < p.Nq;
    for (int j = 0; j < n_tiles; ++j) {
      const int sb = j & 1;
      const int nk = min(kK, p.Nk - j * kK);
      SATB_TRACE(0);
      mbar_wait(&s_full[sb], (j >> 1) & 1);
      tc_fence_after();
      SATB_TRACE(1);
      if (!warp_active) {
        if (j >= 2) mbar_wait(&pv_done[sb], ((j - 2) >> 1) & 1);
        tc_fence_before();
        mbar_arrive(&p_ready[sb]);
        continue;
      }
      if (j == 0) {
        // the first tile fixes the reference max before any exponential is taken
        float mx = -INFINITY;
        for (int c = 0; c * 32 < nk; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(t_lane + kColS + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < nk) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
        m_ref = mx * sc;
      }
      if (j >= 2) {
        mbar_wait(&pv_done[sb], ((j - 2) >> 1) & 1);   // P[sb] V_{j-2} retired: P[sb] is free
        tc_fence_after();
      }
      SATB_TRACE(2);
      float mx_raw = -INFINITY;
      float sum = pass_p(sb, nk, mx_raw);
      SATB_TRACE(3);
      // lazy rescale: only when this tile's max exceeds the reference max by more than 2^8
      const bool need = mx_raw * sc > m_ref + kRescaleThreshold;
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = need ? mx_raw * sc : m_ref;
        const float f = ex2_approx(m_ref - m_new);   // 1 for rows that keep their reference
        m_ref = m_new;
        l *= f;
        if (j > 0) {
          mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);   // every earlier P V has retired
          tc_fence_after();
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
        }
        float dummy = -INFINITY;
        sum = pass_p(sb, nk, dummy);   // P again with the new reference max
      }
      l += sum;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_ready[sb]);
      SATB_TRACE(5);
    }
#undef SATB_TRACE
    // epilogue: O / l -> global (128 B per row)
    mbar_wait(&pv_done[(n_tiles - 1) & 1], ((n_tiles - 1) >> 1) & 1);
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

int make_tmap_rows(CUtensorMap* m, const void* ptr, int cols, int rows, int batches, int64_t ld, int64_t bs,
                   int box_rows) {
  // (cols, rows, batches) 16-bit, box (64, box_rows, 1): reuse the A-operand encoder (phase dim = 1)
  return make_tmap_a(m, ptr, cols, rows, batches, ld, bs, 1, box_rows);
}

}  // namespace

// q / k / v are 16-bit row-major buffers [batch, rows, cols] with row strides ld* and batch strides
// *_bs (elements); head h of q lives at columns q_col + h*64 (k, v likewise with the kv head).
// For the fused QKV buffer pass the same pointer three times with different column offsets.
int launch_attention_tc(const void* q, const void* k, const void* v, void* o, int64_t ldq, int64_t ldk, int64_t ldv,
                        int64_t ldo, int64_t q_bs, int64_t k_bs, int64_t v_bs, int64_t o_bs, int q_cols, int k_cols,
                        int v_cols, int q_col, int k_col, int v_col, int batch, int H, int H_kv, int Nq, int Nk,
                        bool bf16, cudaStream_t stream, unsigned long long* dbg) {
  SATB_REQUIRE(H % H_kv == 0, "num_heads must be a multiple of kv heads");
  SATB_REQUIRE(Nk >= 1 && Nq >= 1, "empty attention problem");
  SATB_REQUIRE(ldo % 8 == 0, "attention output stride must be 16B aligned");
  CUtensorMap tq, tk, tv;
  SATB_PROPAGATE(make_tmap_rows(&tq, q, q_cols, Nq, batch, ldq, q_bs, kQ));
  SATB_PROPAGATE(make_tmap_rows(&tk, k, k_cols, Nk, batch, ldk, k_bs, kK));
  SATB_PROPAGATE(make_tmap_rows(&tv, v, v_cols, Nk, batch, ldv, v_bs, kK));
  AttnTcArgs a;
  a.o = static_cast<uint16_t*>(o);
  a.ldo = ldo; a.o_bs = o_bs;
  a.Nq = Nq; a.Nk = Nk; a.group = H / H_kv;
  a.q_col = q_col; a.k_col = k_col; a.v_col = v_col;
  a.scale_log2 = (1.0f / sqrtf(64.0f)) * 1.4426950408889634f;
  a.dbg = dbg;
  dim3 grid(ceil_div(Nq, kQ), H, batch);
  if (bf16) {
    static bool set = false;
    if (!set) { SATB_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem)); set = true; }
    SATB_CHECK_CUDA(launch_pdl(attn_tc_kernel<true>, grid, dim3(256), kAttnSmem, stream, tq, tk, tv, a));
  } else {
    static bool set = false;
    if (!set) { SATB_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem)); set = true; }
    SATB_CHECK_CUDA(launch_pdl(attn_tc_kernel<false>, grid, dim3(256), kAttnSmem, stream, tq, tk, tv, a));
  }
  count_launch();
  SATB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace satb
