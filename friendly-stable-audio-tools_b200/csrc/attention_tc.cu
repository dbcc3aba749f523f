// tcgen05 attention forward, head dim 64, no mask, non-causal, optional GQA:
//   O = softmax(Q K^T / sqrt(64)) V      (reference models/transformer.py:496-536)
//
// Persistent kernel, two CTAs per SM (2 x 256 TMEM columns, 2 x 85 KB shared memory).  A work unit is 128 query
// rows of one (batch item, head); CTA c processes units c, c + grid, ... without re-initialising anything.  Keys
// are processed in tiles of 128:
//   warp 0 (one thread)  TMA producer: Q of the unit, K / V tiles (2-stage rings)
//   warp 1 (one thread)  S = Q K_j^T -> TMEM (tcgen05.mma, smem operands) as soon as S has been read out
//   warp 3 (one thread)  O += P V_j (A = P from TMEM, B = V tile addressed MN-major)
//   warps 4-7            one query row per thread: P = exp2(S c - m_ref) -> TMEM as packed 16-bit pairs, row sum
//                        (fp32) and raw row max in registers; chunks of 32 columns are software-pipelined
//                        (tcgen05.ld of chunk c+2 in flight while chunk c is exponentiated).  S is handed back to the
//                        MMA thread right after its LAST chunk has been loaded into registers, so Q K_{j+1}^T runs
//                        under the exponentials of that chunk.
//   warp 2               TMEM allocator; afterwards the CUDA-core path for ragged query rows (below)
// TMEM columns: S [0,128)  P [128,192)  O [192,256).  The softmax is SFU-bound (16 ex2 / clk / SM, 202 M exponentials
// per SA-Open layer): what matters is that the two resident CTAs keep the SFU busy, i.e. that each warp's fixed
// per-tile cost (barrier waits, TMEM load / store latency) stays below the exponentiation time of a tile - hence
// 128-key tiles (1024 SFU cycles per warp and tile).
// O stays in TMEM for the whole unit: the reference max m_ref only moves when a tile's row max exceeds it by more
// than 2^8 (lazy rescale: exponentials stay <= 256, sums in fp32), so the O rescale (TMEM load-scale-store) and the
// recomputation of that tile's P are rare.
//
// Ragged query rows: with Nq = 1025 = 8 * 128 + 1 (the prepended conditioning token) a ninth tensor-core tile per
// (item, head) would hold ONE row and cost as much time as a full one.  When Nq % 128 <= kRowPathMax those rows are
// computed by warp 2 on CUDA cores instead (one warp per row: lane-per-key dot products, online softmax over
// blocks of 1024 keys, lane-per-two-dims P V), concurrently with the tensor-core pipeline of the same CTA.
#include "common.cuh"
#include "gemm.cuh"
#include "kernels.h"
#include "ptx.cuh"
#include <cstdlib>

namespace satb {

namespace {

constexpr int kQ = 128;        // query rows per unit
constexpr int kK = 128;        // keys per tile
constexpr int kD = 64;         // head dim
constexpr int kStagesKV = 2;
constexpr int kQBytes = kQ * kD * 2;                         // 16 KB
constexpr int kKVBytes = kK * kD * 2;                        // 16 KB
constexpr int kRowChunk = 1024;                              // keys per block of the CUDA-core row path
constexpr int kRowBatch = 16;                                // independent 16-byte loads in flight per lane (row path)
constexpr int kRowPathMax = 2;                               // Nq % 128 <= this: those rows take the row path
constexpr int kAttnSmem = kQBytes + 2 * kStagesKV * kKVBytes + kRowChunk * 4 + 256 + 1024;   // 85.25 KB: two CTAs per SM
constexpr int kTmemColsAttn = 256;
constexpr uint32_t kColS = 0, kColP = 128, kColO = 192;
constexpr float kRescaleThreshold = 8.0f;                    // log2 units

struct AttnTcArgs {
  uint16_t* o;
  int64_t ldo, o_bs;
  int Nq, Nk, group, H, batch;
  int q_col, k_col, v_col;   // column offsets (elements) of head 0 inside the q / k / v tensors
  int n_qt;                  // tensor-core query tiles per (item, head)
  int n_units;               // batch * H * n_qt
  int row0, n_rows;          // rows [row0, row0 + n_rows) of every (item, head) take the CUDA-core path
  const uint16_t *q, *k, *v; // raw pointers for the row path
  int64_t ldq, ldk, ldv, q_bs, k_bs, v_bs;
  float scale_log2;
  int* sm_slots;             // [>= number of SMs] running counters: the two CTAs of an SM draw consecutive values
  int stagger;               // cycles by which the odd CTA of an SM delays its softmax (see the kernel)
  unsigned long long* dbg;   // optional clock64 trace of CTA 0's first softmax warp (tests / profiles only)
};

// V tile as the MN-major B operand: rows = keys (K dim), 64 contiguous 16-bit d values (128 B, one swizzle atom)
// per row; 8-row groups 1024 B apart.
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

// same instruction, but volatile: the compiler keeps a run of these in program order (see cexp below)
__device__ __forceinline__ float ex2_ordered(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Packed fp32x2 arithmetic (FFMA2 / FADD2 on sm_100): two elements per issue slot for the scale-and-shift and the
// row-sum accumulation.
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

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// --------------------------------------------------------------------------------- CUDA-core row path
// One warp computes one query row of one (item, head): softmax(q K^T / 8) V in fp32.  Eight lanes share a key
// (lane l holds dims 8 (l % 8) .. +7 of q, of the K row and of the V row: one 16-byte load per lane and key),
// so one warp instruction covers four keys and every load of the loop is independent of the previous ones.
template <bool BF16>
__device__ void attn_row_path(const AttnTcArgs& p, int b, int h, int row, float* prow, int lane) {
  const int hk = h / p.group;
  const int sub = lane & 7, grp = lane >> 3;            // dims 8 sub .. 8 sub + 7; key phase grp (0..3)
  const uint16_t* qp = p.q + b * p.q_bs + static_cast<int64_t>(row) * p.ldq + p.q_col + h * kD + 8 * sub;
  const uint16_t* kp = p.k + b * p.k_bs + p.k_col + hk * kD + 8 * sub;
  const uint16_t* vp = p.v + b * p.v_bs + p.v_col + hk * kD + 8 * sub;
  float qf[8];
  {
    const uint4 u = __ldcg(reinterpret_cast<const uint4*>(qp));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = Op16<BF16>::unpack(w[e]);
      qf[2 * e] = f.x * p.scale_log2;
      qf[2 * e + 1] = f.y * p.scale_log2;
    }
  }
  float m_run = -INFINITY, l_run = 0.f;
  float o[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = 0.f;
  for (int k0 = 0; k0 < p.Nk; k0 += kRowChunk) {
    const int nk = min(kRowChunk, p.Nk - k0);
    const int n4 = (nk + 3) >> 2;                       // groups of four keys
    // ---- scores of this block -> prow[], running block max.  Loads are issued in explicit batches of kRowBatch
    // independent 16-byte loads per lane (the compiler will not hoist a global load above the shared-memory store
    // of the previous iteration, which would serialise one L2 round trip per group of four keys).
    float mx = -INFINITY;
    for (int i0 = 0; i0 < n4; i0 += kRowBatch) {
      uint4 u[kRowBatch];
#pragma unroll
      for (int jj = 0; jj < kRowBatch; ++jj) {
        const int key = 4 * (i0 + jj) + grp;
        u[jj] = key < nk ? __ldcg(reinterpret_cast<const uint4*>(kp + static_cast<int64_t>(k0 + key) * p.ldk))
                         : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int jj = 0; jj < kRowBatch; ++jj) {
        const int key = 4 * (i0 + jj) + grp;
        const uint32_t w[4] = {u[jj].x, u[jj].y, u[jj].z, u[jj].w};
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = Op16<BF16>::unpack(w[e]);
          s = fmaf(qf[2 * e], f.x, s);
          s = fmaf(qf[2 * e + 1], f.y, s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (key < nk) {
          if (sub == 0) prow[key] = s;
          mx = fmaxf(mx, s);
        }
      }
    }
    mx = warp_max(mx);
    const float m_new = fmaxf(m_run, mx);
    const float f = ex2_approx(m_run - m_new);          // 0 on the first block (m_run = -inf)
    l_run *= f;
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] *= f;
    m_run = m_new;
    __syncwarp();
    float ls = 0.f;
    for (int i = lane; i < nk; i += 32) {
      const float e = ex2_approx(prow[i] - m_new);
      prow[i] = e;
      ls += e;
    }
    l_run += warp_sum(ls);
    __syncwarp();
    // ---- O += P V: lane accumulates its 8 dims over the keys of its phase (same batching)
    for (int i0 = 0; i0 < n4; i0 += kRowBatch) {
      uint4 u[kRowBatch];
#pragma unroll
      for (int jj = 0; jj < kRowBatch; ++jj) {
        const int key = 4 * (i0 + jj) + grp;
        u[jj] = key < nk ? __ldcg(reinterpret_cast<const uint4*>(vp + static_cast<int64_t>(k0 + key) * p.ldv))
                         : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int jj = 0; jj < kRowBatch; ++jj) {
        const int key = 4 * (i0 + jj) + grp;
        const float e = key < nk ? prow[key] : 0.f;
        const uint32_t w[4] = {u[jj].x, u[jj].y, u[jj].z, u[jj].w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float2 vv = Op16<BF16>::unpack(w[c]);
          o[2 * c] = fmaf(e, vv.x, o[2 * c]);
          o[2 * c + 1] = fmaf(e, vv.y, o[2 * c + 1]);
        }
      }
    }
    __syncwarp();
  }
  // combine the four key phases (lanes l, l + 8, l + 16, l + 24 hold the same dims)
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    o[e] += __shfl_xor_sync(0xffffffffu, o[e], 8);
    o[e] += __shfl_xor_sync(0xffffffffu, o[e], 16);
  }
  if (grp == 0) {
    const float inv = 1.0f / l_run;
    uint16_t* og = p.o + b * p.o_bs + static_cast<int64_t>(row) * p.ldo + static_cast<int64_t>(h) * kD + 8 * sub;
    *reinterpret_cast<uint4*>(og) = make_uint4(Op16<BF16>::pack(o[0] * inv, o[1] * inv), Op16<BF16>::pack(o[2] * inv, o[3] * inv),
                                                Op16<BF16>::pack(o[4] * inv, o[5] * inv), Op16<BF16>::pack(o[6] * inv, o[7] * inv));
  }
}

template <bool BF16>
__global__ void __launch_bounds__(256, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttnTcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // [16 KB] (single: the last Q K^T of a unit is issued a whole
                                                        // tile before the unit ends, which is the time the next Q has to arrive)
  uint8_t* sK = smem + kQBytes;                         // [kStagesKV][16 KB]
  uint8_t* sV = sK + kStagesKV * kKVBytes;
  float* prow = reinterpret_cast<float*>(sV + kStagesKV * kKVBytes);   // [kRowChunk] row-path scratch
  uint64_t* bars = reinterpret_cast<uint64_t*>(prow + kRowChunk);
  uint64_t* q_full = bars;                 // [1] (+1 unused)
  uint64_t* q_empty = bars + 2;            // [1] (+1 unused)
  uint64_t* k_full = bars + 4;             // [kStagesKV]
  uint64_t* k_empty = k_full + kStagesKV;
  uint64_t* v_full = k_empty + kStagesKV;
  uint64_t* v_empty = v_full + kStagesKV;
  uint64_t* s_full = v_empty + kStagesKV;  // MMA -> softmax: S holds Q K_j^T
  uint64_t* s_free = s_full + 1;           // softmax -> MMA: S has been read into registers (128 arrivals)
  uint64_t* p_ready = s_free + 1;          // softmax -> MMA: P written (128 arrivals)
  uint64_t* p_free = p_ready + 1;          // MMA -> softmax: P V_j retired (P reusable, O up to date)
  uint64_t* o_done = p_free + 1;           // MMA -> softmax: last P V of the unit retired
  uint64_t* o_free = o_done + 1;           // softmax -> MMA: O of the previous unit has been read (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 1);
  uint32_t* sm_slot = tmem_slot + 1;       // 0 / 1: which of the SM's two resident CTAs this is

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.Nk + kK - 1) / kK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    for (int i = 0; i < kStagesKV; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(p_ready, 128);
    mbar_init(p_free, 1);
    mbar_init(o_done, 1);
    mbar_init(o_free, 128);
    fence_mbar_init();
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    *sm_slot = p.sm_slots ? static_cast<uint32_t>(atomicAdd(p.sm_slots + smid, 1)) & 1u : 0u;
    if (p.dbg) {   // per-CTA residency record: SM id, slot, start time (ns)
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.dbg[192 + blockIdx.x * 4 + 0] = smid;
      p.dbg[192 + blockIdx.x * 4 + 1] = *sm_slot;
      p.dbg[192 + blockIdx.x * 4 + 2] = t;
    }
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemColsAttn);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool odd_cta = *sm_slot != 0;
  const long long t_start = clock64();
  pdl_launch_dependents();
  pdl_wait();

  // unit u -> (item b, head h, query tile qt); consecutive units share (b, h), i.e. their K / V tiles in L2
  auto unit_coords = [&](int u, int& b, int& h, int& q0) {
    const int bh = u / p.n_qt;
    q0 = (u - bh * p.n_qt) * kQ;
    b = bh / p.H;
    h = bh - b * p.H;
  };

  if (warp == 0) {
    if (elect_one()) {
      // ---------------------------------------------------------------- TMA producer
      int g = 0;   // global key-tile counter of this CTA
      int i = 0;   // local unit counter
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++i) {
        int b, h, q0;
        unit_coords(u, b, h, q0);
        const int hk = h / p.group;
        mbar_wait(&q_empty[0], (i & 1) ^ 1);
        mbar_expect_tx(&q_full[0], kQBytes);
        tma_load_4d(sQ, &tmQ, &q_full[0], p.q_col + h * kD, 0, q0, b);
        for (int j = 0; j < n_tiles; ++j, ++g) {
          const int st = g % kStagesKV;
          const uint32_t ph = ((g / kStagesKV) & 1) ^ 1;
          mbar_wait(&k_empty[st], ph);
          mbar_expect_tx(&k_full[st], kKVBytes);
          tma_load_4d(sK + st * kKVBytes, &tmK, &k_full[st], p.k_col + hk * kD, 0, j * kK, b);
          mbar_wait(&v_empty[st], ph);
          mbar_expect_tx(&v_full[st], kKVBytes);
          tma_load_4d(sV + st * kKVBytes, &tmV, &v_full[st], p.v_col + hk * kD, 0, j * kK, b);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ------------------------------------------------- MMA issuer 1: S = Q K_j^T
      int g = 0, i = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++i) {
        const uint32_t q_addr = smem_u32(sQ);
        mbar_wait(&q_full[0], i & 1);
        tc_fence_after();
        for (int j = 0; j < n_tiles; ++j, ++g) {
          const int st = g % kStagesKV;
          mbar_wait(&k_full[st], (g / kStagesKV) & 1);
          // latency-critical: polled, not a suspending try_wait (the wake-up of a suspended issuer thread was measured
          // at ~1000 cycles between the softmax's arrive and S being full again)
          if (g >= 1) mbar_spin(s_free, (g - 1) & 1);   // S of the previous tile is in the softmax registers
          tc_fence_after();
          const int nk = min(kK, p.Nk - j * kK);
          const int n_mma = (nk + 15) & ~15;
          const uint32_t idesc = make_idesc_f16(kQ, n_mma, BF16);
          const uint32_t k_addr = smem_u32(sK + st * kKVBytes);
#pragma unroll
          for (int ks = 0; ks < kD / 16; ++ks)
            umma_f16_ss(tmem_base + kColS, make_desc_kmajor_sw128(q_addr + ks * 32),
                        make_desc_kmajor_sw128(k_addr + ks * 32), idesc, ks != 0);
          umma_commit(&k_empty[st]);   // the K tile is free as soon as these MMAs retire
          umma_commit(s_full);
        }
        umma_commit(&q_empty[0]);      // every Q K^T of this unit has retired: the Q buffer may be refilled
      }
    }
  } else if (warp == 3) {
    if (elect_one()) {
      // ------------------------------------------------- MMA issuer 2: O += P V_j
      constexpr uint32_t idesc_pv = make_idesc_f16(kQ, kD, BF16, /*b_mn_major=*/true);
      int g = 0, i = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++i) {
        for (int j = 0; j < n_tiles; ++j, ++g) {
          const int st = g % kStagesKV;
          mbar_wait(&v_full[st], (g / kStagesKV) & 1);
          mbar_spin(p_ready, g & 1);                    // polled: see the Q K^T issuer
          if (j == 0 && i >= 1) mbar_wait(o_free, (i - 1) & 1);   // the previous unit's O has been read out
          tc_fence_after();
          const int nk = min(kK, p.Nk - j * kK);
          const int ksteps = (nk + 15) >> 4;
          const uint32_t v_addr = smem_u32(sV + st * kKVBytes);
          if (ksteps == kK / 16) {
#pragma unroll
            for (int ks = 0; ks < kK / 16; ++ks)
              umma_f16_ts(tmem_base + kColO, tmem_base + kColP + ks * 8, make_desc_mnmajor_sw128(v_addr + ks * 2048),
                          idesc_pv, (j | ks) != 0);
          } else {
            for (int ks = 0; ks < ksteps; ++ks)
              umma_f16_ts(tmem_base + kColO, tmem_base + kColP + ks * 8, make_desc_mnmajor_sw128(v_addr + ks * 2048),
                          idesc_pv, (j | ks) != 0);
          }
          umma_commit(p_free);
          umma_commit(&v_empty[st]);
          if (j == n_tiles - 1) umma_commit(o_done);
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------- ragged query rows on CUDA cores
    if (p.n_rows > 0) {
      const int n_tasks = p.batch * p.H * p.n_rows;
      for (int t = blockIdx.x; t < n_tasks; t += gridDim.x) {
        const int bh = t / p.n_rows, r = t - bh * p.n_rows;
        const int b = bh / p.H, h = bh - b * p.H;
        attn_row_path<BF16>(p, b, h, p.row0 + r, prow, lane);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------- softmax + epilogue (1 row / thread)
    const int q = warp - 4;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t s_addr = t_lane + kColS, p_addr = t_lane + kColP, o_addr = t_lane + kColO;
    const float sc = p.scale_log2;
    const bool trace = p.dbg != nullptr && blockIdx.x == 0 && warp == 4 && lane == 0;
    int g = 0, i = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++i) {
      int b, h, q0;
      unit_coords(u, b, h, q0);
      // a warp whose 32 rows are all beyond Nq (partial last query tile) only keeps the barrier protocol going;
      // its P rows are never read back through O
      const bool warp_active = (q0 + q * 32) < p.Nq;
      float m_ref = -INFINITY, l = 0.f;
      for (int j = 0; j < n_tiles; ++j, ++g) {
        const int nk = min(kK, p.Nk - j * kK);
        const int nch = (nk + 31) >> 5;   // 32-column chunks holding real keys
        if (trace && g < 16) p.dbg[g * 12 + 0] = clock64();
        mbar_spin(s_full, g & 1);
        tc_fence_after();
        if (g == 0 && odd_cta) {
          // The two resident CTAs of an SM start together and, being identical, would stay in lockstep: both
          // exponentiate (sharing the SFU) and then both sit in their per-tile bookkeeping with the SFU idle.  The odd
          // one therefore starts its first tile half a period late; the offset persists (neither waits for the other).
          const long long t0 = clock64();
          while (clock64() - t0 < p.stagger) {
          }
        }
        if (trace && g < 16) p.dbg[g * 12 + 1] = clock64();
        if (!warp_active) {
          tc_fence_before();
          mbar_arrive(s_free);
          if (g >= 1) mbar_wait(p_free, (g - 1) & 1);
          mbar_arrive(p_ready);
          continue;
        }
        uint32_t ra[32], rb[32];
        auto cmax = [&](int c, const uint32_t (&r)[32]) -> float {
          const int lim = nk - c * 32;
          if (lim >= 32) {
            // four independent chains of 3-input maxima (a single chain is 16 dependent instructions)
            float m0 = __uint_as_float(r[0]), m1 = __uint_as_float(r[1]), m2 = __uint_as_float(r[2]), m3 = __uint_as_float(r[3]);
#pragma unroll
            for (int e = 4; e < 28; e += 8) {
              m0 = fmaxf(m0, fmaxf(__uint_as_float(r[e]), __uint_as_float(r[e + 1])));
              m1 = fmaxf(m1, fmaxf(__uint_as_float(r[e + 2]), __uint_as_float(r[e + 3])));
              m2 = fmaxf(m2, fmaxf(__uint_as_float(r[e + 4]), __uint_as_float(r[e + 5])));
              m3 = fmaxf(m3, fmaxf(__uint_as_float(r[e + 6]), __uint_as_float(r[e + 7])));
            }
            m0 = fmaxf(m0, fmaxf(__uint_as_float(r[28]), __uint_as_float(r[29])));
            m1 = fmaxf(m1, fmaxf(__uint_as_float(r[30]), __uint_as_float(r[31])));
            return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
          }
          float mx = -INFINITY;
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (e < lim) mx = fmaxf(mx, __uint_as_float(r[e]));
          return mx;
        };
        // P chunk = exp2(S c - m_ref) as packed 16-bit pairs in w[]; returns the fp32 row sum of the chunk
        auto cexp = [&](int c, const uint32_t (&r)[32], uint32_t (&w)[16]) -> float {
          float sum;
          const int lim = nk - c * 32;
          if (lim >= 32) {
            // three phases, so that the 32 MUFU.EX2 are issued back to back and none of their consumers waits on a
            // result that is still in the SFU pipeline (a lone warp per scheduler cannot hide that latency)
            const uint64_t sc2 = pack2(sc, sc), nm2 = pack2(-m_ref, -m_ref);
            float t[32];
#pragma unroll
            for (int e = 0; e < 16; ++e)
              unpack2(ffma2(pack2(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1])), sc2, nm2), t[2 * e], t[2 * e + 1]);
#pragma unroll
            for (int e = 0; e < 32; ++e) t[e] = ex2_ordered(t[e]);
            uint64_t sum2a = pack2(0.f, 0.f), sum2b = pack2(0.f, 0.f);
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              sum2a = fadd2(sum2a, pack2(t[2 * e], t[2 * e + 1]));
              sum2b = fadd2(sum2b, pack2(t[2 * e + 2], t[2 * e + 3]));
              w[e] = Op16<BF16>::pack(t[2 * e], t[2 * e + 1]);
              w[e + 1] = Op16<BF16>::pack(t[2 * e + 2], t[2 * e + 3]);
            }
            float a0, a1;
            unpack2(fadd2(sum2a, sum2b), a0, a1);
            sum = a0 + a1;
          } else {
            sum = 0.f;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float s0 = 2 * e < lim ? __uint_as_float(r[2 * e]) : -INFINITY;
              const float s1 = 2 * e + 1 < lim ? __uint_as_float(r[2 * e + 1]) : -INFINITY;
              const float p0 = ex2_approx(fmaf(s0, sc, -m_ref));
              const float p1 = ex2_approx(fmaf(s1, sc, -m_ref));
              sum += p0 + p1;
              w[e] = Op16<BF16>::pack(p0, p1);
            }
          }
          return sum;
        };
        // ---- pipelined pass: exponentials of chunks 0,1 (kept in registers) while 2,3 load and while P V of the
        // previous tile retires -> store them -> release S -> exponentials of chunks 2,3
        tmem_ld_32x32(s_addr, ra);
        if (nch > 1) tmem_ld_32x32(s_addr + 32, rb);
        tmem_ld_wait();
        if (trace && g < 16) p.dbg[g * 12 + 2] = clock64();
        float mx_raw = cmax(0, ra);
        if (nch > 1) mx_raw = fmaxf(mx_raw, cmax(1, rb));
        // the first tile of a unit fixes the reference max from its first (up to) 64 keys; should a later key of the
        // tile exceed it by more than 2^8 the lazy-rescale path below recomputes the tile
        if (j == 0) m_ref = mx_raw * sc;
        uint32_t w0[16], w1[16];
        float sum = cexp(0, ra, w0);
        if (trace && g < 16) p.dbg[g * 12 + 3] = clock64();
        if (nch > 2) tmem_ld_32x32(s_addr + 64, ra);
        if (nch > 1) sum += cexp(1, rb, w1);
        if (trace && g < 16) p.dbg[g * 12 + 4] = clock64();
        if (nch > 3) tmem_ld_32x32(s_addr + 96, rb);
        if (g >= 1) {
          mbar_spin(p_free, (g - 1) & 1);   // P V of the previous tile retired: P may be overwritten, O is quiescent
          tc_fence_after();
        }
        if (trace && g < 16) p.dbg[g * 12 + 5] = clock64();
        tmem_st_32x16(p_addr, w0);
        if (nch > 1) tmem_st_32x16(p_addr + 16, w1);
        if (trace && g < 16) p.dbg[g * 12 + 6] = clock64();
        if (nch > 2) {
          tmem_ld_wait();
          if (trace && g < 16) p.dbg[g * 12 + 7] = clock64();
          mx_raw = fmaxf(mx_raw, cmax(2, ra));
          if (nch > 3) mx_raw = fmaxf(mx_raw, cmax(3, rb));
        }
        // lazy rescale: only when this tile's max exceeds the reference max by more than 2^8
        const bool need = mx_raw * sc > m_ref + kRescaleThreshold;
        if (!__any_sync(0xffffffffu, need)) {
          tc_fence_before();
          mbar_arrive(s_free);              // every column of S is in registers: Q K_{j+1}^T may overwrite it
          if (trace && g < 16) p.dbg[g * 12 + 8] = clock64();
          if (nch > 2) {
            sum += cexp(2, ra, w0);
            tmem_st_32x16(p_addr + 32, w0);
          }
          if (trace && g < 16) p.dbg[g * 12 + 9] = clock64();
          if (nch > 3) {
            sum += cexp(3, rb, w1);
            tmem_st_32x16(p_addr + 48, w1);
          }
          if (trace && g < 16) p.dbg[g * 12 + 10] = clock64();
        } else {
          const float m_new = need ? mx_raw * sc : m_ref;
          const float f = ex2_approx(m_ref - m_new);   // 1 for rows that keep their reference
          m_ref = m_new;
          l *= f;
          if (j > 0) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t r[32];
              tmem_ld_32x32(o_addr + c * 32, r);
              tmem_ld_wait();
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                uint32_t w[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) w[e] = __float_as_uint(__uint_as_float(r[half * 16 + e]) * f);
                tmem_st_32x16(o_addr + c * 32 + half * 16, w);
              }
            }
          }
          sum = 0.f;                        // P again, all chunks, with the new reference max
          for (int c = 0; c < nch; ++c) {
            tmem_ld_32x32(s_addr + c * 32, ra);
            tmem_ld_wait();
            sum += cexp(c, ra, w0);
            tmem_st_32x16(p_addr + c * 16, w0);
          }
          tc_fence_before();
          mbar_arrive(s_free);
        }
        l += sum;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_ready);
        if (trace && g < 16) p.dbg[g * 12 + 11] = clock64();
      }
      // epilogue of the unit: O / l -> global (128 B per row)
      mbar_spin(o_done, i & 1);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(o_addr, r0);
      tmem_ld_32x32(o_addr + 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(o_free);                  // O is in registers: the next unit's first P V may overwrite it
      const float inv = 1.0f / l;
      const bool valid = (q0 + row) < p.row0;   // row0 = Nq, or the first row of the CUDA-core path
      if (valid) {
        uint16_t* og = p.o + b * p.o_bs + static_cast<int64_t>(q0 + row) * p.ldo + static_cast<int64_t>(h) * kD;
        uint4* dst = reinterpret_cast<uint4*>(og);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          dst[e] = make_uint4(Op16<BF16>::pack(__uint_as_float(r0[8 * e]) * inv, __uint_as_float(r0[8 * e + 1]) * inv),
                              Op16<BF16>::pack(__uint_as_float(r0[8 * e + 2]) * inv, __uint_as_float(r0[8 * e + 3]) * inv),
                              Op16<BF16>::pack(__uint_as_float(r0[8 * e + 4]) * inv, __uint_as_float(r0[8 * e + 5]) * inv),
                              Op16<BF16>::pack(__uint_as_float(r0[8 * e + 6]) * inv, __uint_as_float(r0[8 * e + 7]) * inv));
#pragma unroll
        for (int e = 0; e < 4; ++e)
          dst[4 + e] = make_uint4(Op16<BF16>::pack(__uint_as_float(r1[8 * e]) * inv, __uint_as_float(r1[8 * e + 1]) * inv),
                                  Op16<BF16>::pack(__uint_as_float(r1[8 * e + 2]) * inv, __uint_as_float(r1[8 * e + 3]) * inv),
                                  Op16<BF16>::pack(__uint_as_float(r1[8 * e + 4]) * inv, __uint_as_float(r1[8 * e + 5]) * inv),
                                  Op16<BF16>::pack(__uint_as_float(r1[8 * e + 6]) * inv, __uint_as_float(r1[8 * e + 7]) * inv));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.dbg && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.dbg[192 + blockIdx.x * 4 + 3] = t;
  }
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
  SATB_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && q_col % 8 == 0 && k_col % 8 == 0 && v_col % 8 == 0,
               "attention operand strides / column offsets must be 16B aligned");
  CUtensorMap tq, tk, tv;
  SATB_PROPAGATE(make_tmap_rows(&tq, q, q_cols, Nq, batch, ldq, q_bs, kQ));
  SATB_PROPAGATE(make_tmap_rows(&tk, k, k_cols, Nk, batch, ldk, k_bs, kK));
  SATB_PROPAGATE(make_tmap_rows(&tv, v, v_cols, Nk, batch, ldv, v_bs, kK));
  AttnTcArgs a;
  a.o = static_cast<uint16_t*>(o);
  a.ldo = ldo; a.o_bs = o_bs;
  a.Nq = Nq; a.Nk = Nk; a.group = H / H_kv; a.H = H; a.batch = batch;
  a.q_col = q_col; a.k_col = k_col; a.v_col = v_col;
  // query rows: full 128-row tiles on the tensor cores; a remainder of <= kRowPathMax rows on CUDA cores, a larger
  // remainder as one more (partial) tensor-core tile
  const int rem = Nq % kQ;
  const bool row_path = rem != 0 && rem <= kRowPathMax;
  a.n_qt = Nq / kQ + ((rem != 0 && !row_path) ? 1 : 0);
  a.n_units = batch * H * a.n_qt;
  a.row0 = row_path ? Nq - rem : Nq;
  a.n_rows = row_path ? rem : 0;
  a.q = static_cast<const uint16_t*>(q); a.k = static_cast<const uint16_t*>(k); a.v = static_cast<const uint16_t*>(v);
  a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.q_bs = q_bs; a.k_bs = k_bs; a.v_bs = v_bs;
  a.scale_log2 = (1.0f / sqrtf(64.0f)) * 1.4426950408889634f;
  a.dbg = dbg;
  {
    static int* slots[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!slots[dev]) {   // once per device (first call = warm-up, never inside a graph capture)
      SATB_CHECK_CUDA(cudaMalloc(&slots[dev], 1024 * sizeof(int)));
      SATB_CHECK_CUDA(cudaMemset(slots[dev], 0, 1024 * sizeof(int)));
    }
    a.sm_slots = slots[dev];
    static int stagger = -1;
    if (stagger < 0) {
      const char* e = getenv("SATB_ATTN_STAGGER");   // cycles; tuning / A-B only
      stagger = e ? atoi(e) : 1100;
    }
    a.stagger = stagger;
  }
  const int row_tasks = batch * H * a.n_rows;
  int grid = a.n_units > row_tasks ? a.n_units : row_tasks;
  static int ctas_per_sm = -1;
  if (ctas_per_sm < 0) {
    const char* e = getenv("SATB_ATTN_CTAS_PER_SM");   // 1: one CTA per SM (A/B measurement of the SFU sharing)
    ctas_per_sm = (e && atoi(e) == 1) ? 1 : 2;
  }
  const int slots = ctas_per_sm * device_sm_count();
  if (grid > slots) grid = slots;
  if (grid <= 0) return 0;
  static PerDeviceOnce attr16, attrbf;
  auto prepare = [&](auto kern) -> int {
    SATB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    // two CTAs per SM need 2 x 102 KB: ask for the largest shared-memory carveout
    SATB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    if (getenv("SATB_ATTN_DEBUG")) {
      int nb = -1, dev = 0, sm_smem = 0, rsv = 0, regs = 0;
      cudaGetDevice(&dev);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 256, kAttnSmem);
      cudaDeviceGetAttribute(&sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
      cudaDeviceGetAttribute(&rsv, cudaDevAttrReservedSharedMemoryPerBlock, dev);
      cudaDeviceGetAttribute(&regs, cudaDevAttrMaxRegistersPerMultiprocessor, dev);
      cudaFuncAttributes fa;
      cudaFuncGetAttributes(&fa, kern);
      fprintf(stderr, "[satb] attention: %d resident CTAs per SM (dyn smem %d B + static %zu B per CTA, %d regs/thread; SM: "
              "%d B smem, %d B reserved per block, %d regs)\n", nb, kAttnSmem, fa.sharedSizeBytes, fa.numRegs, sm_smem, rsv, regs);
    }
    return 0;
  };
  if (bf16) {
    if (attrbf.first()) SATB_PROPAGATE(prepare(attn_tc_kernel<true>));
    SATB_CHECK_CUDA(launch_pdl(attn_tc_kernel<true>, dim3(grid), dim3(256), kAttnSmem, stream, tq, tk, tv, a));
  } else {
    if (attr16.first()) SATB_PROPAGATE(prepare(attn_tc_kernel<false>));
    SATB_CHECK_CUDA(launch_pdl(attn_tc_kernel<false>, dim3(grid), dim3(256), kAttnSmem, stream, tq, tk, tv, a));
  }
  count_launch();
  SATB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Debug: resident CTAs per SM the runtime reports for the attention kernel with `dyn_smem` bytes of dynamic shared
// memory and the given carveout preference (percent, -1 = leave unchanged); tests / profiling only.
int debug_attention_occupancy(int dyn_smem, int carveout_pct) {
  auto kern = attn_tc_kernel<false>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem) != cudaSuccess) return -1;
  if (carveout_pct >= 0) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_pct);
  int nb = -1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 256, dyn_smem) != cudaSuccess) return -2;
  return nb;
}

}  // namespace satb
