// tcgen05 attention forward, head dim 64, no mask, non-causal, optional GQA:
//   O = softmax(Q K^T / sqrt(64)) V      (reference models/transformer.py:496-536)
//
// Persistent kernel, two CTAs of 384 threads per SM (2 x 256 TMEM columns, 2 x 86 KB shared memory).  A work unit is
// 128 query rows of one (batch item, head); CTA c processes units c, c + grid, ... without re-initialising anything.
// Keys are processed in tiles of 128:
//   warp 0 (one thread)  TMA producer: Q of the unit, K / V tiles (2-stage rings)
//   warp 1 (one thread)  S = Q K_j^T -> TMEM (tcgen05.mma, smem operands) as soon as S has been read out
//   warp 3 (one thread)  O += P V_j (A = P from TMEM, B = V tile addressed MN-major)
//   warps 4-11           softmax: warps w and w + 4 own the same 32 query rows (TMEM lane quadrant w % 4) and split the
//                        128 key columns of a tile; per 32-column chunk: tcgen05.ld -> row max -> P = exp2(S c - m_ref)
//                        -> packed 16-bit pairs back to TMEM; row sums in fp32 registers.
//   warp 2               TMEM allocator; afterwards the CUDA-core path for ragged query rows (below)
// TMEM columns: S [0,128)  P [128,192)  O [192,256).
// Why eight softmax warps: the softmax is bound by the SFU (16 ex2 / clk / SM = one MUFU.EX2 warp instruction per
// 8 cycles and scheduler; 202 M exponentials per SA-Open layer) only if every scheduler always has a warp with
// exponentials to issue.  A lone warp per scheduler reaches 12.8 cycles per MUFU (in-order issue behind the FFMA2 /
// F2FP / FADD2 of its own stream) and spends as long again in tcgen05.ld / st / mbarrier latencies per tile; with
// two such warps (two CTAs of four softmax warps) the CTAs fall into lockstep - both exponentiate, then both wait -
// (measured, profiles/r02_attention_*.txt).  Four warps per scheduler cover those latencies.
// O stays in TMEM for the whole unit: the reference max m_ref only moves when a tile's row max exceeds it by more
// than 2^8 (lazy rescale: exponentials stay <= 256, sums in fp32), so the O rescale (TMEM load-scale-store) and the
// recomputation of that tile's P are rare.
//
// Ragged shapes (1025 = 8 * 128 + 1 tokens: the prepended conditioning token):
//   * leftover KEYS (Nk % 128 <= kExtraMax) do not get a tile of their own - a whole pipeline step for one column -:
//     every softmax thread computes its row's score against them on CUDA cores (q from the Q tile in shared memory)
//     and adds exp2(.) v to its O row in the epilogue, in fp32;
//   * leftover QUERY rows (Nq % 128 <= kRowPathMax) are computed by warp 2 on CUDA cores (one warp per row:
//     eight lanes per key, online softmax over blocks of 1024 keys), concurrently with the tensor-core pipeline.
#include "common.cuh"
#include "gemm.cuh"
#include "kernels.h"
#include "ptx.cuh"
#include <cmath>
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
constexpr int kRowBatch = 8;                                 // independent 16-byte loads in flight per lane (row path)
constexpr int kRowPathMax = 2;                               // Nq % 128 <= this: those rows take the row path
constexpr int kAttnThreads = 384;                            // 4 control warps + 8 softmax warps
constexpr int kSoftmaxThreads = 256;
constexpr int kAttnPolyDefault = 0;                          // see SATB_ATTN_POLY
constexpr int kAttnSlotsDefault = 2;                         // see SATB_ATTN_SLOTS
constexpr int kExtraMax = 2;                                 // Nk % 128 <= this: those keys are added in the epilogue
constexpr int kXRows = 16;                                   // rows of the leftover-key K / V tiles (TMA box)
constexpr int kXBytes = kXRows * kD * 2;                     // 2 KB
// shared memory of one pipeline ("slot"): Q, K / V rings, leftover-key boxes, row-path scratch, exchange slots, barriers
constexpr int kSlotSmem = (kQBytes + 2 * kStagesKV * kKVBytes + 4 * kXBytes + kRowChunk * 4 + 4 * 2 * kQ * 4 + 256 + 1023) & ~1023;   // 97 KB
constexpr int attn_smem_bytes(int slots) { return slots * kSlotSmem + 1024; }
constexpr int kWarpsPerSlot = kAttnThreads / 32;
constexpr int kAltSpinMax = 400;                             // safety valve of the turn-taking below (never a correctness matter)
constexpr int kTmemColsAttn = 256;
constexpr uint32_t kColS = 0, kColP = 128, kColO = 192;
constexpr float kRescaleThreshold = 8.0f;                    // log2 units

struct AttnTcArgs {
  uint16_t* o;
  int64_t ldo, o_bs;
  int Nq, Nk, group, H, batch;
  int q_col, k_col, v_col;   // column offsets (elements) of head 0 inside the q / k / v tensors
  int n_tiles, n_extra;      // key tiles on the tensor cores; leftover keys (Nk - 128 n_tiles <= kExtraMax) added in the epilogue
  int n_qt;                  // tensor-core query tiles per (item, head)
  int n_units;               // batch * H * n_qt
  int row0, n_rows;          // rows [row0, row0 + n_rows) of every (item, head) take the CUDA-core path
  const uint16_t *q, *k, *v; // raw pointers for the row path
  int64_t ldq, ldk, ldv, q_bs, k_bs, v_bs;
  float scale_log2;
  int alternate;             // SLOTS == 2: the two pipelines of a CTA take turns on the SFU (see attn_tc_kernel)
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

// 2^x for a pair on the FMA / ALU pipes instead of the SFU (x <= ~8; clamped at -126): round-to-nearest split
// x = i + f, |f| <= 0.5 by the magic-number add, 2^f by its degree-4 Taylor polynomial (relative error < 4.2e-5, below
// the 16-bit rounding P gets anyway), 2^i by an integer add into the exponent field.
__device__ __forceinline__ void ex2_poly2(float x0, float x1, float& y0, float& y1) {
  const float kMagic = 12582912.f;   // 1.5 * 2^23
  x0 = fmaxf(x0, -126.f);
  x1 = fmaxf(x1, -126.f);
  const uint64_t x = pack2(x0, x1), mg = pack2(kMagic, kMagic), nmg = pack2(-kMagic, -kMagic);
  const uint64_t t = fadd2(x, mg);                 // integer part in the low mantissa bits
  uint64_t fi = fadd2(t, nmg);                     // float(i)
  float f0, f1, i0, i1;
  unpack2(fi, i0, i1);
  const uint64_t f = fadd2(x, pack2(-i0, -i1));    // f = x - i
  // Horner: ((((c4 f + c3) f + c2) f + c1) f + 1)
  const float c1 = 0.6931471806f, c2 = 0.2402265070f, c3 = 0.0555041087f, c4 = 0.0096181291f;
  uint64_t acc = ffma2(pack2(c4, c4), f, pack2(c3, c3));
  acc = ffma2(acc, f, pack2(c2, c2));
  acc = ffma2(acc, f, pack2(c1, c1));
  acc = ffma2(acc, f, pack2(1.f, 1.f));
  float t0, t1, p0, p1;
  unpack2(t, t0, t1);
  unpack2(acc, p0, p1);
  (void)f0; (void)f1;
  y0 = __uint_as_float(__float_as_uint(p0) + ((__float_as_uint(t0) - 0x4B400000u) << 23));
  y1 = __uint_as_float(__float_as_uint(p1) + ((__float_as_uint(t1) - 0x4B400000u) << 23));
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

// POLY: every fourth pair of exponentials of a full chunk goes through ex2_poly2 (FMA / ALU pipes) instead of the SFU
// SLOTS: pipelines per CTA.  1: a CTA of 384 threads is one pipeline, two CTAs share an SM.  2: ONE CTA of 768 threads
// per SM runs two complete pipelines (warps 0-11 and 12-23, each with its own shared-memory region, barriers, 256 TMEM
// columns and unit sequence - "virtual CTA" 2 blockIdx + slot), which lets the two softmax groups TAKE TURNS on the SFU:
// two independent CTAs fall into lockstep (both exponentiate - at half rate each -, then both sit in tcgen05.ld /
// barrier latencies with the SFU idle); with turns one group's 32-exponential segment runs at the full SFU rate while
// the other group is in its loads / max / exchange.  The turn is a shared-memory word handed over by the eighth
// release of a group; a group whose partner is outside its tile loop (epilogue, finished) does not wait, and every
// wait is bounded, so the scheme can only change timing.
template <bool BF16, bool POLY, int SLOTS>
__global__ void __launch_bounds__(kAttnThreads * SLOTS, 3 - SLOTS)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmKx,
               const __grid_constant__ CUtensorMap tmVx, const AttnTcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ int alt_turn, alt_rel[2], alt_present[2];
  const int warp_g = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = SLOTS == 2 ? warp_g / kWarpsPerSlot : 0;
  const int warp = warp_g - slot * kWarpsPerSlot;                      // role index inside the pipeline
  const int vcta = blockIdx.x * SLOTS + slot, vgrid = gridDim.x * SLOTS;   // this pipeline among all of the grid
  uint8_t* smem0 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem = smem0 + slot * kSlotSmem;
  uint8_t* sQ = smem;                                   // [16 KB] (single: the last Q K^T of a unit is issued a whole
                                                        // tile before the unit ends, which is the time the next Q has to arrive)
  uint8_t* sK = smem + kQBytes;                         // [kStagesKV][16 KB]
  uint8_t* sV = sK + kStagesKV * kKVBytes;
  uint8_t* sX = sV + kStagesKV * kKVBytes;              // [2 units][K | V][2 KB] leftover-key rows (16-row TMA boxes)
  float* prow = reinterpret_cast<float*>(sX + 4 * kXBytes);            // [kRowChunk] row-path scratch
  float* xch = prow + kRowChunk;                        // [4][2][128] exchange between the two column halves of a row:
                                                        // slots 0 / 1 = tile parity, 2 = first-tile max, 3 = row sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch + 4 * 2 * kQ);
  uint64_t* q_full = bars;                 // TMA -> MMA / softmax: Q of the unit has landed
  uint64_t* q_empty = bars + 1;            // MMA (+ softmax, when it reads Q for leftover keys) -> TMA
  uint64_t* k_full = bars + 2;             // [kStagesKV]
  uint64_t* k_empty = k_full + kStagesKV;
  uint64_t* v_full = k_empty + kStagesKV;
  uint64_t* v_empty = v_full + kStagesKV;
  uint64_t* s_full = v_empty + kStagesKV;  // MMA -> softmax: S holds Q K_j^T
  uint64_t* s_free = s_full + 1;           // softmax -> MMA: S has been read (256 arrivals)
  uint64_t* p_ready = s_free + 1;          // softmax -> MMA: P written (256 arrivals)
  uint64_t* p_free = p_ready + 1;          // MMA -> softmax: P V_j retired (P reusable, O up to date)
  uint64_t* o_done = p_free + 1;           // MMA -> softmax: last P V of the unit retired
  uint64_t* o_free = o_done + 1;           // softmax -> MMA: O of the previous unit has been read (256 arrivals)
  uint64_t* x_full = o_free + 1;           // [2] TMA -> softmax: leftover-key K / V rows of the unit have landed
  uint64_t* x_empty = x_full + 2;          // [2] softmax -> TMA: they have been used (256 arrivals)
  // the TMEM base address lands in slot 0's region (one allocation of SLOTS x 256 columns)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(x_empty + 2) - slot * kSlotSmem);

  const int n_tiles = p.n_tiles;           // tensor-core key tiles; the n_extra leftover keys are added in the epilogue
  const int n_extra = p.n_extra;

  if (threadIdx.x == 0) {
    alt_turn = 0;
    alt_rel[0] = alt_rel[1] = 0;
    alt_present[0] = alt_present[1] = 0;
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, n_extra > 0 ? 1 + kSoftmaxThreads : 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], kSoftmaxThreads);
    }
    for (int i = 0; i < kStagesKV; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, kSoftmaxThreads);
    mbar_init(p_ready, kSoftmaxThreads);
    mbar_init(p_free, 1);
    mbar_init(o_done, 1);
    mbar_init(o_free, kSoftmaxThreads);
    fence_mbar_init();
    if (p.dbg && slot == 0) {   // per-CTA residency record: SM id, start time (ns)
      uint32_t smid;
      unsigned long long t;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.dbg[192 + blockIdx.x * 4 + 0] = smid;
      p.dbg[192 + blockIdx.x * 4 + 2] = t;
    }
  }
  if (warp_g == 2) {
    tmem_alloc(tmem_slot, kTmemColsAttn * SLOTS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot + slot * kTmemColsAttn;
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
      for (int u = vcta; u < p.n_units; u += vgrid, ++i) {
        int b, h, q0;
        unit_coords(u, b, h, q0);
        const int hk = h / p.group;
        mbar_wait(q_empty, (i & 1) ^ 1);
        mbar_expect_tx(q_full, kQBytes);
        tma_load_4d(sQ, &tmQ, q_full, p.q_col + h * kD, 0, q0, b);
        if (n_extra > 0) {               // the unit's leftover keys: 16-row boxes (rows past Nk are zero-filled)
          const int xb = i & 1;
          mbar_wait(&x_empty[xb], ((i >> 1) & 1) ^ 1);
          mbar_expect_tx(&x_full[xb], 2 * kXBytes);
          tma_load_4d(sX + xb * 2 * kXBytes, &tmKx, &x_full[xb], p.k_col + hk * kD, 0, n_tiles * kK, b);
          tma_load_4d(sX + xb * 2 * kXBytes + kXBytes, &tmVx, &x_full[xb], p.v_col + hk * kD, 0, n_tiles * kK, b);
        }
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
      for (int u = vcta; u < p.n_units; u += vgrid, ++i) {
        const uint32_t q_addr = smem_u32(sQ);
        mbar_wait(q_full, i & 1);
        tc_fence_after();
        for (int j = 0; j < n_tiles; ++j, ++g) {
          const int st = g % kStagesKV;
          mbar_wait(&k_full[st], (g / kStagesKV) & 1);
          if (g >= 1) mbar_spin(s_free, (g - 1) & 1);   // S of the previous tile has been read by every softmax thread
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
        umma_commit(q_empty);          // every Q K^T of this unit has retired (immediately, if there was none)
      }
    }
  } else if (warp == 3) {
    if (elect_one()) {
      // ------------------------------------------------- MMA issuer 2: O += P V_j
      constexpr uint32_t idesc_pv = make_idesc_f16(kQ, kD, BF16, /*b_mn_major=*/true);
      int g = 0, i = 0;
      for (int u = vcta; u < p.n_units; u += vgrid, ++i) {
        for (int j = 0; j < n_tiles; ++j, ++g) {
          const int st = g % kStagesKV;
          mbar_wait(&v_full[st], (g / kStagesKV) & 1);
          mbar_spin(p_ready, g & 1);
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
      for (int t = vcta; t < n_tasks; t += vgrid) {
        const int bh = t / p.n_rows, r = t - bh * p.n_rows;
        const int b = bh / p.H, h = bh - b * p.H;
        attn_row_path<BF16>(p, b, h, p.row0 + r, prow, lane);
      }
    }
  } else {
    // ------------------------------------------------------- softmax + epilogue
    // Eight warps: warp w may touch TMEM lanes 32 (w % 4) .. +31, so warps w and w + 4 share the 32 query rows of a
    // quadrant and split the key columns of every tile (half 0: keys 0-63, half 1: keys 64-127).  Four warps per
    // scheduler (two CTAs per SM) keep the SFU fed while others sit in TMEM / barrier latencies.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t s_addr = t_lane + kColS, p_addr = t_lane + kColP, o_addr = t_lane + kColO;
    const float sc = p.scale_log2;
    const bool trace = p.dbg != nullptr && vcta == 0 && warp == 4 && lane == 0;
    // exchange slot s: this thread writes xch[s][half][row] and reads xch[s][1 - half][row] after the pair barrier
    auto xput = [&](int slot, float v) { xch[(slot * 2 + half) * kQ + row] = v; };
    auto xget = [&](int slot) -> float { return xch[(slot * 2 + (half ^ 1)) * kQ + row]; };
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + q + 4 * slot) : "memory"); };   // the two warps of a quadrant
    // turn-taking on the SFU between the two pipelines of the CTA (SLOTS == 2)
    const bool alt = SLOTS == 2 && p.alternate != 0;
    auto alt_acquire = [&]() {
      if (alt) {
        if (lane == 0) {
          int spins = 0;
          while (*reinterpret_cast<volatile int*>(&alt_turn) != slot &&
                 *reinterpret_cast<volatile int*>(&alt_present[slot ^ 1]) > 0 && ++spins < kAltSpinMax) {
          }
        }
        __syncwarp();
      }
    };
    auto alt_release = [&]() {
      if (alt && lane == 0) {
        if ((atomicAdd(&alt_rel[slot], 1) & 7) == 7) *reinterpret_cast<volatile int*>(&alt_turn) = slot ^ 1;
      }
    };
    int g = 0, i = 0;
    for (int u = vcta; u < p.n_units; u += vgrid, ++i) {
      int b, h, q0;
      unit_coords(u, b, h, q0);
      const int hk = h / p.group;
      float m_ref = -INFINITY, l = 0.f;
      // ---- leftover keys (Nk % 128 <= kExtraMax): their scores on CUDA cores, from the Q tile in shared memory
      float s_x[kExtraMax];
#pragma unroll
      for (int e = 0; e < kExtraMax; ++e) s_x[e] = -INFINITY;
      const uint8_t* xk = sX + (i & 1) * 2 * kXBytes;   // K rows of the leftover keys; V rows follow at + kXBytes
      if (n_extra > 0) {
        mbar_spin(q_full, i & 1);
        mbar_spin(&x_full[i & 1], (i >> 1) & 1);
#pragma unroll
        for (int e = 0; e < kExtraMax; ++e) {
          if (e < n_extra) {
            float acc = 0.f;
#pragma unroll 2
            for (int c = 0; c < kD / 8; ++c) {
              // 128B-swizzled K-major tiles: 16-byte chunk c of row r sits at chunk position c ^ (r % 8)
              const uint4 qv = *reinterpret_cast<const uint4*>(sQ + row * 128 + ((c ^ (row & 7)) << 4));
              const uint4 kv = *reinterpret_cast<const uint4*>(xk + e * 128 + ((c ^ e) << 4));   // same address in every lane
              const uint32_t qw[4] = {qv.x, qv.y, qv.z, qv.w}, kw[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
              for (int d = 0; d < 4; ++d) {
                const float2 a = Op16<BF16>::unpack(qw[d]), bb = Op16<BF16>::unpack(kw[d]);
                acc = fmaf(a.x, bb.x, fmaf(a.y, bb.y, acc));
              }
            }
            s_x[e] = acc;
          }
        }
        mbar_arrive(q_empty);            // this thread is done with Q
      }
      if (trace && g < 16) p.dbg[g * 12 + 2] = clock64();
      if (alt && lane == 0 && n_tiles > 0) atomicAdd(&alt_present[slot], 1);
      for (int j = 0; j < n_tiles; ++j, ++g) {
        const int nk = min(kK, p.Nk - j * kK);
        const int c0 = 2 * half;                 // this thread's chunks of 32 keys: c0, c0 + 1
        const int lim0 = nk - c0 * 32, lim1 = lim0 - 32;   // valid keys in them (may be <= 0)
        if (trace && g < 16) p.dbg[g * 12 + 0] = clock64();
        mbar_spin(s_full, g & 1);
        tc_fence_after();
        if (trace && g < 16) p.dbg[g * 12 + 1] = clock64();
        uint32_t r[32];
        auto cmax = [&](int lim) -> float {      // raw max of the valid columns of the chunk in r[]
          if (lim >= 32) {
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
        // P chunk = exp2(S c - m_ref) of r[] -> TMEM columns [pcol, pcol + 16); returns the fp32 row sum
        auto cexp_store = [&](int lim, uint32_t pcol) -> float {
          uint32_t w[16];
          float sum;
          if (lim >= 32) {
            const uint64_t sc2 = pack2(sc, sc), nm2 = pack2(-m_ref, -m_ref);
            uint64_t sum2 = pack2(0.f, 0.f);
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              float t0, t1;
              unpack2(ffma2(pack2(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1])), sc2, nm2), t0, t1);
              float p0, p1;
              if (POLY && (e & 3) == 3) {
                ex2_poly2(t0, t1, p0, p1);
              } else {
                p0 = ex2_approx(t0);
                p1 = ex2_approx(t1);
              }
              sum2 = fadd2(sum2, pack2(p0, p1));
              w[e] = Op16<BF16>::pack(p0, p1);
            }
            float a0, a1;
            unpack2(sum2, a0, a1);
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
          tmem_st_32x16(pcol, w);
          return sum;
        };
        const uint32_t sa0 = s_addr + c0 * 32, sa1 = sa0 + 32, pa0 = p_addr + c0 * 16, pa1 = pa0 + 16;
        if (j == 0) {
          // first tile of the unit: the reference max = max over the whole first tile (both halves) and the leftover keys
          float mx = -INFINITY;
          if (lim0 > 0) {
            tmem_ld_32x32(sa0, r);
            tmem_ld_wait();
            mx = cmax(lim0);
          }
          if (lim1 > 0) {
            tmem_ld_32x32(sa1, r);
            tmem_ld_wait();
            mx = fmaxf(mx, cmax(lim1));
          }
#pragma unroll
          for (int e = 0; e < kExtraMax; ++e) mx = fmaxf(mx, s_x[e]);
          xput(2, mx);
          pair_sync();
          m_ref = fmaxf(mx, xget(2)) * sc;
        }
        // ---- hot path: chunk 0: load -> max -> exp2 -> P store; chunk 1: load -> max; exchange of the tile max between
        // the halves; S is released (Q K_{j+1}^T runs under the exponentials of chunk 1); chunk 1: exp2 -> P store
        float mx_raw = -INFINITY, sum = 0.f;
        bool waited = false;
        auto wait_p_free = [&]() {
          if (!waited && g >= 1) {
            mbar_spin(p_free, (g - 1) & 1);      // P V of the previous tile retired: P may be overwritten, O is quiescent
            tc_fence_after();
          }
          waited = true;
        };
        if (lim0 > 0) {
          tmem_ld_32x32(sa0, r);
          tmem_ld_wait();
          mx_raw = cmax(lim0);
          wait_p_free();
          alt_acquire();
          sum = cexp_store(lim0, pa0);
        }
        alt_release();                           // (every warp releases twice per tile, with or without valid keys)
        if (trace && g < 16) p.dbg[g * 12 + 3] = clock64();
        if (lim1 > 0) {
          tmem_ld_32x32(sa1, r);
          tmem_ld_wait();
          mx_raw = fmaxf(mx_raw, cmax(lim1));
        }
        // lazy rescale: only when the tile's row max (over BOTH halves) exceeds the reference max by more than 2^8
        xput(g & 1, mx_raw);
        pair_sync();
        const float mx_tile = fmaxf(mx_raw, xget(g & 1));
        const bool need = mx_tile * sc > m_ref + kRescaleThreshold;
        if (!__any_sync(0xffffffffu, need)) {    // both warps of the quadrant see the same rows, i.e. decide alike
          tc_fence_before();
          mbar_arrive(s_free);                   // this thread holds its last chunk in registers
          if (lim1 > 0) {
            wait_p_free();
            alt_acquire();
            sum += cexp_store(lim1, pa1);
          }
          alt_release();
        } else {
          wait_p_free();
          alt_acquire();
          const float m_new = need ? mx_tile * sc : m_ref;
          const float f = ex2_approx(m_ref - m_new);   // 1 for rows that keep their reference
          m_ref = m_new;
          l *= f;
          sum = 0.f;                             // P again with the new reference max: chunk 1 from registers ...
          if (lim1 > 0) sum = cexp_store(lim1, pa1);
          if (lim0 > 0) {                        // ... chunk 0 from S, which has not been released yet
            tmem_ld_32x32(sa0, r);
            tmem_ld_wait();
            sum += cexp_store(lim0, pa0);
          }
          alt_release();
          tc_fence_before();
          mbar_arrive(s_free);
          if (j > 0) {                           // each half rescales its 32 columns of O
            tmem_ld_32x32(o_addr + half * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t w[16];
#pragma unroll
              for (int e = 0; e < 16; ++e) w[e] = __float_as_uint(__uint_as_float(r[hh * 16 + e]) * f);
              tmem_st_32x16(o_addr + half * 32 + hh * 16, w);
            }
          }
        }
        if (trace && g < 16) p.dbg[g * 12 + 4] = clock64();
        l += sum;
        wait_p_free();                           // (threads without valid keys in this tile have not waited yet)
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_ready);
        if (trace && g < 16) p.dbg[g * 12 + 11] = clock64();
      }
      // ---- epilogue of the unit: (O + leftover keys) / l -> global; each half stores 32 of the 64 columns, in two
      // passes of 16 (the 80-register budget of a 384-thread CTA does not hold a 32-column row plus the extras)
      if (alt && lane == 0 && n_tiles > 0) atomicSub(&alt_present[slot], 1);
      const int gt = g - 1;                      // trace row of the unit's last tile
      if (trace && gt >= 0 && gt < 16) p.dbg[gt * 12 + 5] = clock64();
      if (n_tiles == 0) {
        // no tensor-core tile at all (Nk <= kExtraMax): the reference max comes from the leftover keys alone
        float mx = -INFINITY;
#pragma unroll
        for (int e = 0; e < kExtraMax; ++e) mx = fmaxf(mx, s_x[e]);
        m_ref = mx * sc;
      }
      float p_x[kExtraMax];
#pragma unroll
      for (int e = 0; e < kExtraMax; ++e) {
        p_x[e] = e < n_extra ? ex2_approx(fmaf(s_x[e], sc, -m_ref)) : 0.f;   // fp32: no range issue whatever the score
        if (half == 0) l += p_x[e];
      }
      xput(3, l);
      pair_sync();
      const float inv = 1.0f / (l + xget(3));
      if (n_tiles > 0) {
        mbar_spin(o_done, i & 1);
        tc_fence_after();
      }
      if (trace && gt >= 0 && gt < 16) p.dbg[gt * 12 + 6] = clock64();
      const bool valid = (q0 + row) < p.row0;    // row0 = Nq, or the first row of the CUDA-core path
      uint16_t* og = p.o + b * p.o_bs + static_cast<int64_t>(q0 + row) * p.ldo + static_cast<int64_t>(h) * kD + half * 32;
      const uint8_t* xv = xk + kXBytes;
#pragma unroll 1
      for (int sub = 0; sub < 2; ++sub) {
        uint32_t ro[16];
        if (n_tiles > 0) {
          tmem_ld_32x16(o_addr + half * 32 + sub * 16, ro);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) ro[e] = 0u;
        }
        float ov[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) ov[e] = __uint_as_float(ro[e]);
        if (n_extra > 0) {
#pragma unroll
          for (int e = 0; e < kExtraMax; ++e) {
            if (e < n_extra) {
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                const uint4 vv = *reinterpret_cast<const uint4*>(xv + e * 128 + (((half * 4 + sub * 2 + c) ^ e) << 4));
                const uint32_t vw[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                  const float2 f2 = Op16<BF16>::unpack(vw[d]);
                  ov[c * 8 + 2 * d] = fmaf(p_x[e], f2.x, ov[c * 8 + 2 * d]);
                  ov[c * 8 + 2 * d + 1] = fmaf(p_x[e], f2.y, ov[c * 8 + 2 * d + 1]);
                }
              }
            }
          }
        }
        if (valid) {
          uint4* dst = reinterpret_cast<uint4*>(og + sub * 16);
#pragma unroll
          for (int e = 0; e < 2; ++e)
            dst[e] = make_uint4(Op16<BF16>::pack(ov[8 * e] * inv, ov[8 * e + 1] * inv), Op16<BF16>::pack(ov[8 * e + 2] * inv, ov[8 * e + 3] * inv),
                                Op16<BF16>::pack(ov[8 * e + 4] * inv, ov[8 * e + 5] * inv), Op16<BF16>::pack(ov[8 * e + 6] * inv, ov[8 * e + 7] * inv));
        }
      }
      if (n_tiles > 0) {
        tc_fence_before();
        mbar_arrive(o_free);                     // O has been read: the next unit's first P V may overwrite it
      }
      if (n_extra > 0) mbar_arrive(&x_empty[i & 1]);   // this thread is done with the unit's leftover-key rows
      if (trace && gt >= 0 && gt < 16) p.dbg[gt * 12 + 10] = clock64();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.dbg && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.dbg[192 + blockIdx.x * 4 + 3] = t;
  }
  if (warp_g == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemColsAttn * SLOTS);
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
  CUtensorMap tkx, tvx;   // the leftover keys (rows 128 n_tiles ...) as 16-row boxes
  SATB_PROPAGATE(make_tmap_rows(&tkx, k, k_cols, Nk, batch, ldk, k_bs, kXRows));
  SATB_PROPAGATE(make_tmap_rows(&tvx, v, v_cols, Nk, batch, ldv, v_bs, kXRows));
  AttnTcArgs a;
  a.o = static_cast<uint16_t*>(o);
  a.ldo = ldo; a.o_bs = o_bs;
  a.Nq = Nq; a.Nk = Nk; a.group = H / H_kv; a.H = H; a.batch = batch;
  a.q_col = q_col; a.k_col = k_col; a.v_col = v_col;
  // keys: full 128-key tiles on the tensor cores; a remainder of <= kExtraMax keys in the epilogue, a larger one as one
  // more (partial) tile
  const int krem = Nk % kK;
  const bool extra = krem != 0 && krem <= kExtraMax;
  a.n_tiles = Nk / kK + ((krem != 0 && !extra) ? 1 : 0);
  a.n_extra = extra ? krem : 0;
  // query rows: full 128-row tiles on the tensor cores; a remainder of <= kRowPathMax rows either on CUDA cores (warp 2
  // of the CTAs, concurrently) or as one more, partial, tensor-core tile.  One row task is a serial walk over all keys
  // by a single warp (~110 cycles per key, measured: 65 us for 1025 keys), so it only pays when every CTA has enough
  // tensor-core units to hide it behind - large batches; otherwise (B = 1, 2) the partial tile rides in the slack of
  // the last wave for free.
  const int rem = Nq % kQ;
  const int slots_all = 2 * device_sm_count();
  bool row_path = rem != 0 && rem <= kRowPathMax;
  static int force_rows = -2;
  if (force_rows == -2) {
    const char* e = getenv("SATB_ATTN_ROWPATH");     // 0 / 1: force the partial tile / the CUDA-core rows (A-B switch)
    force_rows = e ? (atoi(e) != 0) : -1;
  }
  if (row_path && force_rows == 0) row_path = false;
  if (row_path && force_rows < 0) {
    const double unit_cycles = a.n_tiles * 2900.0 + 4500.0, row_cycles = 110.0 * Nk;
    const double units_per_cta = static_cast<double>(batch) * H * (Nq / kQ) / slots_all;
    const double rows_per_cta = std::ceil(static_cast<double>(batch) * H * rem / slots_all);
    if (units_per_cta * unit_cycles < rows_per_cta * row_cycles) row_path = false;
  }
  a.n_qt = Nq / kQ + ((rem != 0 && !row_path) ? 1 : 0);
  a.n_units = batch * H * a.n_qt;
  a.row0 = row_path ? Nq - rem : Nq;
  a.n_rows = row_path ? rem : 0;
  a.q = static_cast<const uint16_t*>(q); a.k = static_cast<const uint16_t*>(k); a.v = static_cast<const uint16_t*>(v);
  a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.q_bs = q_bs; a.k_bs = k_bs; a.v_bs = v_bs;
  a.scale_log2 = (1.0f / sqrtf(64.0f)) * 1.4426950408889634f;
  a.dbg = dbg;
  const int row_tasks = batch * H * a.n_rows;
  int grid = a.n_units > row_tasks ? a.n_units : row_tasks;   // pipelines with work
  static int ctas_per_sm = -1, n_slots = -1, alternate = -1;
  if (ctas_per_sm < 0) {
    const char* e = getenv("SATB_ATTN_CTAS_PER_SM");   // 1: one pipeline per SM (A/B measurement of the SFU sharing)
    ctas_per_sm = (e && atoi(e) == 1) ? 1 : 2;
    e = getenv("SATB_ATTN_SLOTS");                     // 1: two CTAs of one pipeline per SM; 2: one CTA of two pipelines
    n_slots = e ? (atoi(e) == 1 ? 1 : 2) : kAttnSlotsDefault;
    e = getenv("SATB_ATTN_ALT");                       // 0: the two pipelines of a CTA do not take turns on the SFU
    alternate = e ? (atoi(e) != 0) : 1;
  }
  a.alternate = alternate;
  const int slots = ctas_per_sm * device_sm_count();
  if (grid > slots) grid = slots;
  if (grid <= 0) return 0;
  if (n_slots == 2) grid = (grid + 1) / 2;             // CTAs of two pipelines
  static int poly = -1;
  if (poly < 0) {
    const char* e = getenv("SATB_ATTN_POLY");        // 0 / 1: A-B of the FMA-pipe exponentials
    poly = e ? (atoi(e) != 0) : kAttnPolyDefault;
  }
  const int smem_bytes = attn_smem_bytes(n_slots);
  auto prepare = [&](auto kern) -> int {
    SATB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    // two CTAs per SM need 2 x 98 KB: ask for the largest shared-memory carveout
    SATB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    return 0;
  };
  auto go = [&](auto kern, PerDeviceOnce& once) -> int {
    if (once.first()) SATB_PROPAGATE(prepare(kern));
    SATB_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(kAttnThreads * n_slots), smem_bytes, stream, tq, tk, tv, tkx, tvx, a));
    return 0;
  };
  static PerDeviceOnce o00, o01, o10, o11, t00, t01, t10, t11;
  if (n_slots == 2) {
    if (bf16) SATB_PROPAGATE(poly ? go(attn_tc_kernel<true, true, 2>, t11) : go(attn_tc_kernel<true, false, 2>, t10));
    else SATB_PROPAGATE(poly ? go(attn_tc_kernel<false, true, 2>, t01) : go(attn_tc_kernel<false, false, 2>, t00));
  } else {
    if (bf16) SATB_PROPAGATE(poly ? go(attn_tc_kernel<true, true, 1>, o11) : go(attn_tc_kernel<true, false, 1>, o10));
    else SATB_PROPAGATE(poly ? go(attn_tc_kernel<false, true, 1>, o01) : go(attn_tc_kernel<false, false, 1>, o00));
  }
  count_launch();
  SATB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Debug: resident CTAs per SM the runtime reports for the attention kernel with `dyn_smem` bytes of dynamic shared
// memory and the given carveout preference (percent, -1 = leave unchanged); tests / profiling only.
int debug_attention_occupancy(int dyn_smem, int carveout_pct) {
  auto kern = attn_tc_kernel<false, false, 1>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem) != cudaSuccess) return -1;
  if (carveout_pct >= 0) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_pct);
  int nb = -1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kAttnThreads, dyn_smem) != cudaSuccess) return -2;
  return nb;
}

}  // namespace satb
