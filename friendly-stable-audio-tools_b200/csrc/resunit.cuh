// Fused Oobleck ResidualUnit for 128-channel stages (models/autoencoders.py:45-68):
//
//   y = x + conv1x1( snake2( conv7_dilated( snake1(x) ) ) )
//
// One kernel per unit, one CTA pair (tcgen05 cta_group::2) per 256 positions x 128 channels tile.
//
//  * k7 convolution = 7-tap implicit GEMM whose A operand is loaded ONCE per tile: TMA brings the
//    128 + 6*dilation activation rows of a 64-channel k-block into a 128B-swizzled smem slot and tap t
//    reads it through a matrix descriptor whose start address is shifted by t*dilation rows (the
//    swizzle is a function of absolute smem address bits, see conv_halo.cuh).  Only the tap weights
//    stream per tile (8 KB per (tap, k-block) and CTA).  L2 -> SM traffic per tile and CTA drops from
//    343 KB (A reloaded per tap) to 161 KB, which is what bounded the two-launch version.
//  * Its fp32 accumulator never leaves the SM: the epilogue warps add the bias, apply snake2, round to
//    the 16-bit operand type and write the tile to smem in the swizzled K-major layout; a second
//    tcgen05 GEMM (K = 128, the 1x1 convolution, weights resident) produces the unit's output in a
//    second TMEM accumulator, whose epilogue is EpiConv: + bias + fp32 skip, fp32 raw stream out, and
//    the consumer's Snake fused into the 16-bit copy.
//
// Pipeline per CTA pair (tile t):  tensor pipe  conv7(t) -> G2(t-1) -> conv7(t+1) -> G2(t) ...
//                                  epilogue     A(t) [acc1 -> smem]   B(t-1) [acc2 -> HBM] ...
// TMEM: acc1 double-buffered at columns [0,256), acc2 double-buffered at [256,512).
// Warps: 0 TMA producer, 1 MMA issuer (CTA 0), 2 TMEM allocator, 3 forwarder, 4-11 epilogue.
#pragma once
#include "gemm.cuh"

namespace satb {

struct ResUnitShape {
  int L;        // positions per batch item
  int batches;
  int dil;      // dilation of the k7 convolution (halo = 6 * dil rows, <= 54)
};

template <bool BF16>
struct ResUnitParams {
  const float* bias7;    // [128]
  const float* sn2_a;    // inner Snake: e^alpha
  const float* sn2_ib;   // inner Snake: 1/(e^beta + 1e-9)
  typename EpiConv<BF16>::Params out;   // 1x1-conv epilogue (bias, skip, raw out, 16-bit out + next Snake)
};

struct ResUnitCfg {
  static constexpr int kC = 128;
  static constexpr int kTaps = 7;
  static constexpr int kMaxDil = 9;
  static constexpr int kKb = kC / kBlockK;                                        // 2 k-blocks of 64 channels
  static constexpr int kSlotA = ((kBlockM + (kTaps - 1) * kMaxDil) * 128 + 1023) / 1024 * 1024;   // 24 KB halo tile
  static constexpr int kStagesA = 3;
  static constexpr int kTileB = (kC / 2) * kBlockK * 2;      // 8 KB: this CTA's 64 output channels of one (tap, k-block)
  static constexpr int kStagesB = 8;
  static constexpr int kTileA2 = kBlockM * kBlockK * 2;      // 16 KB per k-block of the snake2(conv7) tile
  static constexpr int kOffB = kStagesA * kSlotA;
  static constexpr int kOffA2 = kOffB + kStagesB * kTileB;
  static constexpr int kOffW1 = kOffA2 + kKb * kTileA2;
  static constexpr int kOffBars = kOffW1 + kKb * kTileB;
  static constexpr int kOffParams = kOffBars + 512;          // bias7 | sn2_a | sn2_ib, 3 x 128 floats
  static constexpr int kOffEpiStage = kOffParams + 3 * kC * 4;
  static constexpr int kEpiStage = 32 * 36 * 4;
  static constexpr int kSmemBytes = kOffEpiStage + kEpiWarps * kEpiStage + 1024;
  static constexpr int kTmemCols = 512;
  static_assert(kSmemBytes <= 227 * 1024, "smem budget");
  __host__ __device__ static int halo_rows(int dil) { return kBlockM + (kTaps - 1) * dil; }
};

template <bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
resunit_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB7,
                            const __grid_constant__ CUtensorMap tmB1, const ResUnitShape s,
                            const ResUnitParams<BF16> ep) {
  using Cfg = ResUnitCfg;
  using Epi = EpiConv<BF16>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bs = smem + Cfg::kOffB;
  uint8_t* a2s = smem + Cfg::kOffA2;
  uint8_t* w1s = smem + Cfg::kOffW1;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBars);
  uint64_t* a_full = bars;                           // [3] CTA 0 (credited by both CTAs' TMA)
  uint64_t* a_empty = a_full + Cfg::kStagesA;        // [3] per CTA, multicast commit
  uint64_t* b_full = a_empty + Cfg::kStagesA;        // [8] CTA 0
  uint64_t* b_empty = b_full + Cfg::kStagesB;        // [8] per CTA, multicast commit
  uint64_t* acc1_full = b_empty + Cfg::kStagesB;     // [2] per CTA, multicast commit
  uint64_t* a2_full = acc1_full + 2;                 // CTA 0, 2 arrivals (one forwarder thread per CTA)
  uint64_t* a2_local = a2_full + 1;                  // per CTA, 8 arrivals (this CTA's epilogue warps)
  uint64_t* a2_empty = a2_local + 1;                 // per CTA, multicast commit of G2
  uint64_t* acc2_full = a2_empty + 1;                // [2] per CTA, multicast commit
  uint64_t* acc2_empty = acc2_full + 2;              // [2] CTA 0, 16 arrivals
  uint64_t* w1_full = acc2_empty + 2;                // CTA 0
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w1_full + 1);
  float* prm = reinterpret_cast<float*>(smem + Cfg::kOffParams);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  const int m_tiles = (s.L + 2 * kBlockM - 1) / (2 * kBlockM);
  const int total_tiles = m_tiles * s.batches;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB7);
    tma_prefetch_desc(&tmB1);
    for (int i = 0; i < Cfg::kStagesA; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < Cfg::kStagesB; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc1_full[i], 1);
      mbar_init(&acc2_full[i], 1);
      mbar_init(&acc2_empty[i], 2 * kEpiWarps);
    }
    mbar_init(a2_full, 2);
    mbar_init(a2_local, kEpiWarps);
    mbar_init(a2_empty, 1);
    mbar_init(w1_full, 1);
    fence_mbar_init();
  }
  cluster_sync_all();
  if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  // Parameters and weights never depend on the previous kernel: they are fetched before the
  // programmatic-dependency wait (per-channel parameters here, weight tiles by the producer below).
  if (threadIdx.x >= 128 && threadIdx.x < 128 + 3 * 32) {     // 96 threads x float4 = 3 x 128 floats
    const int i = threadIdx.x - 128;
    const float* src = i < 32 ? ep.bias7 : (i < 64 ? ep.sn2_a : ep.sn2_ib);
    reinterpret_cast<float4*>(prm)[i] =
        src ? __ldg(reinterpret_cast<const float4*>(src) + (i & 31)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  if (warp != 0) pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      // ------------------------------------------------------ TMA producer (both CTAs)
      if (rank == 0) mbar_expect_tx(w1_full, 2 * Cfg::kKb * Cfg::kTileB);
      for (int kb = 0; kb < Cfg::kKb; ++kb)
        tma_load_2d_2sm(w1s + kb * Cfg::kTileB, &tmB1, w1_full, kb * kBlockK, rank * (Cfg::kC / 2));
      const int a_tx = Cfg::halo_rows(s.dil) * 128;
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int b_pre = 0;   // conv7 weight tiles of the first tile issued before the dependency wait
      if (cluster_id < total_tiles) {
        for (; b_pre < Cfg::kStagesB && b_pre < Cfg::kKb * Cfg::kTaps; ++b_pre) {
          const int kb = b_pre / Cfg::kTaps, tap = b_pre - kb * Cfg::kTaps;
          if (rank == 0) mbar_expect_tx(&b_full[sb], 2 * Cfg::kTileB);
          tma_load_2d_2sm(bs + sb * Cfg::kTileB, &tmB7, &b_full[sb], kb * kBlockK, tap * Cfg::kC + rank * (Cfg::kC / 2));
          if (++sb == Cfg::kStagesB) {
            sb = 0;
            pb ^= 1;
          }
        }
      }
      pdl_wait();
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
        const int batch = tile / m_tiles;
        const int m0 = (tile - batch * m_tiles) * 2 * kBlockM + rank * kBlockM;
        for (int kb = 0; kb < Cfg::kKb; ++kb) {
          mbar_wait(&a_empty[sa], pa ^ 1);
          if (rank == 0) mbar_expect_tx(&a_full[sa], 2 * a_tx);
          tma_load_4d_2sm(smem + sa * Cfg::kSlotA, &tmA, &a_full[sa], kb * kBlockK, 0, m0 - 3 * s.dil, batch);
          if (++sa == Cfg::kStagesA) {
            sa = 0;
            pa ^= 1;
          }
          for (int tap = 0; tap < Cfg::kTaps; ++tap) {
            if (tile == cluster_id && kb * Cfg::kTaps + tap < b_pre) continue;   // already in flight
            mbar_wait(&b_empty[sb], pb ^ 1);
            if (rank == 0) mbar_expect_tx(&b_full[sb], 2 * Cfg::kTileB);
            tma_load_2d_2sm(bs + sb * Cfg::kTileB, &tmB7, &b_full[sb], kb * kBlockK, tap * Cfg::kC + rank * (Cfg::kC / 2));
            if (++sb == Cfg::kStagesB) {
              sb = 0;
              pb ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      // ------------------------------------------------------ MMA issuer (CTA 0 for the pair)
      constexpr uint32_t idesc = make_idesc_f16(2 * kBlockM, Cfg::kC, BF16);
      const uint32_t a2_addr = smem_u32(a2s), w1_addr = smem_u32(w1s), b_addr0 = smem_u32(bs);
      auto g2 = [&](int j) {   // 1x1 convolution of tile j: acc2[j & 1] = A2 x W1^T
        mbar_wait_cluster(a2_full, j & 1);
        mbar_wait(&acc2_empty[j & 1], ((j >> 1) & 1) ^ 1);
        if (j == 0) mbar_wait(w1_full, 0);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + 256 + (j & 1) * Cfg::kC;
#pragma unroll
        for (int kb = 0; kb < Cfg::kKb; ++kb) {
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t da = make_desc_kmajor_sw128(a2_addr + kb * Cfg::kTileA2 + k * kUmmaK * 2);
            const uint64_t db = make_desc_kmajor_sw128(w1_addr + kb * Cfg::kTileB + k * kUmmaK * 2);
            umma_f16_ss_2sm(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit_2sm(a2_empty);
        umma_commit_2sm(&acc2_full[j & 1]);
      };
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int it = 0;
      const uint32_t tap_bytes = s.dil * 128;
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++it) {
        // acc1[it & 1] is free: phase A of tile it-2 was observed (a2_full) before G2(it-2) was issued
        const uint32_t d_tmem = tmem_base + (it & 1) * Cfg::kC;
        for (int kb = 0; kb < Cfg::kKb; ++kb) {
          mbar_wait(&a_full[sa], pa);
          const uint32_t a_addr = smem_u32(smem + sa * Cfg::kSlotA);
          for (int tap = 0; tap < Cfg::kTaps; ++tap) {
            mbar_wait(&b_full[sb], pb);
            tc_fence_after();
            const uint32_t a_tap = a_addr + tap * tap_bytes;     // row-shifted view of the halo tile
            const uint32_t b_addr = b_addr0 + sb * Cfg::kTileB;
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              const uint64_t da = make_desc_kmajor_sw128(a_tap + k * kUmmaK * 2);
              const uint64_t db = make_desc_kmajor_sw128(b_addr + k * kUmmaK * 2);
              umma_f16_ss_2sm(d_tmem, da, db, idesc, (kb | tap | k) != 0 ? 1u : 0u);
            }
            umma_commit_2sm(&b_empty[sb]);
            if (++sb == Cfg::kStagesB) {
              sb = 0;
              pb ^= 1;
            }
          }
          umma_commit_2sm(&a_empty[sa]);
          if (++sa == Cfg::kStagesA) {
            sa = 0;
            pa ^= 1;
          }
        }
        umma_commit_2sm(&acc1_full[it & 1]);
        if (it > 0) g2(it - 1);
      }
      if (it > 0) g2(it - 1);
    }
  } else if (warp == 3) {
    if (elect_one()) {
      // ------------------------------------------------------ forwarder (both CTAs)
      // The epilogue warps publish their smem tile on a CTA-local barrier; this otherwise idle thread
      // relays it to the MMA thread in CTA 0 with cluster-scope release semantics, so the expensive
      // cluster-wide memory barrier is not executed by warps that have global loads/stores in flight.
      int it = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++it) {
        mbar_wait(a2_local, it & 1);
        mbar_arrive_remote_cluster(a2_full, 0);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (both CTAs)
    const int q = warp & 3;                        // TMEM lane quadrant
    const int half = (warp - 4) >> 2;              // two warps per quadrant: 32-column chunks half, half + 2
    const int row = q * 32 + lane;                 // this thread's row of the CTA's 128 positions
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float* stage_buf = reinterpret_cast<float*>(smem + Cfg::kOffEpiStage + (warp - 4) * Cfg::kEpiStage);

    auto phase_a = [&](int it) {   // acc1 -> + bias7 -> snake2 -> 16-bit -> swizzled smem tile
      mbar_wait(&acc1_full[it & 1], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_row = t_lane + (it & 1) * Cfg::kC;
      uint32_t ra[32], rb[32];
      tmem_ld_32x32(t_row + half * 32, ra);
      tmem_ld_32x32(t_row + (half + 2) * 32, rb);
      if (it > 0) mbar_wait(a2_empty, (it - 1) & 1);   // G2 of the previous tile has consumed the smem tile
      tmem_ld_wait();
      auto chunk = [&](int ci, const uint32_t (&r)[32]) {
        uint32_t o[16];
        const float* pb = prm + ci * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const ulonglong2 b2 = *reinterpret_cast<const ulonglong2*>(pb + 4 * j);
          const ulonglong2 a2 = *reinterpret_cast<const ulonglong2*>(pb + Cfg::kC + 4 * j);
          const ulonglong2 i2 = *reinterpret_cast<const ulonglong2*>(pb + 2 * Cfg::kC + 4 * j);
          const uint64_t acc01 = (static_cast<uint64_t>(r[4 * j + 1]) << 32) | r[4 * j];
          const uint64_t acc23 = (static_cast<uint64_t>(r[4 * j + 3]) << 32) | r[4 * j + 2];
          float v0, v1, v2, v3;
          f2_unpack(snake_fast2(f2_add(acc01, b2.x), a2.x, i2.x), v0, v1);
          f2_unpack(snake_fast2(f2_add(acc23, b2.y), a2.y, i2.y), v2, v3);
          o[2 * j] = Op16<BF16>::pack(v0, v1);
          o[2 * j + 1] = Op16<BF16>::pack(v2, v3);
        }
        uint8_t* rowp = a2s + (ci >> 1) * Cfg::kTileA2 + row * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c16 = ((ci & 1) * 4 + j) ^ (row & 7);
          *reinterpret_cast<uint4*>(rowp + c16 * 16) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        }
      };
      chunk(half, ra);
      chunk(half + 2, rb);
      fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor core's async proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a2_local);
    };

    auto make_ctx = [&](int batch, int m0, int ci) {
      EpiCtx c;
      c.l = m0 + row;
      c.batch = batch;
      c.row = batch * s.L + c.l;
      c.valid = c.l < s.L;
      c.l0 = m0 + q * 32;
      c.L = s.L;
      c.lane = lane;
      c.stage = stage_buf;
      c.col0 = ci * 32;
      return c;
    };
    auto phase_b = [&](int j, int batch, int m0) {   // acc2 -> EpiConv (bias, skip, raw, snake_next)
      const EpiCtx c0 = make_ctx(batch, m0, half), c1 = make_ctx(batch, m0, half + 2);
      float4 rs0[8], rs1[8];
      Epi::prefetch(ep.out, c0, rs0);                // skip values: in flight while we wait for G2
      mbar_wait(&acc2_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t t_row = t_lane + 256 + (j & 1) * Cfg::kC;
      uint32_t r[32];
      tmem_ld_32x32(t_row + half * 32, r);
      Epi::prefetch(ep.out, c1, rs1);
      tmem_ld_wait();
      Epi::finish(ep.out, c0, r, rs0);
      tmem_ld_32x32(t_row + (half + 2) * 32, r);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&acc2_empty[j & 1], 0);   // accumulator fully in registers
      Epi::finish(ep.out, c1, r, rs1);
    };

    int it = 0, prev_batch = 0, prev_m0 = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++it) {
      const int batch = tile / m_tiles;
      const int m0 = (tile - batch * m_tiles) * 2 * kBlockM + rank * kBlockM;
      phase_a(it);
      if (it > 0) phase_b(it - 1, prev_batch, prev_m0);
      prev_batch = batch;
      prev_m0 = m0;
    }
    if (it > 0) phase_b(it - 1, prev_batch, prev_m0);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
  }
}

// tmA must be built with box rows = ResUnitCfg::halo_rows(s.dil); tmB7 / tmB1 with 64-row boxes.
template <bool BF16>
int launch_resunit(const CUtensorMap& tmA, const CUtensorMap& tmB7, const CUtensorMap& tmB1, const ResUnitShape& s,
                   const ResUnitParams<BF16>& ep, cudaStream_t stream) {
  using Cfg = ResUnitCfg;
  SATB_REQUIRE(s.dil >= 1 && s.dil <= Cfg::kMaxDil, "resunit: dilation out of range");
  auto kern = resunit_tcgen05_2cta_kernel<BF16>;
  static PerDeviceOnce attr;
  if (attr.first()) SATB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  const int total = ceil_div(s.L, 2 * kBlockM) * s.batches;
  if (total <= 0) return 0;
  int clusters = device_sm_count() / 2;
  if (clusters > total) clusters = total;
  SATB_CHECK_CUDA(launch_pdl(kern, dim3(2 * clusters), dim3(kGemmThreads), Cfg::kSmemBytes, stream, tmA, tmB7, tmB1, s, ep));
  count_launch();
  return 0;
}


// ------------------------------------------------------------------------------------------------
// 256-channel variant.  Same algorithm; differences forced by the resources of one SM:
//   * both accumulators are 256 TMEM columns, so neither is double-buffered: the tensor pipe idles while the
//     epilogue turns acc1 into the smem tile (phase A), and the epilogue of tile t (phase B) overlaps conv7(t+1);
//   * the 1x1 weights (64 KB per CTA) do not stay resident: they stream through the same B ring, four
//     16 KB tiles per output tile, after the 28 conv7 weight tiles;
//   * A ring 2 x 23 KB, B ring 4 x 16 KB, smem tile 64 KB, 8 epilogue warps with four 32-column chunks each.
struct ResUnit256Cfg {
  static constexpr int kC = 256;
  static constexpr int kTaps = 7;
  static constexpr int kMaxDil = 9;
  static constexpr int kKb = kC / kBlockK;                                         // 4 k-blocks
  static constexpr int kSlotA = ((kBlockM + (kTaps - 1) * kMaxDil) * 128 + 1023) / 1024 * 1024;
  static constexpr int kStagesA = 2;
  static constexpr int kTileB = (kC / 2) * kBlockK * 2;                            // 16 KB
  static constexpr int kStagesB = 4;
  static constexpr int kTileA2 = kBlockM * kBlockK * 2;                            // 16 KB per k-block
  static constexpr int kOffB = kStagesA * kSlotA;
  static constexpr int kOffA2 = kOffB + kStagesB * kTileB;
  static constexpr int kOffBars = kOffA2 + kKb * kTileA2;
  static constexpr int kOffParams = kOffBars + 512;
  static constexpr int kOffEpiStage = kOffParams + 3 * kC * 4;
  static constexpr int kEpiStage = 32 * 36 * 4;
  static constexpr int kSmemBytes = kOffEpiStage + kEpiWarps * kEpiStage + 1024;
  static constexpr int kTmemCols = 512;
  static_assert(kSmemBytes <= 227 * 1024, "smem budget");
  static_assert(kOffB % 1024 == 0 && kOffA2 % 1024 == 0, "swizzled tiles must be 1024-byte aligned");
  __host__ __device__ static int halo_rows(int dil) { return kBlockM + (kTaps - 1) * dil; }
};

template <bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
resunit256_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB7,
                               const __grid_constant__ CUtensorMap tmB1, const ResUnitShape s,
                               const ResUnitParams<BF16> ep) {
  using Cfg = ResUnit256Cfg;
  using Epi = EpiConv<BF16>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bs = smem + Cfg::kOffB;
  uint8_t* a2s = smem + Cfg::kOffA2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBars);
  uint64_t* a_full = bars;                           // [2] CTA 0
  uint64_t* a_empty = a_full + Cfg::kStagesA;        // [2] per CTA
  uint64_t* b_full = a_empty + Cfg::kStagesA;        // [4] CTA 0
  uint64_t* b_empty = b_full + Cfg::kStagesB;        // [4] per CTA
  uint64_t* acc1_full = b_empty + Cfg::kStagesB;     // per CTA, multicast commit
  uint64_t* a2_full = acc1_full + 1;                 // CTA 0, 2 arrivals (forwarders)
  uint64_t* a2_local = a2_full + 1;                  // per CTA, 8 arrivals
  uint64_t* a2_empty = a2_local + 1;                 // per CTA, multicast commit of G2
  uint64_t* acc2_full = a2_empty + 1;                // per CTA, multicast commit
  uint64_t* acc2_empty = acc2_full + 1;              // CTA 0, 16 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_empty + 1);
  float* prm = reinterpret_cast<float*>(smem + Cfg::kOffParams);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int m_tiles = (s.L + 2 * kBlockM - 1) / (2 * kBlockM);
  const int total_tiles = m_tiles * s.batches;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB7);
    tma_prefetch_desc(&tmB1);
    for (int i = 0; i < Cfg::kStagesA; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < Cfg::kStagesB; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    mbar_init(acc1_full, 1);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, 2 * kEpiWarps);
    mbar_init(a2_full, 2);
    mbar_init(a2_local, kEpiWarps);
    mbar_init(a2_empty, 1);
    fence_mbar_init();
  }
  cluster_sync_all();
  if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  if (threadIdx.x >= 128 && threadIdx.x < 128 + 3 * 64) {     // 192 threads x float4 = 3 x 256 floats (static: before the wait)
    const int i = threadIdx.x - 128;
    const float* src = i < 64 ? ep.bias7 : (i < 128 ? ep.sn2_a : ep.sn2_ib);
    reinterpret_cast<float4*>(prm)[i] =
        src ? __ldg(reinterpret_cast<const float4*>(src) + (i & 63)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  if (warp != 0) pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      // ------------------------------------------------------ TMA producer (both CTAs)
      const int a_tx = Cfg::halo_rows(s.dil) * 128;
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      auto load_b = [&](const CUtensorMap* tm, int k0, int row0) {
        mbar_wait(&b_empty[sb], pb ^ 1);
        if (rank == 0) mbar_expect_tx(&b_full[sb], 2 * Cfg::kTileB);
        tma_load_2d_2sm(bs + sb * Cfg::kTileB, tm, &b_full[sb], k0, row0);
        if (++sb == Cfg::kStagesB) {
          sb = 0;
          pb ^= 1;
        }
      };
      int b_pre = 0;   // conv7 weight tiles of the first tile issued before the dependency wait
      if (cluster_id < total_tiles)
        for (; b_pre < Cfg::kStagesB; ++b_pre) load_b(&tmB7, 0, b_pre * Cfg::kC + rank * (Cfg::kC / 2));   // kb 0, taps 0..3
      pdl_wait();
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
        const int batch = tile / m_tiles;
        const int m0 = (tile - batch * m_tiles) * 2 * kBlockM + rank * kBlockM;
        for (int kb = 0; kb < Cfg::kKb; ++kb) {
          mbar_wait(&a_empty[sa], pa ^ 1);
          if (rank == 0) mbar_expect_tx(&a_full[sa], 2 * a_tx);
          tma_load_4d_2sm(smem + sa * Cfg::kSlotA, &tmA, &a_full[sa], kb * kBlockK, 0, m0 - 3 * s.dil, batch);
          if (++sa == Cfg::kStagesA) {
            sa = 0;
            pa ^= 1;
          }
          for (int tap = 0; tap < Cfg::kTaps; ++tap) {
            if (tile == cluster_id && kb * Cfg::kTaps + tap < b_pre) continue;   // already in flight
            load_b(&tmB7, kb * kBlockK, tap * Cfg::kC + rank * (Cfg::kC / 2));
          }
        }
        for (int kb = 0; kb < Cfg::kKb; ++kb) load_b(&tmB1, kb * kBlockK, rank * (Cfg::kC / 2));   // 1x1 weights
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      // ------------------------------------------------------ MMA issuer (CTA 0 for the pair)
      constexpr uint32_t idesc = make_idesc_f16(2 * kBlockM, Cfg::kC, BF16);
      const uint32_t a2_addr = smem_u32(a2s), b_addr0 = smem_u32(bs);
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int it = 0;
      const uint32_t tap_bytes = s.dil * 128;
      const uint32_t acc1 = tmem_base, acc2 = tmem_base + Cfg::kC;
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++it) {
        // acc1 is free: phase A of the previous tile was observed (a2_full) before its G2 was issued
        for (int kb = 0; kb < Cfg::kKb; ++kb) {
          mbar_wait(&a_full[sa], pa);
          const uint32_t a_addr = smem_u32(smem + sa * Cfg::kSlotA);
          for (int tap = 0; tap < Cfg::kTaps; ++tap) {
            mbar_wait(&b_full[sb], pb);
            tc_fence_after();
            const uint32_t a_tap = a_addr + tap * tap_bytes;
            const uint32_t b_addr = b_addr0 + sb * Cfg::kTileB;
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              umma_f16_ss_2sm(acc1, make_desc_kmajor_sw128(a_tap + k * kUmmaK * 2),
                              make_desc_kmajor_sw128(b_addr + k * kUmmaK * 2), idesc, (kb | tap | k) != 0 ? 1u : 0u);
            umma_commit_2sm(&b_empty[sb]);
            if (++sb == Cfg::kStagesB) {
              sb = 0;
              pb ^= 1;
            }
          }
          umma_commit_2sm(&a_empty[sa]);
          if (++sa == Cfg::kStagesA) {
            sa = 0;
            pa ^= 1;
          }
        }
        umma_commit_2sm(acc1_full);
        // G2(it): 1x1 convolution from the smem tile, weights streamed through the B ring
        mbar_wait_cluster(a2_full, it & 1);
        mbar_wait(acc2_empty, (it & 1) ^ 1);
        for (int kb = 0; kb < Cfg::kKb; ++kb) {
          mbar_wait(&b_full[sb], pb);
          tc_fence_after();
          const uint32_t b_addr = b_addr0 + sb * Cfg::kTileB;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            umma_f16_ss_2sm(acc2, make_desc_kmajor_sw128(a2_addr + kb * Cfg::kTileA2 + k * kUmmaK * 2),
                            make_desc_kmajor_sw128(b_addr + k * kUmmaK * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm(&b_empty[sb]);
          if (++sb == Cfg::kStagesB) {
            sb = 0;
            pb ^= 1;
          }
        }
        umma_commit_2sm(a2_empty);
        umma_commit_2sm(acc2_full);
      }
    }
  } else if (warp == 3) {
    if (elect_one()) {
      // ------------------------------------------------------ forwarder (both CTAs), see the 128-channel kernel
      int it = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++it) {
        mbar_wait(a2_local, it & 1);
        mbar_arrive_remote_cluster(a2_full, 0);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (both CTAs)
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;              // chunks half, half + 2, half + 4, half + 6 of the 8
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float* stage_buf = reinterpret_cast<float*>(smem + Cfg::kOffEpiStage + (warp - 4) * Cfg::kEpiStage);

    auto chunk_a = [&](int ci, const uint32_t (&r)[32]) {
      uint32_t o[16];
      const float* pb = prm + ci * 32;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const ulonglong2 b2 = *reinterpret_cast<const ulonglong2*>(pb + 4 * j);
        const ulonglong2 a2 = *reinterpret_cast<const ulonglong2*>(pb + Cfg::kC + 4 * j);
        const ulonglong2 i2 = *reinterpret_cast<const ulonglong2*>(pb + 2 * Cfg::kC + 4 * j);
        const uint64_t acc01 = (static_cast<uint64_t>(r[4 * j + 1]) << 32) | r[4 * j];
        const uint64_t acc23 = (static_cast<uint64_t>(r[4 * j + 3]) << 32) | r[4 * j + 2];
        float v0, v1, v2, v3;
        f2_unpack(snake_fast2(f2_add(acc01, b2.x), a2.x, i2.x), v0, v1);
        f2_unpack(snake_fast2(f2_add(acc23, b2.y), a2.y, i2.y), v2, v3);
        o[2 * j] = Op16<BF16>::pack(v0, v1);
        o[2 * j + 1] = Op16<BF16>::pack(v2, v3);
      }
      uint8_t* rowp = a2s + (ci >> 1) * Cfg::kTileA2 + row * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c16 = ((ci & 1) * 4 + j) ^ (row & 7);
        *reinterpret_cast<uint4*>(rowp + c16 * 16) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
      }
    };
    auto make_ctx = [&](int batch, int m0, int ci) {
      EpiCtx c;
      c.l = m0 + row;
      c.batch = batch;
      c.row = batch * s.L + c.l;
      c.valid = c.l < s.L;
      c.l0 = m0 + q * 32;
      c.L = s.L;
      c.lane = lane;
      c.stage = stage_buf;
      c.col0 = ci * 32;
      return c;
    };

    int it = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++it) {
      const int batch = tile / m_tiles;
      const int m0 = (tile - batch * m_tiles) * 2 * kBlockM + rank * kBlockM;
      // ---- phase A: acc1 -> + bias7 -> snake2 -> 16-bit -> swizzled smem tile
      mbar_wait(acc1_full, it & 1);
      tc_fence_after();
      {
        uint32_t ra[32], rb[32];
        tmem_ld_32x32(t_lane + half * 32, ra);
        tmem_ld_32x32(t_lane + (half + 2) * 32, rb);
        if (it > 0) mbar_wait(a2_empty, (it - 1) & 1);   // G2 of the previous tile has consumed the smem tile
        tmem_ld_wait();
        chunk_a(half, ra);
        tmem_ld_32x32(t_lane + (half + 4) * 32, ra);
        chunk_a(half + 2, rb);
        tmem_ld_32x32(t_lane + (half + 6) * 32, rb);
        tmem_ld_wait();
        chunk_a(half + 4, ra);
        chunk_a(half + 6, rb);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a2_local);
      // ---- phase B: acc2 -> EpiConv (bias, skip, raw, snake_next); overlaps conv7 of the next tile
      {
        float4 rs0[8], rs1[8];
        uint32_t r[32];
        Epi::prefetch(ep.out, make_ctx(batch, m0, half), rs0);
        mbar_wait(acc2_full, it & 1);
        tc_fence_after();
        const uint32_t t_row = t_lane + Cfg::kC;
        tmem_ld_32x32(t_row + half * 32, r);
        Epi::prefetch(ep.out, make_ctx(batch, m0, half + 2), rs1);
        tmem_ld_wait();
        Epi::finish(ep.out, make_ctx(batch, m0, half), r, rs0);
        tmem_ld_32x32(t_row + (half + 2) * 32, r);
        Epi::prefetch(ep.out, make_ctx(batch, m0, half + 4), rs0);
        tmem_ld_wait();
        Epi::finish(ep.out, make_ctx(batch, m0, half + 2), r, rs1);
        tmem_ld_32x32(t_row + (half + 4) * 32, r);
        Epi::prefetch(ep.out, make_ctx(batch, m0, half + 6), rs1);
        tmem_ld_wait();
        Epi::finish(ep.out, make_ctx(batch, m0, half + 4), r, rs0);
        tmem_ld_32x32(t_row + (half + 6) * 32, r);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(acc2_empty, 0);   // accumulator fully in registers
        Epi::finish(ep.out, make_ctx(batch, m0, half + 6), r, rs1);
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
  }
}

// tmA: box rows = ResUnit256Cfg::halo_rows(s.dil); tmB7 / tmB1: 128-row boxes.
template <bool BF16>
int launch_resunit256(const CUtensorMap& tmA, const CUtensorMap& tmB7, const CUtensorMap& tmB1, const ResUnitShape& s,
                      const ResUnitParams<BF16>& ep, cudaStream_t stream) {
  using Cfg = ResUnit256Cfg;
  SATB_REQUIRE(s.dil >= 1 && s.dil <= Cfg::kMaxDil, "resunit256: dilation out of range");
  auto kern = resunit256_tcgen05_2cta_kernel<BF16>;
  static PerDeviceOnce attr;
  if (attr.first()) SATB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  const int total = ceil_div(s.L, 2 * kBlockM) * s.batches;
  if (total <= 0) return 0;
  int clusters = device_sm_count() / 2;
  if (clusters > total) clusters = total;
  SATB_CHECK_CUDA(launch_pdl(kern, dim3(2 * clusters), dim3(kGemmThreads), Cfg::kSmemBytes, stream, tmA, tmB7, tmB1, s, ep));
  count_launch();
  return 0;
}

}  // namespace satb
