// k-tap (dilated) convolution with few output channels as a tcgen05 implicit GEMM that loads every
// activation row ONCE per tile: the A operand of tap t is the same 128B-swizzled shared-memory tile,
// read through a matrix descriptor whose start address is shifted by t * dilation rows (128 B each).
// The swizzle is a function of the absolute smem address bits, which TMA (writer) and the tensor
// core (reader) agree on, so a row-shifted start needs no re-layout.  The generic GEMM path reloads
// the tile once per tap (7x the L2 -> SM traffic), which is what bounds the thin layers.
//
//   out[b, n, l] = bias[n] + sum_{t, k} A[b, l + (t - taps/2) * dil, k] * W[t * cout + n, k]     (N <= 32)
//
// Used for the decoder's final convolution (128 -> 2 channels, k = 7; models/autoencoders.py:190).
// One persistent CTA per SM: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4-7 epilogue.  All tap weights stay resident in shared memory.
#pragma once
#include "gemm.cuh"

namespace satb {

struct ConvHaloShape {
  int L;         // positions per batch item
  int batches;
  int K;         // input channels (multiple of 64, <= 256)
  int n_taps;    // odd
  int dil;
  int cout;      // real output channels (<= 32); B tile rows beyond them are ignored
};

struct ConvHaloCfg {
  static constexpr int kBN = 32;
  static constexpr int kThreads = 256;
  static constexpr int kBTile = kBN * kBlockK * 2;            // 4 KB per (tap, k-block)
  static constexpr int kSmemBudget = 200 * 1024;
  __host__ __device__ static int halo_rows(const ConvHaloShape& s) { return kBlockM + (s.n_taps - 1) * s.dil; }
  __host__ __device__ static int slot_bytes(const ConvHaloShape& s) { return (halo_rows(s) * 128 + 1023) & ~1023; }
  __host__ __device__ static int b_bytes(const ConvHaloShape& s) { return s.n_taps * (s.K / kBlockK) * kBTile; }
  __host__ __device__ static int stages(const ConvHaloShape& s) {
    int n = (kSmemBudget - b_bytes(s) - 1024) / slot_bytes(s);
    return n > 8 ? 8 : n;
  }
};

// Measured on B200: the plain SW128 descriptor (matrix base offset field = 0) is what works for a
// start address on any 128-byte row of the pattern; setting base offset = (addr >> 7) & 7 gives wrong results.

template <class Epi, bool BF16>
__global__ void __launch_bounds__(ConvHaloCfg::kThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvHaloShape s,
                 const typename Epi::Params ep) {
  using Cfg = ConvHaloCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int n_stages = Cfg::stages(s);
  const int slot = Cfg::slot_bytes(s);
  const int kbs = s.K / kBlockK;
  uint8_t* b_res = smem + n_stages * slot;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_res + Cfg::b_bytes(s));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + 8;
  uint64_t* tfull_bar = bars + 16;
  uint64_t* tempty_bar = bars + 18;
  uint64_t* b_full = bars + 20;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (s.L + kBlockM - 1) / kBlockM;
  const int total_tiles = m_tiles * s.batches;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < 8; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    mbar_init(b_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 2 * Cfg::kBN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  if (warp != 0) pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      // ---------------------------------------------------------------- TMA producer
      // the tap weights are static: their loads start before the programmatic-dependency wait
      mbar_expect_tx(b_full, Cfg::b_bytes(s));
      for (int t = 0; t < s.n_taps; ++t)
        for (int kb = 0; kb < kbs; ++kb)
          tma_load_2d(b_res + (t * kbs + kb) * Cfg::kBTile, &tmB, b_full, kb * kBlockK, t * s.cout);
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      const int tx = Cfg::halo_rows(s) * 128;
      const int lead = (s.n_taps / 2) * s.dil;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int batch = tile / m_tiles;
        const int m0 = (tile - batch * m_tiles) * kBlockM;
        for (int kb = 0; kb < kbs; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], tx);
          tma_load_4d(smem + stage * slot, &tmA, &full_bar[stage], kb * kBlockK, 0, m0 - lead, batch);
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ---------------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc = make_idesc_f16(kBlockM, Cfg::kBN, BF16);
      const uint32_t b_addr0 = smem_u32(b_res);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      mbar_wait(b_full, 0);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::kBN;
        for (int kb = 0; kb < kbs; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * slot);
          for (int t = 0; t < s.n_taps; ++t) {
            const uint32_t a_tap = a_addr + t * s.dil * 128;
            const uint32_t b_tap = b_addr0 + (t * kbs + kb) * Cfg::kBTile;
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              const uint64_t da = make_desc_kmajor_sw128(a_tap + k * kUmmaK * 2);
              const uint64_t db = make_desc_kmajor_sw128(b_tap + k * kUmmaK * 2);
              umma_f16_ss(d_tmem, da, db, idesc, (kb | t | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int batch = tile / m_tiles;
      const int m0 = (tile - batch * m_tiles) * kBlockM;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      EpiCtx c;
      c.l = m0 + q * 32 + lane;
      c.batch = batch;
      c.row = batch * s.L + c.l;
      c.valid = c.l < s.L;
      c.l0 = m0 + q * 32;
      c.L = s.L;
      c.lane = lane;
      c.stage = nullptr;
      c.col0 = 0;
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * Cfg::kBN, r);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      Epi::apply(ep, c, r);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * Cfg::kBN);
  }
}

template <class Epi, bool BF16>
int launch_conv_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvHaloShape& s, const typename Epi::Params& ep,
                     cudaStream_t stream) {
  using Cfg = ConvHaloCfg;
  static_assert(Epi::kCols == 32 && Epi::kStageBytes == 0, "conv_halo epilogue: one 32-column chunk, no staging");
  auto kern = conv_halo_kernel<Epi, BF16>;
  const int smem = Cfg::stages(s) * Cfg::slot_bytes(s) + Cfg::b_bytes(s) + 256 + 1024;
  SATB_REQUIRE(s.K % kBlockK == 0 && s.cout <= Cfg::kBN && Cfg::halo_rows(s) <= 256 && Cfg::stages(s) >= 2,
               "conv_halo: unsupported shape");
  SATB_REQUIRE(smem <= Cfg::kSmemBudget + 4096, "conv_halo: shared-memory request too large");
  static PerDeviceOnce attr;   // once per device, to the largest size any shape may ask for
  if (attr.first()) SATB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBudget + 4096));
  const int total = ceil_div(s.L, kBlockM) * s.batches;
  if (total <= 0) return 0;
  int grid = device_sm_count();
  if (grid > total) grid = total;
  SATB_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(Cfg::kThreads), smem, stream, tmA, tmB, s, ep));
  count_launch();
  return 0;
}

}  // namespace satb
