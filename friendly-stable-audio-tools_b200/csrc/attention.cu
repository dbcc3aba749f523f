// Tiled online-softmax attention forward, head dim 64, no mask, non-causal, optional GQA.
// Restates the arithmetic of the reference Attention core (models/transformer.py:496-536):
// softmax(q k^T / sqrt(64)) v with fp32 softmax statistics, 16-bit q/k/v/p operands.
//
// Round-1 implementation: warp-level mma.sync m16n8k16 tiles (64 queries x 64 keys per
// step, cp.async double-buffered K/V).  The attention core is ~8 % of the block FLOPs
// (6.46 + 0.82 of 94.64 GF); the tcgen05 version replaces this kernel next.
#include "common.cuh"
#include "kernels.h"

namespace satb {

namespace {

constexpr int kHd = 64;       // head dim
constexpr int kBq = 64;       // queries per CTA (4 warps x 16)
constexpr int kBk = 64;       // keys per step
constexpr int kAttnThreads = 128;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* smem) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(s));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* smem) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(s));
}
template <bool BF16>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (BF16) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}

// smem tile: rows of 64 x 16-bit = 128 B = 8 chunks of 16 B; chunk index XOR (row & 7)
__device__ __forceinline__ int swz(int row, int chunk) { return row * 64 + ((chunk ^ (row & 7)) << 3); }

struct AttnArgs {
  const uint16_t* q;
  const uint16_t* k;
  const uint16_t* v;
  uint16_t* o;
  int64_t ldq, ldk, ldv, ldo;         // row strides (elements)
  int64_t q_bs, k_bs, v_bs, o_bs;     // batch strides (elements)
  int Nq, Nk, H, group;               // group = H / H_kv
  float scale_log2;                   // softmax scale * log2(e)
};

template <bool BF16>
__global__ void __launch_bounds__(kAttnThreads) attn_fwd_kernel(const AttnArgs p) {
  __shared__ __align__(128) uint16_t sQ[kBq * kHd];
  __shared__ __align__(128) uint16_t sK[2][kBk * kHd];
  __shared__ __align__(128) uint16_t sV[2][kBk * kHd];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * kBq;
  const int h = blockIdx.y, b = blockIdx.z;
  const int hk = h / p.group;
  const uint16_t* qg = p.q + b * p.q_bs + static_cast<int64_t>(h) * kHd;
  const uint16_t* kg = p.k + b * p.k_bs + static_cast<int64_t>(hk) * kHd;
  const uint16_t* vg = p.v + b * p.v_bs + static_cast<int64_t>(hk) * kHd;

  // ---- async loads: Q tile, then K/V tile 0
  for (int c = tid; c < kBq * 8; c += kAttnThreads) {
    const int row = c >> 3, ch = c & 7;
    const bool ok = (q0 + row) < p.Nq;
    cp_async16(&sQ[swz(row, ch)], qg + static_cast<int64_t>(ok ? q0 + row : 0) * p.ldq + ch * 8, ok);
  }
  auto load_kv = [&](int buf, int k0) {
    for (int c = tid; c < kBk * 8; c += kAttnThreads) {
      const int row = c >> 3, ch = c & 7;
      const bool ok = (k0 + row) < p.Nk;
      const int64_t r = ok ? k0 + row : 0;
      cp_async16(&sK[buf][swz(row, ch)], kg + r * p.ldk + ch * 8, ok);
      cp_async16(&sV[buf][swz(row, ch)], vg + r * p.ldv + ch * 8, ok);
    }
  };
  load_kv(0, 0);
  cp_async_commit();

  const int n_tiles = (p.Nk + kBk - 1) / kBk;
  float o_acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o_acc[i][j] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  uint32_t qf[4][4];

  for (int t = 0; t < n_tiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < n_tiles) {
      load_kv(buf ^ 1, (t + 1) * kBk);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (t == 0) {
      // Q fragments (A operand, 16 rows x 16 k per ldmatrix.x4), kept in registers
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int row = warp * 16 + (lane & 15);
        const int ch = ks * 2 + (lane >> 4);
        ldsm_x4(qf[ks], &sQ[swz(row, ch)]);
      }
    }
    // ---- S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of 8-key tiles
        uint32_t kf[4];
        const int row = np * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int ch = ks * 2 + ((lane >> 3) & 1);
        ldsm_x4(kf, &sK[buf][swz(row, ch)]);
        mma16816<BF16>(s[np * 2], qf[ks], kf[0], kf[1]);
        mma16816<BF16>(s[np * 2 + 1], qf[ks], kf[2], kf[3]);
      }
    }
    // ---- mask keys beyond Nk, scale, online softmax (fp32)
    const int kbase = t * kBk + (lane & 3) * 2;
    float m_new[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = kbase + i * 8 + (j & 1);
        const float val = key < p.Nk ? s[i][j] * p.scale_log2 : -INFINITY;
        s[i][j] = val;
        m_new[j >> 1] = fmaxf(m_new[j >> 1], val);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      m_new[r] = fmaxf(m_new[r], __shfl_xor_sync(0xffffffffu, m_new[r], 1));
      m_new[r] = fmaxf(m_new[r], __shfl_xor_sync(0xffffffffu, m_new[r], 2));
    }
    float corr[2], row_sum[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) corr[r] = exp2f(m_run[r] - m_new[r]);
    uint32_t pf[4][4];  // P as A fragments for the PV product (4 k-steps of 16 keys)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float p0 = exp2f(s[i][0] - m_new[0]), p1 = exp2f(s[i][1] - m_new[0]);
      const float p2 = exp2f(s[i][2] - m_new[1]), p3 = exp2f(s[i][3] - m_new[1]);
      row_sum[0] += p0 + p1;
      row_sum[1] += p2 + p3;
      // C-fragment (i = 8-key tile) -> A-fragment: k-step i/2, regs {0,1} for even i, {2,3} for odd i
      pf[i >> 1][(i & 1) * 2 + 0] = Op16<BF16>::pack(p0, p1);
      pf[i >> 1][(i & 1) * 2 + 1] = Op16<BF16>::pack(p2, p3);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      l_run[r] = l_run[r] * corr[r] + row_sum[r];
      m_run[r] = m_new[r];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o_acc[i][0] *= corr[0];
      o_acc[i][1] *= corr[0];
      o_acc[i][2] *= corr[1];
      o_acc[i][3] *= corr[1];
    }
    // ---- O += P V   (V^T fragments through ldmatrix.trans)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {      // 16 keys
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {    // pairs of 8-wide d tiles
        uint32_t vf[4];
        const int row = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        const int ch = dp * 2 + (lane >> 4);
        ldsm_x4_trans(vf, &sV[buf][swz(row, ch)]);
        mma16816<BF16>(o_acc[dp * 2], pf[ks], vf[0], vf[1]);
        mma16816<BF16>(o_acc[dp * 2 + 1], pf[ks], vf[2], vf[3]);
      }
    }
    __syncthreads();
  }
  // ---- finalize: row sums across the quad, normalise, store
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  const int row0 = q0 + warp * 16 + (lane >> 2), row1 = row0 + 8;
  uint16_t* og = p.o + b * p.o_bs + static_cast<int64_t>(h) * kHd + (lane & 3) * 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (row0 < p.Nq)
      *reinterpret_cast<uint32_t*>(og + static_cast<int64_t>(row0) * p.ldo + i * 8) =
          Op16<BF16>::pack(o_acc[i][0] * inv0, o_acc[i][1] * inv0);
    if (row1 < p.Nq)
      *reinterpret_cast<uint32_t*>(og + static_cast<int64_t>(row1) * p.ldo + i * 8) =
          Op16<BF16>::pack(o_acc[i][2] * inv1, o_acc[i][3] * inv1);
  }
}

}  // namespace

int launch_attention(const void* q, const void* k, const void* v, void* o, int64_t ldq, int64_t ldk, int64_t ldv,
                     int64_t ldo, int64_t q_bs, int64_t k_bs, int64_t v_bs, int64_t o_bs, int batch, int H, int H_kv,
                     int Nq, int Nk, int head_dim, bool bf16, cudaStream_t stream) {
  SATB_REQUIRE(head_dim == kHd, "attention kernel supports head_dim 64 only");
  SATB_REQUIRE(H % H_kv == 0, "num_heads must be a multiple of kv heads");
  SATB_REQUIRE(Nk >= 1 && Nq >= 1, "empty attention problem");
  SATB_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0, "attention strides must be 16B aligned");
  AttnArgs a;
  a.q = static_cast<const uint16_t*>(q);
  a.k = static_cast<const uint16_t*>(k);
  a.v = static_cast<const uint16_t*>(v);
  a.o = static_cast<uint16_t*>(o);
  a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo;
  a.q_bs = q_bs; a.k_bs = k_bs; a.v_bs = v_bs; a.o_bs = o_bs;
  a.Nq = Nq; a.Nk = Nk; a.H = H; a.group = H / H_kv;
  a.scale_log2 = (1.0f / sqrtf(static_cast<float>(head_dim))) * 1.4426950408889634f;
  dim3 grid(ceil_div(Nq, kBq), H, batch);
  if (bf16)
    attn_fwd_kernel<true><<<grid, kAttnThreads, 0, stream>>>(a);
  else
    attn_fwd_kernel<false><<<grid, kAttnThreads, 0, stream>>>(a);
  count_launch();
  SATB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace satb
