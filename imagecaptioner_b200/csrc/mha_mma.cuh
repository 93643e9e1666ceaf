// Multi-head self-attention core of AttentionRefinement (reference src/student_model.py:72-118, nn.MultiheadAttention over the
// 49 feature tokens) on the tensor cores, bf16 mode, head_dim = 64, S <= 64.
//
// One CTA of 4 warps per (sample, head); every operand of a head is one 64 x 64 tile (49 real rows, zero padded), so the whole
// problem lives in 3 (forward) / 6-7 (backward) shared-memory tiles and each warp owns 16 rows of every product.
// Warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate) is the right-sized instruction here: a head is 5 products of
// 64 x 64 x 64, far below one tcgen05 128-row tile, the kernel is bound by staging Q/K/V and by the softmax, not by the MMA
// issue rate, and the accumulators are needed in registers for the softmax (no TMEM round trip).  The fp32-mode path and any
// other head size keep the FFMA register-tile kernels in refine_kernels.cuh.
#pragma once
#include "common.cuh"

namespace b2c {

constexpr int MM_THREADS = 128;        // 4 warps x 16 rows
constexpr int MM_R = 64;               // padded tile rows (tokens) and columns (head dim)
constexpr int MM_P = 72;               // row pitch in bf16: 144 bytes, so the 8 rows of an ldmatrix hit 8 different 16-byte bank groups
constexpr int MM_TILE = MM_R * MM_P;   // elements per tile

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const bf16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const bf16* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Fragment addresses inside a [row][MM_P] tile (lane = thread in warp, mi = lane / 8 = which 8x8 matrix this lane addresses).
//   A, stored [m][k]:            matrices (m0-7,k0-7) (m8-15,k0-7) (m0-7,k8-15) (m8-15,k8-15)
__device__ __forceinline__ const bf16* a_addr(const bf16* tile, int m0, int k0, int lane) { return tile + (m0 + (lane & 15)) * MM_P + k0 + (lane >> 4) * 8; }
//   A, stored [k][m] (the transposed operand, read with .trans): same matrix order
__device__ __forceinline__ const bf16* at_addr(const bf16* tile, int m0, int k0, int lane) {
  const int mi = lane >> 3;
  return tile + (k0 + (mi >> 1) * 8 + (lane & 7)) * MM_P + m0 + (mi & 1) * 8;
}
//   B for two adjacent n-tiles, stored [n][k]:   (n0-7,k0-7) (n0-7,k8-15) (n8-15,k0-7) (n8-15,k8-15)  -> b0,b1 | b0,b1
__device__ __forceinline__ const bf16* b_addr(const bf16* tile, int n0, int k0, int lane) {
  const int mi = lane >> 3;
  return tile + (n0 + (mi >> 1) * 8 + (lane & 7)) * MM_P + k0 + (mi & 1) * 8;
}
//   B for two adjacent n-tiles, stored [k][n] (read with .trans): same matrix order
__device__ __forceinline__ const bf16* bt_addr(const bf16* tile, int n0, int k0, int lane) {
  const int mi = lane >> 3;
  return tile + (k0 + (mi & 1) * 8 + (lane & 7)) * MM_P + n0 + (mi >> 1) * 8;
}

// acc[8][4] (16 rows x 64 cols) += A(16 x 64) B^T, A rows m0.. of `ta` ([m][k] or, AT, [k][m]), B = `tb` ([n][k] or, BT, [k][n])
template <bool AT, bool BT>
__device__ __forceinline__ void warp_gemm_64(float (&acc)[8][4], const bf16* ta, const bf16* tb, int m0, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t a[4];
    if (AT) ldsm_x4_t(a, at_addr(ta, m0, kk * 16, lane)); else ldsm_x4(a, a_addr(ta, m0, kk * 16, lane));
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      if (BT) ldsm_x4_t(b, bt_addr(tb, np * 16, kk * 16, lane)); else ldsm_x4(b, b_addr(tb, np * 16, kk * 16, lane));
      mma_bf16(acc[2 * np], a, b[0], b[1]);
      mma_bf16(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
}
__device__ __forceinline__ void zero_acc(float (&acc)[8][4]) {
#pragma unroll
  for (int n = 0; n < 8; ++n) { acc[n][0] = 0.f; acc[n][1] = 0.f; acc[n][2] = 0.f; acc[n][3] = 0.f; }
}
__device__ __forceinline__ float quad_sum(float v) { v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); return v; }
__device__ __forceinline__ float quad_max(float v) { v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1)); v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2)); return v; }

// rows i < S of NT (rows, ld) global matrices (64 columns each) -> NT zero-padded 64 x 64 tiles.  ALL 4 * NT 16-byte loads of a
// thread are issued before the first shared store (ncu on the first version: 45 % of the stall samples sat on the shared stores
// of a load -> store loop, i.e. 12 serialised L2 round trips per CTA).
template <int NT>
__device__ __forceinline__ void mm_load_tiles(bf16* const (&tile)[NT], const bf16* const (&src)[NT], const long (&ld)[NT], int S) {
  constexpr int PER = MM_R * 8 / MM_THREADS;             // 16-byte chunks per thread per tile (4)
  uint4 v[NT][PER];
#pragma unroll
  for (int n = 0; n < NT; ++n) {
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int idx = threadIdx.x + k * MM_THREADS, i = idx >> 3, c = idx & 7;
      v[n][k] = make_uint4(0u, 0u, 0u, 0u);
      if (i < S) v[n][k] = __ldg(reinterpret_cast<const uint4*>(src[n] + (long)i * ld[n] + c * 8));
    }
  }
#pragma unroll
  for (int n = 0; n < NT; ++n) {
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int idx = threadIdx.x + k * MM_THREADS, i = idx >> 3, c = idx & 7;
      *reinterpret_cast<uint4*>(tile[n] + i * MM_P + c * 8) = v[n][k];
    }
  }
}
// accumulator tile (16 rows m0.. x 64 cols) -> bf16 rows of a (rows, ld) global matrix at column offset col0, rows < S only
__device__ __forceinline__ void mm_store_rows(const float (&acc)[8][4], bf16* dst, long ld, int m0, int S, int lane, float mul) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int col = n * 8 + 2 * t;
    if (m0 + g < S) *reinterpret_cast<uint32_t*>(dst + (long)(m0 + g) * ld + col) = pack_bf16(acc[n][0] * mul, acc[n][1] * mul);
    if (m0 + g + 8 < S) *reinterpret_cast<uint32_t*>(dst + (long)(m0 + g + 8) * ld + col) = pack_bf16(acc[n][2] * mul, acc[n][3] * mul);
  }
}

// qkv (B*S, 3E) -> out (B*S, E);  probabilities BEFORE dropout are kept for the backward as (B, heads, S, 64) (row pitch 64 so
// both sides use 16-byte accesses; the FFMA kernels keep (B, heads, S, S)).  grid (B, heads), 128 threads.
__global__ void __launch_bounds__(MM_THREADS)
mha_fwd_mma_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, bf16* __restrict__ probs, int S, int E, int heads, float scale,
                   float drop_p, uint64_t seed, uint32_t drop_site, const unsigned long long* __restrict__ seed_dev) {
  __shared__ __align__(16) bf16 Qs[MM_TILE], Ks[MM_TILE], Vs[MM_TILE];
  if (drop_p > 0.f) seed = drop_seed(seed, seed_dev);
  const int b = blockIdx.x, h = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bf16* base = qkv + (long)b * S * 3 * E + h * MM_R;
  {
    bf16* const tiles[3] = {Qs, Ks, Vs};
    const bf16* const srcs[3] = {base, base + E, base + 2 * E};
    const long lds[3] = {3L * E, 3L * E, 3L * E};
    mm_load_tiles<3>(tiles, srcs, lds, S);
  }
  __syncthreads();
  const int m0 = warp * 16, g = lane >> 2, t = lane & 3;
  float sc[8][4];
  zero_acc(sc);
  warp_gemm_64<false, false>(sc, Qs, Ks, m0, lane);                   // scores = Q K^T
  // softmax over the S real columns; this thread holds rows r0 = m0+g (elements [n][0..1]) and r0+8 ([n][2..3]), cols n*8+2t+{0,1}
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const bool ok = n * 8 + 2 * t + j < S;
      sc[n][j] = ok ? sc[n][j] * scale : -INFINITY;
      sc[n][2 + j] = ok ? sc[n][2 + j] * scale : -INFINITY;
      mx0 = fmaxf(mx0, sc[n][j]); mx1 = fmaxf(mx1, sc[n][2 + j]);
    }
  }
  mx0 = quad_max(mx0); mx1 = quad_max(mx1);
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      sc[n][j] = Math<bf16>::exp_(sc[n][j] - mx0); s0 += sc[n][j];
      sc[n][2 + j] = Math<bf16>::exp_(sc[n][2 + j] - mx1); s1 += sc[n][2 + j];
    }
  }
  const float inv0 = 1.0f / quad_sum(s0), inv1 = 1.0f / quad_sum(s1);
  const float inv_keep = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  const int r0 = m0 + g, r1 = r0 + 8;
  const long pb = (long)(b * heads + h) * S * S;                      // dropout counter base: element (i, j) of the S x S matrix
  bf16* pout = probs + (long)(b * heads + h) * S * MM_R;              // saved with a row pitch of 64 (zeros beyond column S)
  uint32_t pa[4][4];                                                  // P (after dropout) as the A operand of P V, k = keys
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    float p[4];
    const int col0 = n * 8 + 2 * t;
    p[0] = sc[n][0] * inv0; p[1] = sc[n][1] * inv0; p[2] = sc[n][2] * inv1; p[3] = sc[n][3] * inv1;
    if (r0 < S) *reinterpret_cast<uint32_t*>(pout + (long)r0 * MM_R + col0) = pack_bf16(p[0], p[1]);
    if (r1 < S) *reinterpret_cast<uint32_t*>(pout + (long)r1 * MM_R + col0) = pack_bf16(p[2], p[3]);
    if (drop_p > 0.f) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (col0 + j < S) {
          p[j] *= dropout_scale(seed, drop_site, (uint64_t)(pb + (long)min(r0, S - 1) * S + col0 + j), drop_p, inv_keep);
          p[2 + j] *= dropout_scale(seed, drop_site, (uint64_t)(pb + (long)min(r1, S - 1) * S + col0 + j), drop_p, inv_keep);
        }
      }
    }
    pa[n >> 1][(n & 1) * 2] = pack_bf16(p[0], p[1]);                  // rows g:   a0 (k 0-7 of the pair) / a2 (k 8-15)
    pa[n >> 1][(n & 1) * 2 + 1] = pack_bf16(p[2], p[3]);              // rows g+8: a1 / a3
  }
  float o[8][4];
  zero_acc(o);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t bq[4];
      ldsm_x4_t(bq, bt_addr(Vs, np * 16, kk * 16, lane));
      mma_bf16(o[2 * np], pa[kk], bq[0], bq[1]);
      mma_bf16(o[2 * np + 1], pa[kk], bq[2], bq[3]);
    }
  }
  mm_store_rows(o, out + (long)b * S * E + h * MM_R, E, m0, S, lane, 1.0f);
}

// dqkv (B*S, 3E) from dout (B*S, E), the saved probabilities and qkv.  Shared memory: Q K V dO P dS (+ Pd with dropout) tiles.
//   dV = Pd^T dO;  dPd = dO V^T;  dP = dPd * mask;  dS = P (dP - rowsum(dP P)) scale;  dQ = dS K;  dK = dS^T Q
__global__ void __launch_bounds__(MM_THREADS)
mha_bwd_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ probs, const bf16* __restrict__ dout, bf16* __restrict__ dqkv,
                   int S, int E, int heads, float scale, float drop_p, uint64_t seed, uint32_t drop_site, const unsigned long long* __restrict__ seed_dev) {
  extern __shared__ __align__(16) unsigned char mm_smem[];
  if (drop_p > 0.f) seed = drop_seed(seed, seed_dev);
  bf16* Qs = reinterpret_cast<bf16*>(mm_smem);
  bf16* Ks = Qs + MM_TILE; bf16* Vs = Ks + MM_TILE; bf16* dOs = Vs + MM_TILE; bf16* Ps = dOs + MM_TILE; bf16* dSs = Ps + MM_TILE;
  bf16* Pds = drop_p > 0.f ? dSs + MM_TILE : Ps;
  const int b = blockIdx.x, h = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bf16* base = qkv + (long)b * S * 3 * E + h * MM_R;
  {
    bf16* const tiles[3] = {Qs, Ks, Vs};
    const bf16* const srcs[3] = {base, base + E, base + 2 * E};
    const long lds[3] = {3L * E, 3L * E, 3L * E};
    mm_load_tiles<3>(tiles, srcs, lds, S);
  }
  const long pb = (long)(b * heads + h) * S * S;
  const float inv_keep = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  {
    bf16* const tiles[2] = {dOs, Ps};
    const bf16* const srcs[2] = {dout + (long)b * S * E + h * MM_R, probs + (long)(b * heads + h) * S * MM_R};   // P: saved with a row pitch of 64
    const long lds[2] = {(long)E, (long)MM_R};
    mm_load_tiles<2>(tiles, srcs, lds, S);
  }
  if (drop_p > 0.f) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < MM_R * MM_R; idx += MM_THREADS) {
      const int i = idx >> 6, j = idx & 63;
      const float m = (i < S && j < S) ? dropout_scale(seed, drop_site, (uint64_t)(pb + (long)i * S + j), drop_p, inv_keep) : 0.f;
      Pds[i * MM_P + j] = __float2bfloat16(__bfloat162float(Ps[i * MM_P + j]) * m);
    }
  }
  __syncthreads();
  const int m0 = warp * 16, g = lane >> 2, t = lane & 3;
  bf16* dq_out = dqkv + (long)b * S * 3 * E + h * MM_R;
  {
    float dv[8][4];
    zero_acc(dv);
    warp_gemm_64<true, true>(dv, Pds, dOs, m0, lane);                 // dV rows (keys) m0.. = sum_i Pd[i, j] dO[i, :]
    mm_store_rows(dv, dq_out + 2 * E, 3 * E, m0, S, lane, 1.0f);
  }
  {
    float dp[8][4];
    zero_acc(dp);
    warp_gemm_64<false, false>(dp, dOs, Vs, m0, lane);                // dPd rows (queries) m0.. = dO V^T
    const int r0 = m0 + g, r1 = r0 + 8;
    float pr[8][4];
    float dot0 = 0.f, dot1 = 0.f;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int col = n * 8 + 2 * t;
      const uint32_t p01 = *reinterpret_cast<const uint32_t*>(Ps + r0 * MM_P + col), p23 = *reinterpret_cast<const uint32_t*>(Ps + r1 * MM_P + col);
      pr[n][0] = bf16_lo(p01); pr[n][1] = bf16_hi(p01); pr[n][2] = bf16_lo(p23); pr[n][3] = bf16_hi(p23);
      if (drop_p > 0.f) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const bool ok = col + j < S;
          dp[n][j] *= (ok && r0 < S) ? dropout_scale(seed, drop_site, (uint64_t)(pb + (long)r0 * S + col + j), drop_p, inv_keep) : 0.f;
          dp[n][2 + j] *= (ok && r1 < S) ? dropout_scale(seed, drop_site, (uint64_t)(pb + (long)r1 * S + col + j), drop_p, inv_keep) : 0.f;
        }
      }
      dot0 = fmaf(dp[n][0], pr[n][0], fmaf(dp[n][1], pr[n][1], dot0));
      dot1 = fmaf(dp[n][2], pr[n][2], fmaf(dp[n][3], pr[n][3], dot1));
    }
    dot0 = quad_sum(dot0); dot1 = quad_sum(dot1);
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const int col = n * 8 + 2 * t;
      *reinterpret_cast<uint32_t*>(dSs + r0 * MM_P + col) = pack_bf16(pr[n][0] * (dp[n][0] - dot0) * scale, pr[n][1] * (dp[n][1] - dot0) * scale);
      *reinterpret_cast<uint32_t*>(dSs + r1 * MM_P + col) = pack_bf16(pr[n][2] * (dp[n][2] - dot1) * scale, pr[n][3] * (dp[n][3] - dot1) * scale);
    }
  }
  __syncthreads();                                                    // dS complete (dK contracts over every query row)
  {
    float dq[8][4];
    zero_acc(dq);
    warp_gemm_64<false, true>(dq, dSs, Ks, m0, lane);                 // dQ rows m0.. = dS K
    mm_store_rows(dq, dq_out, 3 * E, m0, S, lane, 1.0f);
  }
  {
    float dk[8][4];
    zero_acc(dk);
    warp_gemm_64<true, true>(dk, dSs, Qs, m0, lane);                  // dK rows (keys) m0.. = dS^T Q
    mm_store_rows(dk, dq_out + E, 3 * E, m0, S, lane, 1.0f);
  }
}

}  // namespace b2c
