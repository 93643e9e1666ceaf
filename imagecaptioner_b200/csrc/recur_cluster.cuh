// The T-step forward recurrence of LSTMDecoder.forward (reference src/student_model.py:232-251) as ONE kernel of independent
// thread-block CLUSTERS, one cluster per slice of batch rows.
//
// Why clusters.  The multi-kernel path runs 2 + L dependent launches per time step (u GEMM -> attention -> L fused gate GEMMs), each
// ~1 us of arithmetic inside 6-9 us of launch / prologue / drain latency (31 us per step).  The first persistent attempt
// (recurrent.cuh) replaced the kernel boundaries by GRID barriers and was slower (42 us per step): every phase paid a device-wide
// barrier plus cold operand loads.  But samples are independent: nothing ever has to cross a row slice.  So the batch is cut into
// `ncl` slices of <= 40 rows, each owned for all T steps by one cluster of 8 CTAs that never talks to another cluster, and the
// four exchanges of a step happen inside the cluster through TMA MULTICAST and st.async into the peers' shared memory, signalled
// by mbarrier transaction counts -- no grid barrier, no cluster barrier, no flag polling in global memory.
// (tools/probe_cluster.cu measured the mechanisms on B200: 15 clusters of 8 CTAs with ~220 KB of shared memory are co-resident
// (7 of 16); a store -> proxy fence -> multicast -> all-slices-landed exchange costs ~2000 cycles.)
//
// Work split inside a cluster (H = 512, E = 256, L = 2; CTA rank c = 0..7):
//   * gate contractions, swap-AB:  D^T[gate row, sample] = W[gate row, :] . x[sample, :]^T.  CTA c owns hidden units [64c, 64c+64) of
//     both layers = rows [256c, 256c+256) of the gate-interleaved packed weights W_cat (row 4j+g), two M = 128 tiles per layer;
//     N = 48 sample columns (40 row slots).  The weights stream through a 6-stage TMA ring (they are read once per step and
//     cluster: 7.6 MB x ncl per step from L2); the B operands -- h0, h1 and ctx of the cluster's rows, [40 x K] bf16 -- stay RESIDENT
//     in every CTA's shared memory as SWIZZLE_128B k-blocks of 64, block c being produced by CTA c.
//   * the recurrent halves (W_hh . h_{t-1}) are issued right after h_{t-1} lands and run under the attention phase; only the input
//     halves (W_x . ctx_t, K = 256; W_ih1 . h0_t, K = 512) sit on the critical path.
//   * u_t = h1_{t-1} W_h^T: CTA c computes columns [32c, 32c+32) for all rows of the cluster (M = 128 tile = the resident h tiles
//     used as the A operand, N = 32) and scatters row r to the CTA that owns sample r (st.async + complete_tx).
//   * attention: CTA c owns sample slots [5c, 5c+5): scores from e^{2P} (fp32, written once in the prologue) and e^{2u}
//     (tanh(P+u) = 1 - 2 / (1 + e^{2P} e^{2u}): one MUFU per element), softmax, context; ctx rows go to global (they are a forward
//     save anyway) and are multicast from there into every CTA's ctx tile as 5-row boxes.
//   * LSTM cell in the epilogue: thread = one gate row of one tile (TMEM lane), 48 sample columns in registers; the four gates of a
//     unit sit in four adjacent lanes, so a 4x4 quad transpose (4 shuffles per 4 samples) gives every lane (i,f,g,o) of one
//     (unit, sample); c lives in registers for the whole sequence.  h goes to global (forward saves: recurrent slot, next layer's
//     input, top output) and CTA c multicasts its own 64-unit k-block of the h tile to all 8 CTAs.
//
// Warp roles (384 threads): warp 0 = weight-ring TMA producer, warp 1 = tcgen05.mma issuer, warps 4-11 = 256 workers (u epilogue ->
// attention -> layer-0 cell -> layer-1 cell, TMEM lane quarter = warp % 4, M tile = (warp - 4) / 4).
//
// WAR safety of the resident tiles without any barrier (in-order tensor pipe + data dependencies):
//   * a CTA multicasts ctx_t only after its own early MMAs of step t have committed (bar_early).  Whoever holds all of ctx_t therefore
//     knows every peer is done reading h0_{t-1}, and may overwrite the h0 blocks with h0_t.
//   * h1_t is sent after the sender's layer-1 cell, which needed h0_t from every peer, which each peer sent after its layer-0 cell,
//     i.e. after its W_x MMAs -- issued behind the u and early MMAs that read h1_{t-1}.
//   * ctx_{t+1} / u_{t+1} are sent after the sender has h1_t from every peer, i.e. after every peer's step-t MMAs and attention.
// Every wait is bounded (mbar_wait traps on timeout) so a protocol error ends in a launch failure, not a hung GPU.
#pragma once
#include "gemm.cuh"

namespace b2c {

constexpr int CR_CL = 8;                       // CTAs per cluster (portable maximum)
constexpr int CR_H = 512, CR_E = 256, CR_L = 2;
constexpr int CR_RMAX = 40;                    // row slots per cluster: 5 swizzle atoms of 8 rows
constexpr int CR_N = 48;                       // MMA N: multiple of 16 >= CR_RMAX (columns 40..47 are never read back)
constexpr int CR_SPC = CR_RMAX / CR_CL;        // 5 sample slots per CTA
constexpr int CR_UNITS = CR_H / CR_CL;         // 64 hidden units per CTA and layer
constexpr int CR_UCOLS = CR_E / CR_CL;         // 32 columns of u per CTA
constexpr int CR_STAGES = 6;
constexpr int CR_STAGE_BYTES = 16384;          // one A tile: 128 gate rows x 64 k, SWIZZLE_128B
constexpr int CR_THREADS = 384, CR_WORKERS = 256;
constexpr int CR_HALF = CR_RMAX * 128;         // 5120: one [40 rows x 64 k] tile
constexpr int CR_BLOCK = 2 * CR_HALF;          // k-block c of the h tiles: [h1 half | h0 half]
constexpr int CR_SMAX = 64;                    // tokens (S <= 64)
// shared memory map (bytes from the 1024-aligned base)
constexpr int CR_OFF_BH = 0;                                        // 8 blocks x 10240
constexpr int CR_OFF_CTX = CR_OFF_BH + CR_CL * CR_BLOCK;            // 4 k-blocks x 5120
constexpr int CR_OFF_RING = CR_OFF_CTX + (CR_E / 64) * CR_HALF;     // 102400 (1024-aligned)
constexpr int CR_OFF_HST = CR_OFF_RING + CR_STAGES * CR_STAGE_BYTES;  // h staging [40][64] bf16
constexpr int CR_OFF_UIN = CR_OFF_HST + CR_RMAX * 128;              // u inbox [5][256] fp32
constexpr int CR_OFF_E2U = CR_OFF_UIN + CR_SPC * CR_E * 4;          // e^{2u} [5][256] fp32
constexpr int CR_OFF_CACC = CR_OFF_E2U + CR_SPC * CR_E * 4;         // ctx accumulators [5][8][32] fp32
constexpr int CR_OFF_SC = CR_OFF_CACC + CR_SPC * CR_E * 4;          // scores / weights [5][64] fp32
constexpr int CR_OFF_BAR = CR_OFF_SC + CR_SPC * CR_SMAX * 4;        // mbarriers
constexpr int CR_SMEM_BYTES = CR_OFF_BAR + 256 + 1024;              // + alignment slack
static_assert(CR_OFF_RING % 1024 == 0 && CR_OFF_CTX % 1024 == 0, "swizzled tiles need 1024-byte alignment");
static_assert(CR_SMEM_BYTES <= 227 * 1024, "shared memory budget");
constexpr uint32_t CR_TM_ACC0 = 0, CR_TM_ACC1 = 96, CR_TM_U = 192, CR_TM_COLS = 256;     // TMEM columns: acc[layer][tile] 48 each, u 32

struct ClusterParams {
  int B, T, S, ncl;
  const float* P;                 // (B, S, E) fp32
  float* EP;                      // (B, S, E) fp32 scratch: e^{2P}
  const bf16* F;                  // (B, S, E)
  float* u;                       // (T*B, E) fp32 (forward save)
  float* attw;                    // (T, B, S) fp32
  bf16* xh0; bf16* xh1;           // (T+1, B, E+H) [ctx_t ; h0_{t-1}],  (T+1, B, 2H) [h0_t ; h1_{t-1}]
  const bf16* G0T;                // (T, ncl, 4H, 40) bf16: embedding half of layer 0's pre-activations + b_x, per cluster, sample-contiguous
  const float* bias1;             // (4H) interleaved
  float* c0; float* c1;           // (T+1, B, H)
  bf16* gates0; bf16* gates1;     // (T*B, 4H) interleaved
  bf16* hid_top;                  // (T, B, H)
  unsigned long long* trace;      // optional: per CTA and step 8 clock64() stamps of worker thread 0
};
struct ClusterMaps { CUtensorMap w0, w1, wh, h0, h1, ctx; };

__device__ __forceinline__ uint32_t cr_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cr_cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cr_mapa(uint32_t addr, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r; }
__device__ __forceinline__ void cr_tma_mc(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
               :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void cr_st_async_v4(uint32_t remote_addr, float a, float b, float c, float d, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
               :: "r"(remote_addr), "f"(a), "f"(b), "f"(c), "f"(d), "r"(remote_bar) : "memory");
}
// bounded wait that says what it was waiting for: every stuck waiter reports once (id = which barrier, t = time step), the trap comes later
__device__ __noinline__ void cr_wait_slow(uint64_t* bar, uint32_t parity, int id, int t) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
    if (spins == (threadIdx.x < 128 ? (1u << 22) : (1u << 18)) && (threadIdx.x < 128 || (threadIdx.x & 31) == 0)) printf("b2c cluster kernel: wait %d stuck at step %d (block %d = cluster %d rank %d, thread %d)\n", id, t, blockIdx.x, blockIdx.x / CR_CL, blockIdx.x % CR_CL, threadIdx.x);
    if (spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void cr_wait(uint64_t* bar, uint32_t parity, int id, int t) {
  if (!mbar_try_wait(bar, parity)) cr_wait_slow(bar, parity, id, t);
}
__device__ __forceinline__ void cr_fence_proxy_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void cr_bar_workers() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void cr_tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// exchange within a quad of lanes: every lane holds a[0..3] = its own gate's value for samples 4s..4s+3 and ends with
// a[g] = gate g's value for sample 4s + (lane & 3)
__device__ __forceinline__ void cr_quad_transpose(float (&a)[4], int q) {
  const bool b0 = q & 1, b1 = q & 2;
  { const float s0 = b0 ? a[0] : a[1], s1 = b0 ? a[2] : a[3];
    const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
    if (b0) { a[0] = r0; a[2] = r1; } else { a[1] = r0; a[3] = r1; } }
  { const float s0 = b1 ? a[0] : a[2], s1 = b1 ? a[1] : a[3];
    const float r0 = __shfl_xor_sync(0xffffffffu, s0, 2), r1 = __shfl_xor_sync(0xffffffffu, s1, 2);
    if (b1) { a[0] = r0; a[1] = r1; } else { a[2] = r0; a[3] = r1; } }
}

// rows of cluster i: n_i = B / ncl (+1 for the first B % ncl clusters), starting at r0_i
__host__ __device__ inline void cr_rows(int B, int ncl, int i, int& r0, int& n) {
  const int q = B / ncl, rem = B % ncl;
  n = q + (i < rem ? 1 : 0);
  r0 = i * q + (i < rem ? i : rem);
}

// (T*B, 4H) row-major addend -> (T, ncl, 4H, 40): sample-contiguous per gate row and cluster (what a TMEM-lane thread reads)
__global__ void __launch_bounds__(256) g0_cluster_layout_kernel(const bf16* __restrict__ G0, int B, int T, int ncl, bf16* __restrict__ out) {
  __shared__ bf16 tile[CR_RMAX][64 + 2];
  const int t = blockIdx.z, i = blockIdx.y, g0 = blockIdx.x * 64;       // 64 gate rows per block
  int r0, n; cr_rows(B, ncl, i, r0, n);
  for (int idx = threadIdx.x; idx < CR_RMAX * 64; idx += 256) {
    const int s = idx >> 6, g = idx & 63;
    tile[s][g] = s < n ? G0[((long)t * B + r0 + s) * (4 * CR_H) + g0 + g] : __float2bfloat16(0.f);
  }
  __syncthreads();
  bf16* o = out + (((long)t * ncl + i) * (4 * CR_H) + g0) * CR_RMAX;
  for (int idx = threadIdx.x; idx < 64 * CR_RMAX; idx += 256) {
    const int g = idx / CR_RMAX, s = idx % CR_RMAX;
    o[idx] = tile[s][g];
  }
}

__global__ void __launch_bounds__(CR_THREADS, 1)
recur_cluster_fwd_kernel(const __grid_constant__ ClusterMaps maps, const __grid_constant__ ClusterParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + CR_OFF_BAR);
  uint64_t* full_bar = bars;                        // [CR_STAGES]
  uint64_t* empty_bar = bars + CR_STAGES;           // [CR_STAGES]
  uint64_t* bar_h1 = bars + 2 * CR_STAGES + 0;      // h1 tile landed (tx)
  uint64_t* bar_h0 = bars + 2 * CR_STAGES + 1;      // h0 tile landed (tx)
  uint64_t* bar_ctx = bars + 2 * CR_STAGES + 2;     // ctx tile landed (tx)
  uint64_t* bar_uin = bars + 2 * CR_STAGES + 3;     // u rows of this CTA's samples landed (tx)
  uint64_t* bar_u = bars + 2 * CR_STAGES + 4;       // u MMAs committed
  uint64_t* bar_early = bars + 2 * CR_STAGES + 5;   // recurrent-half MMAs committed
  uint64_t* bar_acc0 = bars + 2 * CR_STAGES + 6;    // layer-0 accumulators complete
  uint64_t* bar_acc1 = bars + 2 * CR_STAGES + 7;    // layer-1 accumulators complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * CR_STAGES + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = (int)cr_cluster_rank();
  const int ci = blockIdx.x / CR_CL;                // cluster index = row slice
  const int B = p.B, T = p.T, S = p.S;
  int r0, n; cr_rows(B, p.ncl, ci, r0, n);
  const int nv = max(0, min(CR_SPC, n - CR_SPC * c));         // valid sample slots of this CTA
  constexpr uint32_t H_TILE_BYTES = CR_CL * CR_HALF;          // 40960: 8 multicasts of one [40 x 64] block
  constexpr uint32_t CTX_TILE_BYTES = CR_CL * (CR_E / 64) * CR_SPC * 128;      // 20480
  const uint32_t uin_bytes = (uint32_t)nv * CR_E * 4;

  if (tid == 0) {
    tma_prefetch_desc(&maps.w0); tma_prefetch_desc(&maps.w1); tma_prefetch_desc(&maps.wh);
    tma_prefetch_desc(&maps.h0); tma_prefetch_desc(&maps.h1); tma_prefetch_desc(&maps.ctx);
    for (int i = 0; i < CR_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 8; ++i) mbar_init(&bars[2 * CR_STAGES + i], 1);
    fence_barrier_init();
    // first phases of the transaction barriers
    mbar_arrive_expect_tx(bar_h1, H_TILE_BYTES);
    mbar_arrive_expect_tx(bar_h0, H_TILE_BYTES);
    mbar_arrive_expect_tx(bar_ctx, CTX_TILE_BYTES);
    mbar_arrive_expect_tx(bar_uin, uin_bytes);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(CR_TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cr_cluster_sync();                               // every peer's barriers are initialised before anything can signal them
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------ weight-ring producer (independent of the data flow)
    if (lane == 0) {
      // initial state h_{-1} = 0: this CTA's k-block of both h tiles, from slot 0 of the operand buffers (zeroed by the prepare step)
      cr_tma_mc(base + CR_OFF_BH + c * CR_BLOCK, &maps.h1, CR_H + 64 * c, r0, bar_h1, 0xFF);
      cr_tma_mc(base + CR_OFF_BH + c * CR_BLOCK + CR_HALF, &maps.h0, CR_E + 64 * c, r0, bar_h0, 0xFF);
      uint32_t it = 0;
      auto acquire = [&]() -> unsigned char* {
        const int s = it % CR_STAGES; const uint32_t ph = (it / CR_STAGES) & 1;
        cr_wait(&empty_bar[s], ph ^ 1, 100 + s, (int)(it / 58));
        mbar_arrive_expect_tx(&full_bar[s], CR_STAGE_BYTES);
        return base + CR_OFF_RING + s * CR_STAGE_BYTES;
      };
      const int row0 = 256 * c;
      for (int t = 0; t < T; ++t) {
        for (int s2 = 0; s2 < 2; ++s2) {                                   // W_h rows [32c, 32c+32): 2 stages x 4 k-blocks of [32 x 64]
          unsigned char* sa = acquire(); uint64_t* fb = &full_bar[it % CR_STAGES];
          for (int i = 0; i < 4; ++i) tma_load_2d(sa + i * 4096, &maps.wh, 64 * (4 * s2 + i), CR_UCOLS * c, fb);
          ++it;
        }
        for (int kb = 0; kb < CR_H / 64; ++kb) for (int g = 0; g < 2; ++g) {     // W_hh1
          unsigned char* sa = acquire(); tma_load_2d(sa, &maps.w1, CR_H + 64 * kb, row0 + 128 * g, &full_bar[it % CR_STAGES]); ++it; }
        for (int kb = 0; kb < CR_H / 64; ++kb) for (int g = 0; g < 2; ++g) {     // W_hh0
          unsigned char* sa = acquire(); tma_load_2d(sa, &maps.w0, CR_E + 64 * kb, row0 + 128 * g, &full_bar[it % CR_STAGES]); ++it; }
        for (int kb = 0; kb < CR_E / 64; ++kb) for (int g = 0; g < 2; ++g) {     // W_x (attention_combine folded in)
          unsigned char* sa = acquire(); tma_load_2d(sa, &maps.w0, 64 * kb, row0 + 128 * g, &full_bar[it % CR_STAGES]); ++it; }
        for (int kb = 0; kb < CR_H / 64; ++kb) for (int g = 0; g < 2; ++g) {     // W_ih1
          unsigned char* sa = acquire(); tma_load_2d(sa, &maps.w1, 64 * kb, row0 + 128 * g, &full_bar[it % CR_STAGES]); ++it; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t IDESC_G = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(CR_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint32_t IDESC_U = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(CR_UCOLS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t bh = smem_u32(base + CR_OFF_BH), bctx = smem_u32(base + CR_OFF_CTX), ring = smem_u32(base + CR_OFF_RING);
      uint32_t it = 0;
      auto stage = [&]() -> uint32_t {
        const int s = it % CR_STAGES; const uint32_t ph = (it / CR_STAGES) & 1;
        cr_wait(&full_bar[s], ph, 200 + s, (int)(it / 58));
        tc_fence_after();
        return ring + s * CR_STAGE_BYTES;
      };
      auto release = [&]() { tc_commit(&empty_bar[it % CR_STAGES]); ++it; };
      // gate tiles: A = weight stage (128 rows), B = resident [40(48) x 64] k-block at `bop`
      auto gate_stage = [&](uint32_t tacc, uint32_t bop, bool first) {
        const uint32_t sa = stage();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_bf16(tacc, make_smem_desc(sa + k * 32, 16, 1024), make_smem_desc(bop + k * 32, 16, 1024), IDESC_G, (first && k == 0) ? 0u : 1u);
        release();
      };
      for (int t = 0; t < T; ++t) {
        // ---- h1_{t-1} has landed
        cr_wait(bar_h1, t & 1, 1, t);
        if (t + 1 < T) mbar_arrive_expect_tx(bar_h1, H_TILE_BYTES);
        tc_fence_after();
        // u_t columns [32c, 32c+32): A = [h1 ; h0 ; ...] rows of k-block kb (M = 128, rows 0..39 are h1), B = W_h stage
        for (int s2 = 0; s2 < 2; ++s2) {
          const uint32_t sa = stage();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int kb = 4 * s2 + i;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_bf16(tmem_base + CR_TM_U, make_smem_desc(bh + kb * CR_BLOCK + k * 32, 16, 1024), make_smem_desc(sa + i * 4096 + k * 32, 16, 1024),
                          IDESC_U, (kb == 0 && k == 0) ? 0u : 1u);
          }
          release();
        }
        tc_commit(bar_u);
        // ---- recurrent halves, off the critical path
        for (int kb = 0; kb < CR_H / 64; ++kb) for (int g = 0; g < 2; ++g)
          gate_stage(tmem_base + CR_TM_ACC1 + g * CR_N, bh + kb * CR_BLOCK, kb == 0);
        if (t == 0) { cr_wait(bar_h0, 0, 2, t); mbar_arrive_expect_tx(bar_h0, H_TILE_BYTES); tc_fence_after(); }
        for (int kb = 0; kb < CR_H / 64; ++kb) for (int g = 0; g < 2; ++g)
          gate_stage(tmem_base + CR_TM_ACC0 + g * CR_N, bh + kb * CR_BLOCK + CR_HALF, kb == 0);
        tc_commit(bar_early);
        // ---- ctx_t has landed: layer-0 input half
        cr_wait(bar_ctx, t & 1, 3, t);
        if (t + 1 < T) mbar_arrive_expect_tx(bar_ctx, CTX_TILE_BYTES);
        tc_fence_after();
        for (int kb = 0; kb < CR_E / 64; ++kb) for (int g = 0; g < 2; ++g)
          gate_stage(tmem_base + CR_TM_ACC0 + g * CR_N, bctx + kb * CR_HALF, false);
        tc_commit(bar_acc0);
        // ---- h0_t has landed: layer-1 input half
        cr_wait(bar_h0, (t + 1) & 1, 4, t);
        if (t + 1 < T) mbar_arrive_expect_tx(bar_h0, H_TILE_BYTES);
        tc_fence_after();
        for (int kb = 0; kb < CR_H / 64; ++kb) for (int g = 0; g < 2; ++g)
          gate_stage(tmem_base + CR_TM_ACC1 + g * CR_N, bh + kb * CR_BLOCK + CR_HALF, false);
        tc_commit(bar_acc1);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ workers
    const int wt = tid - 128, ww = warp - 4;           // worker thread / warp
    const int grp = ww >> 2, qd = ww & 3;              // M tile, TMEM lane quarter
    const int q = lane & 3;                            // gate of this lane (i, f, g, o)
    const int jl = 8 * qd + (lane >> 2);               // unit within the tile
    const int unit = CR_UNITS * c + 32 * grp + jl;     // hidden unit
    const int grow = 256 * c + 128 * grp + 32 * qd + lane;     // packed gate row (= 4 * unit + q)
    const uint32_t lane_bits = (uint32_t)(32 * qd) << 16;
    float* uin = reinterpret_cast<float*>(base + CR_OFF_UIN);
    float* e2u = reinterpret_cast<float*>(base + CR_OFF_E2U);
    float* cacc = reinterpret_cast<float*>(base + CR_OFF_CACC);
    float* sc = reinterpret_cast<float*>(base + CR_OFF_SC);
    bf16* hst = reinterpret_cast<bf16*>(base + CR_OFF_HST);
    const float act_k = (q == 2) ? 2.f : 1.f, act_a = (q == 2) ? 2.f : 1.f, act_b = (q == 2) ? -1.f : 0.f;
    const float bias1 = p.bias1[grow];
    float cst[CR_L][CR_N / 4];                          // cell state of (unit, sample 4s + q), fp32, for the whole sequence
#pragma unroll
    for (int k = 0; k < CR_L; ++k)
#pragma unroll
      for (int s = 0; s < CR_N / 4; ++s) {           // c_{-1}: slot 0 of the cell buffers (zeros, or the caller's initial state)
        const int smp = 4 * s + q;
        cst[k][s] = smp < n ? (k == 0 ? p.c0 : p.c1)[(long)(r0 + smp) * CR_H + unit] : 0.f;
      }
    const long SE = (long)S * CR_E;
    // ---- prologue: e^{2P} of this CTA's samples (read back by this CTA only)
    for (int sm = 0; sm < nv; ++sm) {
      const long off = (long)(r0 + CR_SPC * c + sm) * SE;
      for (long i = wt * 4; i < SE; i += CR_WORKERS * 4) {
        const float4 v = *reinterpret_cast<const float4*>(p.P + off + i);
        float4 o;
        o.x = ex2_ftz(2.8853900817779268f * v.x); o.y = ex2_ftz(2.8853900817779268f * v.y);
        o.z = ex2_ftz(2.8853900817779268f * v.z); o.w = ex2_ftz(2.8853900817779268f * v.w);
        *reinterpret_cast<float4*>(p.EP + off + i) = o;
      }
    }
    __threadfence_block();
    cr_bar_workers();

    for (int t = 0; t < T; ++t) {
      unsigned long long* tr = (p.trace && wt == 0) ? p.trace + ((size_t)blockIdx.x * T + t) * 8 : nullptr;
      if (tr) tr[0] = clock64();
      // ================= u epilogue: rows of the cluster, this CTA's 32 columns -> owner CTA of each row
      cr_wait(bar_u, t & 1, 5, t);
      tc_fence_after();
      if (grp == 0 && qd < 2) {
        const int row = 32 * qd + lane;
        float v[32];
        tmem_ld32(tmem_base + CR_TM_U + lane_bits, v);
        if (row < n) {
          float* ug = p.u + ((long)t * B + r0 + row) * CR_E + CR_UCOLS * c;
          const int dst = row / CR_SPC, slot = row % CR_SPC;
          const uint32_t ra = cr_mapa(smem_u32(uin + slot * CR_E + CR_UCOLS * c), (uint32_t)dst), rb = cr_mapa(smem_u32(bar_uin), (uint32_t)dst);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            cr_st_async_v4(ra + j * 4, v[j], v[j + 1], v[j + 2], v[j + 3], rb);
            *reinterpret_cast<float4*>(ug + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        }
      }
      tc_fence_before();
      if (tr) tr[1] = clock64();
      // ================= attention for this CTA's sample slots
      cr_wait(bar_uin, t & 1, 6, t);
      for (int i = wt; i < nv * CR_E; i += CR_WORKERS) e2u[i] = ex2_ftz(2.8853900817779268f * uin[i]);
      for (int i = wt; i < CR_SPC * CR_E; i += CR_WORKERS) cacc[i] = 0.f;
      cr_bar_workers();
      // next phase armed only after EVERY worker has observed this one: a CTA without samples expects 0 bytes, so its next phase
      // completes at the arrive, and a thread still polling the old parity would then wait forever
      if (wt == 0 && t + 1 < T) mbar_arrive_expect_tx(bar_uin, uin_bytes);
      if (tr) tr[2] = clock64();
      {
        // scores: item (sample, token) per warp; lane covers columns 4*lane..+3 and 128+4*lane..+3
        const int items = nv * S;
        constexpr int BATCH = 8;
        for (int i0 = ww; i0 < items; i0 += 8 * BATCH) {
          float4 pa[BATCH], pb[BATCH];
#pragma unroll
          for (int k = 0; k < BATCH; ++k) {
            const int it = i0 + 8 * k;
            if (it < items) {
              const int sm = it / S, l = it - sm * S;
              const float* ep = p.EP + (long)(r0 + CR_SPC * c + sm) * SE + (long)l * CR_E;
              pa[k] = __ldcg(reinterpret_cast<const float4*>(ep) + lane);
              pb[k] = __ldcg(reinterpret_cast<const float4*>(ep + 128) + lane);
            }
          }
#pragma unroll
          for (int k = 0; k < BATCH; ++k) {
            const int it = i0 + 8 * k;
            if (it < items) {
              const int sm = it / S, l = it - sm * S;
              const float4 ua = *reinterpret_cast<const float4*>(e2u + sm * CR_E + 4 * lane);
              const float4 ub = *reinterpret_cast<const float4*>(e2u + sm * CR_E + 128 + 4 * lane);
              // sum of tanh = 8 - 2 * sum 1 / (1 + e^{2P} e^{2u})
              float a = rcp_ftz_(fmaf(pa[k].x, ua.x, 1.f)) + rcp_ftz_(fmaf(pa[k].y, ua.y, 1.f)) + rcp_ftz_(fmaf(pa[k].z, ua.z, 1.f)) + rcp_ftz_(fmaf(pa[k].w, ua.w, 1.f));
              a += rcp_ftz_(fmaf(pb[k].x, ub.x, 1.f)) + rcp_ftz_(fmaf(pb[k].y, ub.y, 1.f)) + rcp_ftz_(fmaf(pb[k].z, ub.z, 1.f)) + rcp_ftz_(fmaf(pb[k].w, ub.w, 1.f));
              a = warp_sum(a);
              if (lane == 0) sc[sm * CR_SMAX + l] = fmaf(-2.f, a, (float)CR_E);
            }
          }
        }
      }
      cr_bar_workers();
      if (ww < nv) {                                   // softmax over the tokens of sample slot ww
        float* s_ = sc + ww * CR_SMAX;
        const float v0 = lane < S ? s_[lane] : -INFINITY, v1 = lane + 32 < S ? s_[lane + 32] : -INFINITY;
        const float m = warp_max(fmaxf(v0, v1));
        const float e0 = lane < S ? ex2_ftz(1.4426950408889634f * (v0 - m)) : 0.f, e1 = lane + 32 < S ? ex2_ftz(1.4426950408889634f * (v1 - m)) : 0.f;
        const float inv = 1.0f / warp_sum(e0 + e1);
        float* aw = p.attw + ((long)t * B + r0 + CR_SPC * c + ww) * S;
        if (lane < S) { s_[lane] = e0 * inv; aw[lane] = e0 * inv; }
        if (lane + 32 < S) { s_[lane + 32] = e1 * inv; aw[lane + 32] = e1 * inv; }
      }
      cr_bar_workers();
      if (tr) tr[3] = clock64();
      {
        // context: warp ww takes tokens l = ww, ww + 8, ...; lane covers columns 8*lane..+7 of every sample slot
        float acc[CR_SPC][8];
#pragma unroll
        for (int sm = 0; sm < CR_SPC; ++sm)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[sm][j] = 0.f;
        for (int l = ww; l < S; l += 8) {
          uint4 f[CR_SPC];
#pragma unroll
          for (int sm = 0; sm < CR_SPC; ++sm)
            if (sm < nv) f[sm] = ld_nc_v4(p.F + (long)(r0 + CR_SPC * c + sm) * SE + (long)l * CR_E + 8 * lane);
#pragma unroll
          for (int sm = 0; sm < CR_SPC; ++sm)
            if (sm < nv) {
              const float w = sc[sm * CR_SMAX + l];
              acc[sm][0] = fmaf(w, bf16_lo(f[sm].x), acc[sm][0]); acc[sm][1] = fmaf(w, bf16_hi(f[sm].x), acc[sm][1]);
              acc[sm][2] = fmaf(w, bf16_lo(f[sm].y), acc[sm][2]); acc[sm][3] = fmaf(w, bf16_hi(f[sm].y), acc[sm][3]);
              acc[sm][4] = fmaf(w, bf16_lo(f[sm].z), acc[sm][4]); acc[sm][5] = fmaf(w, bf16_hi(f[sm].z), acc[sm][5]);
              acc[sm][6] = fmaf(w, bf16_lo(f[sm].w), acc[sm][6]); acc[sm][7] = fmaf(w, bf16_hi(f[sm].w), acc[sm][7]);
            }
        }
#pragma unroll
        for (int sm = 0; sm < CR_SPC; ++sm)
          if (sm < nv) {
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(cacc + sm * CR_E + j * 32 + lane, acc[sm][j]);     // [slot][j][lane]: conflict-free
          }
      }
      cr_bar_workers();
      for (int i = wt; i < nv * (CR_E / 2); i += CR_WORKERS) {           // ctx rows -> layer 0's operand rows (forward save + multicast source)
        const int sm = i / (CR_E / 2), e = 2 * (i % (CR_E / 2));
        const float a0 = cacc[sm * CR_E + (e & 7) * 32 + (e >> 3)], a1 = cacc[sm * CR_E + ((e + 1) & 7) * 32 + ((e + 1) >> 3)];
        *reinterpret_cast<uint32_t*>(p.xh0 + ((long)t * B + r0 + CR_SPC * c + sm) * (CR_E + CR_H) + e) = pack_bf16(a0, a1);
      }
      cr_fence_proxy_global();
      cr_bar_workers();
      if (wt == 0) {
        cr_wait(bar_early, t & 1, 7, t);                   // this CTA is done reading h0_{t-1} (see the WAR notes on top)
        for (int kb = 0; kb < CR_E / 64; ++kb)
          cr_tma_mc(base + CR_OFF_CTX + kb * CR_HALF + CR_SPC * c * 128, &maps.ctx, 64 * kb, t * B + r0 + CR_SPC * c, bar_ctx, 0xFF);
      }
      if (tr) tr[4] = clock64();
      // ================= LSTM cells (layer 0, then layer 1)
#pragma unroll
      for (int k = 0; k < CR_L; ++k) {
        // addend of this gate row for the 40 sample slots, fetched before the accumulator wait
        uint4 ad[5];
        if (k == 0) {
          const uint4* ap = reinterpret_cast<const uint4*>(p.G0T + (((long)t * p.ncl + ci) * (4 * CR_H) + grow) * CR_RMAX);
#pragma unroll
          for (int j = 0; j < 5; ++j) ad[j] = ld_nc_v4(ap + j);
        }
        cr_wait(k == 0 ? bar_acc0 : bar_acc1, t & 1, 8 + k, t);
        tc_fence_after();
        float v[CR_N];
        const uint32_t tacc = tmem_base + (k == 0 ? CR_TM_ACC0 : CR_TM_ACC1) + grp * CR_N + lane_bits;
        cr_tmem_ld16(tacc, v); cr_tmem_ld16(tacc + 16, v + 16); cr_tmem_ld16(tacc + 32, v + 32);
        tc_fence_before();
        if (k == 0) {
          const uint32_t* aw = reinterpret_cast<const uint32_t*>(ad);
#pragma unroll
          for (int j = 0; j < 20; ++j) { v[2 * j] += bf16_lo(aw[j]); v[2 * j + 1] += bf16_hi(aw[j]); }
        } else {
#pragma unroll
          for (int j = 0; j < CR_RMAX; ++j) v[j] += bias1;
        }
        // activation of this lane's gate: sigmoid(x) or tanh(x) = 2 sigmoid(2x) - 1
#pragma unroll
        for (int j = 0; j < CR_N; ++j) v[j] = fmaf(act_a, rcp_ftz_(1.0f + ex2_ftz_(-1.4426950408889634f * act_k * v[j])), act_b);
        bf16* gout = k == 0 ? p.gates0 : p.gates1;
        float* cout = k == 0 ? p.c0 : p.c1;
#pragma unroll
        for (int s = 0; s < CR_N / 4; ++s) {
          float a[4] = {v[4 * s], v[4 * s + 1], v[4 * s + 2], v[4 * s + 3]};
          cr_quad_transpose(a, q);                     // a = (i, f, g, o) of (unit, sample 4s + q)
          const int smp = 4 * s + q;
          const float cn = fmaf(a[1], cst[k][s], a[0] * a[2]);
          const float hn = a[3] * Math<bf16>::tanh_(cn);
          cst[k][s] = cn;
          if (smp < n) {
            const long row = (long)t * B + r0 + smp;
            *reinterpret_cast<uint2*>(gout + row * (4 * CR_H) + 4 * unit) = make_uint2(pack_bf16(a[0], a[1]), pack_bf16(a[2], a[3]));
            cout[(row + B) * CR_H + unit] = cn;
          }
          if (smp < CR_RMAX) hst[smp * 64 + 32 * grp + jl] = __float2bfloat16_rn(hn);
        }
        cr_bar_workers();
        // staged h rows -> global: 40 rows x 128 bytes, 16 bytes per thread
        for (int i = wt; i < n * 8; i += CR_WORKERS) {
          const int smp = i >> 3, ch = i & 7;
          const uint4 hv = *reinterpret_cast<const uint4*>(hst + smp * 64 + ch * 8);
          const long row = (long)t * B + r0 + smp;
          if (k == 0) {
            *reinterpret_cast<uint4*>(p.xh0 + (row + B) * (CR_E + CR_H) + CR_E + CR_UNITS * c + ch * 8) = hv;       // recurrent slot of step t+1
            *reinterpret_cast<uint4*>(p.xh1 + row * (2 * CR_H) + CR_UNITS * c + ch * 8) = hv;                      // layer 1's input at step t
          } else {
            *reinterpret_cast<uint4*>(p.xh1 + (row + B) * (2 * CR_H) + CR_H + CR_UNITS * c + ch * 8) = hv;
            *reinterpret_cast<uint4*>(p.hid_top + row * CR_H + CR_UNITS * c + ch * 8) = hv;
          }
        }
        cr_fence_proxy_global();
        cr_bar_workers();
        if (wt == 0) {
          if (k == 0) cr_tma_mc(base + CR_OFF_BH + c * CR_BLOCK + CR_HALF, &maps.h0, CR_E + 64 * c, (t + 1) * B + r0, bar_h0, 0xFF);
          else if (t + 1 < T) cr_tma_mc(base + CR_OFF_BH + c * CR_BLOCK, &maps.h1, CR_H + 64 * c, (t + 1) * B + r0, bar_h1, 0xFF);
        }
        if (tr) tr[5 + k] = clock64();
      }
    }
  }
  // nobody leaves while a peer may still write into its shared memory or signal its barriers
  __syncthreads();
  cr_cluster_sync();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(CR_TM_COLS) : "memory");
}

}  // namespace b2c
