// The T-step forward recurrence of LSTMDecoder.forward (reference src/student_model.py:232-251) as ONE persistent cooperative kernel.
//
// Per time step the multi-kernel path launches  u GEMM -> attention -> L fused gate-GEMM + cell kernels, a chain of 2 + L
// dependent launches whose arithmetic is ~1 us each but whose launch / prologue / drain latency is 6-9 us each (profiles/README.md,
// round 1: 31 us per step, 49 % of the KD step together with the reverse recurrence).  Here the whole T loop runs inside one
// kernel: one CTA per SM, co-resident (cooperative launch), each owning
//   * one 128 x 64 tile of every layer's gate contraction (row tile m, 16 hidden units n) for ALL time steps, with the fused
//     LSTM-cell epilogue of gemm.cuh (TMEM -> bias / addend -> gates -> c (fp32) -> h scattered to its consumers),
//   * (the first E/32 * tiles_m CTAs) one 128 x 32 tile of the attention query projection  u_t = h^{top}_{t-1} W_h^T,
//   * ceil(B / grid) samples of the spatial attention (their feature tokens F_b stay resident in shared memory when they fit),
// and the 2 + L kernel boundaries per step become grid-wide barriers (one atomic + a flag spin, ~0.5 us).
//
// Warp roles (320 threads, as in gemm_tc_kernel): warp 0 = TMA producer (one thread), warp 1 = tcgen05.mma issuer (one thread),
// warps 2-9 = epilogue / attention (256 threads, two warps per TMEM lane quarter).
//
// Off-critical-path contraction halves: the operand row of layer k is [input_t ; h^k_{t-1}].  Its recurrent half h^k_{t-1} has been
// final since step t-1, so  acc_k = h^k_{t-1} W_hh^T  is issued right after the barrier that ends step t-1 and runs on the tensor
// pipe UNDER the (SFU-bound) attention phase; only  acc_k += input_t W_in^T  (K = E or H instead of E+H or 2H) sits between the
// barrier that publishes input_t and the cell epilogue.  Every layer has its own TMEM accumulator (64 columns).
//
// Grid barrier: a monotonically increasing arrival counter in global memory (zeroed by the host before the launch); phase p of step
// t is complete when the counter reaches (t * (2 + L) + p + 1) * gridDim.x.  Writers publish with  st.global -> fence.proxy.async
// (the consumers read through TMA, i.e. the async proxy) -> __threadfence -> bar.sync over the 256 epilogue threads -> red.release;
// readers spin with ld.acquire.  Every spin is bounded and traps on timeout instead of hanging the GPU.
#pragma once
#include "gemm.cuh"

namespace b2c {

constexpr int RC_MAXL = B2C_MAX_LAYERS;
constexpr int RC_BN = 64;               // gate-GEMM tile columns (16 hidden units)
constexpr int RC_BNU = 32;              // u-GEMM tile columns
constexpr int RC_STAGES = 4;
constexpr uint32_t RC_STAGE_BYTES = TC_A_BYTES + RC_BN * TC_BK * 2;      // 16 KB + 8 KB
constexpr int RC_EPI_THREADS = 256;
// 12 warps = 3 warpgroups: warpgroup 0 holds the TMA producer (warp 0) and the MMA issuer (warp 1) and gives most of its registers
// away (setmaxnreg.dec), warpgroups 1-2 are the 256 epilogue / attention threads and take them (setmaxnreg.inc): with ~200 KB of
// shared memory the L1 is ~28 KB, so a spilled register is an L2 round trip (first version, 320 threads / 168 registers: 335 MB of
// local-memory traffic per launch, 30 % L1 hit rate -- ncu, profiles/README.md round 2).
constexpr int RC_THREADS = 384;
constexpr int RC_REGS_CTRL = 56, RC_REGS_EPI = 224;       // 128 * 56 + 256 * 224 = 64512 = 384 * 168
constexpr int RC_MAX_RES = 4;           // resident feature-token buffers per CTA

struct RecurParams {
  int B, T, S, E, H, L;
  int tiles_m, tiles_n, u_tiles_n;      // gate tiles: tiles_m x tiles_n (= 4H / 64); u tiles: tiles_m x (E / 32)
  int n_fbuf, f_resident;               // feature-token buffers in shared memory; 1: every sample of this CTA has its own
  int tmem_cols;
  const float* P;                       // (B, S, E) fp32 hoisted projection
  float* EP;                            // (B, S, E) fp32 scratch: e^{2P}, written and read by the CTA that owns the sample
  unsigned long long* trace;            // optional (debug): per CTA, per step, 8 clock64() stamps of the epilogue group
  const bf16* F;                        // (B, S, E) feature tokens
  float* u;                             // (T*B, E) fp32
  float* attw;                          // (T, B, S) fp32
  bf16* xh[RC_MAXL]; int ld[RC_MAXL]; int in[RC_MAXL];       // (T+1, B, ld) operand rows [input ; h]
  const bf16* G0;                       // (T*B, 4H) layer-0 addend (embedding half + b_x), interleaved columns
  const float* bias[RC_MAXL];           // (4H) interleaved, layers >= 1
  float* c[RC_MAXL];                    // (T+1, B, H)
  bf16* gates[RC_MAXL];                 // (T*B, 4H)
  bf16* hid_top;                        // (T, B, H)
  float drop_p; unsigned long long seed; const unsigned long long* seed_dev;
  unsigned int* barrier;                // arrival counter, zeroed before the launch
  unsigned int* flags;                  // one flag word per CTA (128-byte stride), zeroed before the launch
  int early;                            // 1: recurrent halves of the gate contractions issued ahead, under the attention phase
  int ep_nc;                            // 1: e^{2P} rows through the non-coherent (L1) path, 0: L2 only (ld.global.cg)
};
struct RecurMaps { CUtensorMap a[RC_MAXL]; CUtensorMap w[RC_MAXL]; CUtensorMap wh; };

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               :: "r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ unsigned int ld_relaxed_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory"); }

// Grid barrier without a hot polling line.  Arrivals are counted on one line (128 fire-and-return atomics per phase); the CTA
// whose arrival completes phase number `phase` (1-based, monotonically increasing over the launch) publishes it to ONE FLAG LINE
// PER CTA (128-byte stride), and every CTA polls only its own line.  (First version: all CTAs polled the arrival counter itself;
// ~256 pollers on one address serialise in a single L2 slice, and because every TMA tile touches every slice, each operand load of
// every CTA queued behind them: ~1 us per L2 round trip instead of ~0.15 us, profiles/README.md round 2.)
constexpr int RC_FLAG_STRIDE = 32;      // unsigned ints between two CTAs' flag words (128 bytes)
__device__ __forceinline__ void grid_wait(const unsigned int* flags, unsigned int phase) {
  const unsigned int* f = flags + (size_t)blockIdx.x * RC_FLAG_STRIDE;
  for (unsigned int spins = 0; (int)(ld_relaxed_gpu(f) - phase) < 0; ++spins) {
    if (spins > (1u << 26)) { printf("b2c: grid barrier timed out (block %d thread %d phase %u have %u)\n", blockIdx.x, threadIdx.x, phase, ld_relaxed_gpu(f)); __trap(); }
  }
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
// The 256 epilogue threads have finished the stores of this phase.  bar.sync orders every thread's stores before the arriving
// warp's gpu-scope atomic (acq_rel, cumulative, as in cooperative-groups grid.sync); the per-thread fence.proxy.async orders the
// generic-proxy stores for the consumers' TMA (async proxy) reads.  Warp 2 arrives: lane 0 counts, and if this CTA is the last one
// of the phase the 32 lanes publish the phase number to every CTA's flag line.
__device__ __forceinline__ void grid_arrive_epi(unsigned int* ctr, unsigned int* flags, unsigned int phase, int epi_tid) {
  fence_proxy_async_all();
  named_bar_sync(1, RC_EPI_THREADS);
  if (epi_tid < 32) {
    // A CTA without work in a phase arrives for it without waiting for the previous phase to complete (e.g. a CTA without a u tile
    // arrives for "u of step t+1" right behind "last layer of step t"), so consecutive phases must not share a counter: four
    // counters (one line each) used round-robin; the j-th use of a counter is complete at (j + 1) * gridDim.x arrivals.
    const unsigned int p0 = phase - 1;
    unsigned int* slot = ctr + (p0 & 3u) * RC_FLAG_STRIDE;
    unsigned int old = 0;
    if (epi_tid == 0) asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(slot) : "memory");
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old + 1 == ((p0 >> 2) + 1u) * gridDim.x) {
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      for (unsigned int c = epi_tid; c < gridDim.x; c += 32)
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(flags + (size_t)c * RC_FLAG_STRIDE), "r"(phase) : "memory");
    }
  }
}

template <int NQ>       // NQ = float4 per lane per token row (E / 128), 1..3
__global__ void __launch_bounds__(RC_THREADS, 1)
recur_fwd_kernel(const __grid_constant__ RecurMaps maps, const __grid_constant__ RecurParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
  // layout: [ring: RC_STAGES x 24 KB][barriers 512 B][us: RC_MAX_RES x E floats][sc: RC_MAX_RES x 64 floats][F buffers: n_fbuf x S*E bf16]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(base + (size_t)RC_STAGES * RC_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + RC_STAGES;
  uint64_t* tfull_bar = empty_bar + RC_STAGES;          // [RC_MAXL + 1]: accumulator complete (MMA -> epilogue); index L = the u tile
  uint64_t* tempty_bar = tfull_bar + (RC_MAXL + 1);     // [RC_MAXL + 1]: accumulator drained (epilogue -> MMA)
  uint64_t* f_bar = tempty_bar + (RC_MAXL + 1);         // [RC_MAX_RES]: feature-token buffers landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(f_bar + RC_MAX_RES);
  float* us = reinterpret_cast<float*>(base + (size_t)RC_STAGES * RC_STAGE_BYTES + 512);
  float* sc = us + RC_MAX_RES * p.E;
  bf16* Fbuf = reinterpret_cast<bf16*>(sc + RC_MAX_RES * 64);
  const int SE = p.S * p.E;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x, G = gridDim.x;
  const int L = p.L, B = p.B, T = p.T, E = p.E, H = p.H, S = p.S;
  const int n_tiles = p.tiles_m * p.tiles_n, n_utiles = p.tiles_m * p.u_tiles_n;
  const bool has_tile = cta < n_tiles, has_utile = cta < n_utiles;
  const int m0 = has_tile ? (cta / p.tiles_n) * TC_BM : 0, n0 = has_tile ? (cta % p.tiles_n) * RC_BN : 0;
  const int um0 = has_utile ? (cta / p.u_tiles_n) * TC_BM : 0, un0 = has_utile ? (cta % p.u_tiles_n) * RC_BNU : 0;
  const int phases = 2 + L;                              // per step: u | attention | layer 0 .. L-1
  const int inL = p.in[L - 1];

  if (warp == 0 && lane == 0) {
    for (int k = 0; k < L; ++k) { tma_prefetch_desc(&maps.a[k]); tma_prefetch_desc(&maps.w[k]); }
    tma_prefetch_desc(&maps.wh);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < RC_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
      for (int i = 0; i <= RC_MAXL; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], TC_EPI_WARPS); }
      for (int i = 0; i < RC_MAX_RES; ++i) mbar_init(&f_bar[i], 1);
      fence_barrier_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // accumulator of layer k: columns [64 k, 64 k + 64); u tile: columns [64 L, 64 L + 32)

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(RC_REGS_CTRL));
  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      int it = 0;
      auto load_job = [&](const CUtensorMap* ma, int a_k0, int a_m0, int slot, const CUtensorMap* mw, int w_k0, int w_n0, int nkb, uint32_t bbytes) {
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % RC_STAGES; const uint32_t ph = (it / RC_STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          unsigned char* sa = base + (size_t)s * RC_STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[s], TC_A_BYTES + bbytes);
          tma_load_3d(sa, ma, a_k0 + kb * TC_BK, a_m0, slot, &full_bar[s]);            // box {64 k, 128 rows, 1 slot}
          tma_load_2d(sa + TC_A_BYTES, mw, w_k0 + kb * TC_BK, w_n0, &full_bar[s]);     // box {64 k, BN rows}
        }
      };
      for (int t = 0; t < T; ++t) {
        // everything below reads h of step t-1: wait for the barrier that ended step t-1
        if (t > 0) { grid_wait(p.flags, (unsigned)(t * phases)); fence_proxy_async_all(); }
        if (p.trace) p.trace[((size_t)cta * T + t) * 16 + 15] = clock64();
        if (has_utile) load_job(&maps.a[L - 1], inL, um0, t, &maps.wh, 0, un0, H / TC_BK, RC_BNU * TC_BK * 2);
        if (has_tile) {
          for (int k = 0; k < L && p.early; ++k)                                       // recurrent halves: K range [in_k, in_k + H)
            load_job(&maps.a[k], p.in[k], m0, t, &maps.w[k], p.in[k], n0, H / TC_BK, RC_BN * TC_BK * 2);
          for (int k = 0; k < L; ++k) {                                                // input halves, each behind its producer's barrier
            grid_wait(p.flags, (unsigned)(t * phases + 2 + k));                        // k = 0: attention done; k > 0: layer k-1 done
            fence_proxy_async_all();
            load_job(&maps.a[k], 0, m0, t, &maps.w[k], 0, n0, (p.early ? p.in[k] : p.ld[k]) / TC_BK, RC_BN * TC_BK * 2);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_g = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(RC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      constexpr uint32_t idesc_u = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(RC_BNU >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int it = 0;
      auto mma_job = [&](uint32_t tacc, uint32_t idesc, int nkb, bool fresh) {
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % RC_STAGES; const uint32_t ph = (it / RC_STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(base + (size_t)s * RC_STAGE_BYTES);
          const uint32_t sb = sa + TC_A_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            tc_mma_bf16(tacc, make_smem_desc(sa + k * 32, 16, 1024), make_smem_desc(sb + k * 32, 16, 1024), idesc, (!fresh || kb > 0 || k > 0) ? 1u : 0u);
          tc_commit(&empty_bar[s]);
        }
      };
      for (int t = 0; t < T; ++t) {
        const uint32_t par = (uint32_t)(t & 1);
        if (has_utile) {
          mbar_wait(&tempty_bar[L], par ^ 1);
          tc_fence_after();
          mma_job(tmem_base + (uint32_t)(64 * L), idesc_u, H / TC_BK, true);
          tc_commit(&tfull_bar[L]);
        }
        if (p.trace) p.trace[((size_t)cta * T + t) * 16 + 11] = clock64();
        if (has_tile) {
          for (int k = 0; k < L; ++k) {
            mbar_wait(&tempty_bar[k], par ^ 1);                                        // the epilogue of step t-1 has drained this accumulator
            tc_fence_after();
            if (p.early) mma_job(tmem_base + (uint32_t)(64 * k), idesc_g, H / TC_BK, true);
          }
          if (p.trace) p.trace[((size_t)cta * T + t) * 16 + 12] = clock64();
          for (int k = 0; k < L; ++k) {
            mma_job(tmem_base + (uint32_t)(64 * k), idesc_g, (p.early ? p.in[k] : p.ld[k]) / TC_BK, !p.early);
            tc_commit(&tfull_bar[k]);
            if (p.trace) p.trace[((size_t)cta * T + t) * 16 + 13 + (k < 1 ? 0 : 1)] = clock64();
          }
        }
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(RC_REGS_EPI));
    // ------------------------------------------------ epilogue / attention warps (256 threads)
    const int et = threadIdx.x - 128, ew = warp - 4;     // 0..255, 0..7
    const int q = warp & 3, half = ew >> 2;              // TMEM lane quarter, column half
    const int E4 = E >> 2;
    const float inv_keep = p.drop_p > 0.f ? 1.0f / (1.0f - p.drop_p) : 1.0f;
    const uint64_t dseed = p.drop_p > 0.f ? drop_seed(p.seed, p.seed_dev) : 0;
    const int n_mine = cta < B ? (B - cta + G - 1) / G : 0;       // samples b = cta + i G  (<= RC_MAX_RES, all resident)
    const float LOG2E2 = 2.8853900817779268f;            // 2 log2(e):  e^{2x} = 2^{LOG2E2 x}
    // ---- once: the feature tokens of my samples into shared memory (bulk TMA), and EP = e^{2P} of my samples (clamped so that
    // the product with e^{2u} can overflow / underflow but never be inf * 0).  The score sum is
    //   s_l = sum_e tanh(P_le + u_e) = E - 2 sum_e 1 / (1 + e^{2P_le} e^{2u_e}):  ONE MUFU (rcp) per element instead of two (ex2 + rcp).
    if (et == 0) {
      for (int i = 0; i < n_mine; ++i) {
        const uint32_t fb = (uint32_t)(SE * sizeof(bf16));
        mbar_arrive_expect_tx(&f_bar[i], fb);
        bulk_g2s(Fbuf + (size_t)i * SE, p.F + (size_t)(cta + i * G) * SE, fb, &f_bar[i]);
      }
    }
    for (int i = 0; i < n_mine; ++i) {
      const float4* src = reinterpret_cast<const float4*>(p.P + (size_t)(cta + i * G) * SE);
      float4* dst = reinterpret_cast<float4*>(p.EP + (size_t)(cta + i * G) * SE);
      for (int x = et; x < SE / 4; x += RC_EPI_THREADS) {
        const float4 v = __ldg(src + x);
        dst[x] = make_float4(ex2_ftz(fminf(fmaxf(LOG2E2 * v.x, -120.f), 120.f)), ex2_ftz(fminf(fmaxf(LOG2E2 * v.y, -120.f), 120.f)),
                             ex2_ftz(fminf(fmaxf(LOG2E2 * v.z, -120.f), 120.f)), ex2_ftz(fminf(fmaxf(LOG2E2 * v.w, -120.f), 120.f)));
      }
    }
    __threadfence();
    named_bar_sync(1, RC_EPI_THREADS);                   // EP rows are read back by other threads of this CTA (through L2)
    unsigned long long* trace = p.trace ? p.trace + (size_t)cta * T * 16 : nullptr;
    const int n_items = n_mine * S;                      // (sample, token) score items, dealt round-robin to the 8 warps
    const int n_groups = ((n_items + 7) / 8 + 3) / 4;    // items per warp, in register-prefetched groups of 4
    for (int t = 0; t < T; ++t) {
      const uint32_t par = (uint32_t)(t & 1);
      const long slot = (long)t * B;                     // first row of step t in a (T, B, .) buffer
      if (trace && et == 0) trace[t * 16 + 0] = clock64();
      // ---------------- phase 0: u tile epilogue
      if (has_utile) {
        mbar_wait(&tfull_bar[L], par);
        tc_fence_after();
        if (half == 0) {
          float v[32];
          tmem_ld32(tmem_base + (uint32_t)(64 * L) + ((uint32_t)(q * 32) << 16), v);
          const int grow = um0 + q * 32 + lane;
          if (grow < B) {
            float4* dst = reinterpret_cast<float4*>(p.u + (slot + grow) * E + un0);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[L]);
      }
      grid_arrive_epi(p.barrier, p.flags, (unsigned)(t * phases + 1), et);
      if (trace && et == 0) trace[t * 16 + 1] = clock64();
      // ---------------- phase 1: attention of ALL my samples at once (needs u of every column tile)
      // the first group of EP rows does not depend on this step: in flight while the barrier is awaited
      float4 ga[4][NQ], gb[4][NQ];
      auto load_group = [&](int g, float4 (&dst)[4][NQ]) {
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const int it = ew + 8 * (4 * g + d);
          if (it < n_items) {
            const int i = it / S, l = it - i * S;
            const float4* row = reinterpret_cast<const float4*>(p.EP + (size_t)(cta + i * G) * SE) + (size_t)l * E4;
#pragma unroll
            for (int j = 0; j < NQ; ++j) { const int qq = lane + j * 32; if (qq < E4) dst[d][j] = p.ep_nc ? __ldg(row + qq) : __ldcg(row + qq); }
          }
        }
      };
      auto score_group = [&](int g, const float4 (&src)[4][NQ]) {
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const int it = ew + 8 * (4 * g + d);
          if (it < n_items) {
            const int i = it / S, l = it - i * S;
            float a = 0.f;
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
              const int qq = lane + j * 32;
              if (qq < E4) {
                const float4 pv = src[d][j];
                const float4 uu = *reinterpret_cast<const float4*>(us + i * E + qq * 4);         // e^{2u}
                a += rcp_ftz_(fmaf(pv.x, uu.x, 1.0f)) + rcp_ftz_(fmaf(pv.y, uu.y, 1.0f)) + rcp_ftz_(fmaf(pv.z, uu.z, 1.0f)) + rcp_ftz_(fmaf(pv.w, uu.w, 1.0f));
              }
            }
            a = warp_sum(a);
            if (lane == 0) sc[i * 64 + l] = fmaf(-2.0f, a, (float)E);
          }
        }
      };
      if (n_groups > 0) load_group(0, ga);
      if (et == 0) grid_wait(p.flags, (unsigned)(t * phases + 1));
      named_bar_sync(1, RC_EPI_THREADS);
      if (trace && et == 0) trace[t * 16 + 2] = clock64();
      for (int x = et; x < n_mine * E; x += RC_EPI_THREADS) {
        const int i = x / E, e = x - i * E;
        us[x] = ex2_ftz(fminf(fmaxf(LOG2E2 * __ldcg(p.u + (slot + cta + i * G) * E + e), -120.f), 120.f));
      }
      named_bar_sync(1, RC_EPI_THREADS);
      if (trace && et == 0) trace[t * 16 + 3] = clock64();
      for (int g = 0; g < n_groups; g += 2) {
        if (g + 1 < n_groups) load_group(g + 1, gb);
        score_group(g, ga);
        if (g + 2 < n_groups) load_group(g + 2, ga);
        if (g + 1 < n_groups) score_group(g + 1, gb);
      }
      named_bar_sync(1, RC_EPI_THREADS);
      if (trace && et == 0) trace[t * 16 + 4] = clock64();
      if (ew < n_mine) {                                 // softmax over the tokens: one warp per sample
        float* sci = sc + ew * 64;
        const long b = cta + ew * G;
        float m = -INFINITY;
        for (int l = lane; l < S; l += 32) m = fmaxf(m, sci[l]);
        m = warp_max(m);
        float ssum = 0.f;
        for (int l = lane; l < S; l += 32) { const float e = Math<bf16>::exp_(sci[l] - m); sci[l] = e; ssum += e; }
        ssum = warp_sum(ssum);
        const float inv = 1.0f / ssum;
        for (int l = lane; l < S; l += 32) { const float wv = sci[l] * inv; sci[l] = wv; p.attw[(slot + b) * S + l] = wv; }
      }
      named_bar_sync(1, RC_EPI_THREADS);
      if (trace && et == 0) trace[t * 16 + 5] = clock64();
      if (t == 0) for (int i = 0; i < n_mine; ++i) mbar_wait(&f_bar[i], 0);
      for (int w = et; w < n_mine * E4; w += RC_EPI_THREADS) {       // context: 4 columns of one sample per thread
        const int i = w / E4, c4 = w - i * E4;
        const bf16* Fs = Fbuf + (size_t)i * SE + c4 * 4;
        const float* sci = sc + i * 64;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int l = 0; l < S; ++l) {
          const uint2 f = *reinterpret_cast<const uint2*>(Fs + (size_t)l * E);
          const float wv = sci[l];
          a0 = fmaf(wv, bf16_lo(f.x), a0); a1 = fmaf(wv, bf16_hi(f.x), a1); a2 = fmaf(wv, bf16_lo(f.y), a2); a3 = fmaf(wv, bf16_hi(f.y), a3);
        }
        *reinterpret_cast<uint2*>(p.xh[0] + (slot + cta + (long)i * G) * p.ld[0] + c4 * 4) = make_uint2(pack_bf16(a0, a1), pack_bf16(a2, a3));
      }
      if (trace && et == 0) trace[t * 16 + 6] = clock64();
      grid_arrive_epi(p.barrier, p.flags, (unsigned)(t * phases + 2), et);
      // ---------------- phases 2 .. 1 + L: fused LSTM cell of this CTA's tile, layer by layer
      for (int k = 0; k < L; ++k) {
        if (has_tile) {
          const int N4 = 4 * H, ldk = p.ld[k], ink = p.in[k];
          const bf16* addend = (k == 0) ? p.G0 + slot * N4 : nullptr;
          const float* bias = p.bias[k];
          const float* c_prev = p.c[k] + slot * H;
          float* c_out = p.c[k] + (slot + B) * H;
          bf16* gates_out = p.gates[k] + slot * N4;
          bf16* h_rec = p.xh[k] + (slot + B) * ldk + ink;
          bf16* h_next = (k + 1 < L) ? p.xh[k + 1] + slot * p.ld[k + 1] : nullptr;
          const long ld_next = (k + 1 < L) ? p.ld[k + 1] : 0;
          bf16* h_top = (k == L - 1) ? p.hid_top + slot * H : nullptr;
          const int grow = m0 + q * 32 + lane;
          // operands that do not come from the MMA, fetched before the accumulator wait
          const int col0 = n0 + half * 32, u0 = col0 >> 2;
          float4 pre_b[8]; uint4 pre_a[4]; float4 pc0 = make_float4(0.f, 0.f, 0.f, 0.f), pc1 = pc0;
          if (grow < B) {
            if (bias) {
#pragma unroll
              for (int j = 0; j < 8; ++j) pre_b[j] = *reinterpret_cast<const float4*>(bias + col0 + 4 * j);
            }
            if (addend) {
              const uint4* ap = reinterpret_cast<const uint4*>(addend + (long)grow * N4 + col0);
#pragma unroll
              for (int j = 0; j < 4; ++j) pre_a[j] = ap[j];
            }
            pc0 = __ldcg(reinterpret_cast<const float4*>(c_prev + (long)grow * H + u0));
            pc1 = __ldcg(reinterpret_cast<const float4*>(c_prev + (long)grow * H + u0 + 4));
          }
          mbar_wait(&tfull_bar[k], par);
          tc_fence_after();
          float v[32];
          tmem_ld32(tmem_base + (uint32_t)(64 * k + half * 32) + ((uint32_t)(q * 32) << 16), v);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[k]);
          if (grow < B) {
            if (bias) {
#pragma unroll
              for (int j = 0; j < 8; ++j) { v[4 * j] += pre_b[j].x; v[4 * j + 1] += pre_b[j].y; v[4 * j + 2] += pre_b[j].z; v[4 * j + 3] += pre_b[j].w; }
            }
            if (addend) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 a = pre_a[j];
                v[j * 8 + 0] += bf16_lo(a.x); v[j * 8 + 1] += bf16_hi(a.x); v[j * 8 + 2] += bf16_lo(a.y); v[j * 8 + 3] += bf16_hi(a.y);
                v[j * 8 + 4] += bf16_lo(a.z); v[j * 8 + 5] += bf16_hi(a.z); v[j * 8 + 6] += bf16_lo(a.w); v[j * 8 + 7] += bf16_hi(a.w);
              }
            }
            const float cp[8] = {pc0.x, pc0.y, pc0.z, pc0.w, pc1.x, pc1.y, pc1.z, pc1.w};
            float cn[8], hn[8];
#pragma unroll
            for (int uu = 0; uu < 8; ++uu) {
              float pre[4] = {v[4 * uu], v[4 * uu + 1], v[4 * uu + 2], v[4 * uu + 3]}, act[4];
              lstm_cell_unit<bf16>(pre, cp[uu], act, cn[uu], hn[uu]);
              v[4 * uu] = act[0]; v[4 * uu + 1] = act[1]; v[4 * uu + 2] = act[2]; v[4 * uu + 3] = act[3];
            }
            *reinterpret_cast<float4*>(c_out + (long)grow * H + u0) = make_float4(cn[0], cn[1], cn[2], cn[3]);
            *reinterpret_cast<float4*>(c_out + (long)grow * H + u0 + 4) = make_float4(cn[4], cn[5], cn[6], cn[7]);
            uint4* gp = reinterpret_cast<uint4*>(gates_out + (long)grow * N4 + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              gp[j] = make_uint4(pack_bf16(v[j * 8], v[j * 8 + 1]), pack_bf16(v[j * 8 + 2], v[j * 8 + 3]), pack_bf16(v[j * 8 + 4], v[j * 8 + 5]), pack_bf16(v[j * 8 + 6], v[j * 8 + 7]));
            const uint4 hp = make_uint4(pack_bf16(hn[0], hn[1]), pack_bf16(hn[2], hn[3]), pack_bf16(hn[4], hn[5]), pack_bf16(hn[6], hn[7]));
            *reinterpret_cast<uint4*>(h_rec + (long)grow * ldk + u0) = hp;
            if (h_top) *reinterpret_cast<uint4*>(h_top + (long)grow * H + u0) = hp;
            if (h_next) {
              uint4 hd = hp;
              if (p.drop_p > 0.f) {
                float hm[8];
#pragma unroll
                for (int uu = 0; uu < 8; ++uu) hm[uu] = hn[uu] * dropout_scale(dseed, (uint32_t)k, (uint64_t)((slot + grow) * H + u0 + uu), p.drop_p, inv_keep);
                hd = make_uint4(pack_bf16(hm[0], hm[1]), pack_bf16(hm[2], hm[3]), pack_bf16(hm[4], hm[5]), pack_bf16(hm[6], hm[7]));
              }
              *reinterpret_cast<uint4*>(h_next + (long)grow * ld_next + u0) = hd;
            }
          }
        }
        if (trace && et == 0) trace[t * 16 + 7 + (k < 3 ? k : 3)] = clock64();
        grid_arrive_epi(p.barrier, p.flags, (unsigned)(t * phases + 3 + k), et);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------ host side
inline int make_tmap_bf16_3d(CUtensorMap* map, const void* ptr, long inner, long rows, long slots, long ld, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return set_err(B2C_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)slots};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)rows * ld * 2};
  cuuint32_t box[3] = {64u, (cuuint32_t)box_rows, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(B2C_ECUDA, "cuTensorMapEncodeTiled (3-D) failed (%d): ptr=%p inner=%ld rows=%ld slots=%ld ld=%ld", (int)r, ptr, inner, rows, slots, ld);
  return 0;
}

// EXPERIMENTAL, off by default (B2C_PERSISTENT=1 enables it).  Measured on B200 at BASELINE configs[1] (profiles/README.md, round 2):
// 45 us per time step against 31 us for the per-step kernels.  The kernel boundaries it removes cost ~1.5 us each under PDL, but a
// grid-wide barrier costs ~3 us here (arrive: proxy fence + bar.sync + acq_rel atomic = 1.2 us; publish + poll + acquire fence =
// 1.7 us: three dependent L2 round trips at ~0.4 us each under load), and the TMA -> MMA -> epilogue latency of one 128 x 64 tile is
// the same inside and outside a persistent kernel.  Shortening the chain needs synchronisation that does not go through L2 at all:
// thread-block clusters exchanging h / ctx / u through distributed shared memory.
inline bool recur_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("B2C_PERSISTENT"); on = (e && e[0] == '1') ? 1 : 0; }
  return on != 0;
}

struct RecurPlan { bool ok; int grid, tiles_m, tiles_n, u_tiles_n, n_fbuf, f_resident, tmem_cols, nq; size_t smem; };

// Shapes the persistent kernel covers (everything else takes the multi-kernel path): bf16 mode, E and H multiples of 64 (the operand
// halves split on a k-block boundary), E <= 384, S <= 56, one gate tile per CTA (tiles_m * 4H/64 <= SMs).
inline RecurPlan recur_plan(const B2CShape& s) {
  RecurPlan pl{}; pl.ok = false;
  if (!recur_enabled()) return pl;
  if (s.E % 64 || s.H % 64 || s.E > 384 || s.S > 8 * ATT_MAXTOK || s.L > RC_MAXL) return pl;
  const int sms = sm_count();
  pl.tiles_m = cdiv(s.B, TC_BM); pl.tiles_n = 4 * s.H / RC_BN; pl.u_tiles_n = s.E / RC_BNU;
  const int tiles = pl.tiles_m * pl.tiles_n;
  if (tiles > sms || pl.tiles_m * pl.u_tiles_n > sms) return pl;
  const int nper = cdiv(s.B, sms), g_att = cdiv(s.B, nper);
  pl.grid = tiles > g_att ? tiles : g_att;
  const int n_mine = cdiv(s.B, pl.grid);
  const size_t fixed = (size_t)RC_STAGES * RC_STAGE_BYTES + 1024 + 512 + (size_t)RC_MAX_RES * (s.E + 64) * 4;
  const size_t fbytes = (size_t)s.S * s.E * 2;
  const size_t avail = 227 * 1024 - fixed;
  if (n_mine > RC_MAX_RES || (size_t)n_mine * fbytes > avail) return pl;       // the feature tokens of a CTA's samples stay resident
  pl.f_resident = 1; pl.n_fbuf = n_mine;
  pl.smem = fixed + (size_t)pl.n_fbuf * fbytes;
  int cols = 64 * s.L + RC_BNU; pl.tmem_cols = 32; while (pl.tmem_cols < cols) pl.tmem_cols <<= 1;
  if (pl.tmem_cols > 512) return pl;
  pl.nq = cdiv(s.E / 4, 32);
  pl.ok = true;
  return pl;
}

}  // namespace b2c
