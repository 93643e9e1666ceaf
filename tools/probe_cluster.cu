// Hardware probe for the cluster-resident recurrence kernel (csrc/recur_cluster.cuh).  Checks, on one cluster of 16 CTAs:
//   T0  how many clusters of 8 / 16 CTAs (512 threads, ~216 KB shared memory each) the device keeps resident
//   T1  a 4-row TMA box (half a SWIZZLE_128B atom) multicast into every CTA lands with the address-based swizzle pattern
//   T2  64-byte-wide K slices multicast with SWIZZLE_64B, and tcgen05.mma reading them as A (M = 128, two stacked 64-row tiles)
//       against a SWIZZLE_128B B operand, and as B (N = 64) against a SWIZZLE_128B A operand
//   T3  st.async to a peer CTA's shared memory with complete_tx on the peer's mbarrier
//   T4  cycles of one all-to-all exchange (store slice -> proxy fence -> multicast own slice -> wait for all 16 slices)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/probe_cluster.cu -o gpurun_out/probe_cluster -lcuda
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

typedef __nv_bfloat16 bf16;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 22)) { printf("probe: mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x); __trap(); }
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r; }
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
               :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" :: "r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

constexpr int CL = 16, THREADS = 512;
constexpr int R = 64, H = 512, E = 256;
// shared memory map (bytes from the 1024-aligned base)
constexpr int OFF_BH = 0;                      // 16 k-blocks x [h1 4 KB | h0 4 KB]                = 128 KB
constexpr int OFF_CTX = 128 * 1024;            // 4 k-blocks x [64 rows x 128 B] SWIZZLE_128B       =  32 KB
constexpr int OFF_W = 160 * 1024;              // 2 stages x [128 rows x 128 B] SWIZZLE_128B       =  32 KB
constexpr int OFF_WH = 192 * 1024;             // 8 k-blocks x [16 rows x 128 B]                    =  16 KB
constexpr int OFF_MISC = 208 * 1024;           // barriers, u inbox
constexpr int SMEM_BYTES = 216 * 1024 + 1024;

struct Maps { CUtensorMap ctx, h1, h0, w, wh; };
struct Args {
  bf16* ctx_g; bf16* h1_g; bf16* h0_g;        // (64, 256), (64, 512), (64, 512)
  unsigned char* dump_ctx;                      // CL x 32 KB raw image of the ctx tile
  unsigned char* dump_bh;                       // CL x 128 KB raw image
  float* d_u;                                   // CL x 128 x 16
  float* d_g;                                   // CL x 128 x 64
  float* inbox;                                 // CL x CL x 4
  long long* cycles;                            // exchange timing
  int iters;
};

__global__ void __launch_bounds__(THREADS, 1) dummy_kernel(int* p) { extern __shared__ unsigned char s[]; if (p) p[0] = s[0]; }

__global__ void __launch_bounds__(THREADS, 1) probe_kernel(const __grid_constant__ Maps maps, const Args a) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + OFF_MISC);     // [0] ctx full, [1] bh full, [2] w full, [3] mma done, [4] inbox, [5] timing
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 8);
  float* inbox = reinterpret_cast<float*>(base + OFF_MISC + 256);   // CL x 4 floats
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t c = cluster_rank();
  if (tid == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  cluster_sync_all();

  // ---------------- T1: every CTA writes its 4 ctx rows, multicasts them as 4 boxes {64 cols, 4 rows}
  for (int i = tid; i < 4 * E; i += THREADS) {
    const int r = 4 * c + i / E, e = i % E;
    a.ctx_g[(long)r * E + e] = __float2bfloat16((float)((r * 7 + e * 3) % 251) - 125.f);
  }
  asm volatile("fence.proxy.async.global;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar[0], R * E * 2);
    for (int kb = 0; kb < 4; ++kb)
      tma_load_2d_mc(base + OFF_CTX + kb * 8192 + (4 * c / 8) * 1024 + (4 * c % 8) * 128, &maps.ctx, kb * 64, 4 * c, &bar[0], 0xFFFF);
  }
  mbar_wait(&bar[0], 0);
  for (int i = tid; i < 32 * 1024 / 16; i += THREADS)
    reinterpret_cast<uint4*>(a.dump_ctx + (size_t)c * 32 * 1024)[i] = reinterpret_cast<const uint4*>(base + OFF_CTX)[i];

  // ---------------- T2: K slices of h1 / h0 (64 rows x 32 units), SWIZZLE_64B, own slice multicast to everyone
  for (int i = tid; i < R * 32; i += THREADS) {
    const int r = i / 32, j = 32 * c + i % 32;
    a.h1_g[(long)r * H + j] = __float2bfloat16((float)(((r * 13 + j * 5) % 17) - 8) * 0.125f);
    a.h0_g[(long)r * H + j] = __float2bfloat16((float)(((r * 3 + j * 11) % 19) - 9) * 0.0625f);
  }
  asm volatile("fence.proxy.async.global;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar[1], 2 * R * H * 2);
    tma_load_2d_mc(base + OFF_BH + c * 8192, &maps.h1, 32 * c, 0, &bar[1], 0xFFFF);
    tma_load_2d_mc(base + OFF_BH + c * 8192 + 4096, &maps.h0, 32 * c, 0, &bar[1], 0xFFFF);
    // resident W_h slice (rows 16c..16c+15) : 8 boxes {64, 16}
    mbar_arrive_expect_tx(&bar[2], 16 * H * 2);
    for (int kb = 0; kb < 8; ++kb) tma_load_2d(base + OFF_WH + kb * 2048, &maps.wh, kb * 64, 16 * c, &bar[2]);
  }
  mbar_wait(&bar[1], 0);
  mbar_wait(&bar[2], 0);
  for (int i = tid; i < 128 * 1024 / 16; i += THREADS)
    reinterpret_cast<uint4*>(a.dump_bh + (size_t)c * 128 * 1024)[i] = reinterpret_cast<const uint4*>(base + OFF_BH)[i];
  __syncthreads();
  // (a) u-like: D[128 x 16] = [h1 ; h0] (A, SWIZZLE_64B, M = 128) x Wh_slice^T (B, SWIZZLE_128B, N = 16), K = 512
  if (tid == 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int kk = 0; kk < 32; ++kk) {
      const uint64_t ad = make_desc(smem_u32(base + OFF_BH) + (kk >> 1) * 8192 + (kk & 1) * 32, 512, 4);
      const uint64_t bd = make_desc(smem_u32(base + OFF_WH) + (kk >> 2) * 2048 + (kk & 3) * 32, 1024, 2);
      mma_bf16(tmem + 128, ad, bd, idesc, kk > 0);
    }
    tc_commit(&bar[3]);
  }
  mbar_wait(&bar[3], 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp >= 4 && warp < 8) {
    float v[16];
    tmem_ld16(tmem + 128 + ((uint32_t)((warp & 3) * 32) << 16), v);
    for (int j = 0; j < 16; ++j) a.d_u[((size_t)c * 128 + (warp & 3) * 32 + lane) * 16 + j] = v[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  // (b) gate-like: D[128 x 64] = W (A, SWIZZLE_128B, 128 gate rows c*128..) x h1^T (B = SWIZZLE_64B slices, N = 64), K = 512
  if (tid == 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t wph = 1;           // bar[2] already completed phase 0
    for (int kb = 0; kb < 8; ++kb) {
      // single-stage for simplicity: load, wait, 4 MMAs, commit, wait
      mbar_arrive_expect_tx(&bar[2], 128 * 64 * 2);
      tma_load_2d(base + OFF_W, &maps.w, kb * 64, 128 * c, &bar[2]);
      mbar_wait(&bar[2], wph); wph ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int k = 0; k < 4; ++k) {
        const int kk = kb * 4 + k;
        const uint64_t ad = make_desc(smem_u32(base + OFF_W) + k * 32, 1024, 2);
        const uint64_t bd = make_desc(smem_u32(base + OFF_BH) + (kk >> 1) * 8192 + (kk & 1) * 32, 512, 4);
        mma_bf16(tmem, ad, bd, idesc, kk > 0);
      }
      tc_commit(&bar[5]);
      mbar_wait(&bar[5], kb & 1);
    }
    tc_commit(&bar[3]);
  }
  mbar_wait(&bar[3], 1);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp >= 4 && warp < 8) {
    for (int cb = 0; cb < 4; ++cb) {
      float v[16];
      tmem_ld16(tmem + cb * 16 + ((uint32_t)((warp & 3) * 32) << 16), v);
      for (int j = 0; j < 16; ++j) a.d_g[((size_t)c * 128 + (warp & 3) * 32 + lane) * 64 + cb * 16 + j] = v[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();

  // ---------------- T3: st.async 16 bytes to every peer's inbox slot [c], complete_tx on the peer's bar[4]
  if (tid == 0) mbar_arrive_expect_tx(&bar[4], CL * 16);
  cluster_sync_all();          // every inbox barrier is armed (not required by the tx protocol, keeps the test simple)
  if (tid < CL) {
    const uint32_t dst = mapa(smem_u32(inbox + 4 * c), tid), rbar = mapa(smem_u32(&bar[4]), tid);
    const float f0 = (float)(c * 100 + tid), f1 = f0 + 0.25f, f2 = f0 + 0.5f, f3 = f0 + 0.75f;
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
                 :: "r"(dst), "f"(f0), "f"(f1), "f"(f2), "f"(f3), "r"(rbar) : "memory");
  }
  mbar_wait(&bar[4], 0);
  for (int i = tid; i < CL * 4; i += THREADS) a.inbox[(size_t)c * CL * 4 + i] = inbox[i];

  // ---------------- T4: exchange timing.  Each iteration: 128 threads store the CTA's h slice, proxy fence, named barrier,
  // one thread multicasts the slice, everyone waits for all 16 slices.
  cluster_sync_all();
  long long t0 = 0;
  if (tid == 0) t0 = clock64();
  uint32_t ph = 1;
  for (int it = 0; it < a.iters; ++it) {
    if (warp >= 4 && warp < 8) {
      const int w = warp - 4;
      for (int s = 0; s < 16; ++s) a.h1_g[(long)(16 * w + s) * H + 32 * c + lane] = __float2bfloat16((float)(it & 7));
      asm volatile("fence.proxy.async.global;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (tid == 128) {
        mbar_arrive_expect_tx(&bar[1], R * H * 2);
        tma_load_2d_mc(base + OFF_BH + c * 8192, &maps.h1, 32 * c, 0, &bar[1], 0xFFFF);
      }
    }
    mbar_wait(&bar[1], ph); ph ^= 1;
    // WAR guard of this synthetic loop only: nobody may overwrite a slice a slower peer has not yet observed as complete
    cluster_sync_all();
  }
  if (tid == 0) a.cycles[c] = clock64() - t0;
  // cluster barrier alone, for the subtraction
  cluster_sync_all();
  if (tid == 0) t0 = clock64();
  for (int it = 0; it < a.iters; ++it) cluster_sync_all();
  if (tid == 0) a.cycles[CL + c] = clock64() - t0;

  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256) : "memory");
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled enc_fn() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  return (PFN_encodeTiled)fn;
}
static void make_map(CUtensorMap* m, void* ptr, long inner, long outer, long ld, int box_in, int box_out, CUtensorMapSwizzle sw) {
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer}; cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_in, (cuuint32_t)box_out}; cuuint32_t estr[2] = {1, 1};
  CUresult r = enc_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}
static float bf(bf16 v) { return __bfloat162float(v); }

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs\n", prop.name, prop.multiProcessorCount);
  // ---- T0
  for (int cl : {8, 16}) {
    CK(cudaFuncSetAttribute(dummy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    CK(cudaFuncSetAttribute(dummy_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(cl * 8); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM_BYTES;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy_kernel, &cfg);
    printf("T0 cluster size %2d, %d threads, %d B smem: max active clusters = %d (%s)\n", cl, THREADS, SMEM_BYTES, n, cudaGetErrorString(e));
  }
  // ---- buffers
  bf16 *ctx_g, *h1_g, *h0_g, *w_g, *wh_g; unsigned char *dump_ctx, *dump_bh; float *d_u, *d_g, *inbox; long long* cycles;
  CK(cudaMalloc(&ctx_g, R * E * 2)); CK(cudaMalloc(&h1_g, R * H * 2)); CK(cudaMalloc(&h0_g, R * H * 2));
  CK(cudaMalloc(&w_g, 2048 * H * 2)); CK(cudaMalloc(&wh_g, E * H * 2));
  CK(cudaMalloc(&dump_ctx, CL * 32 * 1024)); CK(cudaMalloc(&dump_bh, CL * 128 * 1024));
  CK(cudaMalloc(&d_u, CL * 128 * 16 * 4)); CK(cudaMalloc(&d_g, CL * 128 * 64 * 4)); CK(cudaMalloc(&inbox, CL * CL * 4 * 4)); CK(cudaMalloc(&cycles, 2 * CL * 8));
  std::vector<bf16> w_h(2048 * H), wh_h(E * H);
  for (int i = 0; i < 2048 * H; ++i) w_h[i] = __float2bfloat16((float)(((i * 7 + (i / H) * 3) % 23) - 11) * 0.03125f);
  for (int i = 0; i < E * H; ++i) wh_h[i] = __float2bfloat16((float)(((i * 5 + (i / H)) % 13) - 6) * 0.0625f);
  CK(cudaMemcpy(w_g, w_h.data(), w_h.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(wh_g, wh_h.data(), wh_h.size() * 2, cudaMemcpyHostToDevice));
  Maps maps;
  make_map(&maps.ctx, ctx_g, E, R, E, 64, 4, CU_TENSOR_MAP_SWIZZLE_128B);
  make_map(&maps.h1, h1_g, H, R, H, 32, 64, CU_TENSOR_MAP_SWIZZLE_64B);
  make_map(&maps.h0, h0_g, H, R, H, 32, 64, CU_TENSOR_MAP_SWIZZLE_64B);
  make_map(&maps.w, w_g, H, 2048, H, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
  make_map(&maps.wh, wh_g, H, E, H, 64, 16, CU_TENSOR_MAP_SWIZZLE_128B);
  Args a{ctx_g, h1_g, h0_g, dump_ctx, dump_bh, d_u, d_g, inbox, cycles, 200};
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(CL); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM_BYTES;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, probe_kernel, maps, a));
  CK(cudaDeviceSynchronize());

  // ---- T1 check (the timing loop overwrote h1_g, so read the dumps taken before it)
  std::vector<bf16> ctx_h(R * E); CK(cudaMemcpy(ctx_h.data(), ctx_g, R * E * 2, cudaMemcpyDeviceToHost));
  std::vector<unsigned char> dc(CL * 32 * 1024); CK(cudaMemcpy(dc.data(), dump_ctx, dc.size(), cudaMemcpyDeviceToHost));
  long bad1 = 0;
  for (int cc = 0; cc < CL; ++cc)
    for (int r = 0; r < R; ++r)
      for (int e = 0; e < E; ++e) {
        const int kb = e / 64, ke = e % 64, chunk = (ke * 2) / 16, within = (ke * 2) % 16;
        const size_t off = (size_t)cc * 32768 + kb * 8192 + (r / 8) * 1024 + (r % 8) * 128 + ((chunk ^ (r % 8)) * 16) + within;
        const bf16 got = *reinterpret_cast<const bf16*>(&dc[off]);
        if (bf(got) != bf(ctx_h[r * E + e])) ++bad1;
      }
  printf("T1 partial-atom SWIZZLE_128B multicast (4-row boxes): %ld mismatches of %d\n", bad1, CL * R * E);
  // ---- T2 checks
  std::vector<bf16> h1_h(R * H), h0_h(R * H);
  for (int r = 0; r < R; ++r) for (int j = 0; j < H; ++j) {
    h1_h[r * H + j] = __float2bfloat16((float)(((r * 13 + j * 5) % 17) - 8) * 0.125f);
    h0_h[r * H + j] = __float2bfloat16((float)(((r * 3 + j * 11) % 19) - 9) * 0.0625f);
  }
  std::vector<unsigned char> db(CL * 128 * 1024); CK(cudaMemcpy(db.data(), dump_bh, db.size(), cudaMemcpyDeviceToHost));
  long bad2 = 0;
  for (int cc = 0; cc < CL; ++cc)
    for (int r = 0; r < R; ++r)
      for (int j = 0; j < H; ++j) {
        const int kb = j / 32, ke = j % 32, chunk = (ke * 2) / 16, within = (ke * 2) % 16;
        const size_t off = (size_t)cc * 131072 + kb * 8192 + (r / 8) * 512 + (r % 8) * 64 + ((chunk ^ ((r >> 1) & 3)) * 16) + within;
        if (bf(*reinterpret_cast<const bf16*>(&db[off])) != bf(h1_h[r * H + j])) ++bad2;
        if (bf(*reinterpret_cast<const bf16*>(&db[off + 4096])) != bf(h0_h[r * H + j])) ++bad2;
      }
  printf("T2 SWIZZLE_64B K-slice multicast image: %ld mismatches of %d\n", bad2, 2 * CL * R * H);
  std::vector<float> du(CL * 128 * 16), dg(CL * 128 * 64);
  CK(cudaMemcpy(du.data(), d_u, du.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(dg.data(), d_g, dg.size() * 4, cudaMemcpyDeviceToHost));
  double eu = 0, eg = 0;
  for (int cc = 0; cc < CL; ++cc) {
    for (int r = 0; r < 128; ++r) for (int n = 0; n < 16; ++n) {
      double s = 0; const bf16* hv = r < 64 ? &h1_h[r * H] : &h0_h[(r - 64) * H];
      for (int k = 0; k < H; ++k) s += (double)bf(hv[k]) * bf(wh_h[(16 * cc + n) * H + k]);
      eu = fmax(eu, fabs(s - du[(cc * 128 + r) * 16 + n]));
    }
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
      double s = 0;
      for (int k = 0; k < H; ++k) s += (double)bf(w_h[(size_t)(128 * cc + m) * H + k]) * bf(h1_h[n * H + k]);
      eg = fmax(eg, fabs(s - dg[(cc * 128 + m) * 64 + n]));
    }
  }
  printf("T2a MMA A=[h1;h0] SWIZZLE_64B (M=128) x B=Wh SWIZZLE_128B (N=16): max abs err %.3e\n", eu);
  printf("T2b MMA A=W SWIZZLE_128B (M=128) x B=h1 SWIZZLE_64B (N=64): max abs err %.3e\n", eg);
  // ---- T3
  std::vector<float> ib(CL * CL * 4); CK(cudaMemcpy(ib.data(), inbox, ib.size() * 4, cudaMemcpyDeviceToHost));
  long bad3 = 0;
  for (int dst = 0; dst < CL; ++dst) for (int src = 0; src < CL; ++src) for (int q = 0; q < 4; ++q)
    if (ib[(dst * CL + src) * 4 + q] != (float)(src * 100 + dst) + 0.25f * q) ++bad3;
  printf("T3 st.async + complete_tx to peers: %ld mismatches of %d\n", bad3, CL * CL * 4);
  // ---- T4
  std::vector<long long> cy(2 * CL); CK(cudaMemcpy(cy.data(), cycles, cy.size() * 8, cudaMemcpyDeviceToHost));
  printf("T4 exchange + cluster barrier: %.0f cycles / iteration; cluster barrier alone: %.0f cycles  (CTA 0; %d iterations)\n",
         (double)cy[0] / a.iters, (double)cy[CL] / a.iters, a.iters);
  return 0;
}
