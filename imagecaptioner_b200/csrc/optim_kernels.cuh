// Optimizer side of the KD step on the flat fp32 buffers:
//   GradScaler.unscale_ + clip_grad_norm_ (one norm per clip group) + AdamW (per-segment lr / weight decay) + GradScaler.update
// reference: src/train_student_kd.py:230-236 (three LR groups, wd 0.01), :290-303 (unscale_, two clip_grad_norm_ calls, step, update).
// Two launches: (1) per-block partial squared norms per clip group (+ a non-finite flag), fixed order, no atomics;
//               (2) every block re-reduces the partials (a few KB from L2), then streams p/g/m/v once with 128-bit accesses.
// HBM-bound: 4 reads + 3 writes of 4 bytes per parameter in (2), 1 read in (1).
#pragma once
#include "common.cuh"

namespace b2c {

constexpr int OPT_THREADS = 256;
constexpr int OPT_MAX_SEG = B2C_OPT_MAX_SEG;
constexpr int OPT_MAX_CLIP = B2C_OPT_MAX_CLIP;
constexpr int OPT_NPART = OPT_MAX_CLIP + 1;          // per-block partials: clip-group sums + non-finite count

struct OptSegs {
  long begin[OPT_MAX_SEG], end[OPT_MAX_SEG];
  int lr_index[OPT_MAX_SEG], clip_group[OPT_MAX_SEG];
  float weight_decay[OPT_MAX_SEG];
  int nseg;
};

// partials[block][OPT_NPART]; sums are of the RAW (still loss-scaled) gradients, the consumer applies inv_scale^2.
__global__ void __launch_bounds__(OPT_THREADS)
grad_sqnorm_kernel(const float* __restrict__ g, OptSegs segs, float* __restrict__ partials) {
  float acc[OPT_NPART];
#pragma unroll
  for (int i = 0; i < OPT_NPART; ++i) acc[i] = 0.f;
  for (int s = 0; s < segs.nseg; ++s) {
    const long b4 = segs.begin[s] >> 2, n4 = (segs.end[s] - segs.begin[s]) >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g) + b4;
    float a = 0.f, bad = 0.f;
    for (long i = (long)blockIdx.x * OPT_THREADS + threadIdx.x; i < n4; i += (long)gridDim.x * OPT_THREADS) {
      const float4 v = g4[i];
      a = fmaf(v.x, v.x, a); a = fmaf(v.y, v.y, a); a = fmaf(v.z, v.z, a); a = fmaf(v.w, v.w, a);
    }
    if (blockIdx.x == 0) {                                   // ragged tail (< 4 elements) of the segment
      const long t0 = segs.begin[s] + (n4 << 2);
      for (long i = t0 + threadIdx.x; i < segs.end[s]; i += OPT_THREADS) a = fmaf(g[i], g[i], a);
    }
    if (!(fabsf(a) <= 3.402823466e38f)) bad = 1.f;           // inf or nan anywhere in this thread's share shows up in its sum
    const int cg = segs.clip_group[s];
#pragma unroll
    for (int i = 0; i < OPT_MAX_CLIP; ++i) if (cg == i) acc[i] += a;
    acc[OPT_MAX_CLIP] += bad;
  }
  __shared__ float red[OPT_THREADS / 32][OPT_NPART];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < OPT_NPART; ++i) {
    const float v = warp_sum(acc[i]);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < OPT_NPART) {
    float v = 0.f;
    for (int w = 0; w < OPT_THREADS / 32; ++w) v += red[w][threadIdx.x];
    partials[(long)blockIdx.x * OPT_NPART + threadIdx.x] = v;
  }
}

struct OptHyper {
  double beta1d, beta2d;                      // for the bias corrections 1 - beta^t (fp64, one thread per block)
  float beta2, omb1, omb2, eps, max_norm;     // omb = (float)(1 - beta) formed in fp64 on the host, as torch does
  float grad_scale;                           // applied to every gradient before anything else (1/world after a SUM all-reduce)
  float growth_factor, backoff_factor;
  int growth_interval;
};

// stats (device, OPT_MAX_CLIP + 2 floats): [0..OPT_MAX_CLIP) = unscaled pre-clip gradient norm of each clip group,
// [OPT_MAX_CLIP] = 1 if a non-finite gradient was found (the step was skipped), [OPT_MAX_CLIP+1] = loss scale used.
__global__ void __launch_bounds__(OPT_THREADS)
clip_adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, OptSegs segs,
                  OptHyper hp, const float* __restrict__ lr, int* __restrict__ step, float* __restrict__ loss_scale,
                  int* __restrict__ growth_tracker, const float* __restrict__ partials, int n_part_blocks,
                  float* __restrict__ stats, unsigned int* __restrict__ ticket) {
  __shared__ float red[OPT_THREADS / 32][OPT_NPART];
  __shared__ float coef_s[OPT_MAX_CLIP], norm_s[OPT_MAX_CLIP];
  __shared__ float bc1_s, bc2s_s, bad_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[OPT_NPART];
#pragma unroll
  for (int i = 0; i < OPT_NPART; ++i) acc[i] = 0.f;
  for (int b = threadIdx.x; b < n_part_blocks; b += OPT_THREADS) {
#pragma unroll
    for (int i = 0; i < OPT_NPART; ++i) acc[i] += partials[(long)b * OPT_NPART + i];
  }
#pragma unroll
  for (int i = 0; i < OPT_NPART; ++i) {
    const float s = warp_sum(acc[i]);
    if (lane == 0) red[warp][i] = s;
  }
  __syncthreads();
  const float scale = loss_scale ? *loss_scale : 1.0f;
  const float inv_scale = hp.grad_scale / scale;
  const int t = *step + 1;
  if (threadIdx.x == 0) {
    float tot[OPT_NPART];
    for (int i = 0; i < OPT_NPART; ++i) { float s = 0.f; for (int w = 0; w < OPT_THREADS / 32; ++w) s += red[w][i]; tot[i] = s; }
    float bad = tot[OPT_MAX_CLIP];
    for (int i = 0; i < OPT_MAX_CLIP; ++i) {
      const float nrm = sqrtf(tot[i]) * inv_scale;
      if (!(nrm <= 3.402823466e38f)) bad = 1.f;
      norm_s[i] = nrm;
      // torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1
      coef_s[i] = hp.max_norm > 0.f ? fminf(hp.max_norm / (nrm + 1e-6f), 1.0f) : 1.0f;
    }
    bad_s = bad;
    bc1_s = (float)(1.0 - pow(hp.beta1d, (double)t));
    bc2s_s = (float)sqrt(1.0 - pow(hp.beta2d, (double)t));
  }
  __syncthreads();
  const bool skip = bad_s != 0.f;
  if (!skip) {
    const float bc1 = bc1_s, bc2s = bc2s_s;
    for (int s = 0; s < segs.nseg; ++s) {
      const float lr_s = lr[segs.lr_index[s]];
      const float decay = 1.0f - lr_s * segs.weight_decay[s];
      const float gmul = inv_scale * (segs.clip_group[s] >= 0 ? coef_s[segs.clip_group[s]] : 1.0f);
      const float step_size = lr_s / bc1;
      auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        gg *= gmul;
        pp *= decay;
        mm = fmaf(gg - mm, hp.omb1, mm);                    // lerp, as torch's fused AdamW
        vv = fmaf(vv, hp.beta2, hp.omb2 * gg * gg);
        const float denom = sqrtf(vv) / bc2s + hp.eps;
        pp -= step_size * (mm / denom);
      };
      const long b4 = segs.begin[s] >> 2, n4 = (segs.end[s] - segs.begin[s]) >> 2;
      float4* p4 = reinterpret_cast<float4*>(p) + b4;
      const float4* g4 = reinterpret_cast<const float4*>(g) + b4;
      float4* m4 = reinterpret_cast<float4*>(m) + b4;
      float4* v4 = reinterpret_cast<float4*>(v) + b4;
      for (long i = (long)blockIdx.x * OPT_THREADS + threadIdx.x; i < n4; i += (long)gridDim.x * OPT_THREADS) {
        float4 pp = p4[i]; const float4 gg = g4[i]; float4 mm = m4[i]; float4 vv = v4[i];
        upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
        p4[i] = pp; m4[i] = mm; v4[i] = vv;
      }
      if (blockIdx.x == 0) {
        const long t0 = segs.begin[s] + (n4 << 2);
        for (long i = t0 + threadIdx.x; i < segs.end[s]; i += OPT_THREADS) upd(p[i], g[i], m[i], v[i]);
      }
    }
  }
  // the last block to finish publishes the step count, the statistics and GradScaler.update()'s new scale
  __syncthreads();
  __shared__ unsigned int last_s;
  if (threadIdx.x == 0) { __threadfence(); last_s = atomicAdd(ticket, 1u); }
  __syncthreads();
  if (last_s == gridDim.x - 1 && threadIdx.x == 0) {
    *ticket = 0u;
    for (int i = 0; i < OPT_MAX_CLIP; ++i) stats[i] = norm_s[i];
    stats[OPT_MAX_CLIP] = skip ? 1.f : 0.f;
    stats[OPT_MAX_CLIP + 1] = scale;
    if (!skip) *step = t;
    if (loss_scale) {                                          // torch.amp.GradScaler.update()
      int gt = growth_tracker ? *growth_tracker : 0;
      if (skip) { *loss_scale = scale * hp.backoff_factor; gt = 0; }
      else if (++gt >= hp.growth_interval && hp.growth_interval > 0) { *loss_scale = scale * hp.growth_factor; gt = 0; }
      if (growth_tracker) *growth_tracker = gt;
    }
  }
}

}  // namespace b2c
