// Per-voxel row kernels of the encoder layer (VoxFormerLayer, encoder.py:262-340): one warp per voxel row of C = 32*CPL
// channels, float4 accesses, shuffle reductions.  The dense GEMMs between them run on the tensor cores; these kernels
// carry everything else of the chain so that a level is a short, fixed sequence of launches.
//
//   sgc_layernorm_bwd          gx = rstd * (g*gamma - mean_c(g*gamma) - xhat * mean_c(g*gamma*xhat))   (nn.LayerNorm backward,
//                              encoder.py:325-338 norms) + per-CTA partial sums of (g*xhat, g) for gamma / beta
//   sgc_layernorm_bwd_params   fixed-order reduction of those partials (deterministic; runs on the weight-gradient stream)
#include <cuda_bf16.h>
#include "common.cuh"
#include "../../include/sgcdet_b200.h"

namespace sgc {

constexpr int kRowWarps = 8;          // warps (rows in flight) per CTA
constexpr int kRowMaxBlocks = 148 * 2;

__host__ __device__ inline int rowop_blocks(int R) {
  const int b = (R + kRowWarps - 1) / kRowWarps;
  return b < kRowMaxBlocks ? (b > 0 ? b : 1) : kRowMaxBlocks;
}

template <int CPL>
__global__ void __launch_bounds__(kRowWarps * 32) layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                                       const float* __restrict__ mean,
                                                                       const float* __restrict__ rstd,
                                                                       const float* __restrict__ gamma, int R,
                                                                       float* __restrict__ gx, float* __restrict__ partial) {
  constexpr int C = 32 * CPL;
  __shared__ float s_part[kRowWarps][2 * C];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = lane * CPL;
  float gam[CPL], agam[CPL], abet[CPL];
#pragma unroll
  for (int i = 0; i < CPL; i += 4) {
    const float4 g4 = ldg4(gamma + c0 + i);
    gam[i] = g4.x; gam[i + 1] = g4.y; gam[i + 2] = g4.z; gam[i + 3] = g4.w;
  }
#pragma unroll
  for (int i = 0; i < CPL; ++i) { agam[i] = 0.f; abet[i] = 0.f; }
  for (int r = blockIdx.x * kRowWarps + warp; r < R; r += gridDim.x * kRowWarps) {
    const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
    float xh[CPL], g[CPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; i += 4) {
      const float4 xv = ldg4(x + (size_t)r * C + c0 + i);
      const float4 gv = ldg4(gy + (size_t)r * C + c0 + i);
      xh[i] = (xv.x - mu) * rs; xh[i + 1] = (xv.y - mu) * rs; xh[i + 2] = (xv.z - mu) * rs; xh[i + 3] = (xv.w - mu) * rs;
      g[i] = gv.x; g[i + 1] = gv.y; g[i + 2] = gv.z; g[i + 3] = gv.w;
    }
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      agam[i] += g[i] * xh[i];
      abet[i] += g[i];
      g[i] *= gam[i];
      s1 += g[i];
      s2 += g[i] * xh[i];
    }
    s1 = warp_sum(s1) * (1.f / C);
    s2 = warp_sum(s2) * (1.f / C);
#pragma unroll
    for (int i = 0; i < CPL; i += 4) {
      float4 o;
      o.x = rs * (g[i] - s1 - xh[i] * s2);
      o.y = rs * (g[i + 1] - s1 - xh[i + 1] * s2);
      o.z = rs * (g[i + 2] - s1 - xh[i + 2] * s2);
      o.w = rs * (g[i + 3] - s1 - xh[i + 3] * s2);
      *reinterpret_cast<float4*>(gx + (size_t)r * C + c0 + i) = o;
    }
  }
#pragma unroll
  for (int i = 0; i < CPL; ++i) { s_part[warp][c0 + i] = agam[i]; s_part[warp][C + c0 + i] = abet[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kRowWarps; ++w) t += s_part[w][c];
    partial[(size_t)blockIdx.x * 2 * C + c] = t;
  }
}

// out[c] = sum_b partial[b][c] in a fixed order: 32 columns x 8 block-lanes per CTA
__global__ void __launch_bounds__(256) layernorm_bwd_params_kernel(const float* __restrict__ partial, int nblocks, int C2,
                                                                   float* __restrict__ ggamma, float* __restrict__ gbeta) {
  __shared__ float s[8][33];
  const int cx = threadIdx.x & 31, sy = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float t = 0.f;
  if (c < C2)
    for (int b = sy; b < nblocks; b += 8) t += __ldg(partial + (size_t)b * C2 + c);
  s[sy][cx] = t;
  __syncthreads();
  if (sy == 0 && c < C2) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) a += s[k][cx];
    const int C = C2 / 2;
    if (c < C) ggamma[c] = a; else gbeta[c - C] = a;
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Fused row epilogue / prologue around the tensor-core GEMMs of the layer.  N = 32*CPL channels, lane owns channels
// [lane*CPL, lane*CPL + CPL).
//
// forward:   v = x + bias;  relu;  v *= mask*mscale;  v *= rowscale[r];  v += residual;  [pre = v; v = LayerNorm(v)]
//            -> y (fp32), ysplit (bf16x3 operand image of y for the next GEMM), pre / mean / rstd for the LN backward
// backward:  v = g (+ g2);  [LN backward with pre/mean/rstd/gamma, partial sums for gamma/beta];  gpre = v;
//            v *= mask*mscale;  v = gate > 0 ? v*gscale : 0;  v *= rowscale[r]   -> gx (fp32), gxsplit (bf16x3)
template <int CPL>
__device__ __forceinline__ void load_row(float (&dst)[CPL], const float* p) {
#pragma unroll
  for (int j = 0; j < CPL; j += 4) {
    const float4 t = ldg4(p + j);
    dst[j] = t.x; dst[j + 1] = t.y; dst[j + 2] = t.z; dst[j + 3] = t.w;
  }
}
template <int CPL>
__device__ __forceinline__ void store_row(float* p, const float (&src)[CPL]) {
#pragma unroll
  for (int j = 0; j < CPL; j += 4) *reinterpret_cast<float4*>(p + j) = make_float4(src[j], src[j + 1], src[j + 2], src[j + 3]);
}
template <int CPL>
__device__ __forceinline__ void load_mask(float (&dst)[CPL], const uint8_t* p) {
#pragma unroll
  for (int j = 0; j < CPL; j += 4) {
    const uchar4 t = *reinterpret_cast<const uchar4*>(p + j);
    dst[j] = t.x ? 1.f : 0.f; dst[j + 1] = t.y ? 1.f : 0.f; dst[j + 2] = t.z ? 1.f : 0.f; dst[j + 3] = t.w ? 1.f : 0.f;
  }
}
// bf16x3 image (pattern 0: hi | lo | hi).  heads == 0: out[r][slot*N + c];  heads == H: out[(r*H + h)][slot*dh + d]
template <int CPL>
__device__ __forceinline__ void store_split(__nv_bfloat16* __restrict__ out, size_t r, int lane, int heads, const float (&v)[CPL]) {
  constexpr int N = CPL * 32;
  __align__(16) __nv_bfloat16 h[CPL], l[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    h[j] = __float2bfloat16_rn(v[j]);
    l[j] = __float2bfloat16_rn(v[j] - __bfloat162float(h[j]));
  }
  __nv_bfloat16* o;
  int slot;
  if (heads == 0) {
    o = out + r * 3 * N + lane * CPL;
    slot = N;
  } else {
    const int dh = N / heads, c = lane * CPL, hd = c / dh;
    o = out + (r * heads + hd) * 3 * dh + (c - hd * dh);
    slot = dh;
  }
#pragma unroll
  for (int j = 0; j < CPL; j += 4) {
    const uint2 hv = *reinterpret_cast<const uint2*>(h + j), lv = *reinterpret_cast<const uint2*>(l + j);
    *reinterpret_cast<uint2*>(o + j) = hv;
    *reinterpret_cast<uint2*>(o + slot + j) = lv;
    *reinterpret_cast<uint2*>(o + 2 * slot + j) = hv;
  }
}

template <int CPL>
__global__ void __launch_bounds__(kRowWarps * 32) rowop_fwd_kernel(const sgc_rowop_fwd_args a) {
  constexpr int N = 32 * CPL;
  pdl_sync();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = lane * CPL;
  float bias[CPL], gam[CPL], bet[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) { bias[j] = 0.f; gam[j] = 1.f; bet[j] = 0.f; }
  if (a.bias) load_row<CPL>(bias, a.bias + c0);
  if (a.gamma) { load_row<CPL>(gam, a.gamma + c0); load_row<CPL>(bet, a.beta + c0); }
  for (int r = blockIdx.x * kRowWarps + warp; r < a.R; r += gridDim.x * kRowWarps) {
    float v[CPL];
    if (a.in_heads) {  // x is [H, R, dh]
      const int dh = N / a.in_heads, hd = c0 / dh;
      load_row<CPL>(v, a.x + ((size_t)hd * a.R + r) * dh + (c0 - hd * dh));
    } else {
      load_row<CPL>(v, a.x + (size_t)r * N + c0);
    }
#pragma unroll
    for (int j = 0; j < CPL; ++j) v[j] += bias[j];
    if (a.relu) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (a.mask) {
      float m[CPL];
      load_mask<CPL>(m, a.mask + (size_t)r * N + c0);
#pragma unroll
      for (int j = 0; j < CPL; ++j) v[j] *= m[j] * a.mscale;
    }
    if (a.rowscale) {
      const float rsc = __ldg(a.rowscale + r);
#pragma unroll
      for (int j = 0; j < CPL; ++j) v[j] *= rsc;
    }
    if (a.rowcount) {  // rows no view sees are zeroed (DCA:819-835): scale 1 / 0 from the per-voxel view count
      const float rsc = __ldg(a.rowcount + r) > 0 ? 1.f : 0.f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) v[j] *= rsc;
    }
    if (a.residual) {
      float t[CPL];
      load_row<CPL>(t, a.residual + (size_t)r * N + c0);
#pragma unroll
      for (int j = 0; j < CPL; ++j) v[j] += t[j];
    }
    if (a.gamma) {
      if (a.pre) store_row<CPL>(a.pre + (size_t)r * N + c0, v);
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) s += v[j];
      const float mu = warp_sum(s) * (1.f / N);
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) { const float d = v[j] - mu; q += d * d; }
      const float rs = rsqrtf(warp_sum(q) * (1.f / N) + a.eps);
      if (lane == 0) { a.mean[r] = mu; a.rstd[r] = rs; }
#pragma unroll
      for (int j = 0; j < CPL; ++j) v[j] = (v[j] - mu) * rs * gam[j] + bet[j];
    }
    if (a.y) store_row<CPL>(a.y + (size_t)r * N + c0, v);
    if (a.ysplit) store_split<CPL>(reinterpret_cast<__nv_bfloat16*>(a.ysplit), (size_t)r, lane, a.split_heads, v);
  }
}

template <int CPL>
__global__ void __launch_bounds__(kRowWarps * 32) rowop_bwd_kernel(const sgc_rowop_bwd_args a) {
  constexpr int N = 32 * CPL;
  __shared__ float s_part[kRowWarps][2 * N];
  pdl_sync();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = lane * CPL;
  const bool ln = a.gamma != nullptr;
  float gam[CPL], agam[CPL], abet[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) { gam[j] = 1.f; agam[j] = 0.f; abet[j] = 0.f; }
  if (ln) load_row<CPL>(gam, a.gamma + c0);
  for (int r = blockIdx.x * kRowWarps + warp; r < a.R; r += gridDim.x * kRowWarps) {
    float v[CPL];
    if (a.in_heads) {
      const int dh = N / a.in_heads, hd = c0 / dh;
      load_row<CPL>(v, a.g + ((size_t)hd * a.R + r) * dh + (c0 - hd * dh));
    } else {
      load_row<CPL>(v, a.g + (size_t)r * N + c0);
    }
    if (a.g2) {
      float t[CPL];
      load_row<CPL>(t, a.g2 + (size_t)r * N + c0);
#pragma unroll
      for (int j = 0; j < CPL; ++j) v[j] += t[j];
    }
    if (ln) {
      const float mu = __ldg(a.mean + r), rs = __ldg(a.rstd + r);
      float xh[CPL];
      load_row<CPL>(xh, a.pre + (size_t)r * N + c0);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        xh[j] = (xh[j] - mu) * rs;
        agam[j] += v[j] * xh[j];
        abet[j] += v[j];
        v[j] *= gam[j];
        s1 += v[j];
        s2 += v[j] * xh[j];
      }
      s1 = warp_sum(s1) * (1.f / N);
      s2 = warp_sum(s2) * (1.f / N);
#pragma unroll
      for (int j = 0; j < CPL; ++j) v[j] = rs * (v[j] - s1 - xh[j] * s2);
    }
    if (a.gpre) store_row<CPL>(a.gpre + (size_t)r * N + c0, v);
    if (a.mask) {
      float m[CPL];
      load_mask<CPL>(m, a.mask + (size_t)r * N + c0);
#pragma unroll
      for (int j = 0; j < CPL; ++j) v[j] *= m[j] * a.mscale;
    }
    if (a.gate) {
      float t[CPL];
      load_row<CPL>(t, a.gate + (size_t)r * N + c0);
#pragma unroll
      for (int j = 0; j < CPL; ++j) v[j] = t[j] > 0.f ? v[j] * a.gscale : 0.f;
    }
    if (a.rowscale) {
      const float rsc = __ldg(a.rowscale + r);
#pragma unroll
      for (int j = 0; j < CPL; ++j) v[j] *= rsc;
    }
    if (a.rowcount) {  // rows no view sees are zeroed (DCA:819-835): scale 1 / 0 from the per-voxel view count
      const float rsc = __ldg(a.rowcount + r) > 0 ? 1.f : 0.f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) v[j] *= rsc;
    }
    if (a.gx) store_row<CPL>(a.gx + (size_t)r * N + c0, v);
    if (a.gxsplit) store_split<CPL>(reinterpret_cast<__nv_bfloat16*>(a.gxsplit), (size_t)r, lane, a.split_heads, v);
  }
  if (ln) {
#pragma unroll
    for (int j = 0; j < CPL; ++j) { s_part[warp][c0 + j] = agam[j]; s_part[warp][N + c0 + j] = abet[j]; }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * N; c += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kRowWarps; ++w) t += s_part[w][c];
      a.partial[(size_t)blockIdx.x * 2 * N + c] = t;
    }
  }
}

}  // namespace sgc

extern "C" int sgc_layernorm_bwd_scratch_floats(int R, int C) { return sgc::rowop_blocks(R) * 2 * C; }

extern "C" int sgc_layernorm_bwd(const float* x, const float* gy, const float* mean, const float* rstd, const float* gamma,
                                 int R, int C, float* gx, float* partial, void* stream) {
  if (R <= 0) return 0;
  const int blocks = sgc::rowop_blocks(R);
  cudaStream_t st = (cudaStream_t)stream;
  switch (C) {
    case 128: sgc::layernorm_bwd_kernel<4><<<blocks, sgc::kRowWarps * 32, 0, st>>>(x, gy, mean, rstd, gamma, R, gx, partial); break;
    case 256: sgc::layernorm_bwd_kernel<8><<<blocks, sgc::kRowWarps * 32, 0, st>>>(x, gy, mean, rstd, gamma, R, gx, partial); break;
    default: return (int)cudaErrorInvalidValue;
  }
  SGC_CUDA_CHECK_LAST();
  return 0;
}

extern "C" int sgc_layernorm_bwd_params(const float* partial, int R, int C, float* ggamma, float* gbeta, void* stream) {
  if (R <= 0) return 0;
  const int blocks = sgc::rowop_blocks(R);
  sgc::layernorm_bwd_params_kernel<<<(2 * C + 31) / 32, 256, 0, (cudaStream_t)stream>>>(partial, blocks, 2 * C, ggamma, gbeta);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

extern "C" int sgc_rowop_fwd(const sgc_rowop_fwd_args* args, void* stream) {
  const sgc_rowop_fwd_args a = *args;
  if (a.R <= 0) return 0;
  if (!a.x || (a.gamma && (!a.beta || !a.mean || !a.rstd))) return (int)cudaErrorInvalidValue;
  if (a.in_heads && (a.N % a.in_heads || (a.N / a.in_heads) % (a.N / 32))) return (int)cudaErrorInvalidValue;
  if (a.split_heads && (a.N % a.split_heads || (a.N / a.split_heads) % (a.N / 32))) return (int)cudaErrorInvalidValue;
  const int blocks = sgc::rowop_blocks(a.R);
  cudaStream_t st = (cudaStream_t)stream;
  switch (a.N) {
    case 128: sgc::launch_chain(sgc::rowop_fwd_kernel<4>, dim3(blocks), dim3(sgc::kRowWarps * 32), 0, st, a); break;
    case 256: sgc::launch_chain(sgc::rowop_fwd_kernel<8>, dim3(blocks), dim3(sgc::kRowWarps * 32), 0, st, a); break;
    case 512: sgc::launch_chain(sgc::rowop_fwd_kernel<16>, dim3(blocks), dim3(sgc::kRowWarps * 32), 0, st, a); break;
    default: return (int)cudaErrorInvalidValue;
  }
  SGC_CUDA_CHECK_LAST();
  return 0;
}

extern "C" int sgc_rowop_bwd(const sgc_rowop_bwd_args* args, void* stream) {
  const sgc_rowop_bwd_args a = *args;
  if (a.R <= 0) return 0;
  if (!a.g || (a.gamma && (!a.pre || !a.mean || !a.rstd || !a.partial))) return (int)cudaErrorInvalidValue;
  if (a.in_heads && (a.N % a.in_heads || (a.N / a.in_heads) % (a.N / 32))) return (int)cudaErrorInvalidValue;
  if (a.split_heads && (a.N % a.split_heads || (a.N / a.split_heads) % (a.N / 32))) return (int)cudaErrorInvalidValue;
  const int blocks = sgc::rowop_blocks(a.R);
  cudaStream_t st = (cudaStream_t)stream;
  switch (a.N) {
    case 128: sgc::launch_chain(sgc::rowop_bwd_kernel<4>, dim3(blocks), dim3(sgc::kRowWarps * 32), 0, st, a); break;
    case 256: sgc::launch_chain(sgc::rowop_bwd_kernel<8>, dim3(blocks), dim3(sgc::kRowWarps * 32), 0, st, a); break;
    case 512: sgc::launch_chain(sgc::rowop_bwd_kernel<16>, dim3(blocks), dim3(sgc::kRowWarps * 32), 0, st, a); break;
    default: return (int)cudaErrorInvalidValue;
  }
  SGC_CUDA_CHECK_LAST();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// sgc_dropout_masks: the keep-masks of every dropout of a step (all levels, all layers) in ONE launch.
// nn.Dropout semantics (custom_base_transformer_layer.py / FFN of encoder.py:262-340: x * mask / (1 - p), mask ~
// Bernoulli(1 - p)); the masks are applied inside sgc_rowop_fwd / _bwd.  Philox4x32-10 (Salmon et al., SC'11) keyed by
// `seed`, counter = (16-byte chunk index, job, step): the step number lives in device memory and is advanced by the last
// CTA of the launch, so a CUDA-graph replay draws fresh masks without any host-side RNG bookkeeping.
// state = {step, ticket} (two int64, zero-initialised by the caller, private to one launch site).
namespace sgc {

constexpr int kMaxMaskJobs = 12;
struct MaskJobs {
  unsigned char* out[kMaxMaskJobs];
  long long chunks_end[kMaxMaskJobs];   // running sum of 16-byte chunks
  long long n[kMaxMaskJobs];
  unsigned int thr[kMaxMaskJobs];       // keep iff r <= thr  (thr = keep * 2^32 - 1)
  int njobs;
};

__device__ __forceinline__ void philox4x32_10(unsigned int (&c)[4], unsigned int k0, unsigned int k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    c[0] = hi1 ^ c[1] ^ k0; c[1] = lo1; c[2] = hi0 ^ c[3] ^ k1; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

__global__ void __launch_bounds__(256) dropout_masks_kernel(const __grid_constant__ MaskJobs jobs, unsigned long long seed,
                                                            unsigned long long* __restrict__ state) {
  __shared__ unsigned long long s_step;
  if (threadIdx.x == 0) s_step = *reinterpret_cast<volatile unsigned long long*>(state);
  __syncthreads();
  const unsigned long long step = s_step;
  const long long total = jobs.chunks_end[jobs.njobs - 1];
  for (long long ch = (long long)blockIdx.x * blockDim.x + threadIdx.x; ch < total; ch += (long long)gridDim.x * blockDim.x) {
    int j = 0;
    while (ch >= jobs.chunks_end[j]) ++j;
    const long long local = ch - (j ? jobs.chunks_end[j - 1] : 0);
    const unsigned int thr = jobs.thr[j];
    unsigned int bytes[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      unsigned int c[4] = {(unsigned int)(local * 4 + q), (unsigned int)((unsigned long long)(local * 4 + q) >> 32) ^ ((unsigned int)j << 24),
                           (unsigned int)step, (unsigned int)(step >> 32)};
      philox4x32_10(c, (unsigned int)seed, (unsigned int)(seed >> 32));
      bytes[q] = (c[0] <= thr ? 1u : 0u) | (c[1] <= thr ? 0x100u : 0u) | (c[2] <= thr ? 0x10000u : 0u) | (c[3] <= thr ? 0x1000000u : 0u);
    }
    unsigned char* dst = jobs.out[j] + local * 16;
    if (local * 16 + 16 <= jobs.n[j]) {
      *reinterpret_cast<uint4*>(dst) = make_uint4(bytes[0], bytes[1], bytes[2], bytes[3]);
    } else {
      for (int b = 0; local * 16 + b < jobs.n[j]; ++b) dst[b] = (unsigned char)((bytes[b >> 2] >> (8 * (b & 3))) & 1u);
    }
  }
  // every CTA has read the step before it takes its ticket; the last one advances the step for the next launch / replay
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long t = atomicAdd(state + 1, 1ULL);
    if (t == (unsigned long long)gridDim.x - 1) {
      state[1] = 0;
      state[0] = step + 1;
      __threadfence();
    }
  }
}

}  // namespace sgc

extern "C" int sgc_dropout_masks(const sgc_mask_job* in, int njobs, long long seed, long long* state, void* stream) {
  if (!in || njobs <= 0 || njobs > sgc::kMaxMaskJobs || !state) return (int)cudaErrorInvalidValue;
  sgc::MaskJobs jobs;
  long long chunks = 0;
  for (int j = 0; j < njobs; ++j) {
    if (!in[j].out || in[j].n <= 0 || (reinterpret_cast<uintptr_t>(in[j].out) & 15) || !(in[j].keep > 0.f) || in[j].keep > 1.f)
      return (int)cudaErrorInvalidValue;
    jobs.out[j] = in[j].out;
    jobs.n[j] = in[j].n;
    chunks += (in[j].n + 15) / 16;
    jobs.chunks_end[j] = chunks;
    const double t = (double)in[j].keep * 4294967296.0 - 1.0;
    jobs.thr[j] = t >= 4294967295.0 ? 0xFFFFFFFFu : (unsigned int)t;
  }
  jobs.njobs = njobs;
  long long blocks = (chunks + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  sgc::dropout_masks_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(jobs, (unsigned long long)seed,
                                                                           reinterpret_cast<unsigned long long*>(state));
  SGC_CUDA_CHECK_LAST();
  return 0;
}
