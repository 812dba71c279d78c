// Per-voxel row kernels of the encoder layer (VoxFormerLayer, encoder.py:262-340): one warp per voxel row of C = 32*CPL
// channels, float4 accesses, shuffle reductions.  The dense GEMMs between them run on the tensor cores; these kernels
// carry everything else of the chain so that a level is a short, fixed sequence of launches.
//
//   sgc_layernorm_bwd          gx = rstd * (g*gamma - mean_c(g*gamma) - xhat * mean_c(g*gamma*xhat))   (nn.LayerNorm backward,
//                              encoder.py:325-338 norms) + per-CTA partial sums of (g*xhat, g) for gamma / beta
//   sgc_layernorm_bwd_params   fixed-order reduction of those partials (deterministic; runs on the weight-gradient stream)
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
