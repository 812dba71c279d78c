// sgc_depth.cu -- the step in FRONT of the view transform: the depth-distribution producer.
//
//   sgc_plane_sweep_fwd / _bwd   the plane-sweep cost volume of DepthNet_Fusion
//                                (mmdet3d_plugin/models/im2voxel/depth_utils/depth_est_fusion.py:85-126 homo_warping,
//                                 :209-232 the loop over the neighbour frames):
//        corr[v, d, y, x] = 1/(K sqrt(C)) * sum_k sum_c  f[v, c, y, x] * bilinear(f[nb(v,k)], H_{v,k,d}(x, y))[c]
//     The reference materialises the warped features [V, C, D, H, W] per neighbour (1.2 GB at the ScanNet shape) and
//     reduces them afterwards; here one warp owns a pixel, keeps its reference feature in registers and, for every
//     (neighbour, depth), gathers the four corner rows of the channel-last neighbour map (512 B each, one float4 per lane)
//     and reduces the dot product with shuffles: nothing but the feature maps and the [V, D, H, W] result touches HBM.
//   sgc_depth_pyramid_fwd / _bwd softmax over the depth bins (depth_est_fusion.py:241) fused with the nearest x1/2, x1/4
//                                pyramid of SGCDet.build_volume (detectors/SGCDet.py:83-85) AND the channel-last, cropped
//                                [V, h*w, D] layout the lift kernels read (DenseHead.prepare's permute copies disappear).
//   sgc_nchw_to_nhwc / sgc_nhwc_to_nchw   tiled transposes between the FPN's NCHW maps and the channel-last gather layout.
#include "common.cuh"

namespace sgc {

constexpr int kPsMaxChunks = 4;    // channels <= 512: float4 chunk j of a lane covers channels 4 * (lane + 32 j)
constexpr int kPsMaxDepth = 32;    // lane d keeps the result of depth bin d
constexpr int kPsWarps = 8;

// grid_sample(mode=bilinear, padding_mode=zeros, align_corners=False) of a pixel coordinate that homo_warping normalised
// with (size - 1) / 2: the reference's own mismatch of conventions is part of the contract.
struct Bil {
  int pix[4];     // y * W + x of the NW, NE, SW, SE corners or -1
  float w[4];
};

__device__ __forceinline__ Bil ps_corners(float px, float py, int H, int W) {
  Bil b;
  const float gx = __fsub_rn(__fdiv_rn(px, (float)(W - 1) * 0.5f), 1.f);
  const float gy = __fsub_rn(__fdiv_rn(py, (float)(H - 1) * 0.5f), 1.f);
  const float ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)W), 1.f), 0.5f);
  const float iy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)H), 1.f), 0.5f);
#pragma unroll
  for (int k = 0; k < 4; ++k) { b.pix[k] = -1; b.w[k] = 0.f; }
  // NaN / inf / far-away coordinates sample nothing (every comparison below is false for NaN)
  if (!(ix > -1.f && ix < (float)W && iy > -1.f && iy < (float)H)) return b;
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float tx = ix - fx, ty = iy - fy;
  const bool l = x0 >= 0, r = x0 + 1 <= W - 1, t = y0 >= 0, d = y0 + 1 <= H - 1;
  b.w[0] = (1.f - tx) * (1.f - ty); b.w[1] = tx * (1.f - ty); b.w[2] = (1.f - tx) * ty; b.w[3] = tx * ty;
  if (t && l) b.pix[0] = y0 * W + x0;
  if (t && r) b.pix[1] = y0 * W + x0 + 1;
  if (d && l) b.pix[2] = (y0 + 1) * W + x0;
  if (d && r) b.pix[3] = (y0 + 1) * W + x0 + 1;
  return b;
}

// pixel (x, y) of view v at depth z seen from neighbour (v, k): rt = [rot (9, row-major) | trans (3)] of
// src_proj @ inverse(ref_proj) (depth_est_fusion.py:95-97,106-114)
__device__ __forceinline__ void ps_project(const float* __restrict__ rt, float x, float y, float z, float& px, float& py) {
  const float rx = __ldg(rt + 0) * x + __ldg(rt + 1) * y + __ldg(rt + 2);
  const float ry = __ldg(rt + 3) * x + __ldg(rt + 4) * y + __ldg(rt + 5);
  const float rz = __ldg(rt + 6) * x + __ldg(rt + 7) * y + __ldg(rt + 8);
  const float qx = __fadd_rn(__fmul_rn(rx, z), __ldg(rt + 9));
  const float qy = __fadd_rn(__fmul_rn(ry, z), __ldg(rt + 10));
  const float qz = __fadd_rn(__fmul_rn(rz, z), __ldg(rt + 11));
  px = __fdiv_rn(qx, qz);
  py = __fdiv_rn(qy, qz);
}

// feat: channel-last [V, H*W, C]; nbr [V, K] int32; rt [V, K, 12]; depth [D]; corr [V, D, H, W]
__global__ void __launch_bounds__(kPsWarps * 32) plane_sweep_fwd_kernel(const float* __restrict__ feat,
                                                                        const int* __restrict__ nbr,
                                                                        const float* __restrict__ rt,
                                                                        const float* __restrict__ depth, int V, int K, int D,
                                                                        int H, int W, int C, float scale,
                                                                        float* __restrict__ corr) {
  const int lane = threadIdx.x & 31;
  const int S = H * W;
  const long long total = (long long)V * S;
  const int nch = (C + 127) >> 7;
  for (long long item = (long long)blockIdx.x * kPsWarps + (threadIdx.x >> 5); item < total;
       item += (long long)gridDim.x * kPsWarps) {
    const int v = (int)(item / S), pix = (int)(item - (long long)v * S);
    const float x = (float)(pix % W), y = (float)(pix / W);
    float4 ref[kPsMaxChunks];
#pragma unroll
    for (int j = 0; j < kPsMaxChunks; ++j) {
      const int c = 4 * (lane + 32 * j);
      ref[j] = (j < nch && c < C) ? ldg4(feat + ((size_t)v * S + pix) * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float acc = 0.f;   // lane d: sum over the neighbours of the dot product at depth bin d
    for (int k = 0; k < K; ++k) {
      const int u = __ldg(nbr + v * K + k);
      const float* src = feat + (size_t)u * S * C;
      const float* m = rt + ((size_t)v * K + k) * 12;
      for (int d = 0; d < D; ++d) {
        float px, py;
        ps_project(m, x, y, __ldg(depth + d), px, py);
        const Bil b = ps_corners(px, py, H, W);
        float part = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (b.pix[q] < 0) continue;
          const float* row = src + (size_t)b.pix[q] * C;
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < kPsMaxChunks; ++j) {
            const int c = 4 * (lane + 32 * j);
            if (j < nch && c < C) {
              const float4 t = ldg4(row + c);
              s += ref[j].x * t.x + ref[j].y * t.y + ref[j].z * t.z + ref[j].w * t.w;
            }
          }
          part += b.w[q] * s;
        }
        part = warp_sum(part);
        if (lane == d) acc += part;
      }
    }
    if (lane < D) corr[((size_t)v * D + lane) * S + pix] = acc * scale;
  }
}

// grad_feat (channel-last, ZERO-FILLED by the entry point) += both roles of every feature row: as the reference feature of
// its own pixel (sum over the samples of g * warped) and as a corner of other views' samples (g * w * ref).
__global__ void __launch_bounds__(kPsWarps * 32) plane_sweep_bwd_kernel(const float* __restrict__ feat,
                                                                        const int* __restrict__ nbr,
                                                                        const float* __restrict__ rt,
                                                                        const float* __restrict__ depth,
                                                                        const float* __restrict__ gcorr, int V, int K, int D,
                                                                        int H, int W, int C, float scale,
                                                                        float* __restrict__ gfeat) {
  const int lane = threadIdx.x & 31;
  const int S = H * W;
  const long long total = (long long)V * S;
  const int nch = (C + 127) >> 7;
  for (long long item = (long long)blockIdx.x * kPsWarps + (threadIdx.x >> 5); item < total;
       item += (long long)gridDim.x * kPsWarps) {
    const int v = (int)(item / S), pix = (int)(item - (long long)v * S);
    const float x = (float)(pix % W), y = (float)(pix / W);
    float4 ref[kPsMaxChunks], gref[kPsMaxChunks];
#pragma unroll
    for (int j = 0; j < kPsMaxChunks; ++j) {
      const int c = 4 * (lane + 32 * j);
      ref[j] = (j < nch && c < C) ? ldg4(feat + ((size_t)v * S + pix) * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      gref[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float gmine = lane < D ? __ldg(gcorr + ((size_t)v * D + lane) * S + pix) * scale : 0.f;
    for (int k = 0; k < K; ++k) {
      const int u = __ldg(nbr + v * K + k);
      const float* src = feat + (size_t)u * S * C;
      float* gsrc = gfeat + (size_t)u * S * C;
      const float* m = rt + ((size_t)v * K + k) * 12;
      for (int d = 0; d < D; ++d) {
        const float g = __shfl_sync(SGC_FULL_MASK, gmine, d);
        if (g == 0.f) continue;
        float px, py;
        ps_project(m, x, y, __ldg(depth + d), px, py);
        const Bil b = ps_corners(px, py, H, W);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (b.pix[q] < 0) continue;
          const float gw = g * b.w[q];
          const float* row = src + (size_t)b.pix[q] * C;
          float* grow = gsrc + (size_t)b.pix[q] * C;
#pragma unroll
          for (int j = 0; j < kPsMaxChunks; ++j) {
            const int c = 4 * (lane + 32 * j);
            if (j < nch && c < C) {
              const float4 t = ldg4(row + c);
              gref[j].x += gw * t.x; gref[j].y += gw * t.y; gref[j].z += gw * t.z; gref[j].w += gw * t.w;
              if (gw != 0.f) red_add4(grow + c, gw * ref[j].x, gw * ref[j].y, gw * ref[j].z, gw * ref[j].w);
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kPsMaxChunks; ++j) {
      const int c = 4 * (lane + 32 * j);
      if (j < nch && c < C) red_add4(gfeat + ((size_t)v * S + pix) * C + c, gref[j].x, gref[j].y, gref[j].z, gref[j].w);
    }
  }
}

// ---- tiled transposes: [B, R, Cc] <-> [B, Cc, R] (R = H*W pixels, Cc channels) -------------------------------------------
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows,
                                                        int cols, long long in_batch, long long out_batch) {
  // in[b][r][c] (row stride cols) -> out[b][c][r] (row stride rows)
  __shared__ float tile[32][33];
  const float* src = in + (size_t)blockIdx.z * in_batch;
  float* dst = out + (size_t)blockIdx.z * out_batch;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    if (r < rows && c < cols) tile[i][tx] = __ldg(src + (size_t)r * cols + c);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (r < rows && c < cols) dst[(size_t)c * rows + r] = tile[tx][i];
  }
}

// ---- softmax over D + nearest pyramid + channel-last cropped layouts -----------------------------------------------------
// logits [V, D, H, W] -> prob [V, D, H, W] (optional, the reference's return value for the depth loss) and, for level l
// (stride 1 << l, l = 0, 1, 2), cl[l] [V, h_l * w_l, D] with cl[l][v, y*w_l + x, :] = prob[v, :, y << l, x << l]
// (F.interpolate(mode='nearest', scale_factor=1/2, 1/4) picks the even / every fourth pixel; then the [:h, :w] crop).
struct PyramidArgs {
  float* cl[3];
  int h[3], w[3];
};

__global__ void __launch_bounds__(256) depth_pyramid_fwd_kernel(const float* __restrict__ logits, int V, int D, int H, int W,
                                                                float* __restrict__ prob, const PyramidArgs a) {
  const int S = H * W;
  const long long total = (long long)V * S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i / S), pix = (int)(i - (long long)v * S);
    const int y = pix / W, x = pix - y * W;
    const float* lp = logits + (size_t)v * D * S + pix;
    float mx = -3.0e38f;
    for (int d = 0; d < D; ++d) mx = fmaxf(mx, __ldg(lp + (size_t)d * S));
    float e[kPsMaxDepth];
    float sum = 0.f;
#pragma unroll
    for (int d = 0; d < kPsMaxDepth; ++d) {
      if (d < D) { e[d] = expf(__ldg(lp + (size_t)d * S) - mx); sum += e[d]; }
    }
#pragma unroll
    for (int d = 0; d < kPsMaxDepth; ++d) {
      if (d < D) {
        e[d] = __fdiv_rn(e[d], sum);
        if (prob) prob[(size_t)v * D * S + (size_t)d * S + pix] = e[d];
      }
    }
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      if (!a.cl[l]) continue;
      const int m = (1 << l) - 1;
      if ((x & m) || (y & m)) continue;
      const int xl = x >> l, yl = y >> l;
      if (xl >= a.w[l] || yl >= a.h[l]) continue;
      float* o = a.cl[l] + ((size_t)v * a.h[l] * a.w[l] + (size_t)yl * a.w[l] + xl) * D;
#pragma unroll
      for (int d = 0; d < kPsMaxDepth; ++d)
        if (d < D) o[d] = e[d];
    }
  }
}

struct PyramidGradArgs {
  const float* gcl[3];
  int h[3], w[3];
};

// glogits = p * (g - sum_d p g) with g = gprob (optional) + the pyramid levels' gradients at the pixels they sampled
__global__ void __launch_bounds__(256) depth_pyramid_bwd_kernel(const float* __restrict__ logits, int V, int D, int H, int W,
                                                                const float* __restrict__ gprob, const PyramidGradArgs a,
                                                                float* __restrict__ glogits) {
  const int S = H * W;
  const long long total = (long long)V * S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i / S), pix = (int)(i - (long long)v * S);
    const int y = pix / W, x = pix - y * W;
    const float* lp = logits + (size_t)v * D * S + pix;
    float mx = -3.0e38f;
    for (int d = 0; d < D; ++d) mx = fmaxf(mx, __ldg(lp + (size_t)d * S));
    float p[kPsMaxDepth], g[kPsMaxDepth];
    float sum = 0.f;
#pragma unroll
    for (int d = 0; d < kPsMaxDepth; ++d) {
      if (d < D) {
        p[d] = expf(__ldg(lp + (size_t)d * S) - mx);
        sum += p[d];
        g[d] = gprob ? __ldg(gprob + (size_t)v * D * S + (size_t)d * S + pix) : 0.f;
      }
    }
#pragma unroll
    for (int l = 0; l < 3; ++l) {
      if (!a.gcl[l]) continue;
      const int m = (1 << l) - 1;
      if ((x & m) || (y & m)) continue;
      const int xl = x >> l, yl = y >> l;
      if (xl >= a.w[l] || yl >= a.h[l]) continue;
      const float* o = a.gcl[l] + ((size_t)v * a.h[l] * a.w[l] + (size_t)yl * a.w[l] + xl) * D;
#pragma unroll
      for (int d = 0; d < kPsMaxDepth; ++d)
        if (d < D) g[d] += __ldg(o + d);
    }
    float dot = 0.f;
#pragma unroll
    for (int d = 0; d < kPsMaxDepth; ++d) {
      if (d < D) { p[d] = __fdiv_rn(p[d], sum); dot += p[d] * g[d]; }
    }
#pragma unroll
    for (int d = 0; d < kPsMaxDepth; ++d)
      if (d < D) glogits[(size_t)v * D * S + (size_t)d * S + pix] = p[d] * (g[d] - dot);
  }
}

}  // namespace sgc

static int ps_grid(long long items, int per_cta) {
  long long g = (items + per_cta - 1) / per_cta;
  const long long cap = 148ll * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

static bool ps_args_ok(const void* feat, const void* nbr, const void* rt, const void* depth, int V, int K, int D, int H, int W,
                       int C) {
  return feat && nbr && rt && depth && V > 0 && K > 0 && D > 0 && D <= sgc::kPsMaxDepth && H > 1 && W > 1 && C > 0 && C % 4 == 0 &&
         C <= 128 * sgc::kPsMaxChunks && !(reinterpret_cast<uintptr_t>(feat) & 15);
}

extern "C" int sgc_plane_sweep_fwd(const float* feat_cl, const int* nbr, const float* rt, const float* depth, int V, int K,
                                   int D, int H, int W, int C, float* corr, void* stream) {
  if (!ps_args_ok(feat_cl, nbr, rt, depth, V, K, D, H, W, C) || !corr) return (int)cudaErrorInvalidValue;
  const float scale = 1.f / ((float)K * sqrtf((float)C));
  sgc::plane_sweep_fwd_kernel<<<ps_grid((long long)V * H * W, sgc::kPsWarps), sgc::kPsWarps * 32, 0, (cudaStream_t)stream>>>(
      feat_cl, nbr, rt, depth, V, K, D, H, W, C, scale, corr);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

extern "C" int sgc_plane_sweep_bwd(const float* feat_cl, const int* nbr, const float* rt, const float* depth,
                                   const float* grad_corr, int V, int K, int D, int H, int W, int C, float* grad_feat_cl,
                                   void* stream) {
  if (!ps_args_ok(feat_cl, nbr, rt, depth, V, K, D, H, W, C) || !grad_corr || !grad_feat_cl ||
      (reinterpret_cast<uintptr_t>(grad_feat_cl) & 15))
    return (int)cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(grad_feat_cl, 0, (size_t)V * H * W * C * sizeof(float), (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  const float scale = 1.f / ((float)K * sqrtf((float)C));
  sgc::plane_sweep_bwd_kernel<<<ps_grid((long long)V * H * W, sgc::kPsWarps), sgc::kPsWarps * 32, 0, (cudaStream_t)stream>>>(
      feat_cl, nbr, rt, depth, grad_corr, V, K, D, H, W, C, scale, grad_feat_cl);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

static int transpose_launch(const float* in, float* out, int B, int rows, int cols, void* stream) {
  if (!in || !out || B <= 0 || rows <= 0 || cols <= 0 || B > 65535) return (int)cudaErrorInvalidValue;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, B);
  if (grid.y > 65535) return (int)cudaErrorInvalidValue;
  sgc::transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, rows, cols, (long long)rows * cols, (long long)rows * cols);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

// [B, C, S] -> [B, S, C]
extern "C" int sgc_nchw_to_nhwc(const float* in, int B, int C, int S, float* out, void* stream) {
  return transpose_launch(in, out, B, C, S, stream);
}
// [B, S, C] -> [B, C, S]
extern "C" int sgc_nhwc_to_nchw(const float* in, int B, int C, int S, float* out, void* stream) {
  return transpose_launch(in, out, B, S, C, stream);
}

// cl[l] may be NULL (level not wanted); h[l] <= ceil(H / 2^l), w[l] <= ceil(W / 2^l) is the crop of level l.
extern "C" int sgc_depth_pyramid_fwd(const float* logits, int V, int D, int H, int W, float* prob, float* cl0, int h0, int w0,
                                     float* cl1, int h1, int w1, float* cl2, int h2, int w2, void* stream) {
  if (!logits || V <= 0 || D <= 0 || D > sgc::kPsMaxDepth || H <= 0 || W <= 0) return (int)cudaErrorInvalidValue;
  sgc::PyramidArgs a;
  a.cl[0] = cl0; a.cl[1] = cl1; a.cl[2] = cl2;
  a.h[0] = h0; a.h[1] = h1; a.h[2] = h2;
  a.w[0] = w0; a.w[1] = w1; a.w[2] = w2;
  for (int l = 0; l < 3; ++l) {
    if (!a.cl[l]) continue;
    const int hl = (H + (1 << l) - 1) >> l, wl = (W + (1 << l) - 1) >> l;
    if (a.h[l] <= 0 || a.w[l] <= 0 || a.h[l] > hl || a.w[l] > wl) return (int)cudaErrorInvalidValue;
  }
  sgc::depth_pyramid_fwd_kernel<<<ps_grid((long long)V * H * W, 256), 256, 0, (cudaStream_t)stream>>>(logits, V, D, H, W, prob, a);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

extern "C" int sgc_depth_pyramid_bwd(const float* logits, int V, int D, int H, int W, const float* grad_prob,
                                     const float* gcl0, int h0, int w0, const float* gcl1, int h1, int w1, const float* gcl2,
                                     int h2, int w2, float* grad_logits, void* stream) {
  if (!logits || !grad_logits || V <= 0 || D <= 0 || D > sgc::kPsMaxDepth || H <= 0 || W <= 0) return (int)cudaErrorInvalidValue;
  sgc::PyramidGradArgs a;
  a.gcl[0] = gcl0; a.gcl[1] = gcl1; a.gcl[2] = gcl2;
  a.h[0] = h0; a.h[1] = h1; a.h[2] = h2;
  a.w[0] = w0; a.w[1] = w1; a.w[2] = w2;
  for (int l = 0; l < 3; ++l) {
    if (!a.gcl[l]) continue;
    const int hl = (H + (1 << l) - 1) >> l, wl = (W + (1 << l) - 1) >> l;
    if (a.h[l] <= 0 || a.w[l] <= 0 || a.h[l] > hl || a.w[l] > wl) return (int)cudaErrorInvalidValue;
  }
  sgc::depth_pyramid_bwd_kernel<<<ps_grid((long long)V * H * W, 256), 256, 0, (cudaStream_t)stream>>>(
      logits, V, D, H, W, grad_prob, a, grad_logits);
  SGC_CUDA_CHECK_LAST();
  return 0;
}
