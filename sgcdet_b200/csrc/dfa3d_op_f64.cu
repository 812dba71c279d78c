// fp64 instantiation of the four operator-level DFA3D entry points behind the reference's `_ext` boundary.  The reference
// dispatches its kernels over float and double (AT_DISPATCH_FLOATING_TYPES, csrc/cuda/ms_depth_score_sample_cuda.cu:95,153,
// csrc/cuda/wms_deform_attn_cuda.cu:267,344); the SGCDet path itself is fp32, so these are plain, unfused, scalar kernels:
// same work decomposition as the generic fp32 kernels of dfa3d_op.cu (one thread per sampling point for the depth scores, one
// warp per (b, q, m) with the lanes striding over the head's channels for the attention), all arithmetic in double.
//   ms_depth_score_sample_{forward,backward}   (ms_depth_score_sample_cuda_kernel.cuh:24-327)
//   wms_deform_attn_{forward,backward}         (wms_deform_attn_cuda_kernel.cuh:24-531)
#include <cuda_runtime.h>
#include <stdint.h>

#define SGC_FULL_MASK 0xffffffffu

namespace sgc {
namespace f64 {

struct TapD {
  int pix[4];      // TL, TR, BR, BL (DSK:89-92; WMSK:51,58,65,72) or -1
  double bw[4];
  double lh, lw, ld, hd;
  int d0;
  bool in2d, in3d;
};

__device__ __forceinline__ TapD make_tap(double x, double y, double z, int H, int W, int D) {
  TapD t;
  const double h = y * (double)H - 0.5, w = x * (double)W - 0.5, d = z * (double)D - 0.5;   // DSK:133-135, WMSK:286-287
  t.in2d = (h > -1.0) && (w > -1.0) && (h < (double)H) && (w < (double)W);                  // DSK:137, WMSK:289
  t.in3d = t.in2d && (d > -1.0) && (d < (double)D);
  const double hf = floor(h), wf = floor(w), df = floor(d);
  const int h0 = (int)hf, w0 = (int)wf;
  t.d0 = (int)df;
  t.lh = h - hf; t.lw = w - wf; t.ld = d - df;
  t.hd = 1.0 - t.ld;
  const double hh = 1.0 - t.lh, hw = 1.0 - t.lw;
  t.bw[0] = hh * hw; t.bw[1] = hh * t.lw; t.bw[2] = t.lh * t.lw; t.bw[3] = t.lh * hw;
  const bool top = h0 >= 0, bot = h0 + 1 <= H - 1, lef = w0 >= 0, rig = w0 + 1 <= W - 1;
  t.pix[0] = (t.in2d && top && lef) ? h0 * W + w0 : -1;
  t.pix[1] = (t.in2d && top && rig) ? h0 * W + w0 + 1 : -1;
  t.pix[2] = (t.in2d && bot && rig) ? (h0 + 1) * W + w0 + 1 : -1;
  t.pix[3] = (t.in2d && bot && lef) ? (h0 + 1) * W + w0 : -1;
  return t;
}

__device__ __forceinline__ double depth_score(const TapD& t, const double* __restrict__ dist_px, int D, double& lo, double& hi) {
  lo = (t.d0 >= 0) ? dist_px[t.d0] : 0.0;
  hi = (t.d0 + 1 <= D - 1) ? dist_px[t.d0 + 1] : 0.0;
  return lo * t.hd + hi * t.ld;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SGC_FULL_MASK, v, o);
  return v;
}

__global__ void depth_score_fwd_kernel(const double* __restrict__ dist, const int64_t* __restrict__ shapes3d,
                                       const int64_t* __restrict__ lsi, const double* __restrict__ loc, int B, int S, int M,
                                       int Dch, int L, int Q, int P, double* __restrict__ out) {
  const long long total = (long long)B * Q * M * L * P;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int l = (int)((i / P) % L);
    const int m = (int)((i / ((long long)P * L)) % M);
    const int b = (int)(i / ((long long)P * L * M * Q));
    const int H = (int)shapes3d[l * 3], W = (int)shapes3d[l * 3 + 1], D = (int)shapes3d[l * 3 + 2], start = (int)lsi[l];
    const TapD t = make_tap(loc[i * 3], loc[i * 3 + 1], loc[i * 3 + 2], H, W, D);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      double r = 0.0;
      if (t.in3d && t.pix[k] >= 0) {
        double lo, hi;
        r = depth_score(t, dist + (((size_t)b * S + start + t.pix[k]) * M + m) * Dch, D, lo, hi);
      }
      out[i * 4 + k] = r;
    }
  }
}

__global__ void depth_score_bwd_kernel(const double* __restrict__ dist, const int64_t* __restrict__ shapes3d,
                                       const int64_t* __restrict__ lsi, const double* __restrict__ loc,
                                       const double* __restrict__ grad_out, int B, int S, int M, int Dch, int L, int Q, int P,
                                       double* __restrict__ grad_dist, double* __restrict__ grad_loc) {
  const long long total = (long long)B * Q * M * L * P;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int l = (int)((i / P) % L);
    const int m = (int)((i / ((long long)P * L)) % M);
    const int b = (int)(i / ((long long)P * L * M * Q));
    const int H = (int)shapes3d[l * 3], W = (int)shapes3d[l * 3 + 1], D = (int)shapes3d[l * 3 + 2], start = (int)lsi[l];
    const TapD t = make_tap(loc[i * 3], loc[i * 3 + 1], loc[i * 3 + 2], H, W, D);
    double gz = 0.0;
    if (t.in3d) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (t.pix[k] >= 0) {
          const size_t o = (((size_t)b * S + start + t.pix[k]) * M + m) * Dch;
          const double g = grad_out[i * 4 + k];
          double lo, hi;
          depth_score(t, dist + o, D, lo, hi);
          gz += g * (hi - lo);
          if (t.d0 >= 0) atomicAdd(grad_dist + o + t.d0, t.hd * g);
          if (t.d0 + 1 <= D - 1) atomicAdd(grad_dist + o + t.d0 + 1, t.ld * g);
        }
      }
    }
    grad_loc[i * 3 + 0] = 0.0;  // DSK:238-240
    grad_loc[i * 3 + 1] = 0.0;
    grad_loc[i * 3 + 2] = (double)D * gz;
  }
}

__global__ void __launch_bounds__(256) wms_fwd_kernel(const double* __restrict__ value, const int64_t* __restrict__ shapes,
                                                      const int64_t* __restrict__ lsi, const double* __restrict__ loc,
                                                      const double* __restrict__ attn, const double* __restrict__ ds_in, int B,
                                                      int S, int M, int Cm, int L, int Q, int P, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long items = (long long)B * Q * M;
  const long long w0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long wstride = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long it = w0; it < items; it += wstride) {
    const int m = (int)(it % M);
    const int b = (int)(it / ((long long)M * Q));
    for (int c = lane; c < Cm; c += 32) {
      double acc = 0.0;
      for (int l = 0; l < L; ++l) {
        const int H = (int)shapes[l * 2], W = (int)shapes[l * 2 + 1], start = (int)lsi[l];
        for (int p = 0; p < P; ++p) {
          const long long sp = (it * L + l) * P + p;
          const TapD t = make_tap(loc[sp * 2], loc[sp * 2 + 1], 0.5, H, W, 1);
          if (!t.in2d) continue;
          const double a = attn[sp];
          double val = 0.0;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (t.pix[k] >= 0) val += (t.bw[k] * ds_in[sp * 4 + k]) * value[(((size_t)b * S + start + t.pix[k]) * M + m) * Cm + c];
          acc += val * a;
        }
      }
      out[it * Cm + c] = acc;
    }
  }
}

__global__ void __launch_bounds__(256) wms_bwd_kernel(const double* __restrict__ value, const int64_t* __restrict__ shapes,
                                                      const int64_t* __restrict__ lsi, const double* __restrict__ loc,
                                                      const double* __restrict__ attn, const double* __restrict__ ds_in,
                                                      const double* __restrict__ grad_out, int B, int S, int M, int Cm, int L,
                                                      int Q, int P, double* __restrict__ grad_value, double* __restrict__ grad_loc,
                                                      double* __restrict__ grad_attn, double* __restrict__ grad_ds) {
  const int lane = threadIdx.x & 31;
  const long long items = (long long)B * Q * M;
  const long long w0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long wstride = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long it = w0; it < items; it += wstride) {
    const int m = (int)(it % M);
    const int b = (int)(it / ((long long)M * Q));
    for (int l = 0; l < L; ++l) {
      const int H = (int)shapes[l * 2], W = (int)shapes[l * 2 + 1], start = (int)lsi[l];
      for (int p = 0; p < P; ++p) {
        const long long sp = (it * L + l) * P + p;
        const TapD t = make_tap(loc[sp * 2], loc[sp * 2 + 1], 0.5, H, W, 1);
        const double a = attn[sp];
        double ds[4], dot[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < 4; ++k) ds[k] = ds_in[sp * 4 + k];
        if (t.in2d) {
          for (int c = lane; c < Cm; c += 32) {
            const double g = grad_out[it * Cm + c];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (t.pix[k] >= 0) {
                const size_t o = (((size_t)b * S + start + t.pix[k]) * M + m) * Cm + c;
                dot[k] += value[o] * g;
                const double wgt = t.bw[k] * ds[k] * a;
                if (wgt != 0.0) atomicAdd(grad_value + o, wgt * g);
              }
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) dot[k] = warp_sum(dot[k]);
        }
        if (lane == 0) {
          double g_attn = 0.0, g_w = 0.0, g_h = 0.0, gds[4] = {0.0, 0.0, 0.0, 0.0};
          if (t.in2d) {
            const double hh = 1.0 - t.lh, hw = 1.0 - t.lw;
            const double e0 = ds[0] * dot[0], e1 = ds[1] * dot[1], e2 = ds[2] * dot[2], e3 = ds[3] * dot[3];
            g_attn = t.bw[0] * e0 + t.bw[1] * e1 + t.bw[2] * e2 + t.bw[3] * e3;
            g_w = (double)W * a * (-hh * e0 + hh * e1 + t.lh * e2 - t.lh * e3);      // WMSK:116-158
            g_h = (double)H * a * (-hw * e0 - t.lw * e1 + t.lw * e2 + hw * e3);
#pragma unroll
            for (int k = 0; k < 4; ++k) gds[k] = (t.pix[k] >= 0) ? a * t.bw[k] * dot[k] : 0.0;
          }
          grad_loc[sp * 2 + 0] = g_w;
          grad_loc[sp * 2 + 1] = g_h;
#pragma unroll
          for (int k = 0; k < 4; ++k) grad_ds[sp * 4 + k] = gds[k];
          grad_attn[sp] = g_attn;
        }
      }
    }
  }
}

static inline int grid_for(long long items, int per_block) {
  long long g = (items + per_block - 1) / per_block;
  return (int)(g < 1 ? 1 : (g > 148ll * 16 ? 148ll * 16 : g));
}

}  // namespace f64
}  // namespace sgc

#define SGC_CHECK_LAST()                       \
  do {                                        \
    cudaError_t e__ = cudaGetLastError();     \
    if (e__ != cudaSuccess) return (int)e__;  \
  } while (0)

extern "C" int dfa3d_depth_score_fwd_f64(const double* dist, const int64_t* shapes3d, const int64_t* lsi, const double* loc, int B,
                                         int S, int M, int D, int L, int Q, int P, double* out, void* stream) {
  const long long total = (long long)B * Q * M * L * P;
  if (total == 0) return 0;
  sgc::f64::depth_score_fwd_kernel<<<sgc::f64::grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(dist, shapes3d, lsi, loc, B, S,
                                                                                                 M, D, L, Q, P, out);
  SGC_CHECK_LAST();
  return 0;
}

extern "C" int dfa3d_depth_score_bwd_f64(const double* dist, const int64_t* shapes3d, const int64_t* lsi, const double* loc,
                                         const double* grad_out, int B, int S, int M, int D, int L, int Q, int P,
                                         double* grad_dist, double* grad_loc, void* stream) {
  const long long total = (long long)B * Q * M * L * P;
  if (total == 0) return 0;
  sgc::f64::depth_score_bwd_kernel<<<sgc::f64::grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      dist, shapes3d, lsi, loc, grad_out, B, S, M, D, L, Q, P, grad_dist, grad_loc);
  SGC_CHECK_LAST();
  return 0;
}

extern "C" int dfa3d_wms_fwd_f64(const double* value, const int64_t* shapes, const int64_t* lsi, const double* loc,
                                 const double* attn, const double* depth_score, int B, int S, int M, int Cm, int L, int Q, int P,
                                 double* out, void* stream) {
  const long long items = (long long)B * Q * M;
  if (items == 0) return 0;
  sgc::f64::wms_fwd_kernel<<<sgc::f64::grid_for(items, 8), 256, 0, (cudaStream_t)stream>>>(value, shapes, lsi, loc, attn, depth_score,
                                                                                       B, S, M, Cm, L, Q, P, out);
  SGC_CHECK_LAST();
  return 0;
}

extern "C" int dfa3d_wms_bwd_f64(const double* value, const int64_t* shapes, const int64_t* lsi, const double* loc,
                                 const double* attn, const double* depth_score, const double* grad_out, int B, int S, int M,
                                 int Cm, int L, int Q, int P, double* grad_value, double* grad_loc, double* grad_attn,
                                 double* grad_depth_score, void* stream) {
  const long long items = (long long)B * Q * M;
  if (items == 0) return 0;
  sgc::f64::wms_bwd_kernel<<<sgc::f64::grid_for(items, 8), 256, 0, (cudaStream_t)stream>>>(
      value, shapes, lsi, loc, attn, depth_score, grad_out, B, S, M, Cm, L, Q, P, grad_value, grad_loc, grad_attn, grad_depth_score);
  SGC_CHECK_LAST();
  return 0;
}
