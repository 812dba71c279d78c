// Operator-level DFA3D kernels behind the reference's `_ext` boundary (csrc/pybind.cpp:20-67):
//   ms_depth_score_sample_{forward,backward}   (ms_depth_score_sample_cuda_kernel.cuh:24-327)
//   wms_deform_attn_{forward,backward}         (wms_deform_attn_cuda_kernel.cuh:24-531)
// plus the one-stage fusion used by MultiScale3DDeformableAttnFunction (F3D:277-351) in which the depth
// scores never round-trip through HBM.
//
// Generic over B, Q, M, Cm, L, P, D (any sizes).  One warp per (b,q,m); lanes stride over the head's
// channels, so every corner read is one coalesced row segment of the channel-last value map.
// Accumulation conventions of the reference are kept: forward outputs are fully written; backward
// ACCUMULATES into grad_value / grad_dist (caller zeroes) and WRITES grad_loc / grad_attn / grad_depth_score.
#include "common.cuh"

namespace sgc {

struct LevelInfo { int H, W, D, start; };

__device__ __forceinline__ LevelInfo level_info(const int64_t* shapes, int stride, const int64_t* lsi, int l) {
  LevelInfo li;
  li.H = (int)shapes[l * stride];
  li.W = (int)shapes[l * stride + 1];
  li.D = stride == 3 ? (int)shapes[l * stride + 2] : 1;
  li.start = (int)lsi[l];
  return li;
}

// ---- depth score forward: one thread per (b,q,m,l,p) -----------------------------------------------
__global__ void depth_score_fwd_kernel(const float* __restrict__ dist, const int64_t* __restrict__ shapes3d,
                                       const int64_t* __restrict__ lsi, const float* __restrict__ loc, int B, int S,
                                       int M, int Dch, int L, int Q, int P, float* __restrict__ out) {
  const long long total = (long long)B * Q * M * L * P;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int l = (int)((i / P) % L);
    const int m = (int)((i / ((long long)P * L)) % M);
    const int b = (int)(i / ((long long)P * L * M * Q));
    const LevelInfo li = level_info(shapes3d, 3, lsi, l);
    const float* lp = loc + i * 3;
    const Tap t = make_tap(lp[0], lp[1], lp[2], li.H, li.W, li.D);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t.in3d) {
      float r[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        r[k] = 0.f;
        if (t.pix[k] >= 0) {
          float lo, hi;
          r[k] = depth_score(t, dist + (((size_t)b * S + li.start + t.pix[k]) * M + m) * Dch, li.D, lo, hi);
        }
      }
      o = make_float4(r[0], r[1], r[2], r[3]);
    }
    reinterpret_cast<float4*>(out)[i] = o;
  }
}

// ---- depth score backward: one thread per (b,q,m,l,p) ----------------------------------------------
__global__ void depth_score_bwd_kernel(const float* __restrict__ dist, const int64_t* __restrict__ shapes3d,
                                       const int64_t* __restrict__ lsi, const float* __restrict__ loc,
                                       const float* __restrict__ grad_out, int B, int S, int M, int Dch, int L, int Q,
                                       int P, float* __restrict__ grad_dist, float* __restrict__ grad_loc) {
  const long long total = (long long)B * Q * M * L * P;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int l = (int)((i / P) % L);
    const int m = (int)((i / ((long long)P * L)) % M);
    const int b = (int)(i / ((long long)P * L * M * Q));
    const LevelInfo li = level_info(shapes3d, 3, lsi, l);
    const float* lp = loc + i * 3;
    const Tap t = make_tap(lp[0], lp[1], lp[2], li.H, li.W, li.D);
    float gz = 0.f;
    if (t.in3d) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(grad_out) + i);
      const float g[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (t.pix[k] >= 0) {
          const size_t o = (((size_t)b * S + li.start + t.pix[k]) * M + m) * Dch;
          float lo, hi;
          depth_score(t, dist + o, li.D, lo, hi);
          gz += g[k] * (hi - lo);
          if (t.d0 >= 0) red_add1(grad_dist + o + t.d0, t.hd * g[k]);
          if (t.d0 + 1 <= li.D - 1) red_add1(grad_dist + o + t.d0 + 1, t.ld * g[k]);
        }
      }
    }
    grad_loc[i * 3 + 0] = 0.f;  // DSK:238-240
    grad_loc[i * 3 + 1] = 0.f;
    grad_loc[i * 3 + 2] = (float)li.D * gz;
  }
}

// ---- weighted deformable attention forward (optionally fused with the depth-score sampling) --------
// FUSED: depth scores are computed from `dist` (and optionally stored to ds_io); otherwise read from ds_io.
template <bool FUSED>
__global__ void __launch_bounds__(256) wms_fwd_kernel(const float* __restrict__ value, const float* __restrict__ dist,
                                                      const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                                                      const float* __restrict__ loc, const float* __restrict__ attn,
                                                      float* __restrict__ ds_io, int B, int S, int M, int Cm, int Dch,
                                                      int L, int Q, int P, float* __restrict__ out) {
  constexpr int LS = FUSED ? 3 : 2;  // floats per location / ints per shape row
  const int lane = threadIdx.x & 31;
  const long long items = (long long)B * Q * M;
  const long long w0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long wstride = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long it = w0; it < items; it += wstride) {
    const int m = (int)(it % M);
    const int b = (int)(it / ((long long)M * Q));
    for (int c0 = 0; c0 < Cm; c0 += 32) {
      const int c = c0 + lane;
      float acc = 0.f;
      for (int l = 0; l < L; ++l) {
        const LevelInfo li = level_info(shapes, LS, lsi, l);
        for (int p = 0; p < P; ++p) {
          const long long sp = (it * L + l) * P + p;
          const float* lp = loc + sp * LS;
          const Tap t = make_tap(lp[0], lp[1], FUSED ? lp[2] : 0.5f / (float)1, li.H, li.W, FUSED ? li.D : 1);
          if (!t.in2d) {
            if (FUSED && ds_io && c0 == 0 && lane < 4) ds_io[sp * 4 + lane] = 0.f;
            continue;
          }
          float ds[4];
          if (FUSED) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              ds[k] = 0.f;
              if (t.in3d && t.pix[k] >= 0) {
                float lo, hi;
                ds[k] = depth_score(t, dist + (((size_t)b * S + li.start + t.pix[k]) * M + m) * Dch, li.D, lo, hi);
              }
            }
            if (ds_io && c0 == 0 && lane < 4) ds_io[sp * 4 + lane] = lane == 0 ? ds[0] : lane == 1 ? ds[1] : lane == 2 ? ds[2] : ds[3];
          } else {
            const float4 d4 = __ldg(reinterpret_cast<const float4*>(ds_io) + sp);
            ds[0] = d4.x; ds[1] = d4.y; ds[2] = d4.z; ds[3] = d4.w;
          }
          const float a = __ldg(attn + sp);
          float val = 0.f;
          if (c < Cm) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (t.pix[k] >= 0)
                val += (t.bw[k] * ds[k]) * __ldg(value + (((size_t)b * S + li.start + t.pix[k]) * M + m) * Cm + c);
          }
          acc += val * a;
        }
      }
      if (c < Cm) out[it * Cm + c] = acc;
    }
  }
}

// ---- backward (optionally fused with the depth-score backward) -------------------------------------
template <bool FUSED>
__global__ void __launch_bounds__(256) wms_bwd_kernel(const float* __restrict__ value, const float* __restrict__ dist,
                                                      const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                                                      const float* __restrict__ loc, const float* __restrict__ attn,
                                                      const float* __restrict__ ds_in, const float* __restrict__ grad_out,
                                                      int B, int S, int M, int Cm, int Dch, int L, int Q, int P,
                                                      float* __restrict__ grad_value, float* __restrict__ grad_dist,
                                                      float* __restrict__ grad_loc, float* __restrict__ grad_attn,
                                                      float* __restrict__ grad_ds) {
  constexpr int LS = FUSED ? 3 : 2;
  const int lane = threadIdx.x & 31;
  const long long items = (long long)B * Q * M;
  const long long w0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long wstride = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long it = w0; it < items; it += wstride) {
    const int m = (int)(it % M);
    const int b = (int)(it / ((long long)M * Q));
    for (int l = 0; l < L; ++l) {
      const LevelInfo li = level_info(shapes, LS, lsi, l);
      for (int p = 0; p < P; ++p) {
        const long long sp = (it * L + l) * P + p;
        const float* lp = loc + sp * LS;
        const Tap t = make_tap(lp[0], lp[1], FUSED ? lp[2] : 0.5f, li.H, li.W, FUSED ? li.D : 1);
        const float a = __ldg(attn + sp);
        float ds[4] = {0.f, 0.f, 0.f, 0.f}, dlo[4] = {0.f, 0.f, 0.f, 0.f}, dhi[4] = {0.f, 0.f, 0.f, 0.f};
        if (FUSED) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (t.in3d && t.pix[k] >= 0)
              ds[k] = depth_score(t, dist + (((size_t)b * S + li.start + t.pix[k]) * M + m) * Dch, li.D, dlo[k], dhi[k]);
        } else {
          const float4 d4 = __ldg(reinterpret_cast<const float4*>(ds_in) + sp);
          ds[0] = d4.x; ds[1] = d4.y; ds[2] = d4.z; ds[3] = d4.w;
        }
        float dot[4] = {0.f, 0.f, 0.f, 0.f};
        if (t.in2d) {
          for (int c = lane; c < Cm; c += 32) {
            const float g = __ldg(grad_out + it * Cm + c);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (t.pix[k] >= 0) {
                const size_t o = (((size_t)b * S + li.start + t.pix[k]) * M + m) * Cm + c;
                dot[k] += __ldg(value + o) * g;
                const float wgt = t.bw[k] * ds[k] * a;
                if (wgt != 0.f) red_add1(grad_value + o, wgt * g);
              }
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) dot[k] = warp_sum(dot[k]);
        }
        if (lane == 0) {
          float g_attn = 0.f, g_w = 0.f, g_h = 0.f, g_d = 0.f;
          float gds[4] = {0.f, 0.f, 0.f, 0.f};
          if (t.in2d) {
            const float hh = 1.f - t.lh, hw = 1.f - t.lw;
            const float e0 = ds[0] * dot[0], e1 = ds[1] * dot[1], e2 = ds[2] * dot[2], e3 = ds[3] * dot[3];
            g_attn = t.bw[0] * e0 + t.bw[1] * e1 + t.bw[2] * e2 + t.bw[3] * e3;
            g_w = (float)li.W * a * (-hh * e0 + hh * e1 + t.lh * e2 - t.lh * e3);
            g_h = (float)li.H * a * (-hw * e0 - t.lw * e1 + t.lw * e2 + hw * e3);
#pragma unroll
            for (int k = 0; k < 4; ++k) gds[k] = (t.pix[k] >= 0) ? a * t.bw[k] * dot[k] : 0.f;
          }
          if (FUSED) {
            if (t.in3d) {
              float gz = 0.f;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (t.pix[k] >= 0) {
                  const size_t o = (((size_t)b * S + li.start + t.pix[k]) * M + m) * Dch;
                  gz += gds[k] * (dhi[k] - dlo[k]);
                  if (t.d0 >= 0) red_add1(grad_dist + o + t.d0, t.hd * gds[k]);
                  if (t.d0 + 1 <= li.D - 1) red_add1(grad_dist + o + t.d0 + 1, t.ld * gds[k]);
                }
              }
              g_d = (float)li.D * gz;
            }
            grad_loc[sp * 3 + 0] = g_w; grad_loc[sp * 3 + 1] = g_h; grad_loc[sp * 3 + 2] = g_d;
          } else {
            grad_loc[sp * 2 + 0] = g_w; grad_loc[sp * 2 + 1] = g_h;
            reinterpret_cast<float4*>(grad_ds)[sp] = make_float4(gds[0], gds[1], gds[2], gds[3]);
          }
          grad_attn[sp] = g_attn;
        }
      }
    }
  }
}

// ---- one-stage operator, lane-parallel points (Cm % 4 == 0) -----------------------------------------
// The generic kernels above resolve every sampling point in ALL 32 lanes (make_tap + 8 depth loads, serially over the
// L*P points) and gather one float per lane.  Here a warp still owns one (b, q, m), but the points are spread over the
// lanes for the parameter stage (lane = point: tap, depth scores, weights -- 32 points resolved at once), and the gather
// stage walks the points with the lanes split as 4 corners x 8 four-channel groups: one 16-byte load per lane and point
// covers all four corners of a 32-channel slice, the weights arrive by shuffle.  Backward: vector reductions
// (red.global.add.v4.f32) into grad_value instead of scalar ones.
struct PointTaps {
  long long row[4];   // (b*S + level start + pixel) of the corner, or -1
  float w[4];         // attention weight * bilinear weight * depth score
};

__device__ __forceinline__ void lanes_point(const float* __restrict__ dist, const int64_t* __restrict__ shapes3d,
                                            const int64_t* __restrict__ lsi, const float* __restrict__ loc,
                                            const float* __restrict__ attn, long long it, int pt, int npts, int b, int m, int S,
                                            int M, int Dch, int P, Tap& t, LevelInfo& li, float (&ds)[4], float (&dlo)[4],
                                            float (&dhi)[4], float& a, PointTaps& pt_out) {
#pragma unroll
  for (int k = 0; k < 4; ++k) { ds[k] = 0.f; dlo[k] = 0.f; dhi[k] = 0.f; pt_out.row[k] = -1; pt_out.w[k] = 0.f; }
  a = 0.f;
  t.in2d = t.in3d = false;
  li.H = li.W = li.D = 1; li.start = 0;
  if (pt >= npts) return;
  const int l = pt / P;
  li = level_info(shapes3d, 3, lsi, l);
  const long long sp = it * npts + pt;
  const float* lp = loc + sp * 3;
  t = make_tap(__ldg(lp), __ldg(lp + 1), __ldg(lp + 2), li.H, li.W, li.D);
  a = __ldg(attn + sp);
  if (!t.in2d) return;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (t.pix[k] >= 0) {
      const long long row = (long long)b * S + li.start + t.pix[k];
      pt_out.row[k] = row;
      if (t.in3d) ds[k] = depth_score(t, dist + ((size_t)row * M + m) * Dch, li.D, dlo[k], dhi[k]);
      pt_out.w[k] = a * (t.bw[k] * ds[k]);
    }
  }
}

__global__ void __launch_bounds__(256) fused_fwd_lanes_kernel(const float* __restrict__ value, const float* __restrict__ dist,
                                                              const int64_t* __restrict__ shapes3d,
                                                              const int64_t* __restrict__ lsi, const float* __restrict__ loc,
                                                              const float* __restrict__ attn, float* __restrict__ ds_out, int B,
                                                              int S, int M, int Cm, int Dch, int L, int Q, int P,
                                                              float* __restrict__ out) {
  const int lane = threadIdx.x & 31, k = lane >> 3, cg = lane & 7;
  const int npts = L * P;
  const long long items = (long long)B * Q * M;
  const long long w0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long wstride = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long it = w0; it < items; it += wstride) {
    const int m = (int)(it % M);
    const int b = (int)(it / ((long long)M * Q));
    for (int c0 = 0; c0 < Cm; c0 += 32) {
      const int c = c0 + cg * 4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int p0 = 0; p0 < npts; p0 += 32) {
        Tap t; LevelInfo li; float ds[4], dlo[4], dhi[4], a; PointTaps pt;
        lanes_point(dist, shapes3d, lsi, loc, attn, it, p0 + lane, npts, b, m, S, M, Dch, P, t, li, ds, dlo, dhi, a, pt);
        if (ds_out && c0 == 0 && p0 + lane < npts)
          reinterpret_cast<float4*>(ds_out)[it * npts + p0 + lane] = make_float4(ds[0], ds[1], ds[2], ds[3]);
        const int n = min(32, npts - p0);
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
          // corner k of point j: row and weight from the lane that resolved the point
          long long row = -1;
          float w = 0.f;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const long long r = __shfl_sync(SGC_FULL_MASK, pt.row[kk], j);
            const float ww = __shfl_sync(SGC_FULL_MASK, pt.w[kk], j);
            if (kk == k) { row = r; w = ww; }
          }
          if (row >= 0 && w != 0.f && c < Cm) {
            const float4 x = ldg4(value + ((size_t)row * M + m) * Cm + c);
            acc.x += w * x.x; acc.y += w * x.y; acc.z += w * x.z; acc.w += w * x.w;
          }
        }
      }
      // sum over the four corner groups
#pragma unroll
      for (int o = 8; o < 32; o <<= 1) {
        acc.x += __shfl_xor_sync(SGC_FULL_MASK, acc.x, o); acc.y += __shfl_xor_sync(SGC_FULL_MASK, acc.y, o);
        acc.z += __shfl_xor_sync(SGC_FULL_MASK, acc.z, o); acc.w += __shfl_xor_sync(SGC_FULL_MASK, acc.w, o);
      }
      if (k == 0 && c < Cm) *reinterpret_cast<float4*>(out + it * Cm + c) = acc;
    }
  }
}

__global__ void __launch_bounds__(256) fused_bwd_lanes_kernel(const float* __restrict__ value, const float* __restrict__ dist,
                                                              const int64_t* __restrict__ shapes3d,
                                                              const int64_t* __restrict__ lsi, const float* __restrict__ loc,
                                                              const float* __restrict__ attn, const float* __restrict__ grad_out,
                                                              int B, int S, int M, int Cm, int Dch, int L, int Q, int P,
                                                              float* __restrict__ grad_value, float* __restrict__ grad_dist,
                                                              float* __restrict__ grad_loc, float* __restrict__ grad_attn) {
  const int lane = threadIdx.x & 31, k = lane >> 3, cg = lane & 7;
  const int npts = L * P;
  const long long items = (long long)B * Q * M;
  const long long w0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long wstride = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long it = w0; it < items; it += wstride) {
    const int m = (int)(it % M);
    const int b = (int)(it / ((long long)M * Q));
    for (int p0 = 0; p0 < npts; p0 += 32) {
      Tap t; LevelInfo li; float ds[4], dlo[4], dhi[4], a; PointTaps pt;
      lanes_point(dist, shapes3d, lsi, loc, attn, it, p0 + lane, npts, b, m, S, M, Dch, P, t, li, ds, dlo, dhi, a, pt);
      float dot[4] = {0.f, 0.f, 0.f, 0.f};   // of THIS lane's point, filled while the warp walks the points
      const int n = min(32, npts - p0);
      for (int c0 = 0; c0 < Cm; c0 += 32) {
        const int c = c0 + cg * 4;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < Cm) g = ldg4(grad_out + it * Cm + c);
#pragma unroll 2
        for (int j = 0; j < n; ++j) {
          long long row = -1;
          float w = 0.f;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const long long r = __shfl_sync(SGC_FULL_MASK, pt.row[kk], j);
            const float ww = __shfl_sync(SGC_FULL_MASK, pt.w[kk], j);
            if (kk == k) { row = r; w = ww; }
          }
          float d = 0.f;
          if (row >= 0 && c < Cm) {
            const size_t o = ((size_t)row * M + m) * Cm + c;
            const float4 x = ldg4(value + o);
            d = x.x * g.x + x.y * g.y + x.z * g.z + x.w * g.w;
            if (w != 0.f) red_add4(grad_value + o, w * g.x, w * g.y, w * g.z, w * g.w);
          }
          d += __shfl_xor_sync(SGC_FULL_MASK, d, 1);
          d += __shfl_xor_sync(SGC_FULL_MASK, d, 2);
          d += __shfl_xor_sync(SGC_FULL_MASK, d, 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const float dk = __shfl_sync(SGC_FULL_MASK, d, kk * 8);
            if (lane == j) dot[kk] += dk;
          }
        }
      }
      if (p0 + lane < npts) {
        const long long sp = it * npts + p0 + lane;
        float g_attn = 0.f, g_w = 0.f, g_h = 0.f, g_d = 0.f;
        if (t.in2d) {
          const float hh = 1.f - t.lh, hw = 1.f - t.lw;
          const float e0 = ds[0] * dot[0], e1 = ds[1] * dot[1], e2 = ds[2] * dot[2], e3 = ds[3] * dot[3];
          g_attn = t.bw[0] * e0 + t.bw[1] * e1 + t.bw[2] * e2 + t.bw[3] * e3;
          g_w = (float)li.W * a * (-hh * e0 + hh * e1 + t.lh * e2 - t.lh * e3);
          g_h = (float)li.H * a * (-hw * e0 - t.lw * e1 + t.lw * e2 + hw * e3);
          if (t.in3d) {
            float gz = 0.f;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              if (t.pix[kk] >= 0) {
                const float gds = a * t.bw[kk] * dot[kk];
                const size_t o = ((size_t)pt.row[kk] * M + m) * Dch;
                gz += gds * (dhi[kk] - dlo[kk]);
                if (t.d0 >= 0) red_add1(grad_dist + o + t.d0, t.hd * gds);
                if (t.d0 + 1 <= li.D - 1) red_add1(grad_dist + o + t.d0 + 1, t.ld * gds);
              }
            }
            g_d = (float)li.D * gz;
          }
        }
        grad_loc[sp * 3 + 0] = g_w; grad_loc[sp * 3 + 1] = g_h; grad_loc[sp * 3 + 2] = g_d;
        grad_attn[sp] = g_attn;
      }
    }
  }
}

static inline int op_grid(long long warps_needed) {
  long long blocks = (warps_needed + 7) / 8;
  const long long cap = 148LL * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace sgc

extern "C" int dfa3d_depth_score_fwd(const float* dist, const int64_t* shapes3d, const int64_t* lsi, const float* loc,
                                     int B, int S, int M, int D, int L, int Q, int P, float* out, void* stream) {
  const long long total = (long long)B * Q * M * L * P;
  if (total == 0) return 0;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  sgc::depth_score_fwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(dist, shapes3d, lsi, loc, B, S, M, D, L, Q, P, out);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

extern "C" int dfa3d_depth_score_bwd(const float* dist, const int64_t* shapes3d, const int64_t* lsi, const float* loc,
                                     const float* grad_out, int B, int S, int M, int D, int L, int Q, int P,
                                     float* grad_dist, float* grad_loc, void* stream) {
  const long long total = (long long)B * Q * M * L * P;
  if (total == 0) return 0;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  sgc::depth_score_bwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(dist, shapes3d, lsi, loc, grad_out, B, S, M, D, L, Q, P,
                                                                             grad_dist, grad_loc);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

extern "C" int dfa3d_wms_fwd(const float* value, const int64_t* shapes2d, const int64_t* lsi, const float* loc2d,
                             const float* attn, const float* depth_score, int B, int S, int M, int Cm, int L, int Q,
                             int P, float* out, void* stream) {
  const long long items = (long long)B * Q * M;
  if (items == 0) return 0;
  sgc::wms_fwd_kernel<false><<<sgc::op_grid(items), 256, 0, (cudaStream_t)stream>>>(
      value, nullptr, shapes2d, lsi, loc2d, attn, const_cast<float*>(depth_score), B, S, M, Cm, 0, L, Q, P, out);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

extern "C" int dfa3d_wms_bwd(const float* value, const int64_t* shapes2d, const int64_t* lsi, const float* loc2d,
                             const float* attn, const float* depth_score, const float* grad_out, int B, int S, int M,
                             int Cm, int L, int Q, int P, float* grad_value, float* grad_loc2d, float* grad_attn,
                             float* grad_depth_score, void* stream) {
  const long long items = (long long)B * Q * M;
  if (items == 0) return 0;
  sgc::wms_bwd_kernel<false><<<sgc::op_grid(items), 256, 0, (cudaStream_t)stream>>>(
      value, nullptr, shapes2d, lsi, loc2d, attn, depth_score, grad_out, B, S, M, Cm, 0, L, Q, P, grad_value, nullptr,
      grad_loc2d, grad_attn, grad_depth_score);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

extern "C" int dfa3d_fused_fwd(const float* value, const float* dist, const int64_t* shapes3d, const int64_t* lsi,
                               const float* loc, const float* attn, int B, int S, int M, int Cm, int D, int L, int Q,
                               int P, float* out, float* depth_score_out, void* stream) {
  const long long items = (long long)B * Q * M;
  if (items == 0) return 0;
  if (Cm % 4 == 0 && (reinterpret_cast<uintptr_t>(value) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
    sgc::fused_fwd_lanes_kernel<<<sgc::op_grid(items), 256, 0, (cudaStream_t)stream>>>(value, dist, shapes3d, lsi, loc, attn,
                                                                                      depth_score_out, B, S, M, Cm, D, L, Q, P, out);
  else
    sgc::wms_fwd_kernel<true><<<sgc::op_grid(items), 256, 0, (cudaStream_t)stream>>>(
        value, dist, shapes3d, lsi, loc, attn, depth_score_out, B, S, M, Cm, D, L, Q, P, out);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

extern "C" int dfa3d_fused_bwd(const float* value, const float* dist, const int64_t* shapes3d, const int64_t* lsi,
                               const float* loc, const float* attn, const float* grad_out, int B, int S, int M, int Cm,
                               int D, int L, int Q, int P, float* grad_value, float* grad_dist, float* grad_loc,
                               float* grad_attn, void* stream) {
  const long long items = (long long)B * Q * M;
  if (items == 0) return 0;
  if (Cm % 4 == 0 && (reinterpret_cast<uintptr_t>(value) & 15) == 0 && (reinterpret_cast<uintptr_t>(grad_out) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(grad_value) & 15) == 0)
    sgc::fused_bwd_lanes_kernel<<<sgc::op_grid(items), 256, 0, (cudaStream_t)stream>>>(
        value, dist, shapes3d, lsi, loc, attn, grad_out, B, S, M, Cm, D, L, Q, P, grad_value, grad_dist, grad_loc, grad_attn);
  else
    sgc::wms_bwd_kernel<true><<<sgc::op_grid(items), 256, 0, (cudaStream_t)stream>>>(
        value, dist, shapes3d, lsi, loc, attn, nullptr, grad_out, B, S, M, Cm, D, L, Q, P, grad_value, grad_dist,
        grad_loc, grad_attn, nullptr);
  SGC_CUDA_CHECK_LAST();
  return 0;
}
