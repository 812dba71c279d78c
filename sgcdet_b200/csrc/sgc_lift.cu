// sgc_lift_{fwd,bwd}: the sampling side of the geometry-and-context-aware aggregation, one warp per
// visible (view, voxel) pair, pairs in view-major order so one view's maps stay L2-resident.
//
// Replaces, for every visible pair (v,q) of one DenseHead level:
//   Grid_Sample_3D_Feature            deformable_cross_attention.py:67-116   (ref-point DFA3D sample, M=1,P=1)
//   offset / depth-offset / weight    deformable_cross_attention.py:423-436  (the three Linear layers + softmax)
//   sampling-location arithmetic      deformable_cross_attention.py:445-461
//   MultiScale3DDeformableAttn (M=8, P=4, L=1)   multi_scale_3ddeformable_attn_function.py:277-351
//     = ms_depth_score_sample  (ms_depth_score_sample_cuda_kernel.cuh:24-148, bwd :150-327)
//     + wms_deform_attn        (wms_deform_attn_cuda_kernel.cuh:24-80,240-303, bwd :82-159,305-531)
//
// Data layout (HBM):
//   value  [V,S,ldv]  channel-last projected features WITHOUT the value_proj bias; head m owns channels
//                     [m*Cm,(m+1)*Cm).  The bias is applied here as  b[c] * (sum of tap weights)  which is
//                     what zero padding of (W f + b) gives.
//   G      [V,S,ldg]  the raw feature map pushed through the three small Linear layers (linear maps commute
//                     with the sampling): 4*M*P channels ordered [m][p][off_x, off_y, off_d, logit].
//   dist   [V,S,D]    depth distribution, channel-last (the same for every head, DCA:422).
//   samp   [cap,32,4] per pair: (loc_x, loc_y, loc_z, attn) for (m,p) = (lane>>2, lane&3); saved for the bwd.
//   slots  [cap,C]    per-pair output.
// Lane mapping: lane = 4*m + sub; the lane owns sampling point p=sub of head m in the parameter stage and
// channels [lane*CPL, lane*CPL+CPL) (CPL = C/32 = Cm/4) in the gather stage.
#include "common.cuh"

#include <cstdlib>

namespace sgc {

template <int CPL>
struct Vec { float v[CPL]; };

// Channel ownership inside a row of C = 32*CPL channels (head m = lane >> 2 owns channels [m*4*CPL, (m+1)*4*CPL)):
// register chunk j/4 of a lane is the float4 at  lane_base + (j/4)*16,  lane_base = m*4*CPL + (lane & 3)*4,
// so the four lanes of a head cover 64 CONTIGUOUS bytes per instruction (whole 32-byte sectors for the gathers and,
// above all, for the RED.v4 reductions of the backward: half the L2 atomic sector operations of a
// "CPL contiguous channels per lane" layout).
template <int CPL>
__device__ __forceinline__ int lane_base(int lane) { return (lane >> 2) * (4 * CPL) + (lane & 3) * 4; }

template <int CPL>
__device__ __forceinline__ void load_row(float (&dst)[CPL], const float* p) {
#pragma unroll
  for (int j = 0; j < CPL; j += 4) {
    const float4 t = ldg4(p + j * 4);
    dst[j] = t.x; dst[j + 1] = t.y; dst[j + 2] = t.z; dst[j + 3] = t.w;
  }
}

// Stage 1+2: recompute (or compute) the per-lane sampling point.  Returns loc/attn of (m,p) = (lane>>2, lane&3).
__device__ __forceinline__ float4 lift_params(const float* __restrict__ G, int ldg, const float* __restrict__ dist,
                                              const float* __restrict__ gbias, size_t vS, float rx, float ry, float rz,
                                              int H, int W, int D, int lane) {
  const Tap tr = make_tap(rx, ry, rz, H, W, D);
  float4 g = ldg4(gbias + lane * 4);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (tr.pix[k] >= 0 && tr.in3d) {
      float lo, hi;
      const float ds = depth_score(tr, dist + (vS + tr.pix[k]) * D, D, lo, hi);
      const float wk = tr.bw[k] * ds;
      const float4 t = ldg4(G + (vS + tr.pix[k]) * ldg + lane * 4);
      g.x += wk * t.x; g.y += wk * t.y; g.z += wk * t.z; g.w += wk * t.w;
    }
  }
  // softmax over the 4 points of the head (DCA:431)
  const float mx = quad_max(g.w);
  const float e = expf(g.w - mx);
  const float sum = quad_sum(e);
  float4 r;
  r.x = rx + __fdiv_rn(g.x, (float)W);   // DCA:445-455: offsets / (W,H,D), then + reference point
  r.y = ry + __fdiv_rn(g.y, (float)H);
  r.z = rz + __fdiv_rn(g.z, (float)D);
  r.w = e / sum;
  return r;
}

template <int CPL>
__global__ void __launch_bounds__(256, 4) lift_fwd_kernel(
    const float* __restrict__ value, int ldv, const float* __restrict__ G, int ldg,
    const float* __restrict__ dist, const float* __restrict__ vbias, const float* __restrict__ gbias,
    const int* __restrict__ pair_vq, const int* __restrict__ n_pairs_ptr, const float* __restrict__ ref_cam,
    int S, int H, int W, int D, int Q, float* __restrict__ samp, float* __restrict__ slots) {
  constexpr int C = CPL * 32;
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int n_pairs = __ldg(n_pairs_ptr);
  // the pair id and its reference point are fetched one iteration ahead: two dependent global round trips less in front of
  // every pair's (dependent) G -> depth -> value gathers
  const int stride = gridDim.x * warps_per_block;
  int pair = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  int flat_n = 0;
  float rx_n = 0.f, ry_n = 0.f, rz_n = 0.f;
  if (pair < n_pairs) {
    flat_n = __ldg(pair_vq + pair);
    rx_n = __ldg(ref_cam + (size_t)flat_n * 3); ry_n = __ldg(ref_cam + (size_t)flat_n * 3 + 1);
    rz_n = __ldg(ref_cam + (size_t)flat_n * 3 + 2);
  }
  for (; pair < n_pairs; pair += stride) {
    const int flat = flat_n;
    const float rx = rx_n, ry = ry_n, rz = rz_n;
    if (pair + stride < n_pairs) {
      flat_n = __ldg(pair_vq + pair + stride);
      rx_n = __ldg(ref_cam + (size_t)flat_n * 3); ry_n = __ldg(ref_cam + (size_t)flat_n * 3 + 1);
      rz_n = __ldg(ref_cam + (size_t)flat_n * 3 + 2);
    }
    const int v = flat / Q;
    const size_t vS = (size_t)v * S;
    const float4 sp = lift_params(G, ldg, dist, gbias, vS, rx, ry, rz, H, W, D, lane);
    reinterpret_cast<float4*>(samp)[(size_t)pair * 32 + lane] = sp;

    // this lane's sampling point -> 4 corner weights (bilinear * depth score), zero when invalid
    const Tap t = make_tap(sp.x, sp.y, sp.z, H, W, D);
    float cw[4];
    int px[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      cw[k] = 0.f;
      px[k] = t.pix[k];
      if (t.pix[k] >= 0 && t.in3d) {
        float lo, hi;
        cw[k] = t.bw[k] * depth_score(t, dist + (vS + t.pix[k]) * D, D, lo, hi);
      }
    }
    float acc[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) acc[j] = 0.f;
    float wsum = 0.f;  // sum over the head's taps of attn * cw   (for the value_proj bias)
    const float* vbase = value + lane_base<CPL>(lane);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int src = (lane & ~3) | p;
      const float a = __shfl_sync(SGC_FULL_MASK, sp.w, src);
      int pk[4];
      float wk[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        pk[k] = __shfl_sync(SGC_FULL_MASK, px[k], src);
        wk[k] = __shfl_sync(SGC_FULL_MASK, cw[k], src);
      }
      // the four corner rows of the point are fetched together (4 x CPL/4 independent 16-byte gathers in flight per lane:
      // the kernel is bound by the latency of these L2 gathers, not by their bandwidth)
      float x[4][CPL];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (wk[k] != 0.f) {
          load_row<CPL>(x[k], vbase + (vS + pk[k]) * ldv);
        } else {
#pragma unroll
          for (int j = 0; j < CPL; ++j) x[k][j] = 0.f;
        }
      }
      float val[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) val[j] = 0.f;
      float ws = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) val[j] += wk[k] * x[k][j];
        ws += wk[k];
      }
#pragma unroll
      for (int j = 0; j < CPL; ++j) acc[j] += val[j] * a;
      wsum += ws * a;
    }
    float* out = slots + (size_t)pair * C + lane_base<CPL>(lane);
#pragma unroll
    for (int j = 0; j < CPL; j += 4) {
      const float4 b = ldg4(vbias + lane_base<CPL>(lane) + j * 4);
      float4 o;
      o.x = acc[j] + b.x * wsum; o.y = acc[j + 1] + b.y * wsum;
      o.z = acc[j + 2] + b.z * wsum; o.w = acc[j + 3] + b.w * wsum;
      *reinterpret_cast<float4*>(out + j * 4) = o;
    }
  }
}

// Scatter the depth-score gradient of one corner into grad_dist (DSK:180-236).
__device__ __forceinline__ void scatter_dist(float* gd_px, const Tap& t, int D, float gds) {
  if (t.d0 >= 0) red_add1(gd_px + t.d0, t.hd * gds);
  if (t.d0 + 1 <= D - 1) red_add1(gd_px + t.d0 + 1, t.ld * gds);
}

template <int CPL, int MINB>
__global__ void __launch_bounds__(256, MINB) lift_bwd_kernel(
    const float* __restrict__ value, int ldv, const float* __restrict__ G, int ldg,
    const float* __restrict__ dist, const float* __restrict__ vbias,
    const int* __restrict__ pair_vq, const int* __restrict__ n_pairs_ptr, const float* __restrict__ ref_cam,
    const float* __restrict__ samp, const float* __restrict__ grad_slots,
    int S, int H, int W, int D, int Q,
    float* __restrict__ grad_value, float* __restrict__ grad_G, float* __restrict__ grad_dist,
    float* __restrict__ grad_vbias, float* __restrict__ grad_gbias) {
  constexpr int C = CPL * 32;
  __shared__ float s_part[8][C + 128];
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int n_pairs = __ldg(n_pairs_ptr);
  float gvb[CPL];  // per-warp partial of grad value_proj.bias
#pragma unroll
  for (int j = 0; j < CPL; ++j) gvb[j] = 0.f;
  float4 ggb = make_float4(0.f, 0.f, 0.f, 0.f);  // per-warp partial of grad of the G bias
  float vb[CPL];
  load_row<CPL>(vb, vbias + lane_base<CPL>(lane));

  // (fetching the next pair's id and sampling point one iteration ahead, as lift_fwd does, measured slightly slower here:
  // 597.7 vs 600.6 volumes/s -- the kernel sits at its register limit)
  for (int pair = blockIdx.x * warps_per_block + (threadIdx.x >> 5); pair < n_pairs;
       pair += gridDim.x * warps_per_block) {
    const int flat = __ldg(pair_vq + pair);
    const float4 sp = ldg4(samp + ((size_t)pair * 32 + lane) * 4);
    const int v = flat / Q;
    const size_t vS = (size_t)v * S;
    const Tap t = make_tap(sp.x, sp.y, sp.z, H, W, D);
    float ds[4], dlo[4], dhi[4], cw[4];
    int px[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ds[k] = 0.f; dlo[k] = 0.f; dhi[k] = 0.f;
      px[k] = t.pix[k];
      if (t.pix[k] >= 0 && t.in3d) ds[k] = depth_score(t, dist + (vS + t.pix[k]) * D, D, dlo[k], dhi[k]);
      cw[k] = t.bw[k] * ds[k];
    }
    float g[CPL];
    load_row<CPL>(g, grad_slots + (size_t)pair * C + lane_base<CPL>(lane));
    float gb = 0.f;  // sum_j vbias[j] * g[j] over this lane's channels
#pragma unroll
    for (int j = 0; j < CPL; ++j) gb += vb[j] * g[j];

    // d out / d (tap weight) for this lane's own point, gathered while the head walks its 4 points
    float dot[4] = {0.f, 0.f, 0.f, 0.f};
    float wsum = 0.f;
    const float* vbase = value + lane_base<CPL>(lane);
    float* gvbase = grad_value + lane_base<CPL>(lane);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int src = (lane & ~3) | p;
      const float a = __shfl_sync(SGC_FULL_MASK, sp.w, src);
      int pk[4];
      float wk[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        pk[k] = __shfl_sync(SGC_FULL_MASK, px[k], src);
        wk[k] = __shfl_sync(SGC_FULL_MASK, cw[k], src);
      }
      // the four corner rows of this point are fetched together (4 independent gathers in flight per lane)
      float x[4][CPL];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (pk[k] >= 0) {
          load_row<CPL>(x[k], vbase + (vS + pk[k]) * ldv);
        } else {
#pragma unroll
          for (int j = 0; j < CPL; ++j) x[k][j] = 0.f;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float part = 0.f;
        if (pk[k] >= 0) {  // the corner exists: its value (plus bias) defines d out / d weight even when wk == 0
#pragma unroll
          for (int j = 0; j < CPL; ++j) part += x[k][j] * g[j];
          part += gb;
          const float wa = wk[k] * a;
          if (wa != 0.f) {
            float* dst = gvbase + (vS + pk[k]) * ldv;
#pragma unroll
            for (int j = 0; j < CPL; j += 4) red_add4(dst + j * 4, wa * g[j], wa * g[j + 1], wa * g[j + 2], wa * g[j + 3]);
            wsum += wa;
          }
        }
        part = quad_sum(part);
        if ((lane & 3) == p) dot[k] = part;
      }
    }
#pragma unroll
    for (int j = 0; j < CPL; ++j) gvb[j] += wsum * g[j];

    // gradients of this lane's (m,p) sampling parameters
    float g_attn = 0.f, g_w = 0.f, g_h = 0.f, g_d = 0.f;
    if (t.in3d) {
      const float hh = 1.f - t.lh, hw = 1.f - t.lw;
      // WMSK:116-158 (bilinear), DSK:193-240 (depth)
      g_attn = cw[0] * dot[0] + cw[1] * dot[1] + cw[2] * dot[2] + cw[3] * dot[3];
      const float e0 = ds[0] * dot[0], e1 = ds[1] * dot[1], e2 = ds[2] * dot[2], e3 = ds[3] * dot[3];
      g_w = sp.w * (-hh * e0 + hh * e1 + t.lh * e2 - t.lh * e3) * (float)W;
      g_h = sp.w * (-hw * e0 - t.lw * e1 + t.lw * e2 + hw * e3) * (float)H;
      float gz = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (px[k] >= 0) {
          const float gds = sp.w * t.bw[k] * dot[k];
          gz += gds * (dhi[k] - dlo[k]);
          scatter_dist(grad_dist + (vS + px[k]) * D, t, D, gds);
        }
      }
      g_d = gz * (float)D;
    }
    // softmax backward over the head's 4 points, then offsets = grad_loc / (W,H,D)
    const float sdot = quad_sum(sp.w * g_attn);
    float4 gr;
    gr.x = __fdiv_rn(g_w, (float)W);
    gr.y = __fdiv_rn(g_h, (float)H);
    gr.z = __fdiv_rn(g_d, (float)D);
    gr.w = sp.w * (g_attn - sdot);
    ggb.x += gr.x; ggb.y += gr.y; ggb.z += gr.z; ggb.w += gr.w;

    // backward of the reference-point sample of G (weights 1, M=1, P=1)
    const float rx = __ldg(ref_cam + (size_t)flat * 3), ry = __ldg(ref_cam + (size_t)flat * 3 + 1),
                rz = __ldg(ref_cam + (size_t)flat * 3 + 2);
    const Tap tr = make_tap(rx, ry, rz, H, W, D);
    if (tr.in3d) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (tr.pix[k] >= 0) {
          float lo, hi;
          const float dsr = depth_score(tr, dist + (vS + tr.pix[k]) * D, D, lo, hi);
          const float wk = tr.bw[k] * dsr;
          const float4 x = ldg4(G + (vS + tr.pix[k]) * ldg + lane * 4);
          float dk = x.x * gr.x + x.y * gr.y + x.z * gr.z + x.w * gr.w;
          dk = warp_sum(dk);
          if (wk != 0.f) red_add4(grad_G + (vS + tr.pix[k]) * ldg + lane * 4, wk * gr.x, wk * gr.y, wk * gr.z, wk * gr.w);
          if (lane == 0) scatter_dist(grad_dist + (vS + tr.pix[k]) * D, tr, D, tr.bw[k] * dk);
        }
      }
    }
  }
  // per-CTA reduction of the bias partials in shared memory first (same-address REDs from ~10^4 warps would serialise in L2)
  const int wid = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < CPL; ++j) s_part[wid][lane_base<CPL>(lane) + (j >> 2) * 16 + (j & 3)] = gvb[j];
  s_part[wid][C + lane * 4 + 0] = ggb.x; s_part[wid][C + lane * 4 + 1] = ggb.y;
  s_part[wid][C + lane * 4 + 2] = ggb.z; s_part[wid][C + lane * 4 + 3] = ggb.w;
  __syncthreads();
  // one reduction per CTA and channel straight into the two bias gradients (~1200 per address over the whole kernel,
  // fire-and-forget): a separate reduce launch sat between this kernel and the projection's gradient kernels on the critical
  // path of the backward and waited ~70 us for a free SM slot; a last-CTA reduce of the per-CTA rows was latency-bound
  for (int c = threadIdx.x; c < C + 128; c += blockDim.x) {
    float a = 0.f;
    for (int w = 0; w < warps_per_block; ++w) a += s_part[w][c];
    if (a != 0.f) red_add1(c < C ? grad_vbias + c : grad_gbias + (c - C), a);
  }
}

}  // namespace sgc

static int lift_grid(int cap_pairs) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int want = (cap_pairs + 7) / 8;
  const int full = sms * 8;  // 8 CTAs of 8 warps per SM
  return want < full ? (want > 0 ? want : 1) : full;
}

extern "C" int sgc_lift_fwd(const float* value, int ldv, const float* G, int ldg, const float* dist,
                            const float* vbias, const float* gbias, const int* pair_vq, const int* n_pairs,
                            int cap_pairs, const float* ref_cam, int S, int H, int W, int D, int Q, int C,
                            float* samp, float* slots, void* stream) {
  if (C != 256 && C != 128) return (int)cudaErrorInvalidValue;
  if ((ldv & 3) || (ldg & 3)) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = lift_grid(cap_pairs);
  // 4 CTAs of 8 warps per SM (64 registers) measured better inside the step than 3 (80 registers) or 2: 591.5 / 588.2 / ~585
  // volumes/s (session U), although the kernel alone is fastest with 3
  if (C == 256)
    sgc::launch_chain(sgc::lift_fwd_kernel<8>, dim3(grid), dim3(256), 0, st, value, ldv, G, ldg, dist, vbias, gbias, pair_vq, n_pairs,
                      ref_cam, S, H, W, D, Q, samp, slots);
  else
    sgc::launch_chain(sgc::lift_fwd_kernel<4>, dim3(grid), dim3(256), 0, st, value, ldv, G, ldg, dist, vbias, gbias, pair_vq, n_pairs,
                      ref_cam, S, H, W, D, Q, samp, slots);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

extern "C" int sgc_lift_bwd(const float* value, int ldv, const float* G, int ldg, const float* dist,
                            const float* vbias, const int* pair_vq, const int* n_pairs, int cap_pairs,
                            const float* ref_cam, const float* samp, const float* grad_slots,
                            int S, int H, int W, int D, int Q, int C,
                            float* grad_value, float* grad_G, float* grad_dist, float* grad_vbias,
                            float* grad_gbias, float* scratch, void* stream) {
  if (C != 256 && C != 128) return (int)cudaErrorInvalidValue;
  if ((ldv & 3) || (ldg & 3)) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = lift_grid(cap_pairs);
  (void)scratch;   // kept in the signature (earlier versions reduced per-CTA rows through it)
#define SGC_LIFT_BWD_ARGS value, ldv, G, ldg, dist, vbias, pair_vq, n_pairs, ref_cam, samp, grad_slots, S, H, W, D, Q, \
                          grad_value, grad_G, grad_dist, grad_vbias, grad_gbias
  // 2 CTAs of 8 warps per SM at C = 256 (128 registers, no spills; 3 CTAs with 80 registers + spills measured the same or
  // slower: 589 vs 592 volumes/s), 3 at C = 128
  if (C == 256) sgc::lift_bwd_kernel<8, 2><<<grid, 256, 0, st>>>(SGC_LIFT_BWD_ARGS);
  else sgc::lift_bwd_kernel<4, 3><<<grid, 256, 0, st>>>(SGC_LIFT_BWD_ARGS);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

// floats of scratch sgc_lift_bwd needs for a given pair capacity
extern "C" int sgc_lift_bwd_scratch_floats(int cap_pairs, int C) { return lift_grid(cap_pairs) * (C + 128) + 4; }
