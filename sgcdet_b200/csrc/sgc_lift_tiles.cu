// sgc_lift_bwd_tiles: the backward of sgc_lift_fwd as a GATHER over pixel tiles (round 2), replacing the scatter of
// csrc/sgc_lift.cu (lift_bwd_kernel: ~45 warp-wide RED.v4 per pair into a zero-filled 290 MB grad_vg, bound by the SM's RED
// issue rate).  Reference kernels replaced: wms_deform_attn_cuda_kernel.cuh:82-159,305-531 (col2im with atomics into
// grad_value, shared-memory reductions for grad_sampling_loc / grad_attn_weight) and
// ms_depth_score_sample_cuda_kernel.cuh:150-327.
//
// Every sampling point (pair, head, point) touches a 2x2 block of pixels of its view.  The pixels of a view are cut into
// 4x8 tiles; a point is filed under every tile its block touches (1.3 on average), per head:
//   count  (warp per pair)  -> bin sizes            scan (one CTA) -> bin offsets
//   emit   (warp per pair)  -> per-corner (pixel, weight = attn * bilinear * depth score) of every point, saved once, and the
//                              point's record (pair*32 + head*4 + point) appended to its bins
//   tiles  (CTA per tile, warp per head): the tile's value rows are STAGED IN SHARED MEMORY with bulk async copies
//          (cp.async.bulk, one 1 KB row per pixel); for every record of the bin the warp reads the pair's 128-byte slice of
//          grad_slots ONCE and, per corner inside the tile, (a) dots it with the staged value row -> d out / d weight of that
//          corner (what the scatter kernel re-gathered 16 KB per pair from L2 for) and (b) accumulates weight * grad into
//          a shared-memory tile of grad_value.  The tile is written ONCE: no zero fill, no reductions into grad_vg.
//   params (warp per pair)  -> from the saved dots: gradients of attention logits / offsets / depth (softmax backward, WMSK:
//                              116-158, DSK:193-240), grad_dist (small: scalar REDs), bias partials, and the 128 parameter
//                              gradients `gr` of the pair
//   gtiles (CTA per tile)   -> the folded offset/weight map's gradient: the four reference-point corners of every pair filed
//                              under the tile, gr accumulated in shared memory, columns [C, C+128) of grad_vg written once.
#include "common.cuh"
#include "tc_common.cuh"

namespace sgc {

constexpr int kTH = 4, kTW = 8, kTPx = kTH * kTW;

template <int CPL>
__device__ __forceinline__ int lt_lane_base(int lane) { return (lane >> 2) * (4 * CPL) + (lane & 3) * 4; }

template <int CPL>
__device__ __forceinline__ void lt_load_row(float (&dst)[CPL], const float* p) {
#pragma unroll
  for (int j = 0; j < CPL; j += 4) {
    const float4 t = ldg4(p + j * 4);
    dst[j] = t.x; dst[j + 1] = t.y; dst[j + 2] = t.z; dst[j + 3] = t.w;
  }
}

// Tiles (index within the view) touched by the valid corners of a tap; the SAME function decides the counts and the
// emission, so the two passes agree exactly.
__device__ __forceinline__ int corner_tiles(const Tap& t, int W, int tiles_x, int (&ids)[4]) {
  int n = 0;
  if (!t.in3d) return 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (t.pix[k] >= 0) {
      const int y = t.pix[k] / W, x = t.pix[k] - y * W;
      const int id = (y / kTH) * tiles_x + x / kTW;
      bool dup = false;
#pragma unroll
      for (int q = 0; q < 4; ++q) dup |= (q < n && ids[q] == id);
      if (!dup) ids[n++] = id;
    }
  }
  return n;
}

// cnt: [V*T*8] value bins ((view, tile, head)) followed by [V*T] bins of the reference-point samples.
__global__ void __launch_bounds__(256) lt_count_kernel(const float* __restrict__ samp, const int* __restrict__ pair_vq,
                                                       const int* __restrict__ n_pairs_ptr, const float* __restrict__ ref_cam,
                                                       int H, int W, int D, int Q, int tiles_x, int T, int nb_val,
                                                       int* __restrict__ cnt) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int n_pairs = __ldg(n_pairs_ptr);
  for (int pair = blockIdx.x * wpb + (threadIdx.x >> 5); pair < n_pairs; pair += gridDim.x * wpb) {
    const int flat = __ldg(pair_vq + pair);
    const int v = flat / Q;
    const float4 sp = ldg4(samp + ((size_t)pair * 32 + lane) * 4);
    const Tap t = make_tap(sp.x, sp.y, sp.z, H, W, D);
    int ids[4];
    const int n = corner_tiles(t, W, tiles_x, ids);
    for (int q = 0; q < n; ++q) atomicAdd(cnt + ((size_t)v * T + ids[q]) * 8 + (lane >> 2), 1);
    if (lane == 0) {
      const Tap tr = make_tap(__ldg(ref_cam + (size_t)flat * 3), __ldg(ref_cam + (size_t)flat * 3 + 1),
                              __ldg(ref_cam + (size_t)flat * 3 + 2), H, W, D);
      const int m = corner_tiles(tr, W, tiles_x, ids);
      for (int q = 0; q < m; ++q) atomicAdd(cnt + nb_val + v * T + ids[q], 1);
    }
  }
}

// off[0..nb] = exclusive prefix of cnt[0..nb); cur[0..nb) = 0.  One CTA (nb <= ~150 k bins).
__global__ void __launch_bounds__(1024) lt_scan_kernel(const int* __restrict__ cnt, int nb, int* __restrict__ off,
                                                       int* __restrict__ cur) {
  __shared__ int warp_tot[32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int per = (nb + 1023) / 1024;
  const int b0 = tid * per, b1 = min(nb, b0 + per);
  int s = 0;
  for (int b = b0; b < b1; ++b) s += cnt[b];
  int inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(SGC_FULL_MASK, inc, o);
    if (lane >= o) inc += y;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int t = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(SGC_FULL_MASK, t, o);
      if (lane >= o) t += y;
    }
    warp_tot[lane] = t;
  }
  __syncthreads();
  int run = (wid ? warp_tot[wid - 1] : 0) + inc - s;
  for (int b = b0; b < b1; ++b) {
    off[b] = run;
    cur[b] = 0;
    run += cnt[b];
  }
  if (tid == 1023) off[nb] = warp_tot[31];
}

// Per-corner weights / pixels of every point (and of the reference-point sample), and the records of the bins.
__global__ void __launch_bounds__(256) lt_emit_kernel(const float* __restrict__ samp, const float* __restrict__ dist,
                                                      const int* __restrict__ pair_vq, const int* __restrict__ n_pairs_ptr,
                                                      const float* __restrict__ ref_cam, int S, int H, int W, int D, int Q,
                                                      int tiles_x, int T, int nb_val, const int* __restrict__ off,
                                                      int* __restrict__ cur, int* __restrict__ rec,
                                                      float* __restrict__ wbuf, int* __restrict__ pixbuf,
                                                      float* __restrict__ wkref, int* __restrict__ pixref) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int n_pairs = __ldg(n_pairs_ptr);
  for (int pair = blockIdx.x * wpb + (threadIdx.x >> 5); pair < n_pairs; pair += gridDim.x * wpb) {
    const int flat = __ldg(pair_vq + pair);
    const int v = flat / Q;
    const size_t vS = (size_t)v * S;
    const float4 sp = ldg4(samp + ((size_t)pair * 32 + lane) * 4);
    const Tap t = make_tap(sp.x, sp.y, sp.z, H, W, D);
    float w[4];
    int px[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      px[k] = t.in3d ? t.pix[k] : -1;
      w[k] = 0.f;
      if (px[k] >= 0) {
        float lo, hi;
        w[k] = sp.w * (t.bw[k] * depth_score(t, dist + (vS + px[k]) * D, D, lo, hi));
      }
    }
    reinterpret_cast<float4*>(wbuf)[(size_t)pair * 32 + lane] = make_float4(w[0], w[1], w[2], w[3]);
    reinterpret_cast<int4*>(pixbuf)[(size_t)pair * 32 + lane] = make_int4(px[0], px[1], px[2], px[3]);
    int ids[4];
    const int n = corner_tiles(t, W, tiles_x, ids);
    for (int q = 0; q < n; ++q) {
      const int b = (v * T + ids[q]) * 8 + (lane >> 2);
      const int slot = atomicAdd(cur + b, 1);
      rec[__ldg(off + b) + slot] = pair * 32 + lane;
    }
    // reference-point sample of the folded map (weights 1): lanes 0..3 keep one corner each, lane 0 files the pair
    const Tap tr = make_tap(__ldg(ref_cam + (size_t)flat * 3), __ldg(ref_cam + (size_t)flat * 3 + 1),
                            __ldg(ref_cam + (size_t)flat * 3 + 2), H, W, D);
    if (lane < 4) {
      int pr = -1;
      float wk = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k == lane && tr.in3d && tr.pix[k] >= 0) {
          float lo, hi;
          pr = tr.pix[k];
          wk = tr.bw[k] * depth_score(tr, dist + (vS + tr.pix[k]) * D, D, lo, hi);
        }
      }
      wkref[(size_t)pair * 4 + lane] = wk;
      pixref[(size_t)pair * 4 + lane] = pr;
    }
    if (lane == 0) {
      const int m = corner_tiles(tr, W, tiles_x, ids);
      for (int q = 0; q < m; ++q) {
        const int b = nb_val + v * T + ids[q];
        const int slot = atomicAdd(cur + b, 1);
        rec[__ldg(off + b) + slot] = pair;
      }
    }
  }
}

// CTA = one 4x8 pixel tile of one view, warp = one head.  smem: val[32][C] (staged value rows) | acc[32][C] | mbarrier.
template <int CPL>
__global__ void __launch_bounds__(256) lt_tiles_kernel(const float* __restrict__ value, int ldv,
                                                       const float* __restrict__ grad_slots, const int* __restrict__ off,
                                                       const int* __restrict__ rec, const float* __restrict__ wbuf,
                                                       const int* __restrict__ pixbuf, int S, int H, int W, int tiles_x, int T,
                                                       float* __restrict__ dots, float* __restrict__ grad_value) {
  constexpr int C = CPL * 32, Cm = CPL * 4, CG = CPL;   // channels, channels per head, lanes (of 4 channels) per corner
  extern __shared__ __align__(128) float lt_smem[];
  float* val = lt_smem;
  float* acc = lt_smem + kTPx * C;
  uint64_t* bar = reinterpret_cast<uint64_t*>(lt_smem + 2 * kTPx * C);
  const int tid = threadIdx.x, lane = tid & 31, h = tid >> 5;
  const int v = blockIdx.x / T, tile = blockIdx.x - v * T;
  const int ty0 = (tile / tiles_x) * kTH, tx0 = (tile % tiles_x) * kTW;
  const size_t vS = (size_t)v * S;
  if (tid == 0) {
    tc::mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (h == 0) {
    // one bulk async copy per pixel of the tile (its C value channels are contiguous; the row stride ldv also holds the
    // folded channels): lane = pixel
    const int ly = lane / kTW, lx = lane % kTW;
    const bool ok = (ty0 + ly < H) && (tx0 + lx < W);
    const unsigned okm = __ballot_sync(SGC_FULL_MASK, ok);
    if (lane == 0) tc::mbar_expect_tx(bar, (uint32_t)(__popc(okm) * C * 4));
    __syncwarp();
    if (ok) tc::bulk_g2s(val + lane * C, value + (vS + (size_t)(ty0 + ly) * W + tx0 + lx) * ldv, (uint32_t)(C * 4), bar);
  }
  for (int i = tid; i < kTPx * C / 4; i += 256) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  tc::mbar_wait(bar, 0);
  __syncthreads();

  const int k = lane >> 3, cg = lane & 7;      // corner handled by this lane, 4-channel group inside the head
  const bool chan = cg < CG;
  const int ch = h * Cm + cg * 4;
  const int b = (v * T + tile) * 8 + h;
  const int beg = __ldg(off + b), end = __ldg(off + b + 1);
  for (int base = beg; base < end; base += 32) {
    const int mine = base + lane < end ? __ldg(rec + base + lane) : 0;
    const int nrec = min(32, end - base);
    for (int j = 0; j < nrec; ++j) {
      const int r = __shfl_sync(SGC_FULL_MASK, mine, j);
      const int pair = r >> 5;
      const size_t tb = (size_t)r * 4 + k;
      const int pix = __ldg(pixbuf + tb);
      const float w = __ldg(wbuf + tb);
      float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (chan) g4 = ldg4(grad_slots + (size_t)pair * C + ch);
      int lp = -1;
      if (pix >= 0) {
        const int y = pix / W, x = pix - y * W;
        const int ly = y - ty0, lx = x - tx0;
        if ((unsigned)ly < (unsigned)kTH && (unsigned)lx < (unsigned)kTW) lp = ly * kTW + lx;
      }
      float d = 0.f;
      if (lp >= 0 && chan) {
        const float4 x4 = *reinterpret_cast<const float4*>(val + lp * C + ch);
        d = x4.x * g4.x + x4.y * g4.y + x4.z * g4.z + x4.w * g4.w;
      }
      d += __shfl_xor_sync(SGC_FULL_MASK, d, 1);
      d += __shfl_xor_sync(SGC_FULL_MASK, d, 2);
      d += __shfl_xor_sync(SGC_FULL_MASK, d, 4);
      if (lp >= 0) {
        if (cg == 0) dots[tb] = d;
        if (chan && w != 0.f) {
          float4* a = reinterpret_cast<float4*>(acc + lp * C + ch);
          float4 a4 = *a;
          a4.x += w * g4.x; a4.y += w * g4.y; a4.z += w * g4.z; a4.w += w * g4.w;
          *a = a4;
        }
      }
    }
  }
  __syncthreads();
  // the tile's grad_value rows, written once (coalesced: 1 KB per pixel)
  for (int i = tid; i < kTPx * C / 4; i += 256) {
    const int lp = i / (C / 4), c4 = i - lp * (C / 4);
    const int y = ty0 + lp / kTW, x = tx0 + lp % kTW;
    if (y < H && x < W)
      *reinterpret_cast<float4*>(grad_value + (vS + (size_t)y * W + x) * ldv + c4 * 4) = reinterpret_cast<const float4*>(acc)[i];
  }
}

__device__ __forceinline__ void lt_scatter_dist(float* gd_px, const Tap& t, int D, float gds) {
  if (t.d0 >= 0) red_add1(gd_px + t.d0, t.hd * gds);
  if (t.d0 + 1 <= D - 1) red_add1(gd_px + t.d0 + 1, t.ld * gds);
}

// Per pair, from the saved dots: the gradients of the 128 sampling parameters (gr), grad_dist, bias partials.
template <int CPL>
__global__ void __launch_bounds__(256) lt_params_kernel(const float* __restrict__ G, int ldg, const float* __restrict__ dist,
                                                        const float* __restrict__ vbias, const int* __restrict__ pair_vq,
                                                        const int* __restrict__ n_pairs_ptr, const float* __restrict__ ref_cam,
                                                        const float* __restrict__ samp, const float* __restrict__ grad_slots,
                                                        const float* __restrict__ wbuf, const int* __restrict__ pixbuf,
                                                        const float* __restrict__ dots, int S, int H, int W, int D, int Q,
                                                        float* __restrict__ grad_dist, float* __restrict__ gr_out,
                                                        float* __restrict__ bias_partials) {
  constexpr int C = CPL * 32;
  __shared__ float s_part[8][C + 128];
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int n_pairs = __ldg(n_pairs_ptr);
  float gvb[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) gvb[j] = 0.f;
  float4 ggb = make_float4(0.f, 0.f, 0.f, 0.f);
  float vb[CPL];
  lt_load_row<CPL>(vb, vbias + lt_lane_base<CPL>(lane));
  for (int pair = blockIdx.x * wpb + (threadIdx.x >> 5); pair < n_pairs; pair += gridDim.x * wpb) {
    const int flat = __ldg(pair_vq + pair);
    const int v = flat / Q;
    const size_t vS = (size_t)v * S;
    const float4 sp = ldg4(samp + ((size_t)pair * 32 + lane) * 4);
    const Tap t = make_tap(sp.x, sp.y, sp.z, H, W, D);
    const float4 w4 = ldg4(wbuf + ((size_t)pair * 32 + lane) * 4);
    const int4 p4 = __ldg(reinterpret_cast<const int4*>(pixbuf) + (size_t)pair * 32 + lane);
    const float4 d4 = ldg4(dots + ((size_t)pair * 32 + lane) * 4);
    const int px[4] = {p4.x, p4.y, p4.z, p4.w};
    const float dt[4] = {d4.x, d4.y, d4.z, d4.w};
    float g[CPL];
    lt_load_row<CPL>(g, grad_slots + (size_t)pair * C + lt_lane_base<CPL>(lane));
    float gb = 0.f;
#pragma unroll
    for (int j = 0; j < CPL; ++j) gb += vb[j] * g[j];
    gb = quad_sum(gb);   // value_proj bias of the head dotted with the head's gradient
    // weights of every tap of the head (for grad value_proj.bias): own point's, summed over the head's four points
    const float wsum = quad_sum(w4.x + w4.y + w4.z + w4.w);
#pragma unroll
    for (int j = 0; j < CPL; ++j) gvb[j] += wsum * g[j];
    float ds[4], dlo[4], dhi[4], cw[4], dot[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ds[k] = 0.f; dlo[k] = 0.f; dhi[k] = 0.f; dot[k] = 0.f;
      if (px[k] >= 0) {   // implies t.in3d
        ds[k] = depth_score(t, dist + (vS + px[k]) * D, D, dlo[k], dhi[k]);
        dot[k] = dt[k] + gb;
      }
      cw[k] = t.bw[k] * ds[k];
    }
    float g_attn = 0.f, g_w = 0.f, g_h = 0.f, g_d = 0.f;
    if (t.in3d) {
      const float hh = 1.f - t.lh, hw = 1.f - t.lw;
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
          lt_scatter_dist(grad_dist + (vS + px[k]) * D, t, D, gds);
        }
      }
      g_d = gz * (float)D;
    }
    const float sdot = quad_sum(sp.w * g_attn);
    float4 gr;
    gr.x = __fdiv_rn(g_w, (float)W);
    gr.y = __fdiv_rn(g_h, (float)H);
    gr.z = __fdiv_rn(g_d, (float)D);
    gr.w = sp.w * (g_attn - sdot);
    ggb.x += gr.x; ggb.y += gr.y; ggb.z += gr.z; ggb.w += gr.w;
    reinterpret_cast<float4*>(gr_out)[(size_t)pair * 32 + lane] = gr;
    // backward of the reference-point sample of the folded map w.r.t. the depth distribution
    const Tap tr = make_tap(__ldg(ref_cam + (size_t)flat * 3), __ldg(ref_cam + (size_t)flat * 3 + 1),
                            __ldg(ref_cam + (size_t)flat * 3 + 2), H, W, D);
    if (tr.in3d) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (tr.pix[k] >= 0) {
          const float4 x = ldg4(G + (vS + tr.pix[k]) * ldg + lane * 4);
          float dk = x.x * gr.x + x.y * gr.y + x.z * gr.z + x.w * gr.w;
          dk = warp_sum(dk);
          if (lane == 0) lt_scatter_dist(grad_dist + (vS + tr.pix[k]) * D, tr, D, tr.bw[k] * dk);
        }
      }
    }
  }
  const int wid = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < CPL; ++j) s_part[wid][lt_lane_base<CPL>(lane) + (j >> 2) * 16 + (j & 3)] = gvb[j];
  s_part[wid][C + lane * 4 + 0] = ggb.x; s_part[wid][C + lane * 4 + 1] = ggb.y;
  s_part[wid][C + lane * 4 + 2] = ggb.z; s_part[wid][C + lane * 4 + 3] = ggb.w;
  __syncthreads();
  for (int c = threadIdx.x; c < C + 128; c += blockDim.x) {
    float a = 0.f;
    for (int w = 0; w < wpb; ++w) a += s_part[w][c];
    bias_partials[(size_t)blockIdx.x * (C + 128) + c] = a;
  }
}

// The folded map's gradient: CTA = tile, the pairs filed under it; 4 warps share the tile through shared-memory atomics.
__global__ void __launch_bounds__(128) lt_gtiles_kernel(const int* __restrict__ off, const int* __restrict__ rec,
                                                        const float* __restrict__ wkref, const int* __restrict__ pixref,
                                                        const float* __restrict__ gr, int S, int H, int W, int tiles_x, int T,
                                                        int nb_val, int ldg, float* __restrict__ grad_G) {
  __shared__ __align__(16) float acc[kTPx * 128];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int v = blockIdx.x / T, tile = blockIdx.x - v * T;
  const int ty0 = (tile / tiles_x) * kTH, tx0 = (tile % tiles_x) * kTW;
  const size_t vS = (size_t)v * S;
  for (int i = tid; i < kTPx * 128 / 4; i += 128) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  const int b = nb_val + v * T + tile;
  const int beg = __ldg(off + b), end = __ldg(off + b + 1);
  for (int i = beg + wid; i < end; i += 4) {
    const int pair = __ldg(rec + i);
    const float4 g4 = ldg4(gr + (size_t)pair * 128 + lane * 4);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int pix = __ldg(pixref + (size_t)pair * 4 + k);
      const float wk = __ldg(wkref + (size_t)pair * 4 + k);
      if (pix < 0 || wk == 0.f) continue;
      const int y = pix / W, x = pix - y * W;
      const int ly = y - ty0, lx = x - tx0;
      if ((unsigned)ly >= (unsigned)kTH || (unsigned)lx >= (unsigned)kTW) continue;
      float* a = acc + (ly * kTW + lx) * 128 + lane * 4;
      atomicAdd(a, wk * g4.x); atomicAdd(a + 1, wk * g4.y); atomicAdd(a + 2, wk * g4.z); atomicAdd(a + 3, wk * g4.w);
    }
  }
  __syncthreads();
  for (int i = tid; i < kTPx * 32; i += 128) {
    const int lp = i >> 5, c4 = i & 31;
    const int y = ty0 + lp / kTW, x = tx0 + lp % kTW;
    if (y < H && x < W)
      *reinterpret_cast<float4*>(grad_G + (vS + (size_t)y * W + x) * ldg + c4 * 4) = reinterpret_cast<const float4*>(acc)[i];
  }
}

// grad_vbias[c] = sum_rows partials[row][c] (c < C);  grad_gbias[c-C] = ... (c >= C): assignment, fixed order.
__global__ void __launch_bounds__(256) lt_bias_reduce_kernel(const float* __restrict__ partials, int rows, int C,
                                                            float* __restrict__ grad_vbias, float* __restrict__ grad_gbias) {
  __shared__ float s[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float a = 0.f;
#pragma unroll 4
  for (int r = ry; r < rows; r += 8) a += __ldg(partials + (size_t)r * (C + 128) + c);
  s[ry][cx] = a;
  __syncthreads();
  if (ry == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += s[k][cx];
    if (c < C) grad_vbias[c] = t; else grad_gbias[c - C] = t;
  }
}

struct LtPlan {
  int tiles_x, tiles_y, T, nb_val, nb, grid_pairs;
  size_t o_cnt, o_off, o_cur, o_rec, o_wbuf, o_pixbuf, o_dots, o_gr, o_wkref, o_pixref, o_bias, total;
};

static LtPlan lt_plan(int cap_pairs, int V, int H, int W, int C) {
  LtPlan p;
  p.tiles_x = (W + kTW - 1) / kTW;
  p.tiles_y = (H + kTH - 1) / kTH;
  p.T = p.tiles_x * p.tiles_y;
  p.nb_val = V * p.T * 8;
  p.nb = p.nb_val + V * p.T;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int want = (cap_pairs + 7) / 8, full = sms * 8;
  p.grid_pairs = want < full ? (want > 0 ? want : 1) : full;
  size_t o = 0;
  auto take = [&o](size_t bytes) { const size_t at = o; o += (bytes + 255) & ~(size_t)255; return at; };
  p.o_cnt = take((size_t)p.nb * 4);
  p.o_off = take((size_t)(p.nb + 1) * 4);
  p.o_cur = take((size_t)p.nb * 4);
  p.o_rec = take((size_t)cap_pairs * (32 * 4 + 4) * 4);     // worst case: every point in 4 tiles, every pair in 4
  p.o_wbuf = take((size_t)cap_pairs * 32 * 16);
  p.o_pixbuf = take((size_t)cap_pairs * 32 * 16);
  p.o_dots = take((size_t)cap_pairs * 32 * 16);
  p.o_gr = take((size_t)cap_pairs * 128 * 4);
  p.o_wkref = take((size_t)cap_pairs * 16);
  p.o_pixref = take((size_t)cap_pairs * 16);
  p.o_bias = take((size_t)p.grid_pairs * (C + 128) * 4);
  p.total = o;
  return p;
}

}  // namespace sgc

extern "C" long long sgc_lift_bwd_tiles_workspace_bytes(int cap_pairs, int V, int H, int W, int C) {
  if (cap_pairs <= 0 || V <= 0 || H <= 0 || W <= 0 || (C != 128 && C != 256)) return 0;
  return (long long)sgc::lt_plan(cap_pairs, V, H, W, C).total;
}

extern "C" int sgc_lift_bwd_tiles(const float* value, int ldv, const float* G, int ldg, const float* dist, const float* vbias,
                                  const int* pair_vq, const int* n_pairs, int cap_pairs, const float* ref_cam,
                                  const float* samp, const float* grad_slots, int V, int S, int H, int W, int D, int Q, int C,
                                  float* grad_value, float* grad_G, float* grad_dist, float* grad_vbias, float* grad_gbias,
                                  void* workspace, void* stream) {
  using namespace sgc;
  if ((C != 256 && C != 128) || (ldv & 3) || (ldg & 3) || S != H * W || !workspace) return (int)cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) || (reinterpret_cast<uintptr_t>(value) & 15)) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  const LtPlan p = lt_plan(cap_pairs, V, H, W, C);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  int* cnt = reinterpret_cast<int*>(ws + p.o_cnt);
  int* off = reinterpret_cast<int*>(ws + p.o_off);
  int* cur = reinterpret_cast<int*>(ws + p.o_cur);
  int* rec = reinterpret_cast<int*>(ws + p.o_rec);
  float* wbuf = reinterpret_cast<float*>(ws + p.o_wbuf);
  int* pixbuf = reinterpret_cast<int*>(ws + p.o_pixbuf);
  float* dots = reinterpret_cast<float*>(ws + p.o_dots);
  float* gr = reinterpret_cast<float*>(ws + p.o_gr);
  float* wkref = reinterpret_cast<float*>(ws + p.o_wkref);
  int* pixref = reinterpret_cast<int*>(ws + p.o_pixref);
  float* bias = reinterpret_cast<float*>(ws + p.o_bias);
  cudaError_t e = cudaMemsetAsync(cnt, 0, (size_t)p.nb * 4, st);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(grad_dist, 0, (size_t)V * S * D * 4, st);
  if (e != cudaSuccess) return (int)e;
  lt_count_kernel<<<p.grid_pairs, 256, 0, st>>>(samp, pair_vq, n_pairs, ref_cam, H, W, D, Q, p.tiles_x, p.T, p.nb_val, cnt);
  SGC_CUDA_CHECK_LAST();
  lt_scan_kernel<<<1, 1024, 0, st>>>(cnt, p.nb, off, cur);
  SGC_CUDA_CHECK_LAST();
  lt_emit_kernel<<<p.grid_pairs, 256, 0, st>>>(samp, dist, pair_vq, n_pairs, ref_cam, S, H, W, D, Q, p.tiles_x, p.T, p.nb_val, off, cur,
                                               rec, wbuf, pixbuf, wkref, pixref);
  SGC_CUDA_CHECK_LAST();
  const size_t smem = (size_t)2 * kTPx * C * 4 + 16;
  if (C == 256) {
    e = cudaFuncSetAttribute(lt_tiles_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    lt_tiles_kernel<8><<<V * p.T, 256, smem, st>>>(value, ldv, grad_slots, off, rec, wbuf, pixbuf, S, H, W, p.tiles_x, p.T, dots,
                                                   grad_value);
  } else {
    e = cudaFuncSetAttribute(lt_tiles_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    lt_tiles_kernel<4><<<V * p.T, 256, smem, st>>>(value, ldv, grad_slots, off, rec, wbuf, pixbuf, S, H, W, p.tiles_x, p.T, dots,
                                                   grad_value);
  }
  SGC_CUDA_CHECK_LAST();
  if (C == 256)
    lt_params_kernel<8><<<p.grid_pairs, 256, 0, st>>>(G, ldg, dist, vbias, pair_vq, n_pairs, ref_cam, samp, grad_slots, wbuf, pixbuf,
                                                      dots, S, H, W, D, Q, grad_dist, gr, bias);
  else
    lt_params_kernel<4><<<p.grid_pairs, 256, 0, st>>>(G, ldg, dist, vbias, pair_vq, n_pairs, ref_cam, samp, grad_slots, wbuf, pixbuf,
                                                      dots, S, H, W, D, Q, grad_dist, gr, bias);
  SGC_CUDA_CHECK_LAST();
  lt_gtiles_kernel<<<V * p.T, 128, 0, st>>>(off, rec, wkref, pixref, gr, S, H, W, p.tiles_x, p.T, p.nb_val, ldg, grad_G);
  SGC_CUDA_CHECK_LAST();
  lt_bias_reduce_kernel<<<(C + 128) / 32, 256, 0, st>>>(bias, p.grid_pairs, C, grad_vbias, grad_gbias);
  SGC_CUDA_CHECK_LAST();
  return 0;
}
