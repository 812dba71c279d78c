// sgc_crossview_*: cross-view fusion of the per-pair lifted features, one warp per voxel.
//
// Replaces DeformCrossAttention_DFA3D.forward lines 815-833 (deformable_cross_attention.py):
//   slots scatter + count (815-820), masked mean over views (826), and the 8-head
//   nn.MultiheadAttention over views with key_padding_mask = ~visible (829-833).
//
// The dense projections stay GEMMs of *voxel* count (static shapes), by commuting the linear maps:
//   score[v,h]  = (W_k,h s_v + b_k,h) . q_h / sqrt(dh)  ==  qt[h] . s_v  + const(h)      (const cancels in softmax)
//   out_h       = sum_v alpha[v,h] (W_v,h s_v + b_v,h)  ==  W_v,h t[h] + b_v,h,   t[h] = sum_v alpha[v,h] s_v
// so the kernels here only need   qt [8,Q,C] = scale * W_k,h^T q_h   (host GEMM)   and emit   t [8,Q,C].
//
// Layout: slots [cap,C] pair-major; pair_index [V,Q] (-1 = invisible); count [Q]; alpha [cap,8] saved for bwd.
// Lane l owns channels [l*CPL, l*CPL+CPL), CPL = C/32.
#include <cuda_bf16.h>
#include "common.cuh"

namespace sgc {

constexpr int kMaxViews = 128;  // per-warp score scratch is [kMaxViews][8]
constexpr int kCvWarps = 4;

template <int CPL>
__device__ __forceinline__ void load_row_cv(float (&dst)[CPL], const float* p) {
#pragma unroll
  for (int j = 0; j < CPL; j += 4) {
    const float4 t = ldg4(p + j);
    dst[j] = t.x; dst[j + 1] = t.y; dst[j + 2] = t.z; dst[j + 3] = t.w;
  }
}
template <int CPL>
__device__ __forceinline__ void store_row_cv(float* p, const float (&src)[CPL]) {
#pragma unroll
  for (int j = 0; j < CPL; j += 4) *reinterpret_cast<float4*>(p + j) = make_float4(src[j], src[j + 1], src[j + 2], src[j + 3]);
}

// bf16x3 operand image of one row for the next tensor-core GEMM (pattern 0 of sgc_split_bf16x3: hi | lo | hi along K):
// split[row][slot*C + c], written by the lane that owns channels [lane*CPL, lane*CPL + CPL)
template <int CPL>
__device__ __forceinline__ void store_split_cv(__nv_bfloat16* __restrict__ split, size_t row, int lane, const float (&v)[CPL]) {
  constexpr int C = CPL * 32;
  __align__(16) __nv_bfloat16 h[CPL], l[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    h[j] = __float2bfloat16_rn(v[j]);
    l[j] = __float2bfloat16_rn(v[j] - __bfloat162float(h[j]));
  }
  __nv_bfloat16* o = split + row * 3 * C + lane * CPL;
#pragma unroll
  for (int j = 0; j < CPL; j += 4) {
    const uint2 hv = *reinterpret_cast<const uint2*>(h + j), lv = *reinterpret_cast<const uint2*>(l + j);
    *reinterpret_cast<uint2*>(o + j) = hv;
    *reinterpret_cast<uint2*>(o + C + j) = lv;
    *reinterpret_cast<uint2*>(o + 2 * C + j) = hv;
  }
}

// Rows ids[i .. i+U) of `slots` (this lane's CPL channels), all U gathers issued before any is consumed; rows past n read as 0
// (their loads are aimed at row n-1, so every load is unconditional and the compiler can hoist all of them).
template <int CPL, int U>
__device__ __forceinline__ void load_rows_cv(float (&x)[U][CPL], const float* __restrict__ slots, const int* ids, int i, int n,
                                             int lane) {
  constexpr int C = CPL * 32;
  int id[U];
#pragma unroll
  for (int u = 0; u < U; ++u) id[u] = ids[i + u < n ? i + u : n - 1];
#pragma unroll
  for (int u = 0; u < U; ++u) load_row_cv<CPL>(x[u], slots + (size_t)id[u] * C + lane * CPL);
#pragma unroll
  for (int u = 1; u < U; ++u)
    if (i + u >= n) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) x[u][j] = 0.f;
    }
}

// Collect the pair ids of the views that see voxel q into ids[] (warp-shared), return how many.
__device__ __forceinline__ int gather_views(const int* __restrict__ pair_index, int V, int Q, int q, int lane,
                                            int* ids) {
  int n = 0;
  for (int base = 0; base < V; base += 32) {
    const int v = base + lane;
    const int id = (v < V) ? __ldg(pair_index + (size_t)v * Q + q) : -1;
    const unsigned b = __ballot_sync(SGC_FULL_MASK, id >= 0);
    if (id >= 0) ids[n + __popc(b & ((1u << lane) - 1))] = id;
    n += __popc(b);
  }
  __syncwarp();
  return n;
}

// U = rows fetched together per trip.  The levels with few voxels run a handful of warps per SM, so a warp's time IS the
// kernel's time and it is the chain of dependent row gathers (one per visible view) that sets it: U = 4 there, 1 where
// thousands of warps hide each other's latency (and registers are better spent on occupancy).
template <int CPL, bool DIVIDE = true, int U = 1>
__global__ void __launch_bounds__(kCvWarps * 32) mean_fwd_kernel(const float* __restrict__ slots,
                                                                 const int* __restrict__ pair_index, int V, int Q,
                                                                 float* __restrict__ mean,
                                                                 __nv_bfloat16* __restrict__ split = nullptr) {
  constexpr int C = CPL * 32;
  pdl_sync();
  __shared__ int s_ids[kCvWarps][kMaxViews];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int q = blockIdx.x * kCvWarps + wid;
  if (q >= Q) return;
  int* ids = s_ids[wid];
  const int n = gather_views(pair_index, V, Q, q, lane, ids);
  float acc[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) acc[j] = 0.f;
  for (int i = 0; i < n; i += U) {
    float x[U][CPL];
    load_rows_cv<CPL, U>(x, slots, ids, i, n, lane);
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < CPL; ++j) acc[j] += x[u][j];
  }
  if (DIVIDE && n > 0) {
    const float fn = (float)n;
#pragma unroll
    for (int j = 0; j < CPL; ++j) acc[j] = acc[j] / fn;  // DCA:826
  }
  store_row_cv<CPL>(mean + (size_t)q * C + lane * CPL, acc);
  if (split) store_split_cv<CPL>(split, (size_t)q, lane, acc);
}

template <int CPL, int U = 1>
__global__ void __launch_bounds__(kCvWarps * 32) attn_fwd_kernel(const float* __restrict__ qt,
                                                                 const float* __restrict__ slots,
                                                                 const int* __restrict__ pair_index, int V, int Q,
                                                                 float* __restrict__ t_out, float* __restrict__ alpha,
                                                                 __nv_bfloat16* __restrict__ split = nullptr) {
  constexpr int C = CPL * 32;
  pdl_sync();
  __shared__ int s_ids[kCvWarps][kMaxViews];
  __shared__ float s_sc[kCvWarps][kMaxViews][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int q = blockIdx.x * kCvWarps + wid;
  if (q >= Q) return;
  int* ids = s_ids[wid];
  float(*sc)[8] = s_sc[wid];
  const int n = gather_views(pair_index, V, Q, q, lane, ids);
  const size_t hq = (size_t)Q * C;
  const size_t off = (size_t)q * C + lane * CPL;
  if (n == 0) {
    float z[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) z[j] = 0.f;
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      store_row_cv<CPL>(t_out + h * hq + off, z);
      if (split) store_split_cv<CPL>(split, (size_t)h * Q + q, lane, z);
    }
    return;
  }
  {  // phase 1: scores
    float qv[8][CPL];
#pragma unroll
    for (int h = 0; h < 8; ++h) load_row_cv<CPL>(qv[h], qt + h * hq + off);
    for (int i = 0; i < n; i += U) {
      float x[U][CPL];
      load_rows_cv<CPL, U>(x, slots, ids, i, n, lane);
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          float p = 0.f;
#pragma unroll
          for (int j = 0; j < CPL; ++j) p += qv[h][j] * x[u][j];
          p = warp_sum(p);
          if (lane == h && i + u < n) sc[i + u][h] = p;
        }
      }
    }
  }
  __syncwarp();
  // softmax over the visible views, per head: lane -> (head = lane & 7, views lane>>3, +4, ...)
  {
    const int h = lane & 7;
    float mx = -INFINITY;
    for (int i = lane >> 3; i < n; i += 4) mx = fmaxf(mx, sc[i][h]);
    mx = fmaxf(mx, __shfl_xor_sync(SGC_FULL_MASK, mx, 8));
    mx = fmaxf(mx, __shfl_xor_sync(SGC_FULL_MASK, mx, 16));
    float sum = 0.f;
    for (int i = lane >> 3; i < n; i += 4) {
      const float e = expf(sc[i][h] - mx);
      sc[i][h] = e;
      sum += e;
    }
    sum += __shfl_xor_sync(SGC_FULL_MASK, sum, 8);
    sum += __shfl_xor_sync(SGC_FULL_MASK, sum, 16);
    for (int i = lane >> 3; i < n; i += 4) {
      const float a = sc[i][h] / sum;
      sc[i][h] = a;
      alpha[(size_t)ids[i] * 8 + h] = a;
    }
  }
  __syncwarp();
  {  // phase 2: t[h] = sum_v alpha[v,h] s_v
    float t[8][CPL];
#pragma unroll
    for (int h = 0; h < 8; ++h)
#pragma unroll
      for (int j = 0; j < CPL; ++j) t[h][j] = 0.f;
    for (int i = 0; i < n; i += U) {
      float x[U][CPL];
      load_rows_cv<CPL, U>(x, slots, ids, i, n, lane);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (i + u < n) {
#pragma unroll
          for (int h = 0; h < 8; ++h) {
            const float a = sc[i + u][h];
#pragma unroll
            for (int j = 0; j < CPL; ++j) t[h][j] += a * x[u][j];
          }
        }
      }
    }
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      store_row_cv<CPL>(t_out + h * hq + off, t[h]);
      if (split) store_split_cv<CPL>(split, (size_t)h * Q + q, lane, t[h]);
    }
  }
}

// backward, step 1: g_alpha -> softmax backward -> gscore [cap,8] (stored) and grad_qt [8,Q,C]
template <int CPL>
__global__ void __launch_bounds__(kCvWarps * 32) attn_bwd_qt_kernel(
    const float* __restrict__ slots, const float* __restrict__ alpha, const int* __restrict__ pair_index, int V, int Q,
    const float* __restrict__ grad_t, float* __restrict__ gscore, float* __restrict__ grad_qt,
    __nv_bfloat16* __restrict__ split = nullptr) {
  constexpr int C = CPL * 32;
  pdl_sync();
  __shared__ int s_ids[kCvWarps][kMaxViews];
  __shared__ float s_al[kCvWarps][kMaxViews][8];  // alpha
  __shared__ float s_gs[kCvWarps][kMaxViews][8];  // grad of the scores
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int q = blockIdx.x * kCvWarps + wid;
  if (q >= Q) return;
  int* ids = s_ids[wid];
  float(*al)[8] = s_al[wid];
  float(*gs)[8] = s_gs[wid];
  const int n = gather_views(pair_index, V, Q, q, lane, ids);
  const size_t hq = (size_t)Q * C;
  const size_t off = (size_t)q * C + lane * CPL;
  if (n == 0) {
    float z[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) z[j] = 0.f;
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      store_row_cv<CPL>(grad_qt + h * hq + off, z);
      if (split) store_split_cv<CPL>(split, (size_t)h * Q + q, lane, z);
    }
    return;
  }
  for (int i = lane >> 3; i < n; i += 4) al[i][lane & 7] = __ldg(alpha + (size_t)ids[i] * 8 + (lane & 7));
  {
    float gt[8][CPL];
#pragma unroll
    for (int h = 0; h < 8; ++h) load_row_cv<CPL>(gt[h], grad_t + h * hq + off);
    // pass A: g_alpha[v,h] = grad_t[h] . s_v
    for (int i = 0; i < n; ++i) {
      float x[CPL];
      load_row_cv<CPL>(x, slots + (size_t)ids[i] * C + lane * CPL);
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        float p = 0.f;
#pragma unroll
        for (int j = 0; j < CPL; ++j) p += gt[h][j] * x[j];
        p = warp_sum(p);
        if (lane == h) gs[i][h] = p;
      }
    }
  }
  __syncwarp();
  {  // softmax backward per head
    const int h = lane & 7;
    float d = 0.f;
    for (int i = lane >> 3; i < n; i += 4) d += al[i][h] * gs[i][h];
    d += __shfl_xor_sync(SGC_FULL_MASK, d, 8);
    d += __shfl_xor_sync(SGC_FULL_MASK, d, 16);
    for (int i = lane >> 3; i < n; i += 4) {
      const float g = al[i][h] * (gs[i][h] - d);
      gs[i][h] = g;
      gscore[(size_t)ids[i] * 8 + h] = g;
    }
  }
  __syncwarp();
  // pass B: grad_qt[h] = sum_v gscore[v,h] s_v
  {
    float gq[8][CPL];
#pragma unroll
    for (int h = 0; h < 8; ++h)
#pragma unroll
      for (int j = 0; j < CPL; ++j) gq[h][j] = 0.f;
    for (int i = 0; i < n; ++i) {
      float x[CPL];
      load_row_cv<CPL>(x, slots + (size_t)ids[i] * C + lane * CPL);
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        const float g = gs[i][h];
#pragma unroll
        for (int j = 0; j < CPL; ++j) gq[h][j] += g * x[j];
      }
    }
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      store_row_cv<CPL>(grad_qt + h * hq + off, gq[h]);
      if (split) store_split_cv<CPL>(split, (size_t)h * Q + q, lane, gq[h]);
    }
  }
}

// backward, step 2 (after the host GEMMs produced grad_mean):
//   grad_slots[v] = grad_mean/n + sum_h (alpha[v,h] grad_t[h] + gscore[v,h] qt[h])
template <int CPL>
__global__ void __launch_bounds__(kCvWarps * 32) attn_bwd_slots_kernel(
    const float* __restrict__ qt, const float* __restrict__ alpha, const float* __restrict__ gscore,
    const int* __restrict__ pair_index, int V, int Q, const float* __restrict__ grad_t,
    const float* __restrict__ grad_mean, float* __restrict__ grad_slots, const int* __restrict__ count_override) {
  constexpr int C = CPL * 32;
  pdl_sync();
  __shared__ int s_ids[kCvWarps][kMaxViews];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int q = blockIdx.x * kCvWarps + wid;
  if (q >= Q) return;
  int* ids = s_ids[wid];
  const int n = gather_views(pair_index, V, Q, q, lane, ids);
  if (n == 0) return;
  const size_t hq = (size_t)Q * C;
  const size_t off = (size_t)q * C + lane * CPL;
  float gm[CPL];
  load_row_cv<CPL>(gm, grad_mean + off);
  // view-sharded mode: the mean is over the views of ALL shards
  const float fn = (float)(count_override ? __ldg(count_override + q) : n);
#pragma unroll
  for (int j = 0; j < CPL; ++j) gm[j] = gm[j] / fn;
  float gt[8][CPL], qv[8][CPL];
#pragma unroll
  for (int h = 0; h < 8; ++h) {
    load_row_cv<CPL>(gt[h], grad_t + h * hq + off);
    load_row_cv<CPL>(qv[h], qt + h * hq + off);
  }
  for (int i = 0; i < n; ++i) {
    const float* ap = alpha + (size_t)ids[i] * 8;
    const float* gp = gscore + (size_t)ids[i] * 8;
    const float4 a0 = ldg4(ap), a1 = ldg4(ap + 4), g0 = ldg4(gp), g1 = ldg4(gp + 4);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    float o[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) o[j] = gm[j];
#pragma unroll
    for (int h = 0; h < 8; ++h)
#pragma unroll
      for (int j = 0; j < CPL; ++j) o[j] += a[h] * gt[h][j] + g[h] * qv[h][j];
    store_row_cv<CPL>(grad_slots + (size_t)ids[i] * C + lane * CPL, o);
  }
}

// ---------------------------------------------------------------------------------------------------
// View-sharded variants: the views of a scene are split over shards (GPUs); every softmax
// statistic over views becomes (local partial) -> all-reduce -> (local finish).  Exchange steps live on the host
// (NCCL); these kernels are the local halves.
//   fwd : scores -> [MAX m] -> accum (e = exp(sc - m), s_loc, o_loc) -> [SUM s, o] -> t = o / s
//   bwd : dot (alpha = e/s, g_alpha, D_loc) -> [SUM D] -> qt (gscore, gqt_loc) -> [SUM gqt] -> slots (existing)
constexpr float kNegBig = -3.0e38f;

template <int CPL>
__global__ void __launch_bounds__(kCvWarps * 32) cvs_scores_kernel(const float* __restrict__ qt,
                                                                   const float* __restrict__ slots,
                                                                   const int* __restrict__ pair_index, int V, int Q,
                                                                   float* __restrict__ scores, float* __restrict__ mloc) {
  constexpr int C = CPL * 32;
  __shared__ int s_ids[kCvWarps][kMaxViews];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int q = blockIdx.x * kCvWarps + wid;
  if (q >= Q) return;
  int* ids = s_ids[wid];
  const int n = gather_views(pair_index, V, Q, q, lane, ids);
  const size_t hq = (size_t)Q * C;
  const size_t off = (size_t)q * C + lane * CPL;
  float qv[8][CPL];
#pragma unroll
  for (int h = 0; h < 8; ++h) load_row_cv<CPL>(qv[h], qt + h * hq + off);
  float mx = kNegBig;  // lane h (< 8) tracks the running max of head h
  for (int i = 0; i < n; ++i) {
    float x[CPL];
    load_row_cv<CPL>(x, slots + (size_t)ids[i] * C + lane * CPL);
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      float p = 0.f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) p += qv[h][j] * x[j];
      p = warp_sum(p);
      if (lane == h) { scores[(size_t)ids[i] * 8 + h] = p; mx = fmaxf(mx, p); }
    }
  }
  if (lane < 8) mloc[(size_t)q * 8 + lane] = mx;
}

template <int CPL>
__global__ void __launch_bounds__(kCvWarps * 32) cvs_accum_kernel(const float* __restrict__ scores,
                                                                  const float* __restrict__ mglob,
                                                                  const float* __restrict__ slots,
                                                                  const int* __restrict__ pair_index, int V, int Q,
                                                                  float* __restrict__ e_out, float* __restrict__ sloc,
                                                                  float* __restrict__ oloc) {
  constexpr int C = CPL * 32;
  __shared__ int s_ids[kCvWarps][kMaxViews];
  __shared__ float s_e[kCvWarps][kMaxViews][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int q = blockIdx.x * kCvWarps + wid;
  if (q >= Q) return;
  int* ids = s_ids[wid];
  float(*e)[8] = s_e[wid];
  const int n = gather_views(pair_index, V, Q, q, lane, ids);
  const size_t hq = (size_t)Q * C;
  const size_t off = (size_t)q * C + lane * CPL;
  {
    const int h = lane & 7;
    const float m = __ldg(mglob + (size_t)q * 8 + h);
    float sum = 0.f;
    for (int i = lane >> 3; i < n; i += 4) {
      const float v = expf(__ldg(scores + (size_t)ids[i] * 8 + h) - m);
      e[i][h] = v;
      e_out[(size_t)ids[i] * 8 + h] = v;
      sum += v;
    }
    sum += __shfl_xor_sync(SGC_FULL_MASK, sum, 8);
    sum += __shfl_xor_sync(SGC_FULL_MASK, sum, 16);
    if (lane < 8) sloc[(size_t)q * 8 + lane] = sum;
  }
  __syncwarp();
  float t[8][CPL];
#pragma unroll
  for (int h = 0; h < 8; ++h)
#pragma unroll
    for (int j = 0; j < CPL; ++j) t[h][j] = 0.f;
  for (int i = 0; i < n; ++i) {
    float x[CPL];
    load_row_cv<CPL>(x, slots + (size_t)ids[i] * C + lane * CPL);
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const float a = e[i][h];
#pragma unroll
      for (int j = 0; j < CPL; ++j) t[h][j] += a * x[j];
    }
  }
#pragma unroll
  for (int h = 0; h < 8; ++h) store_row_cv<CPL>(oloc + h * hq + off, t[h]);
}

template <int CPL>
__global__ void __launch_bounds__(kCvWarps * 32) cvs_bwd_dot_kernel(const float* __restrict__ slots,
                                                                    const float* __restrict__ e_in,
                                                                    const float* __restrict__ sglob,
                                                                    const int* __restrict__ pair_index, int V, int Q,
                                                                    const float* __restrict__ grad_t,
                                                                    float* __restrict__ alpha, float* __restrict__ galpha,
                                                                    float* __restrict__ dloc) {
  constexpr int C = CPL * 32;
  __shared__ int s_ids[kCvWarps][kMaxViews];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int q = blockIdx.x * kCvWarps + wid;
  if (q >= Q) return;
  int* ids = s_ids[wid];
  const int n = gather_views(pair_index, V, Q, q, lane, ids);
  const size_t hq = (size_t)Q * C;
  const size_t off = (size_t)q * C + lane * CPL;
  float gt[8][CPL];
#pragma unroll
  for (int h = 0; h < 8; ++h) load_row_cv<CPL>(gt[h], grad_t + h * hq + off);
  const float sg = (lane < 8) ? __ldg(sglob + (size_t)q * 8 + lane) : 1.f;
  float d = 0.f;  // lane h (< 8): partial of sum_v alpha[v,h] g_alpha[v,h]
  for (int i = 0; i < n; ++i) {
    float x[CPL];
    load_row_cv<CPL>(x, slots + (size_t)ids[i] * C + lane * CPL);
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      float p = 0.f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) p += gt[h][j] * x[j];
      p = warp_sum(p);
      if (lane == h) {
        const float a = __ldg(e_in + (size_t)ids[i] * 8 + h) / sg;
        alpha[(size_t)ids[i] * 8 + h] = a;
        galpha[(size_t)ids[i] * 8 + h] = p;
        d += a * p;
      }
    }
  }
  if (lane < 8) dloc[(size_t)q * 8 + lane] = d;
}

template <int CPL>
__global__ void __launch_bounds__(kCvWarps * 32) cvs_bwd_qt_kernel(const float* __restrict__ slots,
                                                                   const float* __restrict__ alpha,
                                                                   const float* __restrict__ galpha,
                                                                   const float* __restrict__ dglob,
                                                                   const int* __restrict__ pair_index, int V, int Q,
                                                                   float* __restrict__ gscore, float* __restrict__ gqt) {
  constexpr int C = CPL * 32;
  __shared__ int s_ids[kCvWarps][kMaxViews];
  __shared__ float s_gs[kCvWarps][kMaxViews][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int q = blockIdx.x * kCvWarps + wid;
  if (q >= Q) return;
  int* ids = s_ids[wid];
  float(*gs)[8] = s_gs[wid];
  const int n = gather_views(pair_index, V, Q, q, lane, ids);
  const size_t hq = (size_t)Q * C;
  const size_t off = (size_t)q * C + lane * CPL;
  {
    const int h = lane & 7;
    const float d = __ldg(dglob + (size_t)q * 8 + h);
    for (int i = lane >> 3; i < n; i += 4) {
      const size_t o = (size_t)ids[i] * 8 + h;
      const float g = __ldg(alpha + o) * (__ldg(galpha + o) - d);
      gs[i][h] = g;
      gscore[o] = g;
    }
  }
  __syncwarp();
  float gq[8][CPL];
#pragma unroll
  for (int h = 0; h < 8; ++h)
#pragma unroll
    for (int j = 0; j < CPL; ++j) gq[h][j] = 0.f;
  for (int i = 0; i < n; ++i) {
    float x[CPL];
    load_row_cv<CPL>(x, slots + (size_t)ids[i] * C + lane * CPL);
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const float g = gs[i][h];
#pragma unroll
      for (int j = 0; j < CPL; ++j) gq[h][j] += g * x[j];
    }
  }
#pragma unroll
  for (int h = 0; h < 8; ++h) store_row_cv<CPL>(gqt + h * hq + off, gq[h]);
}

// Finishing step after an exchange of view sharding: y[r, c] = x[r, c] / max(s[r, c / (C / heads)], smin) (+ bias[c]) --
// the mean over ALL views from the reduced sums and counts (heads = 1, smin = 1; cnt[r] = (int)s[r] is emitted for the row
// masks of the layer), the softmax normalisation of the reduced per-head output partials (heads = 8, smin = 1e-30, bias =
// the value in-projection's), and the matching scaling of the upstream gradient in the backward.
__global__ void __launch_bounds__(256) rows_headscale_kernel(const float* __restrict__ x, const float* __restrict__ s,
                                                             int heads, float smin, const float* __restrict__ bias, int R,
                                                             int C, float* __restrict__ y, int* __restrict__ cnt) {
  const int c4 = C >> 2;
  const long long n4 = (long long)R * c4;
  const int dh = C / heads;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / c4), c = (int)(i - (long long)r * c4) * 4;
    const float d = fmaxf(__ldg(s + (size_t)r * heads + c / dh), smin);
    float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    v.x = __fdiv_rn(v.x, d); v.y = __fdiv_rn(v.y, d); v.z = __fdiv_rn(v.z, d); v.w = __fdiv_rn(v.w, d);
    if (bias) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c));
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    reinterpret_cast<float4*>(y)[i] = v;
    if (cnt && c == 0) cnt[r] = (int)__ldg(s + (size_t)r * heads);
  }
}

}  // namespace sgc

#define SGC_CV_LAUNCH(KERNEL, ...)                                                              \
  do {                                                                                          \
    if (C != 256 && C != 128) return (int)cudaErrorInvalidValue;                                \
    if (V > sgc::kMaxViews) return (int)cudaErrorInvalidValue;                                  \
    const int grid = (Q + sgc::kCvWarps - 1) / sgc::kCvWarps;                                   \
    if (C == 256) sgc::launch_chain(sgc::KERNEL<8>, dim3(grid), dim3(sgc::kCvWarps * 32), 0, (cudaStream_t)stream, __VA_ARGS__); \
    else sgc::launch_chain(sgc::KERNEL<4>, dim3(grid), dim3(sgc::kCvWarps * 32), 0, (cudaStream_t)stream, __VA_ARGS__); \
    SGC_CUDA_CHECK_LAST();                                                                      \
    return 0;                                                                                   \
  } while (0)

// levels of at most kCvSmallQ voxels fetch four rows per trip (see mean_fwd_kernel)
static const int kCvSmallQ = 2048;   // measured (session Z): 621.8 vs 619.4 volumes/s without; 621.8 for every level

static int cv_mean_fwd(const float* slots, const int* pair_index, int V, int Q, int C, float* mean, __nv_bfloat16* split,
                       void* stream) {
  if (C != 256 && C != 128) return (int)cudaErrorInvalidValue;
  if (V > sgc::kMaxViews) return (int)cudaErrorInvalidValue;
  const dim3 grid((Q + sgc::kCvWarps - 1) / sgc::kCvWarps), block(sgc::kCvWarps * 32);
  cudaStream_t st = (cudaStream_t)stream;
  const bool small = Q <= kCvSmallQ;
  if (C == 256 && small) sgc::launch_chain(sgc::mean_fwd_kernel<8, true, 4>, grid, block, 0, st, slots, pair_index, V, Q, mean, split);
  else if (C == 256) sgc::launch_chain(sgc::mean_fwd_kernel<8, true, 1>, grid, block, 0, st, slots, pair_index, V, Q, mean, split);
  else if (small) sgc::launch_chain(sgc::mean_fwd_kernel<4, true, 4>, grid, block, 0, st, slots, pair_index, V, Q, mean, split);
  else sgc::launch_chain(sgc::mean_fwd_kernel<4, true, 1>, grid, block, 0, st, slots, pair_index, V, Q, mean, split);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

static int cv_attn_fwd(const float* qt, const float* slots, const int* pair_index, int V, int Q, int C, float* t_out,
                       float* alpha, __nv_bfloat16* split, void* stream) {
  if (C != 256 && C != 128) return (int)cudaErrorInvalidValue;
  if (V > sgc::kMaxViews) return (int)cudaErrorInvalidValue;
  const dim3 grid((Q + sgc::kCvWarps - 1) / sgc::kCvWarps), block(sgc::kCvWarps * 32);
  cudaStream_t st = (cudaStream_t)stream;
  const bool small = Q <= kCvSmallQ;
  if (C == 256 && small) sgc::launch_chain(sgc::attn_fwd_kernel<8, 4>, grid, block, 0, st, qt, slots, pair_index, V, Q, t_out, alpha, split);
  else if (C == 256) sgc::launch_chain(sgc::attn_fwd_kernel<8, 1>, grid, block, 0, st, qt, slots, pair_index, V, Q, t_out, alpha, split);
  else if (small) sgc::launch_chain(sgc::attn_fwd_kernel<4, 4>, grid, block, 0, st, qt, slots, pair_index, V, Q, t_out, alpha, split);
  else sgc::launch_chain(sgc::attn_fwd_kernel<4, 1>, grid, block, 0, st, qt, slots, pair_index, V, Q, t_out, alpha, split);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

extern "C" int sgc_crossview_mean_fwd(const float* slots, const int* pair_index, int V, int Q, int C, float* mean,
                                      void* stream) {
  return cv_mean_fwd(slots, pair_index, V, Q, C, mean, nullptr, stream);
}

extern "C" int sgc_crossview_attn_fwd(const float* qt, const float* slots, const int* pair_index, int V, int Q, int C,
                                      float* t_out, float* alpha, void* stream) {
  return cv_attn_fwd(qt, slots, pair_index, V, Q, C, t_out, alpha, nullptr, stream);
}

extern "C" int sgc_crossview_attn_bwd_qt(const float* slots, const float* alpha, const int* pair_index, int V, int Q,
                                         int C, const float* grad_t, float* gscore, float* grad_qt, void* stream) {
  SGC_CV_LAUNCH(attn_bwd_qt_kernel, slots, alpha, pair_index, V, Q, grad_t, gscore, grad_qt, nullptr);
}

// Same three kernels, additionally emitting the bf16x3 operand image (sgc_split_bf16x3 pattern 0) of their dense output
// for the tensor-core GEMM that follows: mean [Q,3C], t [8*Q,3C], grad_qt [8*Q,3C].
extern "C" int sgc_crossview_mean_fwd_split(const float* slots, const int* pair_index, int V, int Q, int C, float* mean,
                                            void* split, void* stream) {
  return cv_mean_fwd(slots, pair_index, V, Q, C, mean, (__nv_bfloat16*)split, stream);
}
extern "C" int sgc_crossview_attn_fwd_split(const float* qt, const float* slots, const int* pair_index, int V, int Q, int C,
                                            float* t_out, float* alpha, void* split, void* stream) {
  return cv_attn_fwd(qt, slots, pair_index, V, Q, C, t_out, alpha, (__nv_bfloat16*)split, stream);
}
extern "C" int sgc_crossview_attn_bwd_qt_split(const float* slots, const float* alpha, const int* pair_index, int V, int Q,
                                               int C, const float* grad_t, float* gscore, float* grad_qt, void* split,
                                               void* stream) {
  SGC_CV_LAUNCH(attn_bwd_qt_kernel, slots, alpha, pair_index, V, Q, grad_t, gscore, grad_qt, (__nv_bfloat16*)split);
}

extern "C" int sgc_crossview_attn_bwd_slots(const float* qt, const float* alpha, const float* gscore,
                                            const int* pair_index, int V, int Q, int C, const float* grad_t,
                                            const float* grad_mean, float* grad_slots, void* stream) {
  SGC_CV_LAUNCH(attn_bwd_slots_kernel, qt, alpha, gscore, pair_index, V, Q, grad_t, grad_mean, grad_slots, nullptr);
}

// ---- view-sharded entry points (local halves; the exchange steps are host-side all-reduces) -------------------
#define SGC_CV_LAUNCH2(KERNEL, ...)                                                             \
  do {                                                                                          \
    if (C != 256 && C != 128) return (int)cudaErrorInvalidValue;                                \
    if (V > sgc::kMaxViews) return (int)cudaErrorInvalidValue;                                  \
    const int grid = (Q + sgc::kCvWarps - 1) / sgc::kCvWarps;                                   \
    if (C == 256) sgc::KERNEL<8><<<grid, sgc::kCvWarps * 32, 0, (cudaStream_t)stream>>>(__VA_ARGS__); \
    else sgc::KERNEL<4><<<grid, sgc::kCvWarps * 32, 0, (cudaStream_t)stream>>>(__VA_ARGS__);    \
    SGC_CUDA_CHECK_LAST();                                                                      \
    return 0;                                                                                   \
  } while (0)

extern "C" int sgc_crossview_sum_fwd(const float* slots, const int* pair_index, int V, int Q, int C, float* sum,
                                     void* stream) {
  if (C != 256 && C != 128) return (int)cudaErrorInvalidValue;
  if (V > sgc::kMaxViews) return (int)cudaErrorInvalidValue;
  const int grid = (Q + sgc::kCvWarps - 1) / sgc::kCvWarps;
  if (C == 256) sgc::mean_fwd_kernel<8, false><<<grid, sgc::kCvWarps * 32, 0, (cudaStream_t)stream>>>(slots, pair_index, V, Q, sum);
  else sgc::mean_fwd_kernel<4, false><<<grid, sgc::kCvWarps * 32, 0, (cudaStream_t)stream>>>(slots, pair_index, V, Q, sum);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

extern "C" int sgc_cvs_scores(const float* qt, const float* slots, const int* pair_index, int V, int Q, int C,
                              float* scores, float* m_loc, void* stream) {
  SGC_CV_LAUNCH2(cvs_scores_kernel, qt, slots, pair_index, V, Q, scores, m_loc);
}

extern "C" int sgc_cvs_accum(const float* scores, const float* m_glob, const float* slots, const int* pair_index, int V,
                             int Q, int C, float* e_out, float* s_loc, float* o_loc, void* stream) {
  SGC_CV_LAUNCH2(cvs_accum_kernel, scores, m_glob, slots, pair_index, V, Q, e_out, s_loc, o_loc);
}

extern "C" int sgc_cvs_bwd_dot(const float* slots, const float* e_in, const float* s_glob, const int* pair_index, int V,
                               int Q, int C, const float* grad_t, float* alpha, float* galpha, float* d_loc,
                               void* stream) {
  SGC_CV_LAUNCH2(cvs_bwd_dot_kernel, slots, e_in, s_glob, pair_index, V, Q, grad_t, alpha, galpha, d_loc);
}

extern "C" int sgc_cvs_bwd_qt(const float* slots, const float* alpha, const float* galpha, const float* d_glob,
                              const int* pair_index, int V, int Q, int C, float* gscore, float* gqt_loc, void* stream) {
  SGC_CV_LAUNCH2(cvs_bwd_qt_kernel, slots, alpha, galpha, d_glob, pair_index, V, Q, gscore, gqt_loc);
}

extern "C" int sgc_cvs_bwd_slots(const float* qt, const float* alpha, const float* gscore, const int* pair_index, int V,
                                 int Q, int C, const float* grad_t, const float* grad_mean, const int* count_glob,
                                 float* grad_slots, void* stream) {
  SGC_CV_LAUNCH2(attn_bwd_slots_kernel, qt, alpha, gscore, pair_index, V, Q, grad_t, grad_mean, grad_slots, count_glob);
}

extern "C" int sgc_rows_headscale(const float* x, const float* s, int heads, float smin, const float* bias, int R, int C,
                                  float* y, int* cnt, void* stream) {
  if (R <= 0 || C <= 0 || heads <= 0 || C % heads || (C / heads) % 4 || !x || !s || !y) return (int)cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(bias)) & 15)
    return (int)cudaErrorInvalidValue;
  const long long n4 = (long long)R * (C >> 2);
  long long grid = (n4 + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  sgc::rows_headscale_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(x, s, heads, smin, bias, R, C, y, cnt);
  SGC_CUDA_CHECK_LAST();
  return 0;
}
