// Sparse-volume construction kernels (AdaptiveSparseHead.py:43-98, DenseHead.py:64-83), channel-last volumes.
//
//   sgc_upsample2x_occ_fwd/bwd : F.interpolate(scale_factor=2, 'trilinear', align_corners=False)
//                                (AdaptiveSparseHead.py:64-69) fused with the occupancy head
//                                Linear(C,1)+Sigmoid (AdaptiveSparseHead.py:37-39,71)
//   sgc_topk_select            : topk_wo_grad (AdaptiveSparseHead.py:9-13) + nonzero compaction
//                                (DenseHead.py:66): deterministic, ties broken by lower index
//   sgc_scatter_add_rows / sgc_gather_rows : volume[sel] (+)= y (DenseHead.py:80-81 and the residual add
//                                AdaptiveSparseHead.py:77) and its backward
//
// Volumes are stored [X,Y,Z,C] (== torch.channels_last_3d of the reference's [1,C,X,Y,Z]); voxel flat index
// n = x*(Y*Z) + y*Z + z as in DenseHead.get_voxel_indices (DenseHead.py:32-37).
#include "common.cuh"

namespace sgc {

// source index/weights of one output coordinate, align_corners=False, scale 2 (ATen
// area_pixel_compute_source_index): src = max(0.5*(o+0.5)-0.5, 0)
__device__ __forceinline__ void up_src(int o, int n_in, int& i0, int& i1, float& l0, float& l1) {
  float s = 0.5f * ((float)o + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = (int)s;
  i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  l1 = s - (float)i0;
  l0 = 1.f - l1;
}

constexpr int kUpWarps = 8;

// one warp per output voxel; lane owns channels [lane*CPL, +CPL)
template <int CPL>
__global__ void __launch_bounds__(kUpWarps * 32) upsample_occ_fwd_kernel(const float* __restrict__ in, int X, int Y,
                                                                        int Z, const float* __restrict__ w_occ,
                                                                        const float* __restrict__ b_occ,
                                                                        float* __restrict__ out,
                                                                        float* __restrict__ occ) {
  constexpr int C = CPL * 32;
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int X2 = 2 * X, Y2 = 2 * Y, Z2 = 2 * Z;
  const int n_out = X2 * Y2 * Z2;
  const int n = blockIdx.x * kUpWarps + (threadIdx.x >> 5);
  if (n >= n_out) return;
  const int z = n % Z2, y = (n / Z2) % Y2, x = n / (Z2 * Y2);
  int xi[2], yi[2], zi[2];
  float xl[2], yl[2], zl[2];
  up_src(x, X, xi[0], xi[1], xl[0], xl[1]);
  up_src(y, Y, yi[0], yi[1], yl[0], yl[1]);
  up_src(z, Z, zi[0], zi[1], zl[0], zl[1]);
  float acc[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) acc[j] = 0.f;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float wgt = xl[a] * yl[b] * zl[c];
        const float* p = in + ((size_t)(xi[a] * Y + yi[b]) * Z + zi[c]) * C + lane * CPL;
#pragma unroll
        for (int j = 0; j < CPL; j += 4) {
          const float4 t = ldg4(p + j);
          acc[j] += wgt * t.x; acc[j + 1] += wgt * t.y; acc[j + 2] += wgt * t.z; acc[j + 3] += wgt * t.w;
        }
      }
  float dot = 0.f;
  float* o = out + (size_t)n * C + lane * CPL;
#pragma unroll
  for (int j = 0; j < CPL; j += 4) {
    const float4 wv = ldg4(w_occ + lane * CPL + j);
    dot += acc[j] * wv.x + acc[j + 1] * wv.y + acc[j + 2] * wv.z + acc[j + 3] * wv.w;
    *reinterpret_cast<float4*>(o + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
  }
  dot = warp_sum(dot);
  if (lane == 0) occ[n] = 1.f / (1.f + expf(-(dot + __ldg(b_occ))));
}

// per-voxel pre-activation gradient: gpre[n] = g_occ[n] * occ*(1-occ); also reduces grad_b
__global__ void occ_gpre_kernel(const float* __restrict__ occ, const float* __restrict__ g_occ, int n_out,
                                float* __restrict__ gpre, float* __restrict__ grad_b) {
  pdl_sync();
  float local = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += gridDim.x * blockDim.x) {
    const float s = occ[i];
    const float g = g_occ[i] * s * (1.f - s);
    gpre[i] = g;
    local += g;
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) red_add1(grad_b, local);
}

// grad_w_occ[c] += sum_n gpre[n] * up[n,c], with up recomputed from the coarse volume.
// one warp per output voxel chunk; block partials reduced through shared memory then one RED per channel.
template <int CPL>
__global__ void __launch_bounds__(kUpWarps * 32) occ_gradw_kernel(const float* __restrict__ in, int X, int Y, int Z,
                                                                 const float* __restrict__ gpre,
                                                                 float* __restrict__ grad_w) {
  constexpr int C = CPL * 32;
  __shared__ float part[kUpWarps][C];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int X2 = 2 * X, Y2 = 2 * Y, Z2 = 2 * Z;
  const int n_out = X2 * Y2 * Z2;
  float acc[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) acc[j] = 0.f;
  for (int n = blockIdx.x * kUpWarps + wid; n < n_out; n += gridDim.x * kUpWarps) {
    const float g = __ldg(gpre + n);
    if (g == 0.f) continue;
    const int z = n % Z2, y = (n / Z2) % Y2, x = n / (Z2 * Y2);
    int xi[2], yi[2], zi[2];
    float xl[2], yl[2], zl[2];
    up_src(x, X, xi[0], xi[1], xl[0], xl[1]);
    up_src(y, Y, yi[0], yi[1], yl[0], yl[1]);
    up_src(z, Z, zi[0], zi[1], zl[0], zl[1]);
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float wgt = g * xl[a] * yl[b] * zl[c];
          const float* p = in + ((size_t)(xi[a] * Y + yi[b]) * Z + zi[c]) * C + lane * CPL;
#pragma unroll
          for (int j = 0; j < CPL; j += 4) {
            const float4 t = ldg4(p + j);
            acc[j] += wgt * t.x; acc[j + 1] += wgt * t.y; acc[j + 2] += wgt * t.z; acc[j + 3] += wgt * t.w;
          }
        }
  }
#pragma unroll
  for (int j = 0; j < CPL; ++j) part[wid][lane * CPL + j] = acc[j];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kUpWarps; ++w) s += part[w][c];
    red_add1(grad_w + c, s);
  }
}

// grad_in[i] = sum over the (<=4 per axis) outputs that read input i of  weight * (grad_up[o] + gpre[o]*w_occ)
// (gather form of the transposed upsample: deterministic, no atomics).  One warp per input voxel.
__device__ __forceinline__ int up_adj(int i, int n_in, int (&o)[4], float (&w)[4]) {
  // outputs 2i-1, 2i, 2i+1, 2i+2 can touch input i; recompute their true weights to honour edge clamping
  int cnt = 0;
  for (int d = -1; d <= 2; ++d) {
    const int oo = 2 * i + d;
    if (oo < 0 || oo >= 2 * n_in) continue;
    int i0, i1;
    float l0, l1;
    up_src(oo, n_in, i0, i1, l0, l1);
    float ww = 0.f;
    if (i0 == i) ww += l0;
    if (i1 == i) ww += l1;
    if (ww != 0.f) { o[cnt] = oo; w[cnt] = ww; ++cnt; }
  }
  return cnt;
}

template <int CPL>
__global__ void __launch_bounds__(kUpWarps * 32) upsample_occ_bwd_kernel(const float* __restrict__ grad_up,
                                                                        const float* __restrict__ gpre,
                                                                        const float* __restrict__ w_occ, int X, int Y,
                                                                        int Z, float* __restrict__ grad_in) {
  constexpr int C = CPL * 32;
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int n_in = X * Y * Z;
  const int n = blockIdx.x * kUpWarps + (threadIdx.x >> 5);
  if (n >= n_in) return;
  const int z = n % Z, y = (n / Z) % Y, x = n / (Z * Y);
  const int Y2 = 2 * Y, Z2 = 2 * Z;
  int xo[4], yo[4], zo[4];
  float xw[4], yw[4], zw[4];
  const int nx = up_adj(x, X, xo, xw), ny = up_adj(y, Y, yo, yw), nz = up_adj(z, Z, zo, zw);
  float wv[CPL];
#pragma unroll
  for (int j = 0; j < CPL; j += 4) {
    const float4 t = ldg4(w_occ + lane * CPL + j);
    wv[j] = t.x; wv[j + 1] = t.y; wv[j + 2] = t.z; wv[j + 3] = t.w;
  }
  float acc[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) acc[j] = 0.f;
  for (int a = 0; a < nx; ++a)
    for (int b = 0; b < ny; ++b)
      for (int c = 0; c < nz; ++c) {
        const size_t o = ((size_t)xo[a] * Y2 + yo[b]) * Z2 + zo[c];
        const float wgt = xw[a] * yw[b] * zw[c];
        const float gp = gpre ? __ldg(gpre + o) : 0.f;
        const float* p = grad_up + o * C + lane * CPL;
#pragma unroll
        for (int j = 0; j < CPL; j += 4) {
          const float4 t = ldg4(p + j);
          acc[j] += wgt * (t.x + gp * wv[j]);
          acc[j + 1] += wgt * (t.y + gp * wv[j + 1]);
          acc[j + 2] += wgt * (t.z + gp * wv[j + 2]);
          acc[j + 3] += wgt * (t.w + gp * wv[j + 3]);
        }
      }
  float* g = grad_in + (size_t)n * C + lane * CPL;
#pragma unroll
  for (int j = 0; j < CPL; j += 4) *reinterpret_cast<float4*>(g + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
}

// The same gradient evaluated axis by axis (the x2 trilinear stencil is separable): z, then y, then x, each pass the
// transposed 1-D stencil  dst[a, i, e] = sum_d w_d * src[a, o_d, e]  over the <= 4 outputs o_d = 2i-1 .. 2i+2 that read input i
// (weights recomputed per output, so the clamped borders are honoured).  src [A, 2N, E], dst [A, N, E]; one thread per
// float4 of dst, so every pass streams whole rows: 2.6 x the input bytes in total instead of 64 row gathers per voxel
// (8 x the bytes, latency-bound: 0.73 ms at the 40x40x16 level of the "-L" configs).  The first pass (E == C) also adds the
// occupancy head's contribution gpre[row] * w_occ[c].
__global__ void __launch_bounds__(256) up_bwd_axis_kernel(const float* __restrict__ src, const float* __restrict__ gpre,
                                                          const float* __restrict__ w_occ, int A, int N, int E4,
                                                          float* __restrict__ dst) {
  pdl_sync();
  const long long total = (long long)A * N * E4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int e4 = (int)(idx % E4);
    const long long t = idx / E4;
    const int i = (int)(t % N);
    const long long a = t / N;
    float w[4];
    bool on[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const int oo = 2 * i - 1 + d;
      w[d] = 0.f;
      on[d] = false;
      if (oo >= 0 && oo < 2 * N) {
        int i0, i1;
        float l0, l1;
        up_src(oo, N, i0, i1, l0, l1);
        if (i0 == i) w[d] += l0;
        if (i1 == i) w[d] += l1;
        on[d] = w[d] != 0.f;
      }
    }
    float4 v[4];
    float gp[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      v[d] = make_float4(0.f, 0.f, 0.f, 0.f);
      gp[d] = 0.f;
      if (on[d]) {
        const long long row = a * 2 * N + (2 * i - 1 + d);
        v[d] = ldg4(src + (row * E4 + e4) * 4);
        if (gpre) gp[d] = __ldg(gpre + row);
      }
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gpre) {
      const float4 wv = ldg4(w_occ + e4 * 4);
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        acc.x += w[d] * (v[d].x + gp[d] * wv.x); acc.y += w[d] * (v[d].y + gp[d] * wv.y);
        acc.z += w[d] * (v[d].z + gp[d] * wv.z); acc.w += w[d] * (v[d].w + gp[d] * wv.w);
      }
    } else {
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        acc.x += w[d] * v[d].x; acc.y += w[d] * v[d].y; acc.z += w[d] * v[d].z; acc.w += w[d] * v[d].w;
      }
    }
    reinterpret_cast<float4*>(dst)[idx] = acc;
  }
}

// ---------------------------------------------------------------------------------- rows
// out[sel[i], :] += y[i, :]   (sel entries are unique -> plain read-modify-write)
__global__ void scatter_add_rows_kernel(float* __restrict__ vol, const int* __restrict__ sel, const float* __restrict__ y,
                                        int k, int C4) {
  pdl_sync();
  const int total = k * C4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / C4, c = i - r * C4;
    float4* dst = reinterpret_cast<float4*>(vol) + (size_t)__ldg(sel + r) * C4 + c;
    const float4 a = *dst, b = __ldg(reinterpret_cast<const float4*>(y) + i);
    *dst = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}
__global__ void gather_rows_kernel(const float* __restrict__ vol, const int* __restrict__ sel, float* __restrict__ y,
                                   int k, int C4) {
  pdl_sync();
  const int total = k * C4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / C4, c = i - r * C4;
    reinterpret_cast<float4*>(y)[i] = __ldg(reinterpret_cast<const float4*>(vol) + (size_t)__ldg(sel + r) * C4 + c);
  }
}

// ---------------------------------------------------------------------------------- top-k
// Single-CTA radix select (4 x 8 bits, MSB first) + ordered compaction.  Deterministic:
// the k largest values, ties at the threshold taken in ascending index order; sel is ascending.
__device__ __forceinline__ uint32_t topk_key(float f) {
  f += 0.f;  // -0 -> +0
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(1024) topk_select_kernel(const float* __restrict__ occ, int N, int k,
                                                          int* __restrict__ sel, uint8_t* __restrict__ mask) {
  __shared__ int hist[256];
  __shared__ uint32_t s_prefix;
  __shared__ int s_need;  // how many more are needed among keys matching the prefix so far
  __shared__ int warp_a[32], warp_b[32];
  __shared__ int s_carry_sel, s_carry_eq;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) { s_prefix = 0u; s_need = k; }
  __syncthreads();
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    const uint32_t pmask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (int i = tid; i < 256; i += 1024) hist[i] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    for (int base = 0; base < N; base += 1024) {
      const int i = base + tid;
      bool act = false;
      uint32_t bin = 0;
      if (i < N) {
        const uint32_t key = topk_key(occ[i]);
        act = (key & pmask) == prefix;
        bin = (key >> shift) & 0xffu;
      }
      // warp-aggregated histogram update
      const unsigned am = __ballot_sync(SGC_FULL_MASK, act);
      if (act) {
        const unsigned peers = __match_any_sync(am, bin);
        if ((int)(__ffs(peers) - 1) == lane) atomicAdd(&hist[bin], __popc(peers));
      }
    }
    __syncthreads();
    if (tid == 0) {
      int need = s_need, b = 255;
      for (; b > 0; --b) {
        if (hist[b] >= need) break;
        need -= hist[b];
      }
      s_need = need;  // still needed inside bin b
      s_prefix = prefix | ((uint32_t)b << shift);
    }
    __syncthreads();
  }
  const uint32_t T = s_prefix;  // threshold key
  const int need_eq = s_need;   // number of keys == T to take (lowest indices first)
  if (tid == 0) { s_carry_sel = 0; s_carry_eq = 0; }
  __syncthreads();
  for (int base = 0; base < N; base += 1024) {
    const int i = base + tid;
    uint32_t key = 0;
    if (i < N) key = topk_key(occ[i]);
    const bool gt = (i < N) && key > T;
    const bool eq = (i < N) && key == T;
    const unsigned bg = __ballot_sync(SGC_FULL_MASK, gt), be = __ballot_sync(SGC_FULL_MASK, eq);
    if (lane == 0) { warp_a[wid] = __popc(bg); warp_b[wid] = __popc(be); }
    __syncthreads();
    if (wid == 0) {
      int a = warp_a[lane], b = warp_b[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int ya = __shfl_up_sync(SGC_FULL_MASK, a, o), yb = __shfl_up_sync(SGC_FULL_MASK, b, o);
        if (lane >= o) { a += ya; b += yb; }
      }
      warp_a[lane] = a; warp_b[lane] = b;
    }
    __syncthreads();
    const unsigned lt = (1u << lane) - 1;
    const int gt_before = s_carry_sel + (wid ? warp_a[wid - 1] : 0) + __popc(bg & lt);  // counts only gt so far
    const int eq_before = s_carry_eq + (wid ? warp_b[wid - 1] : 0) + __popc(be & lt);
    const bool take = gt || (eq && eq_before < need_eq);
    if (i < N) {
      mask[i] = take ? 1 : 0;
      if (take) sel[gt_before + (eq_before < need_eq ? eq_before : need_eq)] = i;
    }
    __syncthreads();
    if (tid == 1023) { s_carry_sel = gt_before + (gt ? 1 : 0); s_carry_eq = eq_before + (eq ? 1 : 0); }
    __syncthreads();
  }
}

// Fast path for N <= 32768 (every level of the C=256 configs): the whole key set lives in registers (32 keys per thread),
// 3 radix passes of 11/11/10 bits with a shared 2048-bin histogram and a parallel suffix scan, then the ordered
// compaction from registers.  Same deterministic contract as topk_select_kernel.
constexpr int kTopkKpt = 32;

__device__ __forceinline__ int block_excl_scan_1024(int v, int* warp_tot, int& total) {
  // exclusive prefix of v over the 1024 threads of the block (thread order), total = sum
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(SGC_FULL_MASK, inc, o);
    if (lane >= o) inc += y;
  }
  __syncthreads();
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
  total = warp_tot[31];
  return (wid ? warp_tot[wid - 1] : 0) + inc - v;
}

__global__ void __launch_bounds__(1024) topk_select_small_kernel(const float* __restrict__ occ, int N, int k,
                                                                int* __restrict__ sel, uint8_t* __restrict__ mask) {
  __shared__ int hist[2048];
  __shared__ int warp_tot[32];
  __shared__ uint32_t s_prefix;
  __shared__ int s_need;
  pdl_sync();
  const int tid = threadIdx.x;
  // thread t owns the CONTIGUOUS index range [t*per, (t+1)*per): the ordered compaction then needs one block scan
  const int per = (N + 1023) / 1024;
  uint32_t key[kTopkKpt];
#pragma unroll
  for (int j = 0; j < kTopkKpt; ++j) {
    const int i = tid * per + j;
    key[j] = (j < per && i < N) ? topk_key(__ldg(occ + i)) : 0u;
  }
  __shared__ int warp_cnt[2][32];
  const int lane = tid & 31, wid = tid >> 5;
  // Leading 11 bits of the threshold (the k-th largest key) bit by bit from the top: T keeps a bit whenever at least k
  // keys are >= the candidate.  Occupancy scores are sigmoid outputs, so these bits put nearly all keys into a handful of
  // histogram bins and the shared-memory atomics of a first histogram pass serialise (~25 of the kernel's 33 us); a
  // step of the search is 32 register compares, a hardware warp reduction and ONE barrier (double-buffered counts).
  // Unused key slots are 0 and every candidate is >= 1, so they never count.  The remaining 21 bits are resolved by the
  // two histogram passes below, where only the keys sharing the 11-bit prefix take part.
  const int shifts[3] = {21, 10, 0};
  const int bits[3] = {11, 11, 10};
  uint32_t Tp = 0u;
  int above = 0;            // keys >= the last REJECTED candidate (= Tp + one 11-bit unit), i.e. keys in higher 11-bit bins
  for (int bit = 31; bit >= 21; --bit) {
    const uint32_t cand = Tp | (1u << bit);
    int c = 0;
#pragma unroll
    for (int j = 0; j < kTopkKpt; ++j) c += key[j] >= cand ? 1 : 0;
    c = __reduce_add_sync(SGC_FULL_MASK, c);
    int* buf = warp_cnt[bit & 1];
    if (lane == 0) buf[wid] = c;
    __syncthreads();
    const int tot = __reduce_add_sync(SGC_FULL_MASK, buf[lane]);
    if (tot >= k) Tp = cand; else above = tot;   // the last rejected candidate bounds Tp's bin from above
  }
  if (tid == 0) { s_prefix = Tp; s_need = k - above; }
  uint32_t pmask = 0xFFE00000u;
  for (int pass = 1; pass < 3; ++pass) {
    const int shift = shifts[pass], nb = 1 << bits[pass];
    for (int b = tid; b < 2048; b += 1024) hist[b] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    // plain shared-memory atomics: measured faster than a match_any warp aggregation (33 vs 51 us at N = 25 600) even
    // though sigmoid outputs crowd the leading bits into a handful of bins
#pragma unroll
    for (int j = 0; j < kTopkKpt; ++j) {
      const int i = tid * per + j;
      if (j < per && i < N && (key[j] & pmask) == prefix) atomicAdd(&hist[(key[j] >> shift) & (nb - 1)], 1);
    }
    __syncthreads();
    // suffix counts: thread t handles bins 2t, 2t+1 (descending order = ascending index in the reversed array)
    const int b_hi = 2047 - 2 * tid, b_lo = b_hi - 1;      // reversed: position 2t <-> bin b_hi
    const int c_hi = hist[b_hi], c_lo = hist[b_lo];
    int total;
    const int before = block_excl_scan_1024(c_hi + c_lo, warp_tot, total);  // count of keys in strictly higher bins
    const int need = s_need;
    __syncthreads();
    // the threshold bin is the first (from the top) whose cumulative count reaches `need`
    if (before < need && before + c_hi >= need) { s_prefix = prefix | ((uint32_t)b_hi << shift); s_need = need - before; }
    else if (before + c_hi < need && before + c_hi + c_lo >= need) { s_prefix = prefix | ((uint32_t)b_lo << shift); s_need = need - before - c_hi; }
    __syncthreads();
    pmask |= (uint32_t)(nb - 1) << shift;
  }
  const uint32_t T = s_prefix;
  const int need_eq = s_need;
  int ngt = 0, neq = 0;
#pragma unroll
  for (int j = 0; j < kTopkKpt; ++j) {
    const int i = tid * per + j;
    if (j < per && i < N) { ngt += key[j] > T; neq += key[j] == T; }
  }
  int tot_gt, tot_eq;
  int gt_before = block_excl_scan_1024(ngt, warp_tot, tot_gt);
  int eq_before = block_excl_scan_1024(neq, warp_tot, tot_eq);
#pragma unroll
  for (int j = 0; j < kTopkKpt; ++j) {
    const int i = tid * per + j;
    if (j < per && i < N) {
      const bool gt = key[j] > T, eq = key[j] == T;
      const bool take = gt || (eq && eq_before < need_eq);
      mask[i] = take ? 1 : 0;
      if (take) sel[gt_before + (eq_before < need_eq ? eq_before : need_eq)] = i;
      gt_before += gt;
      eq_before += eq;
    }
  }
}


// ---------------------------------------------------------------------------------- top-k over many CTAs (N > 32 768)
// The finest level of the "-L" configs selects 51 200 of 204 800 voxels; one CTA streaming over them four times took
// ~0.4 ms on the critical path.  Same radix select (11 / 11 / 10 bits, MSB first) and the same deterministic contract, as
// a short sequence of launches: every CTA histograms its 2048-key chunk in shared memory and adds it to a global
// histogram with INTEGER atomics (order-independent, so the result is deterministic); one CTA picks the threshold bin;
// after three rounds a count launch and an ordered-compaction launch write mask / sel.
// scratch (ints): [0,2048) histogram, [2048] prefix, [2049] need, [2050] prefix mask, [2052, 2052+G) keys > T per chunk,
// [2052+G, 2052+2G) keys == T per chunk.
constexpr int kMcChunk = 2048;
constexpr int kMcState = 2048;
constexpr int kMcCounts = 2052;

__global__ void __launch_bounds__(1024) topk_mc_init_kernel(int* __restrict__ scratch, int k) {
  for (int b = threadIdx.x; b < 2048; b += 1024) scratch[b] = 0;
  if (threadIdx.x == 0) { scratch[kMcState] = 0; scratch[kMcState + 1] = k; scratch[kMcState + 2] = 0; }
}

__global__ void __launch_bounds__(1024) topk_mc_hist_kernel(const float* __restrict__ occ, int N, int* __restrict__ scratch,
                                                           int shift, int nb) {
  __shared__ int hist[2048];
  for (int b = threadIdx.x; b < 2048; b += 1024) hist[b] = 0;
  __syncthreads();
  const uint32_t prefix = (uint32_t)scratch[kMcState], pmask = (uint32_t)scratch[kMcState + 2];
#pragma unroll
  for (int j = 0; j < kMcChunk / 1024; ++j) {
    const int i = blockIdx.x * kMcChunk + j * 1024 + threadIdx.x;
    if (i < N) {
      const uint32_t key = topk_key(__ldg(occ + i));
      if ((key & pmask) == prefix) atomicAdd(&hist[(key >> shift) & (nb - 1)], 1);
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < nb; b += 1024) {
    const int c = hist[b];
    if (c) atomicAdd(scratch + b, c);
  }
}

__global__ void __launch_bounds__(1024) topk_mc_pick_kernel(int* __restrict__ scratch, int shift, int nb) {
  __shared__ int warp_tot[32];
  const int tid = threadIdx.x;
  // suffix counts over the (up to 2048) bins: thread t handles bins 2047-2t and 2046-2t, as in topk_select_small_kernel
  const int b_hi = 2047 - 2 * tid, b_lo = b_hi - 1;
  const int c_hi = b_hi < nb ? scratch[b_hi] : 0, c_lo = b_lo < nb ? scratch[b_lo] : 0;
  int total;
  const int before = block_excl_scan_1024(c_hi + c_lo, warp_tot, total);
  const uint32_t prefix = (uint32_t)scratch[kMcState], pmask = (uint32_t)scratch[kMcState + 2];
  const int need = scratch[kMcState + 1];
  __syncthreads();   // every thread has read the state and its bins
  if (before < need && before + c_hi >= need) {
    scratch[kMcState] = (int)(prefix | ((uint32_t)b_hi << shift));
    scratch[kMcState + 1] = need - before;
  } else if (before + c_hi < need && before + c_hi + c_lo >= need) {
    scratch[kMcState] = (int)(prefix | ((uint32_t)b_lo << shift));
    scratch[kMcState + 1] = need - before - c_hi;
  }
  if (tid == 0) scratch[kMcState + 2] = (int)(pmask | ((uint32_t)(nb - 1) << shift));
  scratch[b_hi] = 0;   // histogram cleared for the next round
  scratch[b_lo] = 0;
}

__global__ void __launch_bounds__(1024) topk_mc_count_kernel(const float* __restrict__ occ, int N, int* __restrict__ scratch) {
  __shared__ int warp_a[32], warp_b[32];
  const uint32_t T = (uint32_t)scratch[kMcState];
  int gt = 0, eq = 0;
#pragma unroll
  for (int j = 0; j < kMcChunk / 1024; ++j) {
    const int i = blockIdx.x * kMcChunk + j * 1024 + threadIdx.x;
    if (i < N) {
      const uint32_t key = topk_key(__ldg(occ + i));
      gt += key > T;
      eq += key == T;
    }
  }
  gt = __reduce_add_sync(SGC_FULL_MASK, gt);
  eq = __reduce_add_sync(SGC_FULL_MASK, eq);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { warp_a[wid] = gt; warp_b[wid] = eq; }
  __syncthreads();
  if (wid == 0) {
    const int a = __reduce_add_sync(SGC_FULL_MASK, warp_a[lane]), b = __reduce_add_sync(SGC_FULL_MASK, warp_b[lane]);
    if (lane == 0) { scratch[kMcCounts + blockIdx.x] = a; scratch[kMcCounts + gridDim.x + blockIdx.x] = b; }
  }
}

__global__ void __launch_bounds__(1024) topk_mc_write_kernel(const float* __restrict__ occ, int N, const int* __restrict__ scratch,
                                                            int* __restrict__ sel, uint8_t* __restrict__ mask) {
  __shared__ int warp_a[32], warp_b[32];
  __shared__ int s_carry_sel, s_carry_eq;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t T = (uint32_t)scratch[kMcState];
  const int need_eq = scratch[kMcState + 1];
  // keys > T / == T in the chunks before this one (fixed order: deterministic)
  int a = 0, b = 0;
  for (int c = tid; c < (int)blockIdx.x; c += 1024) { a += scratch[kMcCounts + c]; b += scratch[kMcCounts + gridDim.x + c]; }
  a = __reduce_add_sync(SGC_FULL_MASK, a);
  b = __reduce_add_sync(SGC_FULL_MASK, b);
  if (lane == 0) { warp_a[wid] = a; warp_b[wid] = b; }
  __syncthreads();
  if (wid == 0) {
    const int ta = __reduce_add_sync(SGC_FULL_MASK, warp_a[lane]), tb = __reduce_add_sync(SGC_FULL_MASK, warp_b[lane]);
    if (lane == 0) { s_carry_sel = ta; s_carry_eq = tb; }
  }
  __syncthreads();
  for (int j = 0; j < kMcChunk / 1024; ++j) {
    const int i = blockIdx.x * kMcChunk + j * 1024 + tid;
    uint32_t key = 0;
    if (i < N) key = topk_key(__ldg(occ + i));
    const bool gt = (i < N) && key > T;
    const bool eq = (i < N) && key == T;
    const unsigned bg = __ballot_sync(SGC_FULL_MASK, gt), be = __ballot_sync(SGC_FULL_MASK, eq);
    if (lane == 0) { warp_a[wid] = __popc(bg); warp_b[wid] = __popc(be); }
    __syncthreads();
    if (wid == 0) {
      int x = warp_a[lane], y = warp_b[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int xa = __shfl_up_sync(SGC_FULL_MASK, x, o), yb = __shfl_up_sync(SGC_FULL_MASK, y, o);
        if (lane >= o) { x += xa; y += yb; }
      }
      warp_a[lane] = x; warp_b[lane] = y;
    }
    __syncthreads();
    const unsigned lt = (1u << lane) - 1;
    const int gt_before = s_carry_sel + (wid ? warp_a[wid - 1] : 0) + __popc(bg & lt);
    const int eq_before = s_carry_eq + (wid ? warp_b[wid - 1] : 0) + __popc(be & lt);
    const bool take = gt || (eq && eq_before < need_eq);
    if (i < N) {
      mask[i] = take ? 1 : 0;
      if (take) sel[gt_before + (eq_before < need_eq ? eq_before : need_eq)] = i;
    }
    __syncthreads();
    if (tid == 1023) { s_carry_sel = gt_before + (gt ? 1 : 0); s_carry_eq = eq_before + (eq ? 1 : 0); }
    __syncthreads();
  }
}


// ---------------------------------------------------------------------------------- top-k over a few co-operating CTAs
// One launch for every level size (N <= 229 376): G = ceil(N / 3584) CTAs of 512 threads, at most 7 keys per thread in
// registers, thread g owning the CONTIGUOUS index range [g*per, (g+1)*per).  The threshold (the k-th largest key) is found
// digit by digit, 4 bits per step from the top (8 steps): every thread histograms the digit of its active keys (those
// matching the prefix found so far) into 16 packed nibbles; the counts are reduced over the warp, the CTA and -- for
// G > 1 -- over the grid by ONE 64-bit atomicAdd per word that carries three 19-bit counts AND a 7-bit arrival counter,
// so a step costs one atomic round trip plus a poll (no fence, no separate barrier).  No histogram, no shared-memory
// atomics (sigmoid scores crowd the leading bits into a few bins), fully deterministic.  The ordered compaction publishes
// each CTA's (#keys > T, #keys == T) in one flagged 64-bit word that the higher-ranked CTAs poll.
// The CTAs spin on global memory, which needs every CTA of the (tiny) grid to become resident eventually: all other
// kernels of the library are finite, so at worst the launch waits for SMs to free up; no cluster (a cluster launch next to
// the persistent projection kernels waits for 8 free SMs in ONE GPC).
// scratch (256 x u64, zero-initialised ONCE by the caller, private to a stream): [0] = parity; two halves of
// kTkHalf words are used alternately, and every call clears the half the NEXT call will use.
constexpr int kTkThreads = 512;
constexpr int kTkKpt = 7;        // keys per thread: a nibble of the packed digit histogram holds at most 7
constexpr int kTkSteps = 8;
constexpr int kTkWords = 5;        // 15 counts, three per word
constexpr int kTkMaxG = 64;
constexpr int kTkHalf = kTkSteps * kTkWords + kTkMaxG;

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kTkThreads) topk_select_grid_kernel(const float* __restrict__ occ, int N, int k, int per,
                                                                      int* __restrict__ sel, uint8_t* __restrict__ mask,
                                                                      unsigned long long* __restrict__ scratch) {
  __shared__ int s_warp[2][kTkThreads / 32][4];
  __shared__ int s_tot[2][16];
  __shared__ int s_scan[2][kTkThreads / 32];
  __shared__ int s_carry[2];
  pdl_sync();
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int G = gridDim.x;
  const unsigned long long par = ld_volatile_u64(scratch) & 1ull;
  unsigned long long* mine = scratch + 1 + par * kTkHalf;
  unsigned long long* other = scratch + 1 + (par ^ 1ull) * kTkHalf;
  if (blockIdx.x == 0)
    for (int i = tid; i < kTkHalf; i += kTkThreads) other[i] = 0ull;   // nobody touches this half during this call
  const int g = blockIdx.x * kTkThreads + tid;
  uint32_t key[kTkKpt];
#pragma unroll
  for (int j = 0; j < kTkKpt; ++j) {
    const long long i = (long long)g * per + j;
    key[j] = (j < per && i < N) ? topk_key(__ldg(occ + i)) : 0u;   // unused slots: digit 0 everywhere, never counted
  }
  uint32_t prefix = 0u, pmask = 0u;
  int need = k;
  for (int s = 0; s < kTkSteps; ++s) {
    const int shift = 28 - 4 * s;
    // per-thread histogram of the digit over the active keys, one NIBBLE per digit value (at most 7 keys per thread), split
    // into even / odd digits as bytes so that a warp sum (<= 224 per byte) cannot carry: 4 warp reductions per step
    unsigned long long h = 0ull;
#pragma unroll
    for (int j = 0; j < kTkKpt; ++j) {
      const bool act = (key[j] & pmask) == prefix;
      h += act ? (1ull << (4 * ((key[j] >> shift) & 15u))) : 0ull;   // unused slots: digit 0, never looked at
    }
    const unsigned long long ev = h & 0x0F0F0F0F0F0F0F0Full, od = (h >> 4) & 0x0F0F0F0F0F0F0F0Full;
    const uint32_t r0 = __reduce_add_sync(SGC_FULL_MASK, (uint32_t)ev), r1 = __reduce_add_sync(SGC_FULL_MASK, (uint32_t)(ev >> 32));
    const uint32_t r2 = __reduce_add_sync(SGC_FULL_MASK, (uint32_t)od), r3 = __reduce_add_sync(SGC_FULL_MASK, (uint32_t)(od >> 32));
    if (lane == 0) {   // s_warp[.][wid][q]: bytes = digits (0,2,4,6) (8,10,12,14) (1,3,5,7) (9,11,13,15)
      s_warp[s & 1][wid][0] = (int)r0; s_warp[s & 1][wid][1] = (int)r1; s_warp[s & 1][wid][2] = (int)r2; s_warp[s & 1][wid][3] = (int)r3;
    }
    __syncthreads();
    if (tid < kTkWords) {
      // thread w merges the counts of digits 3w+1 .. 3w+3 over the CTA's warps and, for G > 1, over the grid
      int n3[3] = {0, 0, 0};
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int d = 3 * tid + 1 + q;                         // 1..15
        const int word = ((d & 1) ? 2 : 0) + (d >> 3), byte = (d >> 1) & 3;
        int a = 0;
        for (int w = 0; w < kTkThreads / 32; ++w) a += (s_warp[s & 1][w][word] >> (8 * byte)) & 0xFF;
        n3[q] = a;
      }
      if (G > 1) {
        unsigned long long* word = mine + s * kTkWords + tid;
        const unsigned long long add = (1ull << 57) | ((unsigned long long)n3[2] << 38) | ((unsigned long long)n3[1] << 19) |
                                       (unsigned long long)n3[0];
        unsigned long long v = atomicAdd(word, add) + add;
        while ((v >> 57) != (unsigned long long)G) v = ld_volatile_u64(word);
        n3[0] = (int)(v & 0x7FFFFull); n3[1] = (int)((v >> 19) & 0x7FFFFull); n3[2] = (int)((v >> 38) & 0x7FFFFull);
      }
      s_tot[s & 1][3 * tid] = n3[0]; s_tot[s & 1][3 * tid + 1] = n3[1]; s_tot[s & 1][3 * tid + 2] = n3[2];
    }
    __syncthreads();
    // s_tot[.][i-1] = #active keys with digit == i.  The threshold's digit is the largest d whose count of keys with a digit
    // >= d still reaches `need`; the keys in strictly higher digits are all taken.
    int d = 0, above = 0, acc = 0;
#pragma unroll
    for (int i = 15; i >= 1; --i) {
      const int hi = s_tot[s & 1][i - 1];
      if (d == 0 && acc + hi >= need) { d = i; above = acc; }
      acc += hi;
    }
    need -= d ? above : acc;
    prefix |= (uint32_t)d << shift;
    pmask |= 15u << shift;
  }
  const uint32_t T = prefix;   // threshold key
  const int need_eq = need;    // keys == T to take, lowest indices first
  int ngt = 0, neq = 0;
#pragma unroll
  for (int j = 0; j < kTkKpt; ++j) {
    const long long i = (long long)g * per + j;
    if (j < per && i < N) { ngt += key[j] > T; neq += key[j] == T; }
  }
  // exclusive scan over the CTA's threads (thread order == index order)
  int igt = ngt, ieq = neq;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int a = __shfl_up_sync(SGC_FULL_MASK, igt, o), b = __shfl_up_sync(SGC_FULL_MASK, ieq, o);
    if (lane >= o) { igt += a; ieq += b; }
  }
  if (lane == 31) { s_scan[0][wid] = igt; s_scan[1][wid] = ieq; }
  __syncthreads();
  if (wid == 0) {
    int a = lane < kTkThreads / 32 ? s_scan[0][lane] : 0, b = lane < kTkThreads / 32 ? s_scan[1][lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int ya = __shfl_up_sync(SGC_FULL_MASK, a, o), yb = __shfl_up_sync(SGC_FULL_MASK, b, o);
      if (lane >= o) { a += ya; b += yb; }
    }
    if (lane < kTkThreads / 32) { s_scan[0][lane] = a; s_scan[1][lane] = b; }
  }
  __syncthreads();
  int gt_before = (wid ? s_scan[0][wid - 1] : 0) + igt - ngt;
  int eq_before = (wid ? s_scan[1][wid - 1] : 0) + ieq - neq;
  if (G > 1) {
    unsigned long long* slots = mine + kTkSteps * kTkWords;
    if (tid == 0) {
      const unsigned long long v = (1ull << 63) | ((unsigned long long)s_scan[0][kTkThreads / 32 - 1] << 32) |
                                   (unsigned long long)s_scan[1][kTkThreads / 32 - 1];
      asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(slots + blockIdx.x), "l"(v) : "memory");
    }
    if (wid == 0) {   // totals of the lower-ranked CTAs (at most 63: two per lane)
      int a = 0, b = 0;
      for (int cta = lane; cta < (int)blockIdx.x; cta += 32) {
        unsigned long long v = ld_volatile_u64(slots + cta);
        while (!(v >> 63)) v = ld_volatile_u64(slots + cta);
        a += (int)((v >> 32) & 0x7FFFFFFFull);
        b += (int)(v & 0xFFFFFFFFull);
      }
      a = __reduce_add_sync(SGC_FULL_MASK, a);
      b = __reduce_add_sync(SGC_FULL_MASK, b);
      if (lane == 0) { s_carry[0] = a; s_carry[1] = b; }
    }
    __syncthreads();
    gt_before += s_carry[0];
    eq_before += s_carry[1];
  }
#pragma unroll
  for (int j = 0; j < kTkKpt; ++j) {
    const long long i = (long long)g * per + j;
    if (j < per && i < N) {
      const bool gt = key[j] > T, eq = key[j] == T;
      const bool take = gt || (eq && eq_before < need_eq);
      mask[i] = take ? 1 : 0;
      if (take) sel[gt_before + (eq_before < need_eq ? eq_before : need_eq)] = (int)i;
      gt_before += gt;
      eq_before += eq;
    }
  }
  // CTA 0 only gets here after every CTA arrived at the last step, i.e. after every CTA read the parity
  if (blockIdx.x == 0 && tid == 0) asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(scratch), "l"(par ^ 1ull) : "memory");
}


// ---------------------------------------------------------------------------------- occupancy loss
// AdaptiveSparseHead.occ_loss (AdaptiveSparseHead.py:100-103): nn.BCELoss()(p, t).mean() * 0.5 with torch's clamp of
// the logs at -100, as ONE CTA (N = 28 800 at the ScanNet shape): fixed summation order, no atomics.
__global__ void __launch_bounds__(1024) occ_loss_fwd_kernel(const float* __restrict__ p, const float* __restrict__ t, int N,
                                                            float* __restrict__ loss) {
  __shared__ float s_part[32];
  pdl_sync();
  // eight elements per thread and trip: the loads of a trip are issued together (the kernel sits between the forward and the
  // backward on the step's critical path and is pure latency: 28 dependent trips took 10.7 us), summed in the fixed order
  float a = 0.f;
  for (int base = threadIdx.x; base < N; base += 8 * 1024) {
    float pv[8], tv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + u * 1024;
      pv[u] = i < N ? __ldg(p + i) : 1.f;   // p = t = 1 contributes exactly 0
      tv[u] = i < N ? __ldg(t + i) : 1.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float l1 = fmaxf(logf(pv[u]), -100.f), l0 = fmaxf(log1pf(-pv[u]), -100.f);
      a -= tv[u] * l1 + (1.f - tv[u]) * l0;
    }
  }
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x < 32) {
    float b = warp_sum(s_part[threadIdx.x]);
    if (threadIdx.x == 0) loss[0] = b / (float)N * 0.5f;
  }
}

// grad_p[i] = g * 0.5 / N * (p - t) / max(p (1 - p), 1e-12)   (ATen binary_cross_entropy_backward)
__global__ void __launch_bounds__(256) occ_loss_bwd_kernel(const float* __restrict__ p, const float* __restrict__ t,
                                                           const float* __restrict__ g, int N, float* __restrict__ grad_p) {
  pdl_sync();
  const float gs = __ldg(g) * 0.5f / (float)N;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const float pi = __ldg(p + i), ti = __ldg(t + i);
    grad_p[i] = gs * (pi - ti) / fmaxf((1.f - pi) * pi, 1e-12f);
  }
}

}  // namespace sgc

// Programmatic dependent launch for the chain kernels (see common.cuh); off by default.
static int g_sgc_pdl = 0;
namespace sgc {
int pdl_enabled() { return g_sgc_pdl; }
}  // namespace sgc
extern "C" int sgc_set_pdl(int on) {
  g_sgc_pdl = on ? 1 : 0;
  return 0;
}

extern "C" int sgc_upsample2x_occ_fwd(const float* vol_in, int X, int Y, int Z, int C, const float* w_occ,
                                      const float* b_occ, float* vol_out, float* occ, void* stream) {
  if (C != 256 && C != 128 && C != 64) return (int)cudaErrorInvalidValue;
  const int n_out = 8 * X * Y * Z;
  const int grid = (n_out + sgc::kUpWarps - 1) / sgc::kUpWarps;
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 256) sgc::launch_chain(sgc::upsample_occ_fwd_kernel<8>, dim3(grid), dim3(sgc::kUpWarps * 32), 0, st, vol_in, X, Y, Z, w_occ, b_occ, vol_out, occ);
  else if (C == 128) sgc::launch_chain(sgc::upsample_occ_fwd_kernel<4>, dim3(grid), dim3(sgc::kUpWarps * 32), 0, st, vol_in, X, Y, Z, w_occ, b_occ, vol_out, occ);
  else return (int)cudaErrorInvalidValue;
  SGC_CUDA_CHECK_LAST();
  return 0;
}

// grad_up [2X,2Y,2Z,C], grad_occ [8XYZ] (may be null), occ = forward output.
// Outputs: grad_in [X,Y,Z,C] (written), grad_w [C] and grad_b [1] (accumulated; caller zeroes), gpre scratch [8XYZ].
extern "C" long long sgc_upsample2x_occ_bwd_scratch_floats(int X, int Y, int Z, int C) {
  return 6ll * X * Y * Z * C;   // [2X,2Y,Z,C] after the z pass + [2X,Y,Z,C] after the y pass
}

extern "C" int sgc_upsample2x_occ_bwd(const float* vol_in, int X, int Y, int Z, int C, const float* w_occ,
                                      const float* occ, const float* grad_up, const float* grad_occ, float* gpre,
                                      float* grad_in, float* grad_w, float* grad_b, float* scratch, void* stream) {
  if (C != 256 && C != 128) return (int)cudaErrorInvalidValue;
  const int n_out = 8 * X * Y * Z, n_in = X * Y * Z;
  cudaStream_t st = (cudaStream_t)stream;
  const float* gp = nullptr;
  if (grad_occ) {
    sgc::launch_chain(sgc::occ_gpre_kernel, dim3((n_out + 1023) / 1024 < 148 ? (n_out + 1023) / 1024 : 148), dim3(1024), 0, st, occ, grad_occ, n_out, gpre, grad_b);
    SGC_CUDA_CHECK_LAST();
    if (grad_w) {  // NULL: the caller issues sgc_upsample2x_occ_gradw itself (e.g. on its weight-gradient stream)
      const int g2 = 148 * 2;
      if (C == 256) sgc::occ_gradw_kernel<8><<<g2, sgc::kUpWarps * 32, 0, st>>>(vol_in, X, Y, Z, gpre, grad_w);
      else sgc::occ_gradw_kernel<4><<<g2, sgc::kUpWarps * 32, 0, st>>>(vol_in, X, Y, Z, gpre, grad_w);
      SGC_CUDA_CHECK_LAST();
    }
    gp = gpre;
  }
  if (scratch) {
    // separable evaluation: z, y, x (see up_bwd_axis_kernel)
    float* t1 = scratch;
    float* t2 = scratch + 4ll * X * Y * Z * C;
    const int C4 = C / 4;
    auto pass = [&](const float* src, const float* gpre_, int A, int N, int E4, float* dst) {
      const long long total = (long long)A * N * E4;
      long long blocks = (total + 255) / 256;
      if (blocks > 148 * 16) blocks = 148 * 16;
      sgc::launch_chain(sgc::up_bwd_axis_kernel, dim3((unsigned)blocks), dim3(256), 0, st, src, gpre_, w_occ, A, N, E4, dst);
    };
    pass(grad_up, gp, 4 * X * Y, Z, C4, t1);
    SGC_CUDA_CHECK_LAST();
    pass(t1, nullptr, 2 * X, Y, Z * C4, t2);
    SGC_CUDA_CHECK_LAST();
    pass(t2, nullptr, 1, X, Y * Z * C4, grad_in);
    SGC_CUDA_CHECK_LAST();
    return 0;
  }
  const int grid = (n_in + sgc::kUpWarps - 1) / sgc::kUpWarps;
  if (C == 256) sgc::launch_chain(sgc::upsample_occ_bwd_kernel<8>, dim3(grid), dim3(sgc::kUpWarps * 32), 0, st, grad_up, gp, w_occ, X, Y, Z, grad_in);
  else sgc::launch_chain(sgc::upsample_occ_bwd_kernel<4>, dim3(grid), dim3(sgc::kUpWarps * 32), 0, st, grad_up, gp, w_occ, X, Y, Z, grad_in);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

// grad_w[c] += sum_o gpre[o] * up[o,c] (the occupancy Linear's weight gradient), up recomputed from vol_in.
extern "C" int sgc_upsample2x_occ_gradw(const float* vol_in, int X, int Y, int Z, int C, const float* gpre, float* grad_w,
                                        void* stream) {
  if (C != 256 && C != 128) return (int)cudaErrorInvalidValue;
  const int g2 = 148 * 2;
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 256) sgc::occ_gradw_kernel<8><<<g2, sgc::kUpWarps * 32, 0, st>>>(vol_in, X, Y, Z, gpre, grad_w);
  else sgc::occ_gradw_kernel<4><<<g2, sgc::kUpWarps * 32, 0, st>>>(vol_in, X, Y, Z, gpre, grad_w);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

extern "C" int sgc_scatter_add_rows(float* vol, const int* sel, const float* y, int k, int C, void* stream) {
  if (C & 3) return (int)cudaErrorInvalidValue;
  const int total = k * (C / 4);
  if (total == 0) return 0;
  const int grid = (total + 255) / 256;
  sgc::launch_chain(sgc::scatter_add_rows_kernel, dim3(grid < 148 * 8 ? grid : 148 * 8), dim3(256), 0, (cudaStream_t)stream, vol, sel, y, k, C / 4);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

extern "C" int sgc_gather_rows(const float* vol, const int* sel, float* y, int k, int C, void* stream) {
  if (C & 3) return (int)cudaErrorInvalidValue;
  const int total = k * (C / 4);
  if (total == 0) return 0;
  const int grid = (total + 255) / 256;
  sgc::launch_chain(sgc::gather_rows_kernel, dim3(grid < 148 * 8 ? grid : 148 * 8), dim3(256), 0, (cudaStream_t)stream, vol, sel, y, k, C / 4);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

extern "C" int sgc_topk_select(const float* occ, int N, int k, int* sel, uint8_t* mask, void* stream) {
  if (k < 0 || k > N) return (int)cudaErrorInvalidValue;
  if (k > 0 && N <= 1024 * sgc::kTopkKpt)
    sgc::launch_chain(sgc::topk_select_small_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, occ, N, k, sel, mask);
  else
    sgc::topk_select_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(occ, N, k, sel, mask);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

extern "C" int sgc_topk_scratch_ints(int N) {
  if (N <= 0) return 0;
  const int G = (N + sgc::kMcChunk - 1) / sgc::kMcChunk;
  return sgc::kMcCounts + 2 * G;
}

// The same selection spread over many CTAs (meant for N > 32 768; any N works).  scratch: sgc_topk_scratch_ints(N) ints,
// private to the call (concurrent calls on different streams need their own).
extern "C" int sgc_topk_select_mc(const float* occ, int N, int k, int* sel, uint8_t* mask, int* scratch, void* stream) {
  if (k < 0 || k > N || N <= 0 || !scratch) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  const int G = (N + sgc::kMcChunk - 1) / sgc::kMcChunk;
  sgc::topk_mc_init_kernel<<<1, 1024, 0, st>>>(scratch, k);
  SGC_CUDA_CHECK_LAST();
  const int shifts[3] = {21, 10, 0}, bits[3] = {11, 11, 10};
  for (int pass = 0; pass < 3; ++pass) {
    sgc::topk_mc_hist_kernel<<<G, 1024, 0, st>>>(occ, N, scratch, shifts[pass], 1 << bits[pass]);
    SGC_CUDA_CHECK_LAST();
    sgc::topk_mc_pick_kernel<<<1, 1024, 0, st>>>(scratch, shifts[pass], 1 << bits[pass]);
    SGC_CUDA_CHECK_LAST();
  }
  sgc::topk_mc_count_kernel<<<G, 1024, 0, st>>>(occ, N, scratch);
  SGC_CUDA_CHECK_LAST();
  sgc::topk_mc_write_kernel<<<G, 1024, 0, st>>>(occ, N, scratch, sel, mask);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

// One launch for every level (see topk_select_grid_kernel).  scratch: sgc_topk_grid_scratch_bytes() bytes, zero-filled
// ONCE when allocated and then reused by every call on the same stream (the kernel keeps it consistent itself).
extern "C" int sgc_topk_grid_scratch_bytes() { return 8 * (1 + 2 * sgc::kTkHalf); }
extern "C" int sgc_topk_grid_max_n() { return sgc::kTkMaxG * sgc::kTkThreads * sgc::kTkKpt; }

extern "C" int sgc_topk_select_grid(const float* occ, int N, int k, int* sel, uint8_t* mask, void* scratch, void* stream) {
  if (k <= 0 || k > N || N <= 0 || !scratch || N > sgc_topk_grid_max_n()) return (int)cudaErrorInvalidValue;
  const int G = (N + sgc::kTkThreads * sgc::kTkKpt - 1) / (sgc::kTkThreads * sgc::kTkKpt);
  const int per = (N + G * sgc::kTkThreads - 1) / (G * sgc::kTkThreads);
  sgc::launch_chain(sgc::topk_select_grid_kernel, dim3(G), dim3(sgc::kTkThreads), 0, (cudaStream_t)stream, occ, N, k, per, sel, mask,
                    reinterpret_cast<unsigned long long*>(scratch));
  SGC_CUDA_CHECK_LAST();
  return 0;
}

// occ_loss (ASH:100-103): loss[0] = 0.5 * mean(BCE(p, t)); bwd: grad_p = g[0] * d loss / d p  (g: 1 float on the device).
extern "C" int sgc_occ_loss_fwd(const float* p, const float* t, int N, float* loss, void* stream) {
  if (N <= 0) return (int)cudaErrorInvalidValue;
  sgc::launch_chain(sgc::occ_loss_fwd_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, p, t, N, loss);
  SGC_CUDA_CHECK_LAST();
  return 0;
}
extern "C" int sgc_occ_loss_bwd(const float* p, const float* t, const float* g, int N, float* grad_p, void* stream) {
  if (N <= 0) return (int)cudaErrorInvalidValue;
  const int grid = (N + 255) / 256;
  sgc::launch_chain(sgc::occ_loss_bwd_kernel, dim3(grid < 296 ? grid : 296), dim3(256), 0, (cudaStream_t)stream, p, t, g, N, grad_p);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

// ---------------------------------------------------------------------------------- valid-mask pyramid
// The detection head derives per-level validity masks from the view transform's `valid` volume with
//   nn.Upsample(size=level_shape, mode='trilinear')(valid).round().bool()      (dense_heads/imvoxel_head_v2.py:121-123,256-258)
// for levels of 1, 1/2 and 1/4 of the volume's size.  With align_corners=False a x1/2 output voxel interpolates the two
// source voxels 2d, 2d+1 with weights 1/2 per axis (the mean of a 2x2x2 block), a x1/4 voxel the two central voxels 4d+1,
// 4d+2 of its 4x4x4 block; the mean of eight 0/1 values rounds (half to even: 0.5 -> 0) to 1 iff at least five are set.
// One launch, bit-exact, no float volume in between.
namespace sgc {
__global__ void valid_pyramid_kernel(const long long* __restrict__ valid, int X, int Y, int Z, unsigned char* __restrict__ v0,
                                     unsigned char* __restrict__ v1, unsigned char* __restrict__ v2) {
  const int n0 = X * Y * Z, X1 = X >> 1, Y1 = Y >> 1, Z1 = Z >> 1, X2 = X >> 2, Y2 = Y >> 2, Z2 = Z >> 2;
  const int n1 = v1 ? X1 * Y1 * Z1 : 0, n2 = v2 ? X2 * Y2 * Z2 : 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n0 + n1 + n2; i += gridDim.x * blockDim.x) {
    if (i < n0) {
      if (v0) v0[i] = valid[i] != 0;
      continue;
    }
    const bool half = i < n0 + n1;
    const int j = half ? i - n0 : i - n0 - n1;
    const int Yl = half ? Y1 : Y2, Zl = half ? Z1 : Z2;
    const int z = j % Zl, y = (j / Zl) % Yl, x = j / (Zl * Yl);
    const int sx = half ? 2 * x : 4 * x + 1, sy = half ? 2 * y : 4 * y + 1, sz = half ? 2 * z : 4 * z + 1;
    int cnt = 0;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) cnt += valid[((size_t)(sx + a) * Y + (sy + b)) * Z + (sz + c)] != 0;
    (half ? v1 : v2)[j] = cnt >= 5;
  }
}
}  // namespace sgc

// valid [X,Y,Z] int64 (AdaptiveSparseHead's return value); v0 [X,Y,Z], v1 [X/2,Y/2,Z/2], v2 [X/4,Y/4,Z/4] uint8 0/1 (any may be
// NULL).  X, Y, Z multiples of 4 (2 when v2 is NULL).
extern "C" int sgc_valid_pyramid(const long long* valid, int X, int Y, int Z, unsigned char* v0, unsigned char* v1,
                                 unsigned char* v2, void* stream) {
  if (!valid || X <= 0 || Y <= 0 || Z <= 0) return (int)cudaErrorInvalidValue;
  const int m = v2 ? 3 : (v1 ? 1 : 0);
  if ((X & m) || (Y & m) || (Z & m)) return (int)cudaErrorInvalidValue;
  const int total = X * Y * Z + (v1 ? (X >> 1) * (Y >> 1) * (Z >> 1) : 0) + (v2 ? (X >> 2) * (Y >> 2) * (Z >> 2) : 0);
  const int grid = (total + 255) / 256;
  sgc::valid_pyramid_kernel<<<grid < 148 * 8 ? grid : 148 * 8, 256, 0, (cudaStream_t)stream>>>(valid, X, Y, Z, v0, v1, v2);
  SGC_CUDA_CHECK_LAST();
  return 0;
}
