// fp32 -> (bf16 hi, bf16 lo) operand split for the dense projections.
//
// The value / offset / attention-weight projections (deformable_cross_attention.py:417-436) are fp32 GEMMs in
// the reference; to run them on the tensor cores at fp32-level accuracy each operand is split as
// x = hi + lo (+ O(2^-17 |x|)), hi = bf16(x), lo = bf16(x - hi), and the product is evaluated as
// hi*hi' + lo*hi' + hi*lo' with fp32 accumulation (relative error ~1e-5, inside the rtol 1e-3 budget).
// The three terms are folded into ONE GEMM by concatenating along K:  [hi | lo | hi] x [hi' ; hi' ; lo'].
//
// This kernel writes the [hi | lo | hi] operand in one pass over x:
//   rows r of length `cols` (source row stride `src_stride`), grouped by `rpg` rows:
//   out[((r / rpg) * 3 + slot) * rpg + (r % rpg)][col],  slots = (hi, lo, hi) for pattern 0 and (hi, hi, lo)
//   for pattern 1 (the two operands of one product use opposite patterns).
//   rpg = C  -> feature maps  [V,C,S]   -> [V,3,C,S]     (K = channel concat, S stays contiguous)
//   rpg = 1  -> row-major     [R,N]     -> [R,3N]        (K = column concat)
#include <cuda_bf16.h>
#include "common.cuh"

namespace sgc {

__device__ __forceinline__ void split1(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

template <bool VEC>
__global__ void __launch_bounds__(256) split_bf16x3_kernel(const float* __restrict__ x, long long rows, int cols,
                                                          long long src_stride, int rpg, int pattern,
                                                          __nv_bfloat16* __restrict__ out) {
  const int cpr = VEC ? cols / 4 : cols;  // work items per row
  const long long total = rows * cpr;
  const long long slot_stride = (long long)rpg * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cpr;
    const int c = (int)(i - r * cpr) * (VEC ? 4 : 1);
    const long long base = ((r / rpg) * 3 * rpg + (r % rpg)) * (long long)cols + c;
    if (VEC) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * src_stride + c));
      __nv_bfloat16 h[4], l[4];
      split1(v.x, h[0], l[0]); split1(v.y, h[1], l[1]); split1(v.z, h[2], l[2]); split1(v.w, h[3], l[3]);
      const uint2 hv = *reinterpret_cast<uint2*>(h), lv = *reinterpret_cast<uint2*>(l);
      *reinterpret_cast<uint2*>(out + base) = hv;
      *reinterpret_cast<uint2*>(out + base + slot_stride) = pattern ? hv : lv;
      *reinterpret_cast<uint2*>(out + base + 2 * slot_stride) = pattern ? lv : hv;
    } else {
      __nv_bfloat16 h, l;
      split1(__ldg(x + r * src_stride + c), h, l);
      out[base] = h; out[base + slot_stride] = pattern ? h : l; out[base + 2 * slot_stride] = pattern ? l : h;
    }
  }
}

}  // namespace sgc

extern "C" int sgc_split_bf16x3(const float* x, long long rows, int cols, long long src_stride, int rows_per_group,
                                int pattern, void* out, void* stream) {
  if (rows <= 0 || cols <= 0) return 0;
  if (rows_per_group <= 0 || rows % rows_per_group) return (int)cudaErrorInvalidValue;
  const bool vec = (cols % 4 == 0) && (src_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out) & 7) == 0);
  const long long items = rows * (vec ? cols / 4 : cols);
  long long blocks = (items + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  if (vec)
    sgc::split_bf16x3_kernel<true><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, cols, src_stride, rows_per_group,
                                                                                 pattern, (__nv_bfloat16*)out);
  else
    sgc::split_bf16x3_kernel<false><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, cols, src_stride, rows_per_group,
                                                                                  pattern, (__nv_bfloat16*)out);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// Column sums of a row-major [R,C] matrix (bias gradients of the Linear layers), deterministic:
// every CTA reduces a slab of rows into partial[b][:], the last CTA to finish (atomic ticket) adds the
// partials in slab order.  `counter` must be zero on entry and is reset to zero on exit.
namespace sgc {
constexpr int kColsumRows = 32;

// block = 64 float4-column lanes x 4 row lanes; requires C % 4 == 0
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, int R, int C,
                                                    float* __restrict__ partial, unsigned int* __restrict__ counter,
                                                    float* __restrict__ out) {
  __shared__ float4 s_acc[4][64];
  __shared__ bool is_last;
  const int cx = threadIdx.x & 63, ry = threadIdx.x >> 6;
  const int r0 = blockIdx.x * kColsumRows;
  const int r1 = min(R, r0 + kColsumRows);
  const int C4 = C >> 2;
  for (int c0 = 0; c0 < C4; c0 += 64) {
    const int c = c0 + cx;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C4) {
      for (int r = r0 + ry; r < r1; r += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + (size_t)r * C) + c);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
    }
    s_acc[ry][cx] = a;
    __syncthreads();
    if (ry == 0 && c < C4) {
      float4 t = s_acc[0][cx];
#pragma unroll
      for (int k = 1; k < 4; ++k) { t.x += s_acc[k][cx].x; t.y += s_acc[k][cx].y; t.z += s_acc[k][cx].z; t.w += s_acc[k][cx].w; }
      reinterpret_cast<float4*>(partial + (size_t)blockIdx.x * C)[c] = t;
    }
    __syncthreads();
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f;
    for (int b = 0; b < (int)gridDim.x; ++b) a += partial[(size_t)b * C + c];
    out[c] = a;
  }
  if (threadIdx.x == 0) *counter = 0u;
}
}  // namespace sgc

extern "C" int sgc_colsum_scratch_floats(int R, int C) {
  return ((R + sgc::kColsumRows - 1) / sgc::kColsumRows) * C;
}

extern "C" int sgc_colsum(const float* x, int R, int C, float* out, float* scratch, unsigned int* counter,
                          void* stream) {
  if (R <= 0 || C <= 0 || (C & 3)) return (int)cudaErrorInvalidValue;
  const int blocks = (R + sgc::kColsumRows - 1) / sgc::kColsumRows;
  sgc::colsum_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, R, C, scratch, counter, out);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// Fused: rows-split of g [R,C] (the three bf16 slots stacked along the reduction axis, [3R,C], slots per `pattern`)
// AND the column sums of g (bias gradient) in one pass over g.  Same two-stage deterministic reduction as
// sgc_colsum.  Used for every (weight, bias) gradient pair of the Linear layers: gW = g^T x needs the rows-split
// of g, gb = colsum(g).
namespace sgc {
__global__ void __launch_bounds__(256) split_rows_colsum_kernel(const float* __restrict__ x, int R, int C, int pattern,
                                                               __nv_bfloat16* __restrict__ out, float* __restrict__ partial,
                                                               unsigned int* __restrict__ counter, float* __restrict__ sums) {
  __shared__ float4 s_acc[4][64];
  __shared__ bool is_last;
  const int cx = threadIdx.x & 63, ry = threadIdx.x >> 6;
  const int r0 = blockIdx.x * kColsumRows;
  const int r1 = min(R, r0 + kColsumRows);
  const int C4 = C >> 2;
  const size_t slot = (size_t)R * C;
  for (int c0 = 0; c0 < C4; c0 += 64) {
    const int c = c0 + cx;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C4) {
      for (int r = r0 + ry; r < r1; r += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + (size_t)r * C) + c);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        __nv_bfloat16 h[4], l[4];
        split1(v.x, h[0], l[0]); split1(v.y, h[1], l[1]); split1(v.z, h[2], l[2]); split1(v.w, h[3], l[3]);
        const uint2 hv = *reinterpret_cast<uint2*>(h), lv = *reinterpret_cast<uint2*>(l);
        const size_t o = (size_t)r * C + 4 * c;
        *reinterpret_cast<uint2*>(out + o) = hv;
        *reinterpret_cast<uint2*>(out + o + slot) = pattern ? hv : lv;
        *reinterpret_cast<uint2*>(out + o + 2 * slot) = pattern ? lv : hv;
      }
    }
    s_acc[ry][cx] = a;
    __syncthreads();
    if (ry == 0 && c < C4) {
      float4 t = s_acc[0][cx];
#pragma unroll
      for (int k = 1; k < 4; ++k) { t.x += s_acc[k][cx].x; t.y += s_acc[k][cx].y; t.z += s_acc[k][cx].z; t.w += s_acc[k][cx].w; }
      reinterpret_cast<float4*>(partial + (size_t)blockIdx.x * C)[c] = t;
    }
    __syncthreads();
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f;
    for (int b = 0; b < (int)gridDim.x; ++b) a += partial[(size_t)b * C + c];
    sums[c] = a;
  }
  if (threadIdx.x == 0) *counter = 0u;
}
}  // namespace sgc

extern "C" int sgc_split_rows_colsum(const float* x, int R, int C, int pattern, void* out, float* sums, float* scratch,
                                     unsigned int* counter, void* stream) {
  if (R <= 0 || C <= 0 || (C & 3)) return (int)cudaErrorInvalidValue;
  const int blocks = (R + sgc::kColsumRows - 1) / sgc::kColsumRows;
  sgc::split_rows_colsum_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, R, C, pattern, (__nv_bfloat16*)out, scratch,
                                                                          counter, sums);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Folded projection weights of MSDeformableAttention3D_DFA3D (deformable_cross_attention.py:417-436): the level's dense
// projection is ONE GEMM with  Wcat = [value_proj.weight ; G rows],  G row 4*(m*P + p) + j = sampling_offsets row
// (m*P + p)*2 + j for j < 2, sampling_offsets_depth row m*P + p for j = 2, attention_weights row m*P + p for j = 3 (the
// lift kernels read one float4 (off_x, off_y, off_d, logit) per (head, point)), and likewise for the biases.
// fold: the four weights -> Wcat [C + 4MP, C], the three small biases -> gbias [4MP] (one launch instead of three cats);
// unfold: the gradient of Wcat / gbias -> the seven parameter gradients, each its own contiguous tensor (one launch
// instead of the cat's backward and eight strided copies at the very end of the step).
namespace sgc {

__device__ __forceinline__ void fold_row(int r, int C, const float* const (&w)[4], const float*& src) {
  if (r < C) { src = w[0] + (size_t)r * C; return; }
  const int g = r - C, mp = g >> 2, j = g & 3;
  src = j < 2 ? w[1] + (size_t)(2 * mp + j) * C : j == 2 ? w[2] + (size_t)mp * C : w[3] + (size_t)mp * C;
}

__global__ void __launch_bounds__(256) fold_wcat_kernel(const float* __restrict__ wv, const float* __restrict__ wo,
                                                        const float* __restrict__ wd, const float* __restrict__ wa,
                                                        const float* __restrict__ bo, const float* __restrict__ bd,
                                                        const float* __restrict__ ba, int C, int MP,
                                                        float* __restrict__ wcat, float* __restrict__ gbias) {
  const int N = C + 4 * MP, C4 = C / 4;
  const float* const w[4] = {wv, wo, wd, wa};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * C4; i += gridDim.x * blockDim.x) {
    const int r = i / C4, c = (i - r * C4) * 4;
    const float* src;
    fold_row(r, C, w, src);
    *reinterpret_cast<float4*>(wcat + (size_t)r * C + c) = __ldg(reinterpret_cast<const float4*>(src + c));
  }
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < 4 * MP; g += gridDim.x * blockDim.x) {
    const int mp = g >> 2, j = g & 3;
    gbias[g] = j < 2 ? __ldg(bo + 2 * mp + j) : j == 2 ? __ldg(bd + mp) : __ldg(ba + mp);
  }
}

__global__ void __launch_bounds__(256) unfold_wcat_grad_kernel(const float* __restrict__ gwcat, const float* __restrict__ ggbias,
                                                               int C, int MP, float* __restrict__ gwv, float* __restrict__ gwo,
                                                               float* __restrict__ gwd, float* __restrict__ gwa,
                                                               float* __restrict__ gbo, float* __restrict__ gbd,
                                                               float* __restrict__ gba) {
  const int N = C + 4 * MP, C4 = C / 4;
  float* const w[4] = {gwv, gwo, gwd, gwa};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * C4; i += gridDim.x * blockDim.x) {
    const int r = i / C4, c = (i - r * C4) * 4;
    float* dst;
    if (r < C) dst = w[0] + (size_t)r * C;
    else {
      const int g = r - C, mp = g >> 2, j = g & 3;
      dst = j < 2 ? w[1] + (size_t)(2 * mp + j) * C : j == 2 ? w[2] + (size_t)mp * C : w[3] + (size_t)mp * C;
    }
    *reinterpret_cast<float4*>(dst + c) = __ldg(reinterpret_cast<const float4*>(gwcat + (size_t)r * C + c));
  }
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < 4 * MP; g += gridDim.x * blockDim.x) {
    const int mp = g >> 2, j = g & 3;
    const float v = __ldg(ggbias + g);
    if (j < 2) gbo[2 * mp + j] = v; else if (j == 2) gbd[mp] = v; else gba[mp] = v;
  }
}

}  // namespace sgc

extern "C" int sgc_fold_wcat(const float* value_w, const float* off_w, const float* dep_w, const float* att_w,
                             const float* off_b, const float* dep_b, const float* att_b, int C, int MP, float* wcat,
                             float* gbias, void* stream) {
  if (C <= 0 || C % 4 || MP <= 0) return (int)cudaErrorInvalidValue;
  const int items = (C + 4 * MP) * (C / 4);
  sgc::fold_wcat_kernel<<<(items + 255) / 256, 256, 0, (cudaStream_t)stream>>>(value_w, off_w, dep_w, att_w, off_b, dep_b, att_b, C,
                                                                             MP, wcat, gbias);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

extern "C" int sgc_unfold_wcat_grad(const float* gwcat, const float* ggbias, int C, int MP, float* g_value_w, float* g_off_w,
                                    float* g_dep_w, float* g_att_w, float* g_off_b, float* g_dep_b, float* g_att_b, void* stream) {
  if (C <= 0 || C % 4 || MP <= 0) return (int)cudaErrorInvalidValue;
  const int items = (C + 4 * MP) * (C / 4);
  sgc::unfold_wcat_grad_kernel<<<(items + 255) / 256, 256, 0, (cudaStream_t)stream>>>(gwcat, ggbias, C, MP, g_value_w, g_off_w, g_dep_w,
                                                                                    g_att_w, g_off_b, g_dep_b, g_att_b);
  SGC_CUDA_CHECK_LAST();
  return 0;
}
