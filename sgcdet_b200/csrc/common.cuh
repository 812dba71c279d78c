// Shared device helpers for the sgcdet_b200 kernels (sm_100a).
//
// Sampling conventions follow the DFA3D operator this library replaces:
//   pixel coords  x_im = loc*size - 0.5                      (ms_depth_score_sample_cuda_kernel.cuh:133-135,
//                                                             wms_deform_attn_cuda_kernel.cuh:286-287)
//   range tests   -1 < h < H, -1 < w < W (, -1 < d < D)      (DSK:137, WMSK:289)
//   corner order  TL, TR, BR, BL in depth_score              (DSK:89-92; WMSK:51,58,65,72)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SGC_FULL_MASK 0xffffffffu

namespace sgc {

// One 3-D sampling point resolved against an H x W x D grid.
//   k = 0..3  <->  TL(h0,w0), TR(h0,w0+1), BR(h0+1,w0+1), BL(h0+1,w0)
struct Tap {
  int   pix[4];   // h*W + w of the corner, or -1 when the corner (or the whole sample) is out of range
  float bw[4];    // bilinear weight of the corner (hh*hw, hh*lw, lh*lw, lh*hw)
  float lh, lw;   // fractional parts
  float ld, hd;   // depth lerp weights (ld toward d0+1)
  int   d0;       // floor(d_im)
  bool  in2d, in3d;
};

__device__ __forceinline__ Tap make_tap(float x, float y, float z, int H, int W, int D) {
  Tap t;
  // product rounded first, then the subtract: no FMA contraction (matches the reference's float*int - 0.5)
  const float h = __fsub_rn(__fmul_rn(y, (float)H), 0.5f);
  const float w = __fsub_rn(__fmul_rn(x, (float)W), 0.5f);
  const float d = __fsub_rn(__fmul_rn(z, (float)D), 0.5f);
  t.in2d = (h > -1.f) && (w > -1.f) && (h < (float)H) && (w < (float)W);
  t.in3d = t.in2d && (d > -1.f) && (d < (float)D);
  const float hf = floorf(h), wf = floorf(w), df = floorf(d);
  const int h0 = (int)hf, w0 = (int)wf;
  t.d0 = (int)df;
  t.lh = h - hf; t.lw = w - wf; t.ld = d - df;
  t.hd = 1.f - t.ld;
  const float hh = 1.f - t.lh, hw = 1.f - t.lw;
  t.bw[0] = hh * hw; t.bw[1] = hh * t.lw; t.bw[2] = t.lh * t.lw; t.bw[3] = t.lh * hw;
  const bool top = h0 >= 0, bot = h0 + 1 <= H - 1, lef = w0 >= 0, rig = w0 + 1 <= W - 1;
  t.pix[0] = (t.in2d && top && lef) ? h0 * W + w0 : -1;
  t.pix[1] = (t.in2d && top && rig) ? h0 * W + w0 + 1 : -1;
  t.pix[2] = (t.in2d && bot && rig) ? (h0 + 1) * W + w0 + 1 : -1;
  t.pix[3] = (t.in2d && bot && lef) ? (h0 + 1) * W + w0 : -1;
  return t;
}

// Depth score of corner k: lerp of the depth distribution along d, zero outside [0, D-1] (DSK:53-92).
// dist_px points at the D-vector of that pixel.  v_lo / v_hi are returned for the backward.
__device__ __forceinline__ float depth_score(const Tap& t, const float* __restrict__ dist_px, int D,
                                             float& v_lo, float& v_hi) {
  v_lo = (t.d0 >= 0) ? __ldg(dist_px + t.d0) : 0.f;
  v_hi = (t.d0 + 1 <= D - 1) ? __ldg(dist_px + t.d0 + 1) : 0.f;
  return v_lo * t.hd + v_hi * t.ld;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Vector reduction into global memory (sm_90+: RED.E.ADD.F32x4).  No "memory" clobber on purpose: the gradient buffers
// these accumulate into are never read by the issuing kernel, and the clobber would pin every later gather load behind
// the reduction (the kernels that use them are latency-bound gathers).
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d));
}
__device__ __forceinline__ void red_add1(float* p, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SGC_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_int(int v) { return __reduce_add_sync(SGC_FULL_MASK, v); }
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(SGC_FULL_MASK, v, o));
  return v;
}
// sum over the 4 lanes that share (lane >> 2)
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(SGC_FULL_MASK, v, 1);
  v += __shfl_xor_sync(SGC_FULL_MASK, v, 2);
  return v;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(SGC_FULL_MASK, v, 1));
  v = fmaxf(v, __shfl_xor_sync(SGC_FULL_MASK, v, 2));
  return v;
}

}  // namespace sgc

// ---------------------------------------------------------------------------------------------------------------------
// Programmatic dependent launch (sm_90+) for the latency-bound per-voxel chain: a kernel launched through
// sgc::launch_chain() with the feature enabled (sgc_set_pdl) may start while its predecessor in the stream is still
// running.  Device side, EVERY thread of such a kernel calls pdl_sync() before its first global read or write of data
// another kernel touches: `launch_dependents` lets the successor's CTAs be scheduled (they then sit in their own wait),
// `wait` blocks until the predecessor grid has completed and its memory is visible.  Without the launch attribute both
// instructions are no-ops, so the same kernels run unchanged on ordinary launches.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() {
  pdl_launch_dependents();
  pdl_wait();
}

namespace sgc {
int pdl_enabled();   // csrc/sgc_volume.cu (sgc_set_pdl)

// <<<grid, block, smem, stream>>> with the programmatic-stream-serialization attribute when enabled.  Only for kernels
// that call pdl_sync() / pdl_wait() as described above.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
}  // namespace sgc

#define SGC_CUDA_CHECK_LAST() \
  do {                        \
    cudaError_t e__ = cudaGetLastError(); \
    if (e__ != cudaSuccess) return (int)e__; \
  } while (0)
