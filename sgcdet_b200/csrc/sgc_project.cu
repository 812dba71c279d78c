// sgc_project_compact: per-view voxel-centre projection + visibility mask + view-major pair list.
//
// Replaces VoxFormerEncoder_DFA3D.point_sampling (transformer_utils/encoder.py:179-223) and the V
// host-synchronising nonzero()/rebatch loops of DeformCrossAttention_DFA3D.forward
// (deformable_cross_attention.py:758-773) with two device passes and no host sync.
//
// Arithmetic contract (bit-exact against oracle/path_ref.py:point_sampling): every operation is a
// separately rounded fp32 operation, no FMA:
//   p = ref + origin;  x = ((P0*px + P1*py) + P2*pz) + P3;  u = (x / max(z,eps)) / img_w;  v likewise / img_h;
//   d = (z - dbound0) / dscale;   mask = z>eps & eps<u<1-eps & eps<v<1-eps          (encoder.py:203-219)
//
// Outputs:
//   ref_cam     [V,Q,3]  (u,v,d)
//   mask        [V,Q]    uint8
//   pair_index  [V,Q]    int32, position of the pair in the view-major list, -1 when invisible
//   pair_vq     [cap]    int32, v*Q+q of pair i (ascending q inside a view)
//   view_offsets[V+1]    int32, view v owns pairs [view_offsets[v], view_offsets[v+1]);  [V] = n_pairs
//   count       [Q]      int32, number of views that see voxel q                     (DCA:819-820)
#include "common.cuh"

namespace sgc {

constexpr int kTile = 1024;  // voxels per CTA tile (== blockDim)

__device__ __forceinline__ bool project_one(const float* __restrict__ P, float px, float py, float pz, float eps,
                                            float hi, float img_w, float img_h, float db0, float dscale, float& u,
                                            float& v, float& d) {
  float r[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float a = __fmul_rn(P[i * 4 + 0], px);
    a = __fadd_rn(a, __fmul_rn(P[i * 4 + 1], py));
    a = __fadd_rn(a, __fmul_rn(P[i * 4 + 2], pz));
    r[i] = __fadd_rn(a, P[i * 4 + 3]);
  }
  const float z = r[2];
  const float zc = fmaxf(z, eps);
  u = __fdiv_rn(__fdiv_rn(r[0], zc), img_w);
  v = __fdiv_rn(__fdiv_rn(r[1], zc), img_h);
  d = __fdiv_rn(__fsub_rn(z, db0), dscale);
  return (z > eps) && (u > eps) && (u < hi) && (v > eps) && (v < hi);
}

// pass 1: project, write ref_cam/mask, count visible voxels per (view, tile); zero count[]
__global__ void __launch_bounds__(kTile) project_kernel(const float* __restrict__ proj, const float* __restrict__ ref3d,
                                                       const int* __restrict__ sel, int Q, float ox, float oy, float oz,
                                                       float eps, float hi, float img_w, float img_h, float db0, float dscale,
                                                       float* __restrict__ ref_cam, uint8_t* __restrict__ mask,
                                                       int* __restrict__ tile_counts, int* __restrict__ count) {
  __shared__ float P[12];
  pdl_sync();
  const int v = blockIdx.y, tile = blockIdx.x;
  if (threadIdx.x < 12) P[threadIdx.x] = proj[v * 12 + threadIdx.x];
  __syncthreads();
  const int q = tile * kTile + threadIdx.x;
  bool vis = false;
  if (q < Q) {
    const int n = sel ? sel[q] : q;
    const float px = __fadd_rn(ref3d[n * 3 + 0], ox), py = __fadd_rn(ref3d[n * 3 + 1], oy),
                pz = __fadd_rn(ref3d[n * 3 + 2], oz);
    float u, w, d;
    vis = project_one(P, px, py, pz, eps, hi, img_w, img_h, db0, dscale, u, w, d);
    const size_t o = ((size_t)v * Q + q);
    ref_cam[o * 3 + 0] = u; ref_cam[o * 3 + 1] = w; ref_cam[o * 3 + 2] = d;
    mask[o] = vis ? 1 : 0;
    if (v == 0) count[q] = 0;
  }
  const int c = __syncthreads_count(vis);
  if (threadIdx.x == 0) tile_counts[v * gridDim.x + tile] = c;
}

// pass 2: every CTA sums the counts of the (view-major) tiles in front of its own -- at most a few thousand integers, one
// block reduction, instead of a scan launch of its own between the two passes (the three launches sat on the critical
// path of every level's chain) --, ranks its voxels inside the tile and emits the pair ids
__global__ void __launch_bounds__(kTile) emit_kernel(const uint8_t* __restrict__ mask, const int* __restrict__ tile_counts,
                                                    int Q, int* __restrict__ pair_index, int* __restrict__ pair_vq,
                                                    int* __restrict__ count, int* __restrict__ view_offsets) {
  __shared__ int warp_tot[32];
  __shared__ int warp_pre[32];
  pdl_sync();
  const int v = blockIdx.y, tile = blockIdx.x;
  const int q = tile * kTile + threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int me = v * gridDim.x + tile, n_tiles = gridDim.x * gridDim.y;
  int before = 0;
  for (int i = threadIdx.x; i < me; i += kTile) before += __ldg(tile_counts + i);
  before = warp_sum_int(before);
  const bool vis = (q < Q) && mask[(size_t)v * Q + q];
  const unsigned b = __ballot_sync(SGC_FULL_MASK, vis);
  if (lane == 0) { warp_tot[wid] = __popc(b); warp_pre[wid] = before; }
  __syncthreads();
  if (wid == 0) {
    int t = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(SGC_FULL_MASK, t, o);
      if (lane >= o) t += y;
    }
    warp_tot[lane] = t;
    const int pre = warp_sum_int(warp_pre[lane]);
    if (lane == 0) warp_pre[0] = pre;
  }
  __syncthreads();
  const int tile_offset = warp_pre[0];
  if (threadIdx.x == 0) {
    if (tile == 0) view_offsets[v] = tile_offset;
    if (me == n_tiles - 1) view_offsets[gridDim.y] = tile_offset + warp_tot[31];
  }
  if (q < Q) {
    int id = -1;
    if (vis) {
      id = tile_offset + (wid ? warp_tot[wid - 1] : 0) + __popc(b & ((1u << lane) - 1));
      pair_vq[id] = v * Q + q;
      atomicAdd(count + q, 1);
    }
    pair_index[(size_t)v * Q + q] = id;
  }
}

}  // namespace sgc

extern "C" int sgc_project_scratch_ints(int V, int Q) {
  const int tiles = (Q + sgc::kTile - 1) / sgc::kTile;
  return V * tiles;
}

extern "C" int sgc_project_compact(const float* proj, const float* ref3d, const int* sel, int V, int Q, float ox,
                                   float oy, float oz, float eps, float one_minus_eps, float img_w, float img_h,
                                   float dbound0, float dscale, float* ref_cam, uint8_t* mask, int* pair_index, int* pair_vq, int* view_offsets,
                                   int* count, int* scratch, void* stream) {
  if (V <= 0 || Q <= 0) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles = (Q + sgc::kTile - 1) / sgc::kTile;
  int* tile_counts = scratch;
  dim3 grid(tiles, V);
  sgc::launch_chain(sgc::project_kernel, grid, dim3(sgc::kTile), 0, st, proj, ref3d, sel, Q, ox, oy, oz, eps, one_minus_eps, img_w,
                    img_h, dbound0, dscale, ref_cam, mask, tile_counts, count);
  SGC_CUDA_CHECK_LAST();
  sgc::launch_chain(sgc::emit_kernel, grid, dim3(sgc::kTile), 0, st, (const uint8_t*)mask, (const int*)tile_counts, Q, pair_index,
                    pair_vq, count, view_offsets);
  SGC_CUDA_CHECK_LAST();
  return 0;
}
