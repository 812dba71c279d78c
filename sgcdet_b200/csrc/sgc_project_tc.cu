// sgc_project_tc_fwd: the dense feature projection of one level on the 5th-gen tensor cores (tcgen05 / TMEM),
// fused with the NCHW -> channel-last layout change and the fp32 -> bf16 hi/lo operand split.
//
//   vg[v, s, n] = sum_c feat[v, c, s] * Wcat[n, c]        (value_proj + the folded offset / depth-offset /
//                                                          attention-weight rows, deformable_cross_attention.py:417-436;
//                                                          replaces the flatten/permute copy of transformer.py:151-170)
//
// fp32-level accuracy from bf16 tensor cores: x = hi + lo, x*y ~= hi*hi' + lo*hi' + hi*lo' (fp32 accumulate in TMEM,
// relative error ~1e-5).  Unlike the library path (sgc_split_bf16x3 + bf16 GEMM) the split never touches HBM:
// the fp32 map is read once, converted in shared memory, and the fp32 result is written once.
//
// One persistent CTA per SM, 12 warps, warp-specialised; tile = 128 pixels x all N output channels:
//   warp  5    F producer  : one thread issues TMA tile loads (cp.async.bulk.tensor.2d) of feat[c0..c0+31][s0..s0+127] fp32
//   warps 0-3  A converters: fp32 staging tile -> bf16 hi/lo -> K-major core-matrix tiles (generic proxy + proxy fence)
//   warp  4    B producer  : one thread issues cp.async.bulk (TMA bulk copy) of pre-packed weight slabs (L2 resident)
//   warp  6    MMA issuer  : one thread issues tcgen05.mma (M=128, N=N/parts, K=16), accumulators in TMEM
//   warp  7    TMEM alloc / dealloc
//   warps 8-11 epilogue    : tcgen05.ld 32x32b.x32 -> registers -> 128B-swizzled smem tile -> TMA tensor store
//                            (cp.async.bulk.tensor.3d) into the channel-last vg; rows beyond S are clipped by the map
// Pipelines: A stages (a_full/a_empty), B stages (b_full/b_empty), one TMEM accumulator (tmem_full/tmem_empty),
// all mbarrier based; tcgen05.commit releases the shared-memory stages.
//
// Shared-memory operand layout: canonical UMMA K-major, no swizzle: 8x8 bf16 core matrices (128 contiguous bytes),
// LBO (next core matrix along K) = 128 B, SBO (next 8 rows) = (BK/8)*128 B.  The weight slabs are pre-packed into
// exactly this image by sgc_pack_weight_tc so the B producer is a plain bulk copy.
#include <cuda.h>
#include <cstdlib>
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/sgcdet_b200.h"

namespace sgc {
namespace tc {

constexpr int NA = 2;        // A stages (hi+lo)
constexpr int NB = 4;        // B stages (one of hi / lo per stage)
constexpr int NF = 3;        // fp32 staging stages filled by TMA (BK x BM floats = 16 KB each), forward
constexpr int NF_BWD = 5;    // data gradient: no epilogue staging buffers, so more loads can be in flight
constexpr int NF_MAX = 6;
constexpr int NE = 2;        // epilogue staging buffers (BM rows x 32 floats, 128B-swizzled, TMA-stored)
constexpr int kThreads = 384;

// shared memory carve-up (dynamic): [A stages: hi 8 KB | lo 8 KB] x NA, [B stages: N*BK*2 bytes] x NB, barriers
struct Smem {
  uint64_t f_full[NF_MAX], f_empty[NF_MAX], a_full[NA], a_empty[NA], b_full[NB], b_empty[NB], tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
};

// BWD = false: forward projection   vg[v,s,n]    = sum_c feat[v,c,s] W[n,c]   (A tile arrives [k][m], TMA-store epilogue)
// BWD = true : data gradient        gfeat[v,c,s] = sum_n gvg[v,s,n]  W[n,c]   (A tile arrives [m][k] 128B-swizzled,
//              transposed epilogue: lanes = consecutive pixels -> coalesced NCHW rows).  In both cases "C" is the
//              reduction length, "N" the output width, and wpack the [N x C] operand packed by sgc_pack_weight_tc.
template <bool BWD>
__global__ void __launch_bounds__(kThreads, 1)
project_tc_kernel(const __grid_constant__ CUtensorMap fmap, const __grid_constant__ CUtensorMap omap, int V, int C, int S,
                  const __nv_bfloat16* __restrict__ wpack, int N, float* __restrict__ vg, long long out_pitch, int dbg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a_stage_bytes = 2 * BM * BK * 2;          // hi + lo
  const int b_stage_bytes = N * BK * 2;
  constexpr int f_stage_bytes = BK * BM * 4;
  uint8_t* f_base = smem_raw;
  constexpr int nf = BWD ? NF_BWD : NF;
  constexpr int ne = BWD ? 0 : NE;
  uint8_t* a_base = f_base + nf * f_stage_bytes;
  uint8_t* b_base = a_base + NA * a_stage_bytes;
  uint8_t* e_base = b_base + NB * b_stage_bytes;  // 1024-byte aligned (all stage sizes are multiples of 1 KB)
  Smem* sm = reinterpret_cast<Smem*>(e_base + ne * BM * 128);

  const int n_parts = N > 256 ? 2 : 1;
  const int n_mma = N / n_parts;                      // 192 (C=256) or 256 (C=128): multiple of 16, <= 256
  // two TMEM accumulators (columns [0,256) and [256,512)) whenever the output tile fits twice: the MMAs of tile i+1
  // overlap the epilogue of tile i
  const int n_bufs = N <= 256 ? 2 : 1;
  const int k_slabs = C / BK;
  const int m_tiles = (S + BM - 1) / BM;
  const int tiles = V * m_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < nf; ++i) { mbar_init(&sm->f_full[i], 1); mbar_init(&sm->f_empty[i], 128); }
    for (int i = 0; i < NA; ++i) { mbar_init(&sm->a_full[i], 128); mbar_init(&sm->a_empty[i], 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(&sm->b_full[i], 1); mbar_init(&sm->b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sm->tmem_full[i], 1); mbar_init(&sm->tmem_empty[i], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 7) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;

  if (warp == 5) {
    // ===================== F producer: TMA tile loads of the fp32 NCHW map =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&fmap) : "memory");
      Pipe pf(nf);
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int v = tile / m_tiles, s0 = (tile - v * m_tiles) * BM;
        for (int j = 0; j < k_slabs; ++j) {
          mbar_wait(&sm->f_empty[pf.stage], pf.phase ^ 1);
          mbar_expect_tx(&sm->f_full[pf.stage], (uint32_t)f_stage_bytes);
          if (BWD) tma_load_3d(f_base + pf.stage * f_stage_bytes, &fmap, j * BK, s0, v, &sm->f_full[pf.stage]);
          else tma_load_2d(f_base + pf.stage * f_stage_bytes, &fmap, s0, v * C + j * BK, &sm->f_full[pf.stage]);
          pf.next();
        }
      }
    }
  } else if (warp < 4) {
    // ===================== A converters: fp32 staging -> bf16 hi/lo core-matrix tiles =====================
    const int m = threadIdx.x;  // row of the tile (pixel)
    Pipe pa(NA), pf(nf);
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      for (int j = 0; j < k_slabs; ++j) {
        mbar_wait(&sm->f_full[pf.stage], pf.phase);
        float x[BK];
        if (BWD) {  // staging tile is [m][k], 128 B per row, 16-byte chunks XOR-swizzled with (m & 7)
          const uint8_t* rowp = f_base + pf.stage * f_stage_bytes + m * 128;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 t4 = *reinterpret_cast<const float4*>(rowp + ((i ^ (m & 7)) << 4));
            x[4 * i] = t4.x; x[4 * i + 1] = t4.y; x[4 * i + 2] = t4.z; x[4 * i + 3] = t4.w;
          }
        } else {    // staging tile is [k][m]
          const float* src = reinterpret_cast<const float*>(f_base + pf.stage * f_stage_bytes) + m;
#pragma unroll
          for (int k = 0; k < BK; ++k) x[k] = src[k * BM];
        }
        mbar_wait(&sm->a_empty[pa.stage], pa.phase ^ 1);
        uint8_t* hi = a_base + pa.stage * a_stage_bytes;
        uint8_t* lo = hi + BM * BK * 2;
        const uint32_t off = (m >> 3) * SBO + (m & 7) * 16;
#pragma unroll
        for (int kc = 0; kc < ((dbg & 4) ? 0 : BK / 8); ++kc) {
          __align__(16) __nv_bfloat16 h[8], l[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            h[i] = __float2bfloat16_rn(x[kc * 8 + i]);
            l[i] = __float2bfloat16_rn(x[kc * 8 + i] - __bfloat162float(h[i]));
          }
          *reinterpret_cast<uint4*>(hi + off + kc * LBO) = *reinterpret_cast<const uint4*>(h);
          *reinterpret_cast<uint4*>(lo + off + kc * LBO) = *reinterpret_cast<const uint4*>(l);
        }
        // the staging slot is released only after every value read from it has been consumed by the conversion
        // (an arrive right after the ld.shared could overtake the loads and let the next TMA tile land early)
        mbar_arrive(&sm->f_empty[pf.stage]);
        pf.next();
        fence_proxy_async();
        mbar_arrive(&sm->a_full[pa.stage]);
        pa.next();
      }
    }
  } else if (warp == 4) {
    // ===================== B producer: bulk copies of the pre-packed weight slabs =====================
    if (lane == 0) {
      Pipe pb(NB);
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        for (int q = 0; q < 2 * k_slabs; ++q) {  // q = 2*slab + (0: hi, 1: lo)
          mbar_wait(&sm->b_empty[pb.stage], pb.phase ^ 1);
          mbar_expect_tx(&sm->b_full[pb.stage], (uint32_t)b_stage_bytes);
          bulk_g2s(b_base + pb.stage * b_stage_bytes, reinterpret_cast<const uint8_t*>(wpack) + (size_t)q * b_stage_bytes,
                   (uint32_t)b_stage_bytes, &sm->b_full[pb.stage]);
          pb.next();
        }
      }
    }
  } else if (warp == 6) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // cute::UMMA::InstrDescriptor: c_format F32 (1<<4), a/b format BF16 (1<<7, 1<<10), K-major A and B,
      // n_dim = N>>3 at [17,23), m_dim = M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      Pipe pa(NA), pb(NB);
      int it = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int buf = n_bufs == 2 ? (it & 1) : 0;
        const uint32_t tphase = n_bufs == 2 ? ((it >> 1) & 1) : (it & 1);
        const uint32_t acc = tmem + buf * 256;
        mbar_wait(&sm->tmem_empty[buf], tphase ^ 1);
        tc_fence_after();
        for (int j = 0; j < k_slabs; ++j) {
          mbar_wait(&sm->a_full[pa.stage], pa.phase);
          const uint32_t a_hi = smem_u32(a_base + pa.stage * a_stage_bytes);
          const uint32_t a_lo = a_hi + BM * BK * 2;
          // --- B_hi stage: hi*hi + lo*hi
          mbar_wait(&sm->b_full[pb.stage], pb.phase);
          tc_fence_after();
          uint32_t b_s = smem_u32(b_base + pb.stage * b_stage_bytes);
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            if (p < n_parts) {
#pragma unroll
              for (int ks = 0; ks < BK / 16; ++ks) {
                const uint64_t bd = umma_desc(b_s + p * (n_mma / 8) * SBO + ks * 2 * LBO);
                if (!(dbg & 2)) {
                  umma_bf16(acc + p * n_mma, umma_desc(a_hi + ks * 2 * LBO), bd, idesc, (j | ks) ? 1u : 0u);
                  umma_bf16(acc + p * n_mma, umma_desc(a_lo + ks * 2 * LBO), bd, idesc, 1u);
                }
              }
            }
          }
          tc_commit(&sm->b_empty[pb.stage]);
          pb.next();
          // --- B_lo stage: hi*lo
          mbar_wait(&sm->b_full[pb.stage], pb.phase);
          tc_fence_after();
          b_s = smem_u32(b_base + pb.stage * b_stage_bytes);
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            if (p < n_parts) {
#pragma unroll
              for (int ks = 0; ks < BK / 16; ++ks)
                if (!(dbg & 2))
                  umma_bf16(acc + p * n_mma, umma_desc(a_hi + ks * 2 * LBO), umma_desc(b_s + p * (n_mma / 8) * SBO + ks * 2 * LBO),
                            idesc, 1u);
            }
          }
          tc_commit(&sm->b_empty[pb.stage]);
          pb.next();
          tc_commit(&sm->a_empty[pa.stage]);
          pa.next();
        }
        tc_commit(&sm->tmem_full[buf]);
      }
    }
  } else if (warp >= 8) {
    // ===================== epilogue: TMEM -> registers -> swizzled smem -> TMA store =====================
    const int lane_base = (warp & 3) * 32;
    const int row = lane_base + lane;           // row of the tile == TMEM lane
    const bool issuer = (threadIdx.x == 8 * 32);
    int chunk = 0;                               // running chunk counter -> staging buffer parity
    int it = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const int v = tile / m_tiles, s0 = (tile - v * m_tiles) * BM;
      const int buf = n_bufs == 2 ? (it & 1) : 0;
      const uint32_t ephase = n_bufs == 2 ? ((it >> 1) & 1) : (it & 1);
      const uint32_t acc = tmem + buf * 256;
      mbar_wait(&sm->tmem_full[buf], ephase);
      tc_fence_after();
      for (int c0 = 0; c0 < N; c0 += 32, ++chunk) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(acc + ((uint32_t)lane_base << 16) + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (BWD) {
          const int sg = s0 + row;
          if (sg < S) {
            float* dstc = vg + ((size_t)v * N + c0) * out_pitch + sg;
#pragma unroll
            for (int i = 0; i < 32; ++i) dstc[(size_t)i * out_pitch] = __uint_as_float(r[i]);
          }
          continue;
        }
        uint8_t* buf = e_base + (chunk & 1) * (BM * 128);
        // the TMA store that read this buffer two chunks ago must have finished reading it
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (!(dbg & 1)) {
#pragma unroll
          for (int i = 0; i < 8; ++i)  // 16-byte chunk i of the 128-byte row, CU_TENSOR_MAP_SWIZZLE_128B pattern
            *reinterpret_cast<uint4*>(buf + row * 128 + ((i ^ (row & 7)) << 4)) = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (issuer) {
          if (!(dbg & 1)) tma_store_3d(&omap, c0, s0, v, buf);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      tc_fence_before();
      mbar_arrive(&sm->tmem_empty[buf]);
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 7) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

// Pack Wcat [N,C] fp32 into the shared-memory image the kernel above bulk-copies:
//   out[(2*slab + part)][row-group n/8][k-chunk kc][n%8][8 k]  bf16, part 0 = hi, 1 = lo.
__global__ void pack_weight_kernel(const float* __restrict__ w, int N, int C, __nv_bfloat16* __restrict__ out) {
  const int total = N * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n = i / C, c = i - n * C;
    const float x = w[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    const int slab = c / BK, k = c - slab * BK;
    const size_t stage_elems = (size_t)N * BK;
    const size_t o = (size_t)(n >> 3) * (SBO / 2) + (size_t)(k >> 3) * (LBO / 2) + (n & 7) * 8 + (k & 7);
    out[(size_t)(2 * slab) * stage_elems + o] = h;
    out[(size_t)(2 * slab + 1) * stage_elems + o] = l;
  }
}


// One launch for ALL per-step operand preparations of a level's weights (replaces ~35 tiny launches):
// every job reads a logical [rows, cols] fp32 matrix through (row_stride, col_stride) -- so transposed operands need
// no copy --, multiplies by `scale`, and writes either the bf16x3 K-concatenated split of sgc_split_bf16x3 (kind 0)
// or the tcgen05 slab image of pack_weight_kernel (kind 1, rows = N, cols = C).
struct WeightJobs {
  sgc_weight_job job[SGC_MAX_WEIGHT_JOBS];
};

__global__ void __launch_bounds__(256) prepare_weights_kernel(const __grid_constant__ WeightJobs jobs) {
  const sgc_weight_job& j = jobs.job[blockIdx.y];
  const int total = j.rows * j.cols;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(j.out);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / j.cols, c = i - r * j.cols;
    const float x = __ldg(j.src + (long long)r * j.row_stride + (long long)c * j.col_stride) * j.scale;
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    if (j.kind == 0) {
      const int rpg = j.rows_per_group;
      const size_t slot = (size_t)rpg * j.cols;
      const size_t base = ((size_t)(r / rpg) * 3 * rpg + (r % rpg)) * j.cols + c;
      out[base] = h;
      out[base + slot] = j.pattern ? h : l;
      out[base + 2 * slot] = j.pattern ? l : h;
    } else {
      const int slab = c / BK, k = c - slab * BK;
      const size_t stage_elems = (size_t)j.rows * BK;
      const size_t o = (size_t)(r >> 3) * (SBO / 2) + (size_t)(k >> 3) * (LBO / 2) + (r & 7) * 8 + (k & 7);
      out[(size_t)(2 * slab) * stage_elems + o] = h;
      out[(size_t)(2 * slab + 1) * stage_elems + o] = l;
    }
  }
}

}  // namespace tc
}  // namespace sgc

// Upper bound on the CTAs (= SMs, the kernels are persistent with one CTA per SM) the tensor-core projection kernels
// occupy; 0 = all SMs.  The kernels run on side streams next to the latency-bound per-voxel chain: leaving a few SMs
// free lets the chain's kernels start immediately instead of waiting for a persistent CTA to retire.
static int g_tc_max_ctas = 0;
extern "C" int sgc_project_tc_set_max_ctas(int n) {
  if (n < 0) return (int)cudaErrorInvalidValue;
  g_tc_max_ctas = n;
  return 0;
}
// Separate cap for the FORWARD projection only: in the forward the projections of the finer levels run beside the voxel
// chain of the coarser ones and are far off the critical path, so they can leave more SMs to the chain than the backward
// kernels (which ARE the critical path of the backward).  0 = use the common cap.
static int g_tc_max_ctas_fwd = 0;
extern "C" int sgc_project_tc_set_max_ctas_fwd(int n) {
  if (n < 0) return (int)cudaErrorInvalidValue;
  g_tc_max_ctas_fwd = n;
  return 0;
}
static int g_tc_tiles_per_cta = 0;  // 0 = persistent (grid = SM cap); n > 0: short-lived CTAs of n tiles each
extern "C" int sgc_project_tc_set_tiles_per_cta(int n) {
  if (n < 0) return (int)cudaErrorInvalidValue;
  g_tc_tiles_per_cta = n;
  return 0;
}
static inline int tc_grid(int tiles, int sms);
static inline int tc_cta_cap(int sms) { return (g_tc_max_ctas > 0 && g_tc_max_ctas < sms) ? g_tc_max_ctas : sms; }

static inline int tc_grid(int tiles, int sms) {
  if (g_tc_tiles_per_cta > 0) return (tiles + g_tc_tiles_per_cta - 1) / g_tc_tiles_per_cta;
  const int cap = tc_cta_cap(sms);
  return tiles < cap ? tiles : cap;
}

extern "C" int sgc_prepare_weights(const sgc_weight_job* jobs, int njobs, void* stream) {
  if (njobs <= 0) return 0;
  if (njobs > SGC_MAX_WEIGHT_JOBS) return (int)cudaErrorInvalidValue;
  sgc::tc::WeightJobs wj;
  for (int i = 0; i < njobs; ++i) {
    const sgc_weight_job& j = jobs[i];
    if (j.rows <= 0 || j.cols <= 0 || !j.src || !j.out) return (int)cudaErrorInvalidValue;
    if (j.kind == 0 && (j.rows_per_group <= 0 || j.rows % j.rows_per_group)) return (int)cudaErrorInvalidValue;
    if (j.kind == 1 && (j.rows % 16 || j.cols % sgc::tc::BK)) return (int)cudaErrorInvalidValue;
    if (j.kind != 0 && j.kind != 1) return (int)cudaErrorInvalidValue;
    wj.job[i] = j;
  }
  sgc::tc::prepare_weights_kernel<<<dim3(96, njobs), 256, 0, (cudaStream_t)stream>>>(wj);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

extern "C" int sgc_pack_weight_tc(const float* w, int N, int C, void* out, void* stream) {
  if (N % 16 || C % sgc::tc::BK) return (int)cudaErrorInvalidValue;
  sgc::tc::pack_weight_kernel<<<(N * C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, N, C, (__nv_bfloat16*)out);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

// feat: [V*C] channel planes of S_full floats each (element (v,c,s) at feat[(v*C+c)*chan_stride + s]); only s < S is used.
// wpack: sgc_pack_weight_tc output (2*N*C bf16).  vg: [V,S,N] fp32, fully written.
using sgc::tc::PFN_encodeTiled;

extern "C" int sgc_project_tc_fwd(const float* feat, long long view_stride, long long chan_stride, int V, int C, int S,
                                  const void* wpack, int N, float* vg, void* stream) {
  using namespace sgc::tc;
  if (C % BK || N % 32 || N > 512 || (N > 256 && (N / 2) % 16) || V <= 0 || S <= 0) return (int)cudaErrorInvalidValue;
  // the TMA descriptor needs a regular [V*C, chan_stride] matrix with a 16-byte aligned pitch
  if (view_stride != (long long)C * chan_stride || (chan_stride * 4) % 16 || (reinterpret_cast<uintptr_t>(feat) & 15) ||
      chan_stride < S)
    return (int)cudaErrorInvalidValue;
  static int sms = 0;
  static PFN_encodeTiled encode = nullptr;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return (int)cudaErrorNotSupported;
    encode = (PFN_encodeTiled)fn;
  }
  CUtensorMap fmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)chan_stride, (cuuint64_t)V * C};
  const cuuint64_t gstr[1] = {(cuuint64_t)chan_stride * 4};
  const cuuint32_t box[2] = {(cuuint32_t)BM, (cuuint32_t)BK};
  const cuuint32_t estr[2] = {1, 1};
  if (encode(&fmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(feat), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return (int)cudaErrorInvalidValue;
  CUtensorMap omap;
  const cuuint64_t odim[3] = {(cuuint64_t)N, (cuuint64_t)S, (cuuint64_t)V};
  const cuuint64_t ostr[2] = {(cuuint64_t)N * 4, (cuuint64_t)S * N * 4};
  const cuuint32_t obox[3] = {32, (cuuint32_t)BM, 1};
  const cuuint32_t oestr[3] = {1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(vg) & 15) ||
      encode(&omap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, vg, odim, ostr, obox, oestr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return (int)cudaErrorInvalidValue;
  const size_t smem = (size_t)NF * BK * BM * 4 + (size_t)NA * 2 * BM * BK * 2 + (size_t)NB * N * BK * 2 + (size_t)NE * BM * 128 +
                      sizeof(Smem) + 64;
  cudaError_t e = cudaFuncSetAttribute(project_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int tiles = V * ((S + BM - 1) / BM);
  int grid = tc_grid(tiles, sms);
  if (g_tc_tiles_per_cta == 0 && g_tc_max_ctas_fwd > 0 && grid > g_tc_max_ctas_fwd) grid = g_tc_max_ctas_fwd;
  project_tc_kernel<false><<<grid, kThreads, smem, (cudaStream_t)stream>>>(fmap, omap, V, C, S, (const __nv_bfloat16*)wpack, N, vg, 0,
                                                                           getenv("SGC_TC_DBG") ? atoi(getenv("SGC_TC_DBG")) : 0);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

// Data gradient of the projection on the tensor cores: gfeat[v,c,s] = sum_n gvg[v,s,n] * W[n,c], written straight into
// the NCHW gradient (pitch chan_stride per channel plane; only s < S is written).  wpack_t = sgc_pack_weight_tc(W^T [C,N]).
extern "C" int sgc_project_tc_bwd_data(const float* gvg, int V, int S, int N, const void* wpack_t, int C, float* gfeat,
                                       long long chan_stride, void* stream) {
  using namespace sgc::tc;
  if (N % BK || C % 32 || C > 256 || V <= 0 || S <= 0 || chan_stride < S) return (int)cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(gvg) & 15) || (N * 4) % 16) return (int)cudaErrorInvalidValue;
  static int sms = 0;
  static PFN_encodeTiled encode = nullptr;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return (int)cudaErrorNotSupported;
    encode = (PFN_encodeTiled)fn;
  }
  CUtensorMap fmap;
  const cuuint64_t gdim[3] = {(cuuint64_t)N, (cuuint64_t)S, (cuuint64_t)V};
  const cuuint64_t gstr[2] = {(cuuint64_t)N * 4, (cuuint64_t)S * N * 4};
  const cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BM, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  if (encode(&fmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(gvg), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return (int)cudaErrorInvalidValue;
  const size_t smem = (size_t)NF_BWD * BK * BM * 4 + (size_t)NA * 2 * BM * BK * 2 + (size_t)NB * C * BK * 2 +
                      sizeof(Smem) + 64;
  cudaError_t e = cudaFuncSetAttribute(project_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int tiles = V * ((S + BM - 1) / BM);
  const int grid = tc_grid(tiles, sms);
  // kernel convention: reduction length = N (channels of gvg), output width = C
  project_tc_kernel<true><<<grid, kThreads, smem, (cudaStream_t)stream>>>(fmap, fmap, V, N, S, (const __nv_bfloat16*)wpack_t, C, gfeat,
                                                                          chan_stride, 0);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

// =====================================================================================================================
// Weight gradient of the projection on the tensor cores:
//     gw[n,c] = sum_{v,s} gvg[v,s,n] * feat[v,c,s]            (reduction over all pixels of all views)
// Both operands are fp32 in HBM and are split to bf16 hi/lo in shared memory (no split round trip):
//   A[m = n][k = s]  <- gvg tile  [32 s][128 n]  (TMA 3-D, rows = s)           -> "[k][m]" converter (as the forward)
//   B[c][k = s]      <- feat tile [C rows][32 s] (TMA 2-D, 128B-swizzled rows) -> "[m][k]" converter (as the data grad)
// Split-K: CTA (m-tile, k-chunk) accumulates its slab range in TMEM and writes a partial [128, C] tile; a second kernel
// sums the partials in a fixed order (deterministic).
namespace sgc {
namespace tc {

constexpr int WG_THREADS = 480;  // warps 0-3: A converters (+ epilogue), 4-11: B converters, 12: TMA, 13: MMA, 14: TMEM
constexpr int WG_CONV = 384;     // converter threads (arrivals on f_empty / op_full)
constexpr int WG_ST = 2;  // pipeline stages (staging and operand)

struct SmemW {
  uint64_t f_full[WG_ST], f_empty[WG_ST], op_full[WG_ST], op_empty[WG_ST], tmem_full;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap, int V, int S, int C,
                int m_tiles, int slabs_per_cta, float* __restrict__ partial) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a_src = BK * BM * 4;              // 16 KB  [32 s][128 n] fp32
  const int b_src = C * 128;                  // [C][32 s] fp32, swizzled
  const int f_stage = a_src + b_src;
  const int a_op = 2 * BM * BK * 2;           // hi + lo, 16 KB
  const int b_op = 2 * C * BK * 2;            // hi + lo
  const int op_stage = a_op + b_op;
  uint8_t* f_base = smem_raw;
  uint8_t* op_base = f_base + WG_ST * f_stage;
  SmemW* sm = reinterpret_cast<SmemW*>(op_base + WG_ST * op_stage);

  const int mt = blockIdx.x % m_tiles, kc = blockIdx.x / m_tiles;
  const int spv = (S + BK - 1) / BK;          // slabs per view
  const int total = V * spv;
  const int s_begin = kc * slabs_per_cta;
  const int s_end = min(total, s_begin + slabs_per_cta);
  const int n_slabs = max(0, s_end - s_begin);

  if (threadIdx.x == 0) {
    for (int i = 0; i < WG_ST; ++i) {
      mbar_init(&sm->f_full[i], 1); mbar_init(&sm->f_empty[i], WG_CONV);
      mbar_init(&sm->op_full[i], WG_CONV); mbar_init(&sm->op_empty[i], 1);
    }
    mbar_init(&sm->tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 14) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_base)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;

  if (warp == 12) {
    if (lane == 0) {
      Pipe pf(WG_ST);
      for (int i = 0; i < n_slabs; ++i) {
        const int idx = s_begin + i, v = idx / spv, s0 = (idx - v * spv) * BK;
        mbar_wait(&sm->f_empty[pf.stage], pf.phase ^ 1);
        mbar_expect_tx(&sm->f_full[pf.stage], (uint32_t)f_stage);
        uint8_t* st = f_base + pf.stage * f_stage;
        tma_load_3d(st, &amap, mt * BM, s0, v, &sm->f_full[pf.stage]);
        tma_load_2d(st + a_src, &bmap, s0, v * C, &sm->f_full[pf.stage]);
        pf.next();
      }
    }
  } else if (warp < 12) {
    // converters: warps 0-3 -> A (gvg, [k][m] tile), warps 4-11 -> B (feat, swizzled [row][k] tile; one row per thread
    // for C = 256, so that the two operands take the same time per slab)
    const bool is_b = warp >= 4;
    const int t = is_b ? (int)threadIdx.x - 128 : (int)threadIdx.x;   // A: 0..127, B: 0..255
    Pipe pf(WG_ST), po(WG_ST);
    for (int i = 0; i < n_slabs; ++i) {
      mbar_wait(&sm->f_full[pf.stage], pf.phase);
      const uint8_t* st = f_base + pf.stage * f_stage;
      mbar_wait(&sm->op_empty[po.stage], po.phase ^ 1);
      uint8_t* op = op_base + po.stage * op_stage;
      if (!is_b) {
        const float* src = reinterpret_cast<const float*>(st) + t;
        float x[BK];
#pragma unroll
        for (int k = 0; k < BK; ++k) x[k] = src[k * BM];
        uint8_t* hi = op;
        uint8_t* lo = op + BM * BK * 2;
        const uint32_t off = (t >> 3) * SBO + (t & 7) * 16;
#pragma unroll
        for (int kcx = 0; kcx < BK / 8; ++kcx) {
          __align__(16) __nv_bfloat16 h[8], l[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            h[q] = __float2bfloat16_rn(x[kcx * 8 + q]);
            l[q] = __float2bfloat16_rn(x[kcx * 8 + q] - __bfloat162float(h[q]));
          }
          *reinterpret_cast<uint4*>(hi + off + kcx * LBO) = *reinterpret_cast<const uint4*>(h);
          *reinterpret_cast<uint4*>(lo + off + kcx * LBO) = *reinterpret_cast<const uint4*>(l);
        }
      } else {
        uint8_t* hi = op + a_op;
        uint8_t* lo = hi + C * BK * 2;
        for (int r = t; r < C; r += 256) {
          const uint8_t* rowp = st + a_src + r * 128;
          const uint32_t off = (r >> 3) * SBO + (r & 7) * 16;
#pragma unroll
          for (int kcx = 0; kcx < BK / 8; ++kcx) {
            const float4 u0 = *reinterpret_cast<const float4*>(rowp + (((2 * kcx) ^ (r & 7)) << 4));
            const float4 u1 = *reinterpret_cast<const float4*>(rowp + (((2 * kcx + 1) ^ (r & 7)) << 4));
            const float xv[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
            __align__(16) __nv_bfloat16 h[8], l[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              h[q] = __float2bfloat16_rn(xv[q]);
              l[q] = __float2bfloat16_rn(xv[q] - __bfloat162float(h[q]));
            }
            *reinterpret_cast<uint4*>(hi + off + kcx * LBO) = *reinterpret_cast<const uint4*>(h);
            *reinterpret_cast<uint4*>(lo + off + kcx * LBO) = *reinterpret_cast<const uint4*>(l);
          }
        }
      }
      mbar_arrive(&sm->f_empty[pf.stage]);   // after the staged values were consumed (see the forward kernel)
      pf.next();
      fence_proxy_async();
      mbar_arrive(&sm->op_full[po.stage]);
      po.next();
    }
    if (!is_b) {
      // epilogue by warps 0-3: partial[kc][mt*128 + row][0..C)
      const int lane_base = (warp & 3) * 32;
      const int row = lane_base + lane;
      float* dst = partial + ((size_t)kc * m_tiles * BM + (size_t)mt * BM + row) * C;
      if (n_slabs > 0) {
        mbar_wait(&sm->tmem_full, 0);
        tc_fence_after();
      }
      for (int c0 = 0; c0 < C; c0 += 32) {
        uint32_t r[32];
        if (n_slabs > 0) {
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
              : "r"(tmem + ((uint32_t)lane_base << 16) + (uint32_t)c0));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) r[q] = 0u;
        }
#pragma unroll
        for (int q = 0; q < 32; q += 4)
          *reinterpret_cast<uint4*>(dst + c0 + q) = make_uint4(r[q], r[q + 1], r[q + 2], r[q + 3]);
      }
    }
  } else if (warp == 13) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      Pipe po(WG_ST);
      for (int i = 0; i < n_slabs; ++i) {
        mbar_wait(&sm->op_full[po.stage], po.phase);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(op_base + po.stage * op_stage);
        const uint32_t a_lo = a_hi + BM * BK * 2;
        const uint32_t b_hi = a_hi + a_op;
        const uint32_t b_lo = b_hi + C * BK * 2;
#pragma unroll
        for (int ks = 0; ks < BK / 16; ++ks) {
          const uint32_t o = ks * 2 * LBO;
          umma_bf16(tmem, umma_desc(a_hi + o), umma_desc(b_hi + o), idesc, (i | ks) ? 1u : 0u);
          umma_bf16(tmem, umma_desc(a_lo + o), umma_desc(b_hi + o), idesc, 1u);
          umma_bf16(tmem, umma_desc(a_hi + o), umma_desc(b_lo + o), idesc, 1u);
        }
        tc_commit(&sm->op_empty[po.stage]);
        po.next();
      }
      if (n_slabs > 0) tc_commit(&sm->tmem_full);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 14) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
  }
}

// gw[i] = sum_k partial[k][i]   (fixed order)
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int kch, int elems, float* __restrict__ gw) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= elems) return;
  // the last kernel of the step: 8 partials in flight per thread (the additions stay in k order: deterministic); one load per
  // iteration made it a chain of ~49 dependent L2 round trips (23 us for 98 K outputs)
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  int k = 0;
  for (; k + 8 <= kch; k += 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(partial + (size_t)(k + u) * elems + i));
#pragma unroll
    for (int u = 0; u < 8; ++u) { a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w; }
  }
  for (; k < kch; ++k) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(partial + (size_t)k * elems + i));
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  *reinterpret_cast<float4*>(gw + i) = a;
}

}  // namespace tc
}  // namespace sgc

extern "C" int sgc_project_tc_wgrad_scratch_floats(int N, int C) {
  const int m_tiles = N / sgc::tc::BM;
  const int kch = 148 / (m_tiles > 0 ? m_tiles : 1);
  return kch * N * C;
}

// gw [N,C] = sum_{v,s} gvg[v,s,n] feat[(v*C+c)*chan_stride + s]; scratch: sgc_project_tc_wgrad_scratch_floats(N,C) floats.
extern "C" int sgc_project_tc_wgrad(const float* gvg, const float* feat, long long chan_stride, int V, int S, int N, int C,
                                    float* gw, float* scratch, void* stream) {
  using namespace sgc::tc;
  if (N % BM || C % 32 || C > 256 || V <= 0 || S <= 0 || chan_stride < S) return (int)cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(gvg) & 15) || (reinterpret_cast<uintptr_t>(feat) & 15) || (chan_stride * 4) % 16)
    return (int)cudaErrorInvalidValue;
  static PFN_encodeTiled encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return (int)cudaErrorNotSupported;
    encode = (PFN_encodeTiled)fn;
  }
  CUtensorMap amap, bmap;
  {
    const cuuint64_t gdim[3] = {(cuuint64_t)N, (cuuint64_t)S, (cuuint64_t)V};
    const cuuint64_t gstr[2] = {(cuuint64_t)N * 4, (cuuint64_t)S * N * 4};
    const cuuint32_t box[3] = {(cuuint32_t)BM, (cuuint32_t)BK, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (encode(&amap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(gvg), gdim, gstr, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return (int)cudaErrorInvalidValue;
  }
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)chan_stride, (cuuint64_t)V * C};
    const cuuint64_t gstr[1] = {(cuuint64_t)chan_stride * 4};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)C};
    const cuuint32_t estr[2] = {1, 1};
    if (encode(&bmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(feat), gdim, gstr, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return (int)cudaErrorInvalidValue;
  }
  const int m_tiles = N / BM;
  const int kch = tc_cta_cap(148) / m_tiles > 0 ? tc_cta_cap(148) / m_tiles : 1;
  const int spv = (S + BK - 1) / BK;
  const int total = V * spv;
  const int slabs_per_cta = (total + kch - 1) / kch;
  const size_t smem = (size_t)WG_ST * (BK * BM * 4 + C * 128) + (size_t)WG_ST * (2 * BM * BK * 2 + 2 * C * BK * 2) + sizeof(SmemW) + 64;
  cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  wgrad_tc_kernel<<<m_tiles * kch, WG_THREADS, smem, (cudaStream_t)stream>>>(amap, bmap, V, S, C, m_tiles, slabs_per_cta, scratch);
  SGC_CUDA_CHECK_LAST();
  const int elems = N * C;
  wgrad_reduce_kernel<<<(elems / 4 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(scratch, kch, elems, gw);
  SGC_CUDA_CHECK_LAST();
  return 0;
}
