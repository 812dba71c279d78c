// sgc_rows_chain_bwd_tc: the backward of the row-local tail of one encoder layer (the mirror of csrc/sgc_rows_chain_tc.cu)
// as ONE launch instead of three row-kernel + three GEMM launches of functional.EncoderLayerRows.backward:
//
//   R0   v = LayerNorm2'(gy);  gpre2 = v;  gf = v * mask2*s2                      (+ partial sums for gamma2 / beta2)
//   G1   ghdn = gf @ W_2;      gh = hdn > 0 ? ghdn * s1 : 0                       (ReLU gate and dropout mask of the FFN)
//   G2   gx1 = gh @ W_1
//   R2   v = LayerNorm1'(gx1 + gpre2);  gout = v * mask0*s0 * [view count > 0]    (+ partial sums for gamma1 / beta1)
//   G3   go2 = gout @ W_o
//
// A CTA owns a 128-row tile.  The GEMMs use the pipeline of rows_gemm_tc_kernel (TMA loads of fp32 rows, bf16 hi/lo split in
// shared memory, packed weight slabs, TMEM accumulators); the two LayerNorm backward steps are the arithmetic of
// sgc_rowop_bwd (warp per row, lanes over channels) executed by the four epilogue warps on the tile's rows, between the
// GEMMs: their results go to global memory with ordinary stores, are fenced (device scope + async proxy) and the TMA
// producer is released through an mbarrier.  gf, gh and gout are the operands of the weight-gradient kernels and are
// written exactly as the unfused path writes them; the per-CTA gamma / beta partials use the layout of
// sgc_layernorm_bwd_params (one row of 2C floats per CTA; the caller zero-fills the rows no CTA writes).
//
// STATUS (end of round 1): compiled, never run on a GPU (the round's GPU budget was spent); NOT on the product path.
// tests/test_gpu_rows_chain.py::test_rows_chain_bwd_* only run with SGC_TEST_CHAIN_BWD=1.
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/sgcdet_b200.h"

namespace sgc {
namespace tc {

constexpr int RB_NF = 4, RB_NA = 2, RB_NB = 4, RB_NE = 2;
constexpr int RB_THREADS = 384;
constexpr int RB_MAX_ITEMS = 4;
constexpr int RB_B_STAGE = 256 * BK * 2;

struct SmemRB {
  uint64_t f_full[RB_NF], f_empty[RB_NF], a_full[RB_NA], a_empty[RB_NA], b_full[RB_NB], b_empty[RB_NB], tmem_full[2],
      tmem_empty[2], ready[3];
  uint32_t tmem_base;
};

// LayerNorm backward + mask / row-count scaling over the rows of the tile (the arithmetic of rowop_bwd_kernel)
struct RowStage {
  const float* g;              // [R, C] incoming gradient
  const float* g2;             // [R, C] second gradient added first, or null
  const float* pre;            // [R, C] LayerNorm input saved by the forward
  const float* mean;           // [R]
  const float* rstd;           // [R]
  const float* gamma;          // [C]
  const unsigned char* mask;   // [R, C] keep-mask or null
  const int* rowcount;         // [R] or null
  float* gpre;                 // [R, C] LayerNorm-input gradient (before mask / row scaling) or null
  float* gx;                   // [R, C] output
  float* partial;              // [grid, 2C] gamma / beta partial sums
  float mscale;
};

struct BwdItem {
  const uint8_t* wpack;
  long long pack_stage_bytes;
  const float* gate;           // mode 1: [R, ldg] forward activation (hdn)
  float gscale;
  int ldg;
  int a_map, o_map, k_slabs, n_cta, col0;
  int mode;                    // 0: plain store, 1: v = gate > 0 ? acc * gscale : 0
  int rowstage;                // row stage run by the epilogue warps BEFORE this item (its output is this item's operand), -1: none
  int wait_ready;              // producer waits for ready[wait_ready] before its first load, -1: none
  int signal_ready;            // epilogue signals ready[signal_ready] once this item's stores completed, -1: none
};

struct BwdParams {
  BwdItem item[RB_MAX_ITEMS];
  RowStage rs[2];
  int n_items, R, C;
};

struct BwdMaps {
  CUtensorMap a[3];   // gf, gh, gout as operands
  CUtensorMap o[3];   // gh, gx1, go2 as outputs
};

__device__ __forceinline__ void ld_tmem_32b(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Rows [row0, row0+128) of the tile, 32 per epilogue warp, lanes over CPL channels each.  `scratch` = 4 x 2C floats of shared
// memory (the idle epilogue staging buffers).  Called by all 128 epilogue threads.
template <int CPL>
__device__ __forceinline__ void row_stage(const RowStage& s, int row0, int R, int ewarp, int lane, float* scratch) {
  constexpr int N = 32 * CPL;
  const int c0 = lane * CPL;
  float gam[CPL], agam[CPL], abet[CPL];
#pragma unroll
  for (int j = 0; j < CPL; j += 4) {
    const float4 t = ldg4(s.gamma + c0 + j);
    gam[j] = t.x; gam[j + 1] = t.y; gam[j + 2] = t.z; gam[j + 3] = t.w;
  }
#pragma unroll
  for (int j = 0; j < CPL; ++j) { agam[j] = 0.f; abet[j] = 0.f; }
  for (int rr = 0; rr < 32; ++rr) {
    const int r = row0 + ewarp * 32 + rr;
    if (r >= R) break;
    float v[CPL], xh[CPL];
#pragma unroll
    for (int j = 0; j < CPL; j += 4) {
      // plain loads: g (and g2) may have been written by this very kernel a moment ago
      const float4 t = *reinterpret_cast<const float4*>(s.g + (size_t)r * N + c0 + j);
      v[j] = t.x; v[j + 1] = t.y; v[j + 2] = t.z; v[j + 3] = t.w;
    }
    if (s.g2) {
#pragma unroll
      for (int j = 0; j < CPL; j += 4) {
        const float4 t = *reinterpret_cast<const float4*>(s.g2 + (size_t)r * N + c0 + j);
        v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
      }
    }
    const float mu = __ldg(s.mean + r), rs = __ldg(s.rstd + r);
#pragma unroll
    for (int j = 0; j < CPL; j += 4) {
      const float4 t = ldg4(s.pre + (size_t)r * N + c0 + j);
      xh[j] = t.x; xh[j + 1] = t.y; xh[j + 2] = t.z; xh[j + 3] = t.w;
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      xh[j] = (xh[j] - mu) * rs;
      agam[j] += v[j] * xh[j];
      abet[j] += v[j];
      v[j] *= gam[j];
      s1 += v[j];
      s2 += v[j] * xh[j];
    }
    s1 = warp_sum(s1) * (1.f / N);
    s2 = warp_sum(s2) * (1.f / N);
#pragma unroll
    for (int j = 0; j < CPL; ++j) v[j] = rs * (v[j] - s1 - xh[j] * s2);
    if (s.gpre) {
#pragma unroll
      for (int j = 0; j < CPL; j += 4)
        *reinterpret_cast<float4*>(s.gpre + (size_t)r * N + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    if (s.mask) {
#pragma unroll
      for (int j = 0; j < CPL; j += 4) {
        const uchar4 m = *reinterpret_cast<const uchar4*>(s.mask + (size_t)r * N + c0 + j);
        v[j] *= (m.x ? 1.f : 0.f) * s.mscale; v[j + 1] *= (m.y ? 1.f : 0.f) * s.mscale;
        v[j + 2] *= (m.z ? 1.f : 0.f) * s.mscale; v[j + 3] *= (m.w ? 1.f : 0.f) * s.mscale;
      }
    }
    if (s.rowcount) {
      const float rsc = __ldg(s.rowcount + r) > 0 ? 1.f : 0.f;
#pragma unroll
      for (int j = 0; j < CPL; ++j) v[j] *= rsc;
    }
#pragma unroll
    for (int j = 0; j < CPL; j += 4)
      *reinterpret_cast<float4*>(s.gx + (size_t)r * N + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  }
  // per-CTA gamma / beta partials: the four warps' sums combined in a fixed order
#pragma unroll
  for (int j = 0; j < CPL; ++j) { scratch[ewarp * 2 * N + c0 + j] = agam[j]; scratch[ewarp * 2 * N + N + c0 + j] = abet[j]; }
  asm volatile("bar.sync 1, 128;" ::: "memory");
  for (int c = ewarp * 32 + lane; c < 2 * N; c += 128)
    s.partial[(size_t)blockIdx.x * 2 * N + c] = scratch[c] + scratch[2 * N + c] + scratch[4 * N + c] + scratch[6 * N + c];
  // the rows written above are TMA-loaded by this CTA next: make them visible to the async proxy
  __threadfence();
  asm volatile("fence.proxy.async;" ::: "memory");
  asm volatile("bar.sync 1, 128;" ::: "memory");
}

__global__ void __launch_bounds__(RB_THREADS, 1)
rows_chain_bwd_tc_kernel(const __grid_constant__ BwdMaps maps, const __grid_constant__ BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int f_stage_bytes = BK * BM * 4;
  constexpr int a_stage_bytes = 2 * BM * BK * 2;
  uint8_t* f_base = smem_raw;
  uint8_t* a_base = f_base + RB_NF * f_stage_bytes;
  uint8_t* b_base = a_base + RB_NA * a_stage_bytes;
  uint8_t* e_base = b_base + RB_NB * RB_B_STAGE;
  SmemRB* sm = reinterpret_cast<SmemRB*>(e_base + RB_NE * BM * 128);
  const int row0 = blockIdx.x * BM;

  if (threadIdx.x == 0) {
    for (int i = 0; i < RB_NF; ++i) { mbar_init(&sm->f_full[i], 1); mbar_init(&sm->f_empty[i], 128); }
    for (int i = 0; i < RB_NA; ++i) { mbar_init(&sm->a_full[i], 128); mbar_init(&sm->a_empty[i], 1); }
    for (int i = 0; i < RB_NB; ++i) { mbar_init(&sm->b_full[i], 1); mbar_init(&sm->b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sm->tmem_full[i], 1); mbar_init(&sm->tmem_empty[i], 128); }
    for (int i = 0; i < 3; ++i) mbar_init(&sm->ready[i], 1);
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
    if (lane == 0) {
      Pipe pf(RB_NF);
      for (int it = 0; it < p.n_items; ++it) {
        const BwdItem& I = p.item[it];
        if (I.wait_ready >= 0) mbar_wait(&sm->ready[I.wait_ready], 0);
        for (int j = 0; j < I.k_slabs; ++j) {
          mbar_wait(&sm->f_empty[pf.stage], pf.phase ^ 1);
          mbar_expect_tx(&sm->f_full[pf.stage], (uint32_t)f_stage_bytes);
          tma_load_3d(f_base + pf.stage * f_stage_bytes, &maps.a[I.a_map], j * BK, row0, 0, &sm->f_full[pf.stage]);
          pf.next();
        }
      }
    }
  } else if (warp < 4) {
    const int m = threadIdx.x;
    Pipe pa(RB_NA), pf(RB_NF);
    for (int it = 0; it < p.n_items; ++it) {
      const int k_slabs = p.item[it].k_slabs;
      for (int j = 0; j < k_slabs; ++j) {
        mbar_wait(&sm->f_full[pf.stage], pf.phase);
        float x[BK];
        const uint8_t* rowp = f_base + pf.stage * f_stage_bytes + m * 128;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t4 = *reinterpret_cast<const float4*>(rowp + ((i ^ (m & 7)) << 4));
          x[4 * i] = t4.x; x[4 * i + 1] = t4.y; x[4 * i + 2] = t4.z; x[4 * i + 3] = t4.w;
        }
        mbar_wait(&sm->a_empty[pa.stage], pa.phase ^ 1);
        uint8_t* hi = a_base + pa.stage * a_stage_bytes;
        uint8_t* lo = hi + BM * BK * 2;
        const uint32_t off = (m >> 3) * SBO + (m & 7) * 16;
#pragma unroll
        for (int kc = 0; kc < BK / 8; ++kc) {
          __align__(16) __nv_bfloat16 h[8], l[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            h[i] = __float2bfloat16_rn(x[kc * 8 + i]);
            l[i] = __float2bfloat16_rn(x[kc * 8 + i] - __bfloat162float(h[i]));
          }
          *reinterpret_cast<uint4*>(hi + off + kc * LBO) = *reinterpret_cast<const uint4*>(h);
          *reinterpret_cast<uint4*>(lo + off + kc * LBO) = *reinterpret_cast<const uint4*>(l);
        }
        mbar_arrive(&sm->f_empty[pf.stage]);
        pf.next();
        fence_proxy_async();
        mbar_arrive(&sm->a_full[pa.stage]);
        pa.next();
      }
    }
  } else if (warp == 4) {
    if (lane == 0) {
      Pipe pb(RB_NB);
      for (int it = 0; it < p.n_items; ++it) {
        const BwdItem& I = p.item[it];
        const uint32_t bytes = (uint32_t)I.n_cta * BK * 2;
        const uint8_t* src = I.wpack + (size_t)I.col0 * (BK * 2);
        for (int q = 0; q < 2 * I.k_slabs; ++q) {
          mbar_wait(&sm->b_empty[pb.stage], pb.phase ^ 1);
          mbar_expect_tx(&sm->b_full[pb.stage], bytes);
          bulk_g2s(b_base + pb.stage * RB_B_STAGE, src + (size_t)q * I.pack_stage_bytes, bytes, &sm->b_full[pb.stage]);
          pb.next();
        }
      }
    }
  } else if (warp == 6) {
    if (lane == 0) {
      Pipe pa(RB_NA), pb(RB_NB);
      for (int it = 0; it < p.n_items; ++it) {
        const BwdItem& I = p.item[it];
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(I.n_cta >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        const int buf = it & 1;
        const uint32_t tphase = (it >> 1) & 1;
        const uint32_t acc = tmem + buf * 256;
        mbar_wait(&sm->tmem_empty[buf], tphase ^ 1);
        tc_fence_after();
        for (int j = 0; j < I.k_slabs; ++j) {
          mbar_wait(&sm->a_full[pa.stage], pa.phase);
          const uint32_t a_hi = smem_u32(a_base + pa.stage * a_stage_bytes);
          const uint32_t a_lo = a_hi + BM * BK * 2;
          mbar_wait(&sm->b_full[pb.stage], pb.phase);
          tc_fence_after();
          uint32_t b_s = smem_u32(b_base + pb.stage * RB_B_STAGE);
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks) {
            const uint64_t bd = umma_desc(b_s + ks * 2 * LBO);
            umma_bf16(acc, umma_desc(a_hi + ks * 2 * LBO), bd, idesc, (j | ks) ? 1u : 0u);
            umma_bf16(acc, umma_desc(a_lo + ks * 2 * LBO), bd, idesc, 1u);
          }
          tc_commit(&sm->b_empty[pb.stage]);
          pb.next();
          mbar_wait(&sm->b_full[pb.stage], pb.phase);
          tc_fence_after();
          b_s = smem_u32(b_base + pb.stage * RB_B_STAGE);
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks)
            umma_bf16(acc, umma_desc(a_hi + ks * 2 * LBO), umma_desc(b_s + ks * 2 * LBO), idesc, 1u);
          tc_commit(&sm->b_empty[pb.stage]);
          pb.next();
          tc_commit(&sm->a_empty[pa.stage]);
          pa.next();
        }
        tc_commit(&sm->tmem_full[buf]);
      }
    }
  } else if (warp >= 8) {
    const int ewarp = warp & 3;
    const int lane_base = ewarp * 32;
    const int row = lane_base + lane;
    const int rg = row0 + row;
    const bool row_ok = rg < p.R;
    const bool issuer = (threadIdx.x == 8 * 32);
    int chunk = 0;
    for (int it = 0; it < p.n_items; ++it) {
      const BwdItem& I = p.item[it];
      if (I.rowstage >= 0) {
        // the row stage whose output this item's GEMM consumes; every earlier TMA store of this CTA has completed
        // (the previous item signalled) so the staging buffers are free to serve as scratch
        float* scratch = reinterpret_cast<float*>(e_base);
        if (p.C == 256) row_stage<8>(p.rs[I.rowstage], row0, p.R, ewarp, lane, scratch);
        else row_stage<4>(p.rs[I.rowstage], row0, p.R, ewarp, lane, scratch);
        if (issuer) mbar_arrive(&sm->ready[I.wait_ready]);
      }
      const int buf = it & 1;
      const uint32_t ephase = (it >> 1) & 1;
      const uint32_t acc = tmem + buf * 256 + ((uint32_t)lane_base << 16);
      mbar_wait(&sm->tmem_full[buf], ephase);
      tc_fence_after();
      for (int c0 = 0; c0 < I.n_cta; c0 += 32, ++chunk) {
        uint32_t r[32];
        ld_tmem_32b(acc + (uint32_t)c0, r);
        if (I.mode == 1) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 gt = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row_ok) gt = ldg4(I.gate + (size_t)rg * I.ldg + I.col0 + c0 + i);
            r[i] = gt.x > 0.f ? __float_as_uint(__uint_as_float(r[i]) * I.gscale) : 0u;
            r[i + 1] = gt.y > 0.f ? __float_as_uint(__uint_as_float(r[i + 1]) * I.gscale) : 0u;
            r[i + 2] = gt.z > 0.f ? __float_as_uint(__uint_as_float(r[i + 2]) * I.gscale) : 0u;
            r[i + 3] = gt.w > 0.f ? __float_as_uint(__uint_as_float(r[i + 3]) * I.gscale) : 0u;
          }
        }
        uint8_t* ebuf = e_base + (chunk & 1) * (BM * 128);
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<uint4*>(ebuf + row * 128 + ((i ^ (row & 7)) << 4)) = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
        fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (issuer) {
          tma_store_3d(&maps.o[I.o_map], I.col0 + c0, row0, 0, ebuf);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      tc_fence_before();
      mbar_arrive(&sm->tmem_empty[buf]);
      if (I.signal_ready >= 0 || (it + 1 < p.n_items && p.item[it + 1].rowstage >= 0)) {
        // the rows this item stored are read next (by the TMA producer or by a row stage): complete the stores first
        if (issuer) {
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
          asm volatile("fence.proxy.async;" ::: "memory");
          if (I.signal_ready >= 0) mbar_arrive(&sm->ready[I.signal_ready]);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 7) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

static inline bool make_bwd_map(PFN_encodeTiled encode, CUtensorMap* map, const float* base, int cols, int rows) {
  if ((reinterpret_cast<uintptr_t>(base) & 15) || cols % 32) return false;
  const cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 1};
  const cuuint64_t gstr[2] = {(cuuint64_t)cols * 4, (cuuint64_t)cols * 4 * (cuuint64_t)rows};
  const cuuint32_t box[3] = {32, (cuuint32_t)BM, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tc
}  // namespace sgc

extern "C" int sgc_rows_chain_bwd_tc(const sgc_rows_chain_bwd_args* args, void* stream) {
  using namespace sgc::tc;
  const sgc_rows_chain_bwd_args a = *args;
  const int C = a.C, F = a.F, R = a.R;
  if (R <= 0 || (C != 128 && C != 256) || F % 256 || F <= 0 || F > 512) return (int)cudaErrorInvalidValue;
  if (!a.gy || !a.p_w2_t || !a.p_w1_t || !a.p_wo_t || !a.pre1 || !a.mean1 || !a.rstd1 || !a.g1 || !a.pre2 || !a.mean2 ||
      !a.rstd2 || !a.g2 || !a.hdn || !a.gf || !a.gpre2 || !a.gh || !a.gx1 || !a.gout || !a.go2 || !a.partial1 || !a.partial2)
    return (int)cudaErrorInvalidValue;
  PFN_encodeTiled encode = get_encode_tiled();
  if (!encode) return (int)cudaErrorNotSupported;
  BwdMaps maps;
  if (!make_bwd_map(encode, &maps.a[0], a.gf, C, R) || !make_bwd_map(encode, &maps.a[1], a.gh, F, R) ||
      !make_bwd_map(encode, &maps.a[2], a.gout, C, R) || !make_bwd_map(encode, &maps.o[0], a.gh, F, R) ||
      !make_bwd_map(encode, &maps.o[1], a.gx1, C, R) || !make_bwd_map(encode, &maps.o[2], a.go2, C, R))
    return (int)cudaErrorInvalidValue;
  BwdParams p = {};
  p.R = R; p.C = C;
  // R0: LayerNorm 2 backward + dropout mask of the second FFN layer
  p.rs[0] = RowStage{a.gy, nullptr, a.pre2, a.mean2, a.rstd2, a.g2, a.mask2, nullptr, a.gpre2, a.gf, a.partial2, a.mscale2};
  // R2: (+ identity branch) LayerNorm 1 backward + attention dropout mask + view-count mask
  p.rs[1] = RowStage{a.gx1, a.gpre2, a.pre1, a.mean1, a.rstd1, a.g1, a.mask0, a.rowcount, nullptr, a.gout, a.partial1, a.mscale0};
  int n = 0;
  for (int c0 = 0; c0 < F; c0 += 256) {  // G1: ghdn = gf @ W_2 (packed W_2^T [F, C]), ReLU / dropout gate
    BwdItem& I = p.item[n++];
    I.wpack = (const uint8_t*)a.p_w2_t; I.pack_stage_bytes = (long long)F * BK * 2; I.gate = a.hdn; I.gscale = a.gscale1; I.ldg = F;
    I.a_map = 0; I.o_map = 0; I.k_slabs = C / BK; I.n_cta = 256; I.col0 = c0; I.mode = 1;
    I.rowstage = c0 == 0 ? 0 : -1; I.wait_ready = c0 == 0 ? 0 : -1; I.signal_ready = c0 + 256 >= F ? 1 : -1;
  }
  {  // G2: gx1 = gh @ W_1 (packed W_1^T [C, F])
    BwdItem& I = p.item[n++];
    I.wpack = (const uint8_t*)a.p_w1_t; I.pack_stage_bytes = (long long)C * BK * 2; I.gate = nullptr; I.gscale = 1.f; I.ldg = 0;
    I.a_map = 1; I.o_map = 1; I.k_slabs = F / BK; I.n_cta = C; I.col0 = 0; I.mode = 0;
    I.rowstage = -1; I.wait_ready = 1; I.signal_ready = -1;
  }
  {  // G3: go2 = gout @ W_o (packed W_o^T [C, C]); gout comes from row stage R2
    BwdItem& I = p.item[n++];
    I.wpack = (const uint8_t*)a.p_wo_t; I.pack_stage_bytes = (long long)C * BK * 2; I.gate = nullptr; I.gscale = 1.f; I.ldg = 0;
    I.a_map = 2; I.o_map = 2; I.k_slabs = C / BK; I.n_cta = C; I.col0 = 0; I.mode = 0;
    I.rowstage = 1; I.wait_ready = 2; I.signal_ready = -1;
  }
  p.n_items = n;
  const size_t smem = (size_t)RB_NF * BK * BM * 4 + (size_t)RB_NA * 2 * BM * BK * 2 + (size_t)RB_NB * RB_B_STAGE +
                      (size_t)RB_NE * BM * 128 + sizeof(SmemRB) + 64;
  cudaError_t e = cudaFuncSetAttribute(rows_chain_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int grid = (R + BM - 1) / BM;
  rows_chain_bwd_tc_kernel<<<grid, RB_THREADS, smem, (cudaStream_t)stream>>>(maps, p);
  SGC_CUDA_CHECK_LAST();
  return 0;
}
