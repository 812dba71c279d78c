// sgc_rows_chain_tc: the row-local tail of one encoder layer (encoder.py:310-338 after the cross-view attention of
// deformable_cross_attention.py:827-837) as ONE launch instead of three GEMM + three row-kernel launches:
//
//   stage 1   v = (o2 @ W_o^T + b_o) * mask0*s0 * [view count > 0]          pre1 = v      x1  = LayerNorm1(v)
//   stage 2   hdn = relu(x1 @ W_1^T + b_1) * mask1*s1
//   stage 3   v = (hdn @ W_2^T + b_2) * mask2*s2 + x1                       pre2 = v      y   = LayerNorm2(v)
//
// Everything is local to a voxel row, so a CTA takes one 128-row tile through all stages with the pipeline of
// rows_gemm_tc_kernel (TMA loads of the fp32 rows, bf16 hi/lo split in shared memory, packed weight slabs by bulk copies,
// fp32 accumulation in TMEM).  The CTA owns whole rows (n_cta = C for stages 1 and 3), so the LayerNorm statistics are
// computed by the epilogue thread of a row from its own accumulator (two passes like sgc_rowop_fwd).  Stage outputs that a
// later stage consumes (x1, hdn) leave through the swizzled staging tile + TMA store and are TMA-loaded back from L2 after
// the stores have completed (an mbarrier per stage orders the producer warp behind them); everything the backward saves
// (pre1, pre2, the row statistics, hdn) is written exactly as the unfused path writes it.
//
// STATUS (end of round 1): parity-green on the B200 against the separate launches (tests/test_gpu_rows_chain.py, 6 shapes)
// and at module level against the oracle with SGC_ROWS_CHAIN=1 (tests/test_gpu_path.py), but the GPU budget ran out before
// it could be benchmarked: functional.EncoderLayerRows.forward uses it only when SGC_ROWS_CHAIN=1 (default 0).
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/sgcdet_b200.h"

namespace sgc {
namespace tc {

constexpr int RC_NF = 4;        // fp32 staging stages (TMA, 16 KB each)
constexpr int RC_NA = 2;        // converted A stages (hi 8 KB + lo 8 KB)
constexpr int RC_NB = 4;        // weight stages (one of hi / lo per stage, up to 256 rows x 64 bytes)
constexpr int RC_NE = 2;        // epilogue staging buffers (128 rows x 32 floats)
constexpr int RC_THREADS = 384;
constexpr int RC_MAX_ITEMS = 4;
constexpr int RC_B_STAGE = 256 * BK * 2;   // bytes reserved per weight stage

struct SmemRC {
  uint64_t f_full[RC_NF], f_empty[RC_NF], a_full[RC_NA], a_empty[RC_NA], b_full[RC_NB], b_empty[RC_NB], tmem_full[2],
      tmem_empty[2], stage_ready[2];
  uint32_t tmem_base;
};

// One work item of a CTA: a GEMM over its 128-row tile producing n_cta output columns starting at col0.
struct ChainItem {
  const uint8_t* wpack;        // packed [rows, K] weight (sgc_pack_weight_tc image)
  long long pack_stage_bytes;  // rows * BK * 2
  const float* bias;           // [N of the stage]
  const unsigned char* mask;   // dropout keep-mask [R, ldm] or null
  float mscale;
  int ldm;
  int a_map, o_map;            // indices into the tensor-map arrays
  int k_slabs, n_cta, col0;
  int mode;                    // 0: bias, ReLU, mask -> store;  1: bias, mask, row count, residual, LayerNorm -> pre / stats / store
  int wait_stage;              // producer waits for stage_ready[wait_stage] before its first load (-1: none)
  int signal_stage;            // epilogue signals stage_ready[signal_stage] after this item's stores completed (-1: none)
  // mode 1 only
  const int* rowcount;         // [R] or null
  const float* residual;       // [R, n_cta] or null
  const float* gamma;
  const float* beta;
  float eps;
  float* pre;                  // [R, n_cta]
  float* mean;                 // [R]
  float* rstd;                 // [R]
};

struct ChainParams {
  ChainItem item[RC_MAX_ITEMS];
  int n_items, R;
};

struct ChainMaps {
  CUtensorMap a[3];   // o2, x1, hdn as operands
  CUtensorMap o[3];   // x1, hdn, y as outputs
};

__device__ __forceinline__ void ld_tmem_32(uint32_t taddr, uint32_t (&r)[32]) {
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

__global__ void __launch_bounds__(RC_THREADS, 1)
rows_chain_tc_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int f_stage_bytes = BK * BM * 4;
  constexpr int a_stage_bytes = 2 * BM * BK * 2;
  uint8_t* f_base = smem_raw;
  uint8_t* a_base = f_base + RC_NF * f_stage_bytes;
  uint8_t* b_base = a_base + RC_NA * a_stage_bytes;
  uint8_t* e_base = b_base + RC_NB * RC_B_STAGE;
  SmemRC* sm = reinterpret_cast<SmemRC*>(e_base + RC_NE * BM * 128);
  const int mt = blockIdx.x;          // this CTA's 128-row tile
  const int row0 = mt * BM;

  if (threadIdx.x == 0) {
    for (int i = 0; i < RC_NF; ++i) { mbar_init(&sm->f_full[i], 1); mbar_init(&sm->f_empty[i], 128); }
    for (int i = 0; i < RC_NA; ++i) { mbar_init(&sm->a_full[i], 128); mbar_init(&sm->a_empty[i], 1); }
    for (int i = 0; i < RC_NB; ++i) { mbar_init(&sm->b_full[i], 1); mbar_init(&sm->b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm->tmem_full[i], 1); mbar_init(&sm->tmem_empty[i], 128); mbar_init(&sm->stage_ready[i], 1);
    }
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
    // ===================== producer: TMA loads of the operand rows of every item =====================
    if (lane == 0) {
      Pipe pf(RC_NF);
      for (int it = 0; it < p.n_items; ++it) {
        const ChainItem& I = p.item[it];
        // operand rows written by an earlier stage of this CTA: wait until its TMA stores have completed
        if (I.wait_stage >= 0) mbar_wait(&sm->stage_ready[I.wait_stage], 0);
        for (int j = 0; j < I.k_slabs; ++j) {
          mbar_wait(&sm->f_empty[pf.stage], pf.phase ^ 1);
          mbar_expect_tx(&sm->f_full[pf.stage], (uint32_t)f_stage_bytes);
          tma_load_3d(f_base + pf.stage * f_stage_bytes, &maps.a[I.a_map], j * BK, row0, 0, &sm->f_full[pf.stage]);
          pf.next();
        }
      }
    }
  } else if (warp < 4) {
    // ===================== converters: swizzled fp32 [m][k] tile -> bf16 hi/lo core-matrix tiles =====================
    const int m = threadIdx.x;
    Pipe pa(RC_NA), pf(RC_NF);
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
    // ===================== weight producer =====================
    if (lane == 0) {
      Pipe pb(RC_NB);
      for (int it = 0; it < p.n_items; ++it) {
        const ChainItem& I = p.item[it];
        const uint32_t bytes = (uint32_t)I.n_cta * BK * 2;
        const uint8_t* src = I.wpack + (size_t)I.col0 * (BK * 2);
        for (int q = 0; q < 2 * I.k_slabs; ++q) {
          mbar_wait(&sm->b_empty[pb.stage], pb.phase ^ 1);
          mbar_expect_tx(&sm->b_full[pb.stage], bytes);
          bulk_g2s(b_base + pb.stage * RC_B_STAGE, src + (size_t)q * I.pack_stage_bytes, bytes, &sm->b_full[pb.stage]);
          pb.next();
        }
      }
    }
  } else if (warp == 6) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      Pipe pa(RC_NA), pb(RC_NB);
      for (int it = 0; it < p.n_items; ++it) {
        const ChainItem& I = p.item[it];
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
          uint32_t b_s = smem_u32(b_base + pb.stage * RC_B_STAGE);
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
          b_s = smem_u32(b_base + pb.stage * RC_B_STAGE);
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
    // ===================== epilogue =====================
    const int lane_base = (warp & 3) * 32;
    const int row = lane_base + lane;            // row of the tile == TMEM lane
    const int rg = row0 + row;                   // global row
    const bool row_ok = rg < p.R;
    const bool issuer = (threadIdx.x == 8 * 32);
    int chunk = 0;
    for (int it = 0; it < p.n_items; ++it) {
      const ChainItem& I = p.item[it];
      const int buf = it & 1;
      const uint32_t ephase = (it >> 1) & 1;
      const uint32_t acc = tmem + buf * 256 + ((uint32_t)lane_base << 16);
      const int n_cta = I.n_cta;
      mbar_wait(&sm->tmem_full[buf], ephase);
      tc_fence_after();
      float mu = 0.f, rs = 0.f;
      if (I.mode == 1) {
        // ---- pass 1: v = (acc + bias) * mask*mscale * [count > 0] + residual;  pre = v;  row sum
        const float has = (I.rowcount && row_ok) ? (__ldg(I.rowcount + rg) > 0 ? 1.f : 0.f) : 1.f;
        float s = 0.f;
        for (int c0 = 0; c0 < n_cta; c0 += 32) {
          uint32_t r[32];
          ld_tmem_32(acc + (uint32_t)c0, r);
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) + __ldg(I.bias + c0 + i);
          if (I.mask && row_ok) {
            const uint4 m0 = *reinterpret_cast<const uint4*>(I.mask + (size_t)rg * I.ldm + c0);
            const uint4 m1 = *reinterpret_cast<const uint4*>(I.mask + (size_t)rg * I.ldm + c0 + 16);
            const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= (((mw[i >> 2] >> ((i & 3) * 8)) & 0xffu) ? 1.f : 0.f) * I.mscale;
          }
          if (I.rowcount) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= has;
          }
          if (I.residual && row_ok) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 t4 = *reinterpret_cast<const float4*>(I.residual + (size_t)rg * n_cta + c0 + i);
              v[i] += t4.x; v[i + 1] += t4.y; v[i + 2] += t4.z; v[i + 3] += t4.w;
            }
          }
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<float4*>(I.pre + (size_t)rg * n_cta + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) s += v[i];
        }
        // the accumulator is not needed any more: passes 2 and 3 re-read `pre` (written by this very thread)
        tc_fence_before();
        mbar_arrive(&sm->tmem_empty[buf]);
        mu = s * (1.f / n_cta);
        float q = 0.f;
        if (row_ok) {
          for (int c = 0; c < n_cta; c += 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(I.pre + (size_t)rg * n_cta + c);
            const float d0 = t4.x - mu, d1 = t4.y - mu, d2 = t4.z - mu, d3 = t4.w - mu;
            q += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
          }
        }
        rs = rsqrtf(q * (1.f / n_cta) + I.eps);
        if (row_ok) { I.mean[rg] = mu; I.rstd[rg] = rs; }
      }
      for (int c0 = 0; c0 < n_cta; c0 += 32, ++chunk) {
        float v[32];
        if (I.mode == 1) {
          // ---- pass 3: y = (pre - mean) * rstd * gamma + beta
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row_ok) t4 = *reinterpret_cast<const float4*>(I.pre + (size_t)rg * n_cta + c0 + i);
            v[i] = t4.x; v[i + 1] = t4.y; v[i + 2] = t4.z; v[i + 3] = t4.w;
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = (v[i] - mu) * rs * __ldg(I.gamma + c0 + i) + __ldg(I.beta + c0 + i);
        } else {
          uint32_t r[32];
          ld_tmem_32(acc + (uint32_t)c0, r);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(__uint_as_float(r[i]) + __ldg(I.bias + I.col0 + c0 + i), 0.f);
          if (I.mask && row_ok) {
            const uint4 m0 = *reinterpret_cast<const uint4*>(I.mask + (size_t)rg * I.ldm + I.col0 + c0);
            const uint4 m1 = *reinterpret_cast<const uint4*>(I.mask + (size_t)rg * I.ldm + I.col0 + c0 + 16);
            const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= (((mw[i >> 2] >> ((i & 3) * 8)) & 0xffu) ? 1.f : 0.f) * I.mscale;
          }
        }
        uint8_t* ebuf = e_base + (chunk & 1) * (BM * 128);
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<uint4*>(ebuf + row * 128 + ((i ^ (row & 7)) << 4)) =
              make_uint4(__float_as_uint(v[4 * i]), __float_as_uint(v[4 * i + 1]), __float_as_uint(v[4 * i + 2]),
                         __float_as_uint(v[4 * i + 3]));
        fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (issuer) {
          tma_store_3d(&maps.o[I.o_map], I.col0 + c0, row0, 0, ebuf);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (I.mode == 0) {
        tc_fence_before();
        mbar_arrive(&sm->tmem_empty[buf]);
      }
      if (I.signal_stage >= 0) {
        // the rows this stage produced are operands of a later item: complete the stores, then release the producer
        if (issuer) {
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
          fence_proxy_async();
          mbar_arrive(&sm->stage_ready[I.signal_stage]);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // later generic-proxy reads (the residual) come after the stores
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

static inline bool make_chain_map(PFN_encodeTiled encode, CUtensorMap* map, const float* base, int cols, int rows) {
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

extern "C" int sgc_rows_chain_tc(const sgc_rows_chain_args* args, void* stream) {
  using namespace sgc::tc;
  const sgc_rows_chain_args a = *args;
  const int C = a.C, F = a.F, R = a.R;
  if (R <= 0 || (C != 128 && C != 256) || F % 256 || F <= 0 || F > 512) return (int)cudaErrorInvalidValue;
  if (!a.o2 || !a.p_wo || !a.p_w1 || !a.p_w2 || !a.bo || !a.b1 || !a.b2 || !a.g1 || !a.be1 || !a.g2 || !a.be2 || !a.x1 ||
      !a.pre1 || !a.mean1 || !a.rstd1 || !a.hdn || !a.y || !a.pre2 || !a.mean2 || !a.rstd2)
    return (int)cudaErrorInvalidValue;
  PFN_encodeTiled encode = get_encode_tiled();
  if (!encode) return (int)cudaErrorNotSupported;
  ChainMaps maps;
  if (!make_chain_map(encode, &maps.a[0], a.o2, C, R) || !make_chain_map(encode, &maps.a[1], a.x1, C, R) ||
      !make_chain_map(encode, &maps.a[2], a.hdn, F, R) || !make_chain_map(encode, &maps.o[0], a.x1, C, R) ||
      !make_chain_map(encode, &maps.o[1], a.hdn, F, R) || !make_chain_map(encode, &maps.o[2], a.y, C, R))
    return (int)cudaErrorInvalidValue;
  ChainParams p = {};
  p.R = R;
  int n = 0;
  {  // stage 1: W_o, LayerNorm 1
    ChainItem& I = p.item[n++];
    I.wpack = (const uint8_t*)a.p_wo; I.pack_stage_bytes = (long long)C * BK * 2; I.bias = a.bo;
    I.mask = a.mask0; I.mscale = a.mscale0; I.ldm = C;
    I.a_map = 0; I.o_map = 0; I.k_slabs = C / BK; I.n_cta = C; I.col0 = 0; I.mode = 1; I.wait_stage = -1; I.signal_stage = 0;
    I.rowcount = a.rowcount; I.residual = nullptr; I.gamma = a.g1; I.beta = a.be1; I.eps = a.eps1;
    I.pre = a.pre1; I.mean = a.mean1; I.rstd = a.rstd1;
  }
  for (int c0 = 0; c0 < F; c0 += 256) {  // stage 2: W_1, ReLU (256-column parts)
    ChainItem& I = p.item[n++];
    I.wpack = (const uint8_t*)a.p_w1; I.pack_stage_bytes = (long long)F * BK * 2; I.bias = a.b1;
    I.mask = a.mask1; I.mscale = a.mscale1; I.ldm = F;
    I.a_map = 1; I.o_map = 1; I.k_slabs = C / BK; I.n_cta = 256; I.col0 = c0; I.mode = 0;
    I.wait_stage = c0 == 0 ? 0 : -1; I.signal_stage = c0 + 256 >= F ? 1 : -1;
  }
  {  // stage 3: W_2, residual, LayerNorm 2
    ChainItem& I = p.item[n++];
    I.wpack = (const uint8_t*)a.p_w2; I.pack_stage_bytes = (long long)C * BK * 2; I.bias = a.b2;
    I.mask = a.mask2; I.mscale = a.mscale2; I.ldm = C;
    I.a_map = 2; I.o_map = 2; I.k_slabs = F / BK; I.n_cta = C; I.col0 = 0; I.mode = 1; I.wait_stage = 1; I.signal_stage = -1;
    I.rowcount = nullptr; I.residual = a.x1; I.gamma = a.g2; I.beta = a.be2; I.eps = a.eps2;
    I.pre = a.pre2; I.mean = a.mean2; I.rstd = a.rstd2;
  }
  p.n_items = n;
  const size_t smem = (size_t)RC_NF * BK * BM * 4 + (size_t)RC_NA * 2 * BM * BK * 2 + (size_t)RC_NB * RC_B_STAGE +
                      (size_t)RC_NE * BM * 128 + sizeof(SmemRC) + 64;
  cudaError_t e = cudaFuncSetAttribute(rows_chain_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int grid = (R + BM - 1) / BM;
  rows_chain_tc_kernel<<<grid, RC_THREADS, smem, (cudaStream_t)stream>>>(maps, p);
  SGC_CUDA_CHECK_LAST();
  return 0;
}
