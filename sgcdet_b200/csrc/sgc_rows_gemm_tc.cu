// sgc_rows_gemm_tc: the voxel-count GEMMs of one encoder layer on the 5th-gen tensor cores (tcgen05 / TMEM / TMA):
//
//   y[b, r, n] = sum_k x[b, r, k] * W_b[n, k]  (+ bias_b[n])        r < R voxel rows, b < B batches (attention heads)
//
// i.e. every nn.Linear applied to the selected voxel rows: output_proj and the in/out projections of attention_pooling
// (deformable_cross_attention.py:815-833, per head for the key / value projections), the two FFN layers
// (encoder.py:335-338), and the data gradients of all of them (the same product with the transposed weight).
//
// Same numerics as the pixel-count projection kernels (csrc/sgc_project_tc.cu): both operands are split into bf16
// hi + lo, hi*hi' + lo*hi' + hi*lo' accumulates in fp32 in TMEM (relative error ~1e-5).  The fp32 activations are read
// once with TMA (128B-swizzled [row][k] tiles) and split in shared memory; the weights come pre-packed
// (sgc_pack_weight_tc / sgc_prepare_weights kind 1) and are streamed with bulk copies; the fp32 result leaves through a
// 128B-swizzled staging tile and a TMA tensor store, which also clips the rows beyond R.  No bf16x3 operand image ever
// touches HBM and -- unlike the library's small-GEMM kernels -- no thread-block cluster is needed, so a launch starts as
// soon as ONE SM is free next to the persistent projection kernels.
//
// One CTA per SM, 12 warps, warp-specialised exactly like project_tc_kernel<true> + the forward kernel's epilogue:
//   warp 5: TMA producer (fp32 x tiles)      warps 0-3: converters (fp32 -> bf16 hi/lo core-matrix tiles)
//   warp 4: weight-slab producer             warp 6: MMA issuer          warp 7: TMEM alloc
//   warps 8-11: epilogue (tcgen05.ld -> +bias -> swizzled smem -> TMA store)
// Work item = (batch b, 128-row tile, column part of n_cta columns); column parts of the same row tile are adjacent in
// the work order so that their x tile is shared through L2.  Two TMEM accumulators: the MMAs of work item i+1 overlap
// the epilogue of item i.
#include <cuda.h>
#include <cstdlib>
#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/sgcdet_b200.h"

namespace sgc {
namespace tc {

constexpr int RG_NF = 3;        // fp32 staging stages filled by TMA (BM x BK floats = 16 KB each)
constexpr int RG_NA = 4;        // converted A stages (hi 8 KB + lo 8 KB): the converters run up to 4 k-slabs ahead of the MMAs
constexpr int RG_NB = 8;        // weight stages at most (one of hi / lo per stage, n_cta * 64 bytes): 8 up to 128 columns per CTA, else 4
constexpr int RG_NE = 2;        // epilogue staging buffers (BM rows x 32 floats)
constexpr int RG_THREADS = 384;

struct SmemRG {
  uint64_t f_full[RG_NF], f_empty[RG_NF], a_full[RG_NA], a_empty[RG_NA], b_full[RG_NB], b_empty[RG_NB], tmem_full[2],
      tmem_empty[2];
  uint32_t tmem_base;
  float bias[256];            // the work item's bias slice (read per 32-column chunk of the epilogue)
};

struct RowsGemmParams {
  const uint8_t* wpack;
  const float* bias;
  long long pack_stage_bytes;   // bytes between consecutive (k-slab, hi/lo) stages of one packed matrix = pack_rows*BK*2
  long long pack_batch_bytes;   // bytes between the packed matrices of consecutive batches (0: one matrix for all)
  int pack_batch_rows;          // row offset of batch b inside a packed matrix: b * pack_batch_rows
  int bias_batch;               // bias of batch b starts at bias + b * bias_batch
  int m_tiles, nsplit, works, k_slabs, n_cta, tmem_cols, nb;   // nb = weight stages in use (<= RG_NB)
  int dual;                     // 0: the epilogue warps never convert (SGC_ROWS_DUAL=0, for A/B measurements)
  int a_swap, o_swap;           // tensor-map coordinate order: 0 = (col, row, batch), 1 = (col, batch, row)
  int rows;                     // R: voxel rows of one batch (TMA clips the stores; fully clipped 32-row slices are skipped)
  int a_bcast;                  // 1: every batch reads the SAME activation rows (batch coordinate 0)
  int kpb;                      // > 0: the reduction runs over the batches of A too, kpb k-slabs per batch (K-concatenation)
  long long* dbg;               // optional [16] clock64 stamps of CTA 0 (sgc_rows_gemm_tc_set_debug): where a launch's time goes
};

#define RG_STAMP(i)                                                         \
  do {                                                                      \
    if (p.dbg && blockIdx.x == 0) p.dbg[i] = clock64();                     \
  } while (0)

__global__ void __launch_bounds__(RG_THREADS, 1)
rows_gemm_tc_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap omap,
                    const __grid_constant__ RowsGemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int f_stage_bytes = BK * BM * 4;
  constexpr int a_stage_bytes = 2 * BM * BK * 2;
  const int b_stage_bytes = p.n_cta * BK * 2;
  uint8_t* f_base = smem_raw;
  uint8_t* a_base = f_base + RG_NF * f_stage_bytes;
  uint8_t* b_base = a_base + RG_NA * a_stage_bytes;
  uint8_t* e_base = b_base + p.nb * b_stage_bytes;    // 1 KB aligned: every stage size is a multiple of 2 KB
  SmemRG* sm = reinterpret_cast<SmemRG*>(e_base + RG_NE * BM * 128);
  const int n_cta = p.n_cta, k_slabs = p.k_slabs, works = p.works;
  // One work item per CTA (the latency-bound launches of the per-voxel chain): the four epilogue warps, idle until the
  // accumulator is complete, convert every second k-slab.  Measured with the clock stamps below (tools/rows_gemm_timeline.py):
  // 4 converter warps needed ~0.55 us per slab = 4.2 of the 8.4 us a CTA lived at K = 256.
  const bool dual = p.dual && works <= (int)gridDim.x;

  // ---- producers keep their position in (work item, k-slab) order so that they can start BEFORE the CTA-wide sync ----
  int a_w = blockIdx.x, a_j = 0;
  Pipe a_pf(RG_NF);
  auto produce_a = [&](int limit) {   // fp32 activation tiles [128 rows][32 k] by TMA
    for (int cnt = 0; a_w < works && cnt < limit; ++cnt) {
      const int t = a_w / p.nsplit, mt = t % p.m_tiles, b = t / p.m_tiles;
      mbar_wait(&sm->f_empty[a_pf.stage], a_pf.phase ^ 1);
      mbar_expect_tx(&sm->f_full[a_pf.stage], (uint32_t)f_stage_bytes);
      int col = a_j * BK, bc = p.a_bcast ? 0 : b;
      if (p.kpb) { bc = a_j / p.kpb; col = (a_j - bc * p.kpb) * BK; }
      if (p.a_swap) tma_load_3d(f_base + a_pf.stage * f_stage_bytes, &amap, col, bc, mt * BM, &sm->f_full[a_pf.stage]);
      else tma_load_3d(f_base + a_pf.stage * f_stage_bytes, &amap, col, mt * BM, bc, &sm->f_full[a_pf.stage]);
      if (a_j == 0 && a_w == (int)blockIdx.x) RG_STAMP(3);
      a_pf.next();
      if (++a_j == k_slabs) { a_j = 0; a_w += gridDim.x; }
    }
  };
  int b_w = blockIdx.x, b_q = 0;
  Pipe b_pb(p.nb);
  auto produce_b = [&](int limit) {   // n_cta rows of every packed (slab, hi/lo) weight stage by bulk copy; q = 2*slab + (0 hi, 1 lo)
    for (int cnt = 0; b_w < works && cnt < limit; ++cnt) {
      const int np = b_w % p.nsplit, b = (b_w / p.nsplit) / p.m_tiles;
      const uint8_t* src = p.wpack + (size_t)b * p.pack_batch_bytes +
                           ((size_t)b * p.pack_batch_rows + (size_t)np * n_cta) * (BK * 2);
      mbar_wait(&sm->b_empty[b_pb.stage], b_pb.phase ^ 1);
      mbar_expect_tx(&sm->b_full[b_pb.stage], (uint32_t)b_stage_bytes);
      bulk_g2s(b_base + b_pb.stage * b_stage_bytes, src + (size_t)b_q * p.pack_stage_bytes, (uint32_t)b_stage_bytes,
               &sm->b_full[b_pb.stage]);
      b_pb.next();
      if (++b_q == 2 * k_slabs) { b_q = 0; b_w += gridDim.x; }
    }
  };

  pdl_launch_dependents();   // programmatic dependent launch (common.cuh): the successor may be scheduled from here on
  if (threadIdx.x == 0) {
    RG_STAMP(0);
    for (int i = 0; i < RG_NF; ++i) { mbar_init(&sm->f_full[i], 1); mbar_init(&sm->f_empty[i], 128); }
    for (int i = 0; i < RG_NA; ++i) { mbar_init(&sm->a_full[i], 128); mbar_init(&sm->a_empty[i], 1); }
    for (int i = 0; i < RG_NB; ++i) { mbar_init(&sm->b_full[i], 1); mbar_init(&sm->b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sm->tmem_full[i], 1); mbar_init(&sm->tmem_empty[i], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    RG_STAMP(1);
  } else if (threadIdx.x == 5 * 32) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&amap) : "memory");   // ~0.5 us to fetch: started at kernel entry
  } else if (threadIdx.x == 8 * 32) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&omap) : "memory");
  }
  __syncthreads();           // barriers are live; producers and converters go ahead, they do not need the TMEM address
  uint32_t tmem = 0;
  if (warp >= 6) {
    // TMEM allocation concerns the MMA issuer, the allocating warp and the epilogue warps only (192 threads)
    if (warp == 7) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_base)),
                   "r"(p.tmem_cols));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    asm volatile("bar.sync 2, 192;" ::: "memory");
    tc_fence_after();
    tmem = sm->tmem_base;
    if (threadIdx.x == 6 * 32) RG_STAMP(2);
  }
  // everything below reads the predecessor's output (or is ordered behind threads that do)
  pdl_wait();

  // ===================== converters: swizzled fp32 [m][k] tile -> bf16 hi/lo core-matrix tiles =====================
  // group 0 = warps 0-3, group 1 = warps 8-11 (dual mode only): slab s of the CTA's running slab count goes to group s & 1
  auto convert_item = [&](int g, int w, Pipe& pa, Pipe& pf, int& s) {
    const int m = threadIdx.x & 127;  // row of the tile
    for (int j = 0; j < k_slabs; ++j, ++s) {
      if (!dual || (s & 1) == g) {
        mbar_wait(&sm->f_full[pf.stage], pf.phase);
        if (threadIdx.x == 0 && j == 0 && w == (int)blockIdx.x) RG_STAMP(4);
        float x[BK];
        const uint8_t* rowp = f_base + pf.stage * f_stage_bytes + m * 128;
#pragma unroll
        for (int i = 0; i < 8; ++i) {   // 16-byte chunk i of the 128-byte row sits at chunk (i ^ (m & 7))
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
        // released only after every staged value has been consumed by the conversion (see project_tc_kernel)
        mbar_arrive(&sm->f_empty[pf.stage]);
        fence_proxy_async();
        mbar_arrive(&sm->a_full[pa.stage]);
        if (threadIdx.x == 0 && w == (int)blockIdx.x) { if (j == 0) RG_STAMP(5); if (j >= k_slabs - 2) RG_STAMP(6); }
      }
      pf.next();
      pa.next();
    }
  };

  if (warp == 5) {
    if (lane == 0) produce_a(0x7fffffff);
  } else if (warp < 4) {
    Pipe pa(RG_NA), pf(RG_NF);
    int s = 0;
    for (int w = blockIdx.x; w < works; w += gridDim.x) convert_item(0, w, pa, pf, s);
  } else if (warp == 4) {
    if (lane == 0) produce_b(0x7fffffff);
  } else if (warp == 6) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_cta >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      Pipe pa(RG_NA), pb(p.nb);
      int it = 0;
      for (int w = blockIdx.x; w < works; w += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t tphase = (it >> 1) & 1;
        const uint32_t acc = tmem + buf * n_cta;
        mbar_wait(&sm->tmem_empty[buf], tphase ^ 1);
        tc_fence_after();
        for (int j = 0; j < k_slabs; ++j) {
          mbar_wait(&sm->a_full[pa.stage], pa.phase);
          const uint32_t a_hi = smem_u32(a_base + pa.stage * a_stage_bytes);
          const uint32_t a_lo = a_hi + BM * BK * 2;
          // --- W_hi stage: hi*hi + lo*hi
          mbar_wait(&sm->b_full[pb.stage], pb.phase);
          tc_fence_after();
          uint32_t b_s = smem_u32(b_base + pb.stage * b_stage_bytes);
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks) {
            const uint64_t bd = umma_desc(b_s + ks * 2 * LBO);
            umma_bf16(acc, umma_desc(a_hi + ks * 2 * LBO), bd, idesc, (j | ks) ? 1u : 0u);
            umma_bf16(acc, umma_desc(a_lo + ks * 2 * LBO), bd, idesc, 1u);
          }
          tc_commit(&sm->b_empty[pb.stage]);
          pb.next();
          // --- W_lo stage: hi*lo
          mbar_wait(&sm->b_full[pb.stage], pb.phase);
          tc_fence_after();
          b_s = smem_u32(b_base + pb.stage * b_stage_bytes);
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks)
            umma_bf16(acc, umma_desc(a_hi + ks * 2 * LBO), umma_desc(b_s + ks * 2 * LBO), idesc, 1u);
          tc_commit(&sm->b_empty[pb.stage]);
          pb.next();
          tc_commit(&sm->a_empty[pa.stage]);
          if (w == (int)blockIdx.x && j == 0) RG_STAMP(7);
          pa.next();
        }
        tc_commit(&sm->tmem_full[buf]);
        if (w == (int)blockIdx.x) RG_STAMP(8);
      }
    }
  } else if (warp >= 8) {
    // ===================== epilogue: TMEM -> registers (+bias) -> swizzled smem -> TMA store =====================
    // Every warp owns the 32 rows of its TMEM lane quarter end to end: its own slice of the staging tile, its own TMA stores
    // (box = 32 columns x 32 rows) and bulk groups -- no CTA-wide barrier and no single issuing thread per 32-column chunk.
    const int lane_base = (warp & 3) * 32;
    const int row = lane_base + lane;           // row of the tile == TMEM lane
    const bool stamp = (threadIdx.x == 8 * 32);
    int chunk = 0;                               // running chunk counter -> staging buffer parity
    int it = 0;
    Pipe pa(RG_NA), pf(RG_NF);
    int s = 0;
    for (int w = blockIdx.x; w < works; w += gridDim.x, ++it) {
      const int np = w % p.nsplit, t = w / p.nsplit, mt = t % p.m_tiles, b = t / p.m_tiles;
      const float* bias = p.bias ? p.bias + (size_t)b * p.bias_batch + np * n_cta : nullptr;
      if (bias) {
        // the bias slice goes to shared memory while the accumulator is still being formed: read from global per chunk it
        // cost a load round trip (~0.5 us) per 32 columns behind tcgen05.wait::ld
        asm volatile("bar.sync 1, 128;" ::: "memory");     // the previous item's chunks have read their bias
        for (int c = threadIdx.x - 8 * 32; c < n_cta; c += 128) sm->bias[c] = __ldg(bias + c);
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      if (dual) convert_item(1, w, pa, pf, s);
      const int buf = it & 1;
      const uint32_t ephase = (it >> 1) & 1;
      const uint32_t acc = tmem + buf * n_cta;
      const bool rows_live = mt * BM + lane_base < p.rows;   // a slice entirely behind the last row stores nothing
      mbar_wait(&sm->tmem_full[buf], ephase);
      tc_fence_after();
      if (stamp && w == (int)blockIdx.x) RG_STAMP(9);
      for (int c0 = 0; c0 < n_cta; c0 += 32, ++chunk) {
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
        if (bias) {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + sm->bias[c0 + i]);
        }
        uint8_t* ebuf = e_base + (chunk & 1) * (BM * 128);
        // the store that read this warp's slice of the buffer two chunks ago must have finished reading it
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i)  // 16-byte chunk i of the 128-byte row, CU_TENSOR_MAP_SWIZZLE_128B pattern
          *reinterpret_cast<uint4*>(ebuf + row * 128 + ((i ^ (row & 7)) << 4)) =
              make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (rows_live) {
            if (p.o_swap) tma_store_3d(&omap, np * n_cta + c0, b, mt * BM + lane_base, ebuf + lane_base * 128);
            else tma_store_3d(&omap, np * n_cta + c0, mt * BM + lane_base, b, ebuf + lane_base * 128);
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      tc_fence_before();
      mbar_arrive(&sm->tmem_empty[buf]);
      if (stamp && w == (int)blockIdx.x) RG_STAMP(10);
    }
    // the staging buffers only have to outlive the stores' READS; the writes are complete (and visible) at kernel end
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (stamp) RG_STAMP(11);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 7) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols));
    if (lane == 0) RG_STAMP(12);
  }
}

// [cols, rows, batch] fp32 view with a [32, box_rows, 1] box (or [cols, batch, rows] with a [32, 1, box_rows] box when the
// batch stride is the smaller one, e.g. the heads of a [R, H*dh] matrix): strides stay ascending for the descriptor.
// box_rows = 128 for the activation tiles, 32 for the per-warp stores of the epilogue.
static inline bool make_rows_map(PFN_encodeTiled encode, CUtensorMap* map, const float* base, int cols, int rows, int batch,
                                 long long ld, long long batch_stride, int* swapped, int box_rows = BM) {
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * 4) % 16 || ld < cols) return false;
  if (batch > 1 && ((batch_stride * 4) % 16 || batch_stride <= 0)) return false;
  const bool swap = batch > 1 && batch_stride < ld;
  *swapped = swap ? 1 : 0;
  const cuuint64_t bs = batch > 1 ? (cuuint64_t)batch_stride * 4 : (cuuint64_t)ld * 4 * (cuuint64_t)rows;
  cuuint64_t gdim[3], gstr[2];
  cuuint32_t box[3];
  const cuuint32_t estr[3] = {1, 1, 1};
  gdim[0] = (cuuint64_t)cols;
  box[0] = 32;
  if (swap) {
    gdim[1] = (cuuint64_t)batch; gdim[2] = (cuuint64_t)rows;
    gstr[0] = bs; gstr[1] = (cuuint64_t)ld * 4;
    box[1] = 1; box[2] = (cuuint32_t)box_rows;
  } else {
    gdim[1] = (cuuint64_t)rows; gdim[2] = (cuuint64_t)batch;
    gstr[0] = (cuuint64_t)ld * 4; gstr[1] = bs;
    box[1] = (cuuint32_t)box_rows; box[2] = 1;
  }
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tc
}  // namespace sgc

// Columns per CTA chosen by sgc_rows_gemm_tc for (R, N, B) when n_cta == 0: the widest of {256,128,64,32} dividing N that
// still yields enough work items to spread over the SMs (the problems are latency-, not throughput-bound).
extern "C" int sgc_rows_gemm_tc_auto_ncta(int R, int N, int B) {
  using namespace sgc::tc;
  if (R <= 0 || N <= 0 || N % 32 || B <= 0) return 0;
  const int m_tiles = (R + BM - 1) / BM;
  int n_cta = 256;
  // Problems of at most 4 row tiles (the coarsest level) settle for 16 work items: in the forward their chain runs beside the
  // persistent projection kernel of the finest level, which leaves 16 SMs free, and a CTA's lifetime hardly depends on its
  // column count (the conversion of the shared A tile dominates).  Measured (session V; split as far as possible / 4 / 8 / 16 work items):
  // 597.7 / 605.1 / 609.9 / 613.5 volumes/s; extending it to 8 row tiles (the middle level) gave 611.9.
  constexpr int small_works = 16, small_tiles = 4;
  if (m_tiles <= small_tiles) {
    while (n_cta > 32 && (N % n_cta || (long long)m_tiles * B * (N / n_cta) < small_works)) n_cta >>= 1;
    while (N % n_cta) n_cta >>= 1;
    return n_cta;
  }
  while (n_cta > 32 && (N % n_cta || (long long)m_tiles * B * (N / n_cta) * 2 <= 148)) n_cta >>= 1;
  while (N % n_cta) n_cta >>= 1;
  return n_cta;
}

static long long* g_rows_gemm_dbg = nullptr;
// Debug aid: `stamps` = device buffer of 16 int64 (or NULL to switch off): CTA 0 of every following sgc_rows_gemm_tc launch
// records clock64() at 13 points of its life (0 entry, 1 barriers initialised, 2 TMEM allocated + CTA sync, 3 first TMA issued,
// 4 first tile landed, 5 first / 6 last k-slab converted, 7 first slab's MMAs issued, 8 accumulator committed, 9 epilogue sees
// it, 10 last store issued, 11 stores drained, 12 TMEM freed).  tools/rows_gemm_timeline.py prints the differences.
extern "C" int sgc_rows_gemm_tc_set_debug(long long* stamps) {
  g_rows_gemm_dbg = stamps;
  return 0;
}

static int rows_gemm_tc_launch(const float* x, long long ldx, long long batch_x, int R, int K, int B, const void* wpack,
                               int pack_rows, long long pack_batch_elems, int pack_batch_rows, const float* bias,
                               int bias_batch, int N, float* y, long long ldy, long long batch_y, int n_cta, int a_mode,
                               int a_batches, void* stream) {
  using namespace sgc::tc;
  if (R <= 0 || B <= 0 || K <= 0 || K % BK || N <= 0 || N % 32 || !x || !y || !wpack) return (int)cudaErrorInvalidValue;
  if (a_mode < 0 || a_mode > 2) return (int)cudaErrorInvalidValue;
  if (a_mode == 2 && (B != 1 || a_batches <= 0 || K % a_batches || (K / a_batches) % BK)) return (int)cudaErrorInvalidValue;
  if (n_cta == 0) n_cta = sgc_rows_gemm_tc_auto_ncta(R, N, B);
  if (n_cta < 32 || n_cta > 256 || (n_cta & (n_cta - 1)) || N % n_cta) return (int)cudaErrorInvalidValue;
  if (pack_rows < N || pack_rows % 8 || pack_batch_rows % 8 || (reinterpret_cast<uintptr_t>(wpack) & 15) ||
      (pack_batch_elems * 2) % 16 || pack_batch_elems < 0 || pack_batch_rows < 0)
    return (int)cudaErrorInvalidValue;
  if (pack_batch_elems == 0 && (long long)(B - 1) * pack_batch_rows + N > pack_rows) return (int)cudaErrorInvalidValue;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  PFN_encodeTiled encode = get_encode_tiled();
  if (!encode) return (int)cudaErrorNotSupported;
  CUtensorMap amap, omap;
  RowsGemmParams p;
  p.a_bcast = a_mode == 1 ? 1 : 0;
  p.kpb = a_mode == 2 ? (K / a_batches) / BK : 0;
  if (a_mode == 1) {
    if (!make_rows_map(encode, &amap, x, K, R, 1, ldx, 0, &p.a_swap)) return (int)cudaErrorInvalidValue;
  } else if (a_mode == 2) {
    if (!make_rows_map(encode, &amap, x, K / a_batches, R, a_batches, ldx, batch_x, &p.a_swap)) return (int)cudaErrorInvalidValue;
  } else if (!make_rows_map(encode, &amap, x, K, R, B, ldx, batch_x, &p.a_swap)) return (int)cudaErrorInvalidValue;
  if (!make_rows_map(encode, &omap, y, N, R, B, ldy, batch_y, &p.o_swap, 32)) return (int)cudaErrorInvalidValue;
  p.wpack = reinterpret_cast<const uint8_t*>(wpack);
  p.bias = bias;
  p.pack_stage_bytes = (long long)pack_rows * BK * 2;
  p.pack_batch_bytes = pack_batch_elems * 2;
  p.pack_batch_rows = pack_batch_rows;
  p.bias_batch = bias_batch;
  p.m_tiles = (R + BM - 1) / BM;
  p.nsplit = N / n_cta;
  p.works = p.m_tiles * p.nsplit * B;
  p.k_slabs = K / BK;
  p.n_cta = n_cta;
  p.rows = R;
  p.dbg = g_rows_gemm_dbg;
  // dual converter groups: -25 % conversion time per k-slab in isolation, but an unexplained launch failure after a few hundred
  // graph replays of the whole step (with 4 converted stages): off
  static const int env_dual = getenv("SGC_ROWS_DUAL") ? atoi(getenv("SGC_ROWS_DUAL")) : 0;
  static const int env_nb = getenv("SGC_ROWS_NB") ? atoi(getenv("SGC_ROWS_NB")) : RG_NB;
  p.dual = env_dual;
  p.nb = n_cta <= 128 ? (env_nb >= 2 && env_nb <= RG_NB ? env_nb : RG_NB) : 4;
  p.tmem_cols = 2 * n_cta < 32 ? 32 : 2 * n_cta;
  const size_t smem = (size_t)RG_NF * BK * BM * 4 + (size_t)RG_NA * 2 * BM * BK * 2 + (size_t)p.nb * n_cta * BK * 2 +
                      (size_t)RG_NE * BM * 128 + sizeof(SmemRG) + 64;
  // the opt-in limit is raised once to the widest configuration (n_cta = 256): launches captured into a CUDA graph with
  // different column splits must not depend on which of them set the attribute last
  const size_t smem_max = (size_t)RG_NF * BK * BM * 4 + (size_t)RG_NA * 2 * BM * BK * 2 + (size_t)4 * 256 * BK * 2 +
                          (size_t)RG_NE * BM * 128 + sizeof(SmemRG) + 64;
  cudaError_t e = cudaFuncSetAttribute(rows_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
  if (e != cudaSuccess) return (int)e;
  const int grid = p.works < sms ? p.works : sms;
  sgc::launch_chain(rows_gemm_tc_kernel, dim3(grid), dim3(RG_THREADS), smem, (cudaStream_t)stream, amap, omap, p);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

extern "C" int sgc_rows_gemm_tc(const float* x, long long ldx, long long batch_x, int R, int K, int B, const void* wpack,
                                int pack_rows, long long pack_batch_elems, int pack_batch_rows, const float* bias,
                                int bias_batch, int N, float* y, long long ldy, long long batch_y, int n_cta, void* stream) {
  return rows_gemm_tc_launch(x, ldx, batch_x, R, K, B, wpack, pack_rows, pack_batch_elems, pack_batch_rows, bias, bias_batch, N, y,
                             ldy, batch_y, n_cta, 0, 0, stream);
}

// a_mode 1: the B batches share ONE activation matrix x [R, K] (batch_x ignored) -- per-head products whose heads are narrower
// than a k-slab, with the head's weights zero-extended to all K columns.  a_mode 2: B == 1 and the reduction also runs over
// the a_batches matrices x[a][R, K / a_batches] (batch stride batch_x): y = sum_a x[a] W_a^T with wpack the packed
// [N, K] = [W_0 | W_1 | ...] -- the per-head output products of narrow heads as one K-concatenated GEMM.
extern "C" int sgc_rows_gemm_tc_ex(const float* x, long long ldx, long long batch_x, int R, int K, int B, const void* wpack,
                                   int pack_rows, long long pack_batch_elems, int pack_batch_rows, const float* bias,
                                   int bias_batch, int N, float* y, long long ldy, long long batch_y, int n_cta, int a_mode,
                                   int a_batches, void* stream) {
  return rows_gemm_tc_launch(x, ldx, batch_x, R, K, B, wpack, pack_rows, pack_batch_elems, pack_batch_rows, bias, bias_batch, N, y,
                             ldy, batch_y, n_cta, a_mode, a_batches, stream);
}

// =====================================================================================================================
// sgc_rows_wgrad_tc / sgc_rows_wgrad_group_tc: the weight gradients of the same layers (reduction over the voxel rows):
//
//     out_b[m, n] = scale * sum_r A_b[r, m] * B_b[r, n]          bias[m] = sum_r A[r, m]   (or over B's columns)
//
// For y = x W^T + b with upstream gradient g:  gW = g^T x  (A = g, B = x, bias gradient = column sums of A).  The per-head
// key / value weights use the transposed product (A = t[h] / gqt[h] with M = C, B = the head's dh columns of go / qv) and
// the reduce kernel writes the result transposed, because a UMMA tile needs M = 128 rows.
// Both operands are fp32 in HBM, TMA-loaded as [32 rows][cols] tiles and split to bf16 hi/lo in shared memory (the
// "[k][m]" converter of wgrad_tc_kernel for both).  Split-K over the rows: CTA (m-tile, k-chunk, column part, batch)
// accumulates its rows in TMEM and writes a partial tile; the reduce kernel sums the partials in a fixed order
// (deterministic), applies `scale` and writes through arbitrary output strides (so gradients land directly inside
// in_proj_weight's [3C,C] gradient).  Column sums are accumulated by the converter threads per CTA and reduced alike.
//
// Round 2: the kernel takes a TABLE of jobs, so ALL weight gradients of an encoder layer (output_proj, the three
// in-projections incl. the per-head key / value products, out_proj, both FFN layers: 7 jobs) are ONE launch + ONE reduce
// launch instead of 14 + 14.  The single launches were latency-bound (TMEM allocation, pipeline fill, 128 KB partial
// tile per CTA after as little as 4 row slabs: ~0.5 ms of kernel time per step for ~35 GFLOP); in the grouped launch the
// CTAs of all jobs fill the SMs together, so each CTA can own a long run of rows (few k-chunks, 8x less partial traffic).
namespace sgc {
namespace tc {

constexpr int RW_THREADS = 480;  // warps 0-3: A converters (+ epilogue), 4-11: B converters, 12: TMA, 13: MMA, 14: TMEM
constexpr int RW_CONV = 384;     // converter threads (arrivals on f_empty / op_full)
constexpr int RW_ST = 2;          // converted-operand stages
constexpr int RW_MAX_F = 6;       // fp32 staging stages filled by TMA: as many as the job's tile width leaves room for
constexpr int RW_MAX_JOBS = 8;

struct SmemRW {
  uint64_t f_full[RW_MAX_F], f_empty[RW_MAX_F], op_full[RW_ST], op_empty[RW_ST], tmem_full;
  uint32_t tmem_base;
};

struct alignas(64) RowsWgradJob {
  CUtensorMap amap, bmap;
  float* partial;        // [kch][B][M][N]
  float* bias_partial;   // [kch][B][M or N]
  int M, N, n_cta, m_tiles, n_parts, nb, kch, slabs_per_cta, total_slabs;
  int a_swap, b_swap, bias_from;
  int cta_begin;         // first CTA of the job in the grouped grid
};

struct RowsWgradJobs {
  RowsWgradJob job[RW_MAX_JOBS];
  int njobs;
  int stage_bytes;       // shared memory available for the operand stages (the same for every CTA of the launch)
};

__global__ void __launch_bounds__(RW_THREADS, 1)
rows_wgrad_tc_kernel(const __grid_constant__ RowsWgradJobs jobs) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int ji = 0;
  for (int j = 1; j < jobs.njobs; ++j)
    if ((int)blockIdx.x >= jobs.job[j].cta_begin) ji = j;
  const RowsWgradJob& p = jobs.job[ji];
  const int n_cta = p.n_cta;
  const int a_src = BK * BM * 4;              // 16 KB  [32 r][128 m] fp32
  const int b_src = BK * n_cta * 4;           // [32 r][n_cta] fp32
  const int f_stage = a_src + b_src;
  const int a_op = 2 * BM * BK * 2;           // hi + lo, 16 KB
  const int b_op = 2 * n_cta * BK * 2;        // hi + lo
  const int op_stage = a_op + b_op;
  // the TMA round trip (operands mostly come from HBM: they were produced many kernels ago) is hidden by as many fp32
  // staging stages as fit beside the two converted-operand stages: 2 for 256-column tiles, 4 for 128, 6 for the heads' 16 / 32
  int nf = (jobs.stage_bytes - RW_ST * op_stage) / f_stage;
  nf = nf > RW_MAX_F ? RW_MAX_F : nf;
  uint8_t* op_base = smem_raw;
  uint8_t* f_base = op_base + RW_ST * op_stage;
  SmemRW* sm = reinterpret_cast<SmemRW*>(smem_raw + jobs.stage_bytes);

  int w = (int)blockIdx.x - p.cta_begin;
  const int mt = w % p.m_tiles; w /= p.m_tiles;
  const int kc = w % p.kch; w /= p.kch;
  const int np = w % p.n_parts;
  const int bt = w / p.n_parts, nb = p.nb;
  const int s_begin = kc * p.slabs_per_cta;
  const int s_end = min(p.total_slabs, s_begin + p.slabs_per_cta);
  const int n_slabs = max(0, s_end - s_begin);

  if (threadIdx.x == 0) {
    for (int i = 0; i < RW_MAX_F; ++i) { mbar_init(&sm->f_full[i], 1); mbar_init(&sm->f_empty[i], RW_CONV); }
    for (int i = 0; i < RW_ST; ++i) { mbar_init(&sm->op_full[i], RW_CONV); mbar_init(&sm->op_empty[i], 1); }
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
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.amap) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.bmap) : "memory");
      Pipe pf(nf);
      for (int i = 0; i < n_slabs; ++i) {
        const int r0 = (s_begin + i) * BK;
        mbar_wait(&sm->f_empty[pf.stage], pf.phase ^ 1);
        mbar_expect_tx(&sm->f_full[pf.stage], (uint32_t)f_stage);
        uint8_t* st = f_base + pf.stage * f_stage;
        if (p.a_swap) tma_load_3d(st, &p.amap, mt * BM, bt, r0, &sm->f_full[pf.stage]);
        else tma_load_3d(st, &p.amap, mt * BM, r0, bt, &sm->f_full[pf.stage]);
        if (p.b_swap) tma_load_3d(st + a_src, &p.bmap, np * n_cta, bt, r0, &sm->f_full[pf.stage]);
        else tma_load_3d(st + a_src, &p.bmap, np * n_cta, r0, bt, &sm->f_full[pf.stage]);
        pf.next();
      }
    }
  } else if (warp < 12) {
    // converters: warps 0-3 -> A tile [32 r][128 m], warps 4-11 -> B tile [32 r][n_cta <= 256]; thread = column of the tile
    const bool is_b = warp >= 4;
    const int t = is_b ? (int)threadIdx.x - 128 : (int)threadIdx.x;   // A: 0..127, B: 0..255
    const int width = is_b ? n_cta : BM;
    float csum0 = 0.f;                  // column sum of this thread's column over the CTA's rows
    Pipe pf(nf), po(RW_ST);
    for (int i = 0; i < n_slabs; ++i) {
      mbar_wait(&sm->f_full[pf.stage], pf.phase);
      const uint8_t* st = f_base + pf.stage * f_stage + (is_b ? a_src : 0);
      mbar_wait(&sm->op_empty[po.stage], po.phase ^ 1);
      uint8_t* op = op_base + po.stage * op_stage + (is_b ? a_op : 0);
      uint8_t* hi = op;
      uint8_t* lo = op + width * BK * 2;
      if (t < width) {
        const int c = t;
        const float* src = reinterpret_cast<const float*>(st) + c;
        float x[BK];
#pragma unroll
        for (int k = 0; k < BK; ++k) x[k] = src[k * width];
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < BK; ++k) s += x[k];
        csum0 += s;
        const uint32_t off = (c >> 3) * SBO + (c & 7) * 16;
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
      }
      mbar_arrive(&sm->f_empty[pf.stage]);   // after the staged values were consumed
      pf.next();
      fence_proxy_async();
      mbar_arrive(&sm->op_full[po.stage]);
      po.next();
    }
    // column sums (bias gradients): one CTA per (k-chunk, batch, column) writes them
    if (p.bias_from == 1 && !is_b && np == 0) {
      p.bias_partial[((size_t)kc * nb + bt) * p.M + mt * BM + t] = csum0;
    } else if (p.bias_from == 2 && is_b && mt == 0) {
      float* dst = p.bias_partial + ((size_t)kc * nb + bt) * p.N + np * n_cta;
      if (t < n_cta) dst[t] = csum0;
    }
    if (!is_b) {
      // epilogue by warps 0-3: partial[kc][bt][mt*128 + row][np*n_cta .. +n_cta)
      const int lane_base = (warp & 3) * 32;
      const int row = lane_base + lane;
      float* dst = p.partial + (((size_t)kc * nb + bt) * p.M + (size_t)mt * BM + row) * p.N + np * n_cta;
      if (n_slabs > 0) {
        mbar_wait(&sm->tmem_full, 0);
        tc_fence_after();
      }
      for (int c0 = 0; c0 < n_cta; c0 += 16) {
        uint32_t r[16];
        if (n_slabs > 0) {
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
              : "r"(tmem + ((uint32_t)lane_base << 16) + (uint32_t)c0));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int q = 0; q < 16; ++q) r[q] = 0u;
        }
#pragma unroll
        for (int q = 0; q < 16; q += 4)
          *reinterpret_cast<uint4*>(dst + c0 + q) = make_uint4(r[q], r[q + 1], r[q + 2], r[q + 3]);
      }
    }
  } else if (warp == 13) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_cta >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      Pipe po(RW_ST);
      for (int i = 0; i < n_slabs; ++i) {
        mbar_wait(&sm->op_full[po.stage], po.phase);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(op_base + po.stage * op_stage);
        const uint32_t a_lo = a_hi + BM * BK * 2;
        const uint32_t b_hi = a_hi + a_op;
        const uint32_t b_lo = b_hi + n_cta * BK * 2;
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

// out[b*ob + m*om + n*on] = scale * sum_k partial[k][b][m][n]  and  bias_out[i] = sum_k bias_partial[k][i]   (fixed order),
// for every job of the table; block -> job through block_begin.
struct RowsWgradReduceJob {
  const float* partial;
  const float* bias_partial;
  float* out;
  float* bias_out;
  long long ob, om, on;
  int kch, B, M, N, bias_len, block_begin;
  float scale;
};
struct RowsWgradReduceJobs {
  RowsWgradReduceJob job[RW_MAX_JOBS];
  int njobs;
};

__global__ void __launch_bounds__(256) rows_wgrad_reduce_kernel(const __grid_constant__ RowsWgradReduceJobs jobs) {
  int ji = 0;
  for (int j = 1; j < jobs.njobs; ++j)
    if ((int)blockIdx.x >= jobs.job[j].block_begin) ji = j;
  const RowsWgradReduceJob& q = jobs.job[ji];
  const long long elems = (long long)q.B * q.M * q.N;
  const long long i = (long long)((int)blockIdx.x - q.block_begin) * blockDim.x + threadIdx.x;
  if (i < elems) {
    float a = 0.f;
    for (int k = 0; k < q.kch; ++k) a += __ldg(q.partial + (size_t)k * elems + i);
    const int n = (int)(i % q.N);
    const long long t = i / q.N;
    const int m = (int)(t % q.M), b = (int)(t / q.M);
    q.out[b * q.ob + m * q.om + n * q.on] = a * q.scale;
  }
  if (q.bias_out && i < q.bias_len) {
    float a = 0.f;
    for (int k = 0; k < q.kch; ++k) a += __ldg(q.bias_partial + (size_t)k * q.bias_len + i);
    q.bias_out[i] = a;
  }
}

// [cols, rows, batch] fp32 view with a [box_cols, 32, 1] box, no swizzle (or [cols, batch, rows] / [box_cols, 1, 32] when the
// batch stride is the smaller one: the heads of a [R, H*dh] matrix).
static inline bool make_wgrad_map(PFN_encodeTiled encode, CUtensorMap* map, const float* base, int cols, int rows, int batch,
                                  long long ld, long long batch_stride, int box_cols, int* swapped) {
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * 4) % 16 || ld < cols || box_cols > 256 || (box_cols * 4) % 16) return false;
  if (batch > 1 && ((batch_stride * 4) % 16 || batch_stride <= 0)) return false;
  const bool swap = batch > 1 && batch_stride < ld;
  *swapped = swap ? 1 : 0;
  const cuuint64_t bs = batch > 1 ? (cuuint64_t)batch_stride * 4 : (cuuint64_t)ld * 4 * (cuuint64_t)rows;
  cuuint64_t gdim[3], gstr[2];
  cuuint32_t box[3];
  const cuuint32_t estr[3] = {1, 1, 1};
  gdim[0] = (cuuint64_t)cols;
  box[0] = (cuuint32_t)box_cols;
  if (swap) {
    gdim[1] = (cuuint64_t)batch; gdim[2] = (cuuint64_t)rows;
    gstr[0] = bs; gstr[1] = (cuuint64_t)ld * 4;
    box[1] = 1; box[2] = (cuuint32_t)BK;
  } else {
    gdim[1] = (cuuint64_t)rows; gdim[2] = (cuuint64_t)batch;
    gstr[0] = (cuuint64_t)ld * 4; gstr[1] = bs;
    box[1] = (cuuint32_t)BK; box[2] = 1;
  }
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// column part of a job's gradient tile per CTA: 256 for a single job (round-1 behaviour), 128 in a grouped launch (4 instead
// of 2 fp32 staging stages; the CTAs of the other jobs keep the SMs busy, so the wider tile's lower operand traffic matters
// less than the latency it cannot hide)
static inline int rows_wgrad_ncta(int N, int njobs) {
  int nc = njobs > 1 ? 128 : 256;
  if (N < nc) nc = N;
  while (N % nc) nc >>= 1;
  return nc;
}

// Split-K plan of a job table: every job gets the same number of k-chunks, chosen so that the CTAs of all jobs together
// are about two per SM (the jobs differ in cost per row slab, two waves even that out) while a CTA keeps at least 8 row
// slabs (256 rows) -- a single job is spread over the SMs once, with at least 4 slabs per CTA, as in round 1.
static inline int rows_wgrad_group_kch(const sgc_wgrad_job* jobs, int njobs, int R) {
  long long tiles = 0;
  for (int j = 0; j < njobs; ++j) tiles += (long long)(jobs[j].M / BM) * (jobs[j].N / rows_wgrad_ncta(jobs[j].N, njobs)) * jobs[j].B;
  const int total_slabs = (R + BK - 1) / BK;
  int k, cap;
  if (njobs == 1) { k = total_slabs / 4; cap = tiles < 148 ? (int)(148 / tiles) : 1; }
  else { k = total_slabs / 8; cap = tiles < 296 ? (int)(296 / tiles) : 1; }
  if (k > cap) k = cap;
  if (k < 1) k = 1;
  return k;
}

static inline bool rows_wgrad_job_ok(const sgc_wgrad_job& j) {
  return j.a && j.b && j.out && j.M > 0 && j.M % BM == 0 && j.N > 0 && j.N % 16 == 0 && j.B > 0 && j.bias_from >= 0 &&
         j.bias_from <= 2 && (!j.bias_from || j.bias_out);
}

}  // namespace tc
}  // namespace sgc

extern "C" long long sgc_rows_wgrad_group_scratch_floats(const sgc_wgrad_job* jobs, int njobs, int R) {
  using namespace sgc::tc;
  if (!jobs || njobs <= 0 || njobs > RW_MAX_JOBS || R <= 0) return 0;
  for (int j = 0; j < njobs; ++j)
    if (!rows_wgrad_job_ok(jobs[j])) return 0;
  const long long kch = rows_wgrad_group_kch(jobs, njobs, R);
  long long total = 0;
  for (int j = 0; j < njobs; ++j) {
    const long long M = jobs[j].M, N = jobs[j].N, B = jobs[j].B;
    total += kch * B * (M * N + (M > N ? M : N));
  }
  return total;
}

extern "C" int sgc_rows_wgrad_group_tc(const sgc_wgrad_job* jobs, int njobs, int R, float* scratch, void* stream) {
  using namespace sgc::tc;
  if (!jobs || njobs <= 0 || njobs > RW_MAX_JOBS || R <= 0 || !scratch) return (int)cudaErrorInvalidValue;
  PFN_encodeTiled encode = get_encode_tiled();
  if (!encode) return (int)cudaErrorNotSupported;
  RowsWgradJobs tab;
  RowsWgradReduceJobs red;
  tab.njobs = red.njobs = njobs;
  const int kch = rows_wgrad_group_kch(jobs, njobs, R);
  int cta = 0, blocks = 0, max_ncta = 0;
  float* sp = scratch;
  for (int j = 0; j < njobs; ++j) {
    const sgc_wgrad_job& in = jobs[j];
    if (!rows_wgrad_job_ok(in)) return (int)cudaErrorInvalidValue;
    RowsWgradJob& p = tab.job[j];
    p.n_cta = rows_wgrad_ncta(in.N, njobs);
    if (p.n_cta < 16 || p.n_cta % 16) return (int)cudaErrorInvalidValue;
    if (!make_wgrad_map(encode, &p.amap, in.a, in.M, R, in.B, in.lda, in.batch_a, BM, &p.a_swap)) return (int)cudaErrorInvalidValue;
    if (!make_wgrad_map(encode, &p.bmap, in.b, in.N, R, in.B, in.ldb, in.batch_b, p.n_cta, &p.b_swap)) return (int)cudaErrorInvalidValue;
    p.M = in.M; p.N = in.N;
    p.m_tiles = in.M / BM;
    p.n_parts = in.N / p.n_cta;
    p.nb = in.B;
    p.kch = kch;
    p.total_slabs = (R + BK - 1) / BK;
    p.slabs_per_cta = (p.total_slabs + kch - 1) / kch;
    p.bias_from = in.bias_from;
    p.partial = sp;
    sp += (size_t)kch * in.B * in.M * in.N;
    p.bias_partial = sp;
    sp += (size_t)kch * in.B * (in.M > in.N ? in.M : in.N);
    p.cta_begin = cta;
    cta += p.m_tiles * kch * p.n_parts * in.B;
    if (p.n_cta > max_ncta) max_ncta = p.n_cta;
    RowsWgradReduceJob& q = red.job[j];
    q.partial = p.partial; q.bias_partial = p.bias_partial;
    q.out = in.out; q.bias_out = in.bias_from ? in.bias_out : nullptr;
    q.ob = in.out_b; q.om = in.out_m; q.on = in.out_n;
    q.kch = kch; q.B = in.B; q.M = in.M; q.N = in.N;
    q.bias_len = in.bias_from == 1 ? in.B * in.M : in.bias_from == 2 ? in.B * in.N : 0;
    q.scale = in.scale;
    q.block_begin = blocks;
    const long long elems = (long long)in.B * in.M * in.N;
    const long long work = elems > q.bias_len ? elems : q.bias_len;
    blocks += (int)((work + 255) / 256);
  }
  // operand stages: two converted stages + at least two fp32 staging stages of the widest tile; always the full budget of
  // the widest configuration, so that narrower jobs get their extra staging stages and graph-captured launches agree
  const int stage_bytes = RW_ST * (2 * BM * BK * 2 + 2 * 256 * BK * 2) + 2 * (BK * BM * 4 + BK * 256 * 4);
  (void)max_ncta;
  tab.stage_bytes = stage_bytes;
  const size_t smem = (size_t)stage_bytes + sizeof(SmemRW) + 64;
  cudaError_t e = cudaFuncSetAttribute(rows_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  rows_wgrad_tc_kernel<<<cta, RW_THREADS, smem, (cudaStream_t)stream>>>(tab);
  SGC_CUDA_CHECK_LAST();
  rows_wgrad_reduce_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(red);
  SGC_CUDA_CHECK_LAST();
  return 0;
}

static inline sgc_wgrad_job rows_wgrad_single_job(const float* a, long long lda, long long batch_a, int M, const float* b,
                                                  long long ldb, long long batch_b, int N, int B, float* out, long long out_b,
                                                  long long out_m, long long out_n, float scale, float* bias_out, int bias_from) {
  sgc_wgrad_job j;
  j.a = a; j.lda = lda; j.batch_a = batch_a; j.M = M;
  j.b = b; j.ldb = ldb; j.batch_b = batch_b; j.N = N; j.B = B;
  j.out = out; j.out_b = out_b; j.out_m = out_m; j.out_n = out_n;
  j.scale = scale; j.bias_out = bias_out; j.bias_from = bias_from;
  return j;
}

extern "C" int sgc_rows_wgrad_tc_scratch_floats(int M, int N, int R, int B) {
  if (M <= 0 || N <= 0 || R <= 0 || B <= 0 || M % sgc::tc::BM || N % 16) return 0;
  float dummy = 0.f;
  const sgc_wgrad_job j = rows_wgrad_single_job(&dummy, M, 0, M, &dummy, N, 0, N, B, &dummy, 0, N, 1, 1.f, nullptr, 0);
  return (int)sgc_rows_wgrad_group_scratch_floats(&j, 1, R);
}

extern "C" int sgc_rows_wgrad_tc(const float* a, long long lda, long long batch_a, int M, const float* b, long long ldb,
                                 long long batch_b, int N, int R, int B, float* out, long long out_b, long long out_m,
                                 long long out_n, float scale, float* bias_out, int bias_from, float* scratch, void* stream) {
  const sgc_wgrad_job j = rows_wgrad_single_job(a, lda, batch_a, M, b, ldb, batch_b, N, B, out, out_b, out_m, out_n, scale,
                                                bias_out, bias_from);
  return sgc_rows_wgrad_group_tc(&j, 1, R, scratch, stream);
}
