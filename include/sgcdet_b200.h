/* sgcdet_b200 -- C ABI of the B200-native SGCDet view-transform hot path.
 *
 * One shared library (sgcdet_b200/_C/libsgcdet_b200.so, sm_100a only).  Conventions for EVERY entry point:
 *   - plain device pointers + sizes, no torch types; fp32 data, int32 indices unless stated (int64 for the
 *     reference's spatial_shapes / level_start_index tensors);
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued on it; nothing allocates, nothing syncs;
 *   - the current device must already be the one that owns the pointers;
 *   - return value is 0 on success, otherwise a cudaError_t (launch/argument error) -- the reference only
 *     printf()s launch failures (wms_deform_attn_cuda.cu:45-48,207-210); here they are returned and the Python
 *     shim raises RuntimeError.
 * Reference file abbreviations:
 *   CSRC = packages/3D-deformable-attention/DFA3D/dfa3D/ops/csrc
 *   DSK  = CSRC/common/cuda/ms_depth_score_sample_cuda_kernel.cuh     DSL = CSRC/cuda/ms_depth_score_sample_cuda.cu
 *   WMSK = CSRC/common/cuda/wms_deform_attn_cuda_kernel.cuh           WMSL = CSRC/cuda/wms_deform_attn_cuda.cu
 *   F3D  = mmdet3d_plugin/models/im2voxel/transformer_utils/multi_scale_3ddeformable_attn_function.py
 *   DCA  = mmdet3d_plugin/models/im2voxel/transformer_utils/deformable_cross_attention.py
 *   ENC  = mmdet3d_plugin/models/im2voxel/transformer_utils/encoder.py
 *   ASH  = mmdet3d_plugin/models/im2voxel/AdaptiveSparseHead.py       DH = mmdet3d_plugin/models/im2voxel/DenseHead.py
 */
#ifndef SGCDET_B200_H
#define SGCDET_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------------------------------------------
 * B2: DFA3D operator boundary  (replaces the four functions of dfa3D._ext, CSRC/pybind.cpp:42-67)
 *   value [B,S,M,Cm]  dist [B,S,M,D]  shapes3d [L,3] i64 (H,W,D)  shapes2d [L,2] i64  lsi [L] i64
 *   loc [B,Q,M,L,P,3] (x=w,y=h,z=d in [0,1])  loc2d [...,2]  attn [B,Q,M,L,P]  depth_score [B,Q,M,L,P,4]
 *   (corner order TL,TR,BR,BL: DSK:89-92).  im2col_step batching (WMSL:250-283) is a launch detail of the
 *   reference and has no numerical effect; it is accepted and ignored by the Python shim.
 * ------------------------------------------------------------------------------------------------- */

/* ms_depth_score_sample_forward (DSL:49-111, kernel DSK:94-148). out fully written. */
int dfa3d_depth_score_fwd(const float* dist, const int64_t* shapes3d, const int64_t* lsi, const float* loc,
                          int B, int S, int M, int D, int L, int Q, int P, float* out, void* stream);
/* ms_depth_score_sample_backward (DSL:138-201, kernel DSK:242-327). grad_dist accumulated (atomics),
 * grad_loc [B,Q,M,L,P,3] written: (0, 0, D * sum_corner g*(v_hi - v_lo)) (DSK:238-240). */
int dfa3d_depth_score_bwd(const float* dist, const int64_t* shapes3d, const int64_t* lsi, const float* loc,
                          const float* grad_out, int B, int S, int M, int D, int L, int Q, int P,
                          float* grad_dist, float* grad_loc, void* stream);
/* wms_deform_attn_forward (WMSL:213-288, kernel WMSK:240-303). out [B,Q,M*Cm] fully written. */
int dfa3d_wms_fwd(const float* value, const int64_t* shapes2d, const int64_t* lsi, const float* loc2d,
                  const float* attn, const float* depth_score, int B, int S, int M, int Cm, int L, int Q, int P,
                  float* out, void* stream);
/* wms_deform_attn_backward (WMSL:291-370, kernels WMSK:305-531). grad_value accumulated; grad_loc2d,
 * grad_attn, grad_depth_score written. */
int dfa3d_wms_bwd(const float* value, const int64_t* shapes2d, const int64_t* lsi, const float* loc2d,
                  const float* attn, const float* depth_score, const float* grad_out, int B, int S, int M, int Cm,
                  int L, int Q, int P, float* grad_value, float* grad_loc2d, float* grad_attn,
                  float* grad_depth_score, void* stream);
/* fp64 instantiation of the four entry points above (the reference dispatches over float and double:
 * csrc/cuda/ms_depth_score_sample_cuda.cu:95,153, csrc/cuda/wms_deform_attn_cuda.cu:267,344); plain scalar kernels
 * (csrc/dfa3d_op_f64.cu), same argument meaning and accumulation rules. */
int dfa3d_depth_score_fwd_f64(const double* dist, const int64_t* shapes3d, const int64_t* lsi, const double* loc, int B, int S,
                              int M, int D, int L, int Q, int P, double* out, void* stream);
int dfa3d_depth_score_bwd_f64(const double* dist, const int64_t* shapes3d, const int64_t* lsi, const double* loc,
                              const double* grad_out, int B, int S, int M, int D, int L, int Q, int P, double* grad_dist,
                              double* grad_loc, void* stream);
int dfa3d_wms_fwd_f64(const double* value, const int64_t* shapes2d, const int64_t* lsi, const double* loc2d, const double* attn,
                      const double* depth_score, int B, int S, int M, int Cm, int L, int Q, int P, double* out, void* stream);
int dfa3d_wms_bwd_f64(const double* value, const int64_t* shapes2d, const int64_t* lsi, const double* loc2d, const double* attn,
                      const double* depth_score, const double* grad_out, int B, int S, int M, int Cm, int L, int Q, int P,
                      double* grad_value, double* grad_loc2d, double* grad_attn, double* grad_depth_score, void* stream);
/* One-stage operator = MultiScale3DDeformableAttnFunction_fp32.forward/backward (F3D:277-351) without the
 * depth-score round trip.  depth_score_out may be NULL. */
int dfa3d_fused_fwd(const float* value, const float* dist, const int64_t* shapes3d, const int64_t* lsi,
                    const float* loc, const float* attn, int B, int S, int M, int Cm, int D, int L, int Q, int P,
                    float* out, float* depth_score_out, void* stream);
int dfa3d_fused_bwd(const float* value, const float* dist, const int64_t* shapes3d, const int64_t* lsi,
                    const float* loc, const float* attn, const float* grad_out, int B, int S, int M, int Cm, int D,
                    int L, int Q, int P, float* grad_value, float* grad_dist, float* grad_loc, float* grad_attn,
                    void* stream);

/* ---------------------------------------------------------------------------------------------------
 * B1: path entry points (one DenseHead level of AdaptiveSparseHead; M = 8 heads, P = 4 points, L = 1)
 * ------------------------------------------------------------------------------------------------- */

/* Projection + visibility + view-major pair list.  Replaces VoxFormerEncoder_DFA3D.point_sampling
 * (ENC:179-223) and the per-view nonzero()/rebatch loops (DCA:758-773).  proj [V,3,4] is built on the host
 * exactly as ENC:168-177 does.  sel [Q] = selected voxel ids (NULL = identity), ref3d [N,3] (DH:44-45).
 * Outputs: ref_cam [V,Q,3], mask [V,Q] u8, pair_index [V,Q] (-1 invisible), pair_vq [>= #pairs],
 * view_offsets [V+1] (last = #pairs), count [Q].  scratch: sgc_project_scratch_ints(V,Q) int32. */
int sgc_project_scratch_ints(int V, int Q);
int sgc_project_compact(const float* proj, const float* ref3d, const int* sel, int V, int Q, float ox, float oy,
                        float oz, float eps, float one_minus_eps, float img_w, float img_h, float dbound0,
                        float dscale, float* ref_cam, uint8_t* mask, int* pair_index, int* pair_vq,
                        int* view_offsets, int* count, int* scratch, void* stream);

/* fp32 -> [bf16 hi | bf16 lo | bf16 hi] operand split for the tensor-core projections (value_proj /
 * sampling_offsets / sampling_offsets_depth / attention_weights, DCA:417-436): rows of length cols (source row
 * stride src_stride), out[((r / rpg) * 3 + slot) * rpg + r % rpg][col] as bf16; slots (hi,lo,hi) for pattern 0,
 * (hi,hi,lo) for pattern 1.  See csrc/sgc_gemm_prep.cu. */
int sgc_split_bf16x3(const float* x, long long rows, int cols, long long src_stride, int rows_per_group, int pattern,
                     void* out, void* stream);

/* Tensor-core (tcgen05/TMEM) feature projection, fused with the NCHW -> channel-last layout change and the bf16 hi/lo
 * split: vg[v,s,n] = sum_c feat[v*view_stride + c*chan_stride + s] * W[n,c]  (value_proj + folded offset/weight rows,
 * DCA:417-436; replaces the flatten/permute of transformer.py:151-170).  wpack = sgc_pack_weight_tc(W [N,C]) is
 * 2*N*C bf16 (hi/lo slabs in the kernel's shared-memory image).  C % 32 == 0, N % 32 == 0, N <= 512. */
int sgc_pack_weight_tc(const float* w, int N, int C, void* out, void* stream);

/* All per-step operand preparations of one level's weights in ONE launch (the nn.Linear weights of DCA:417-436,
 * DCA:691-702 and the FFN, ENC:262-340, in the layouts the tensor-core GEMMs consume).  Each job reads a logical
 * [rows, cols] fp32 matrix at src[r*row_stride + c*col_stride] (so transposed operands need no copy), multiplies by
 * `scale` and writes kind 0: the sgc_split_bf16x3 image (rows_per_group, pattern as there), or kind 1: the
 * sgc_pack_weight_tc image (rows = N, cols = C).  `jobs` is a HOST array of njobs <= SGC_MAX_WEIGHT_JOBS entries. */
#define SGC_MAX_WEIGHT_JOBS 48
typedef struct sgc_weight_job {
  const float* src;
  void* out;
  long long row_stride, col_stride;
  int rows, cols;
  int rows_per_group, pattern;
  float scale;
  int kind;
} sgc_weight_job;
int sgc_prepare_weights(const sgc_weight_job* jobs, int njobs, void* stream);
int sgc_project_tc_fwd(const float* feat, long long view_stride, long long chan_stride, int V, int C, int S,
                       const void* wpack, int N, float* vg, void* stream);

/* Data gradient of the projection on the tensor cores (same kernel family): gfeat[(v*C+c)*chan_stride + s] =
 * sum_n gvg[v,s,n] W[n,c] for s < S (NCHW gradient written in place).  wpack_t = sgc_pack_weight_tc(W^T [C,N]). */
int sgc_project_tc_bwd_data(const float* gvg, int V, int S, int N, const void* wpack_t, int C, float* gfeat,
                            long long chan_stride, void* stream);

/* Weight gradient of the projection on the tensor cores: gw[n,c] = sum_{v,s} gvg[v,s,n] feat[(v*C+c)*chan_stride+s],
 * both fp32 operands split to bf16 hi/lo in shared memory, split-K partials summed in a fixed order. */
int sgc_project_tc_wgrad_scratch_floats(int N, int C);
/* Cap on the SMs the three projection kernels occupy (0 = all; they run beside the latency-bound voxel chain). */
int sgc_project_tc_set_max_ctas(int n);
/* The same cap for the forward projection alone (0 = the common cap): in the forward the projections of the finer levels
 * only run beside the voxel chain of the coarser levels and may leave it more SMs. */
int sgc_project_tc_set_max_ctas_fwd(int n);
/* n > 0: the forward / data-gradient kernels run as short-lived CTAs of n tiles each instead of persistent ones. */
int sgc_project_tc_set_tiles_per_cta(int n);
/* Programmatic dependent launch (sm_90+) for the kernels of the per-voxel chain (projection / pair list, lift forward,
 * cross-view kernels, row kernels, voxel-count GEMM, upsample / top-k / scatter / gather): with on != 0 they are launched
 * with the programmatic-stream-serialization attribute and overlap their launch + prologue with the tail of their
 * predecessor in the stream (every such kernel waits with griddepcontrol.wait before touching global data).  Off by
 * default. */
int sgc_set_pdl(int on);
int sgc_project_tc_wgrad(const float* gvg, const float* feat, long long chan_stride, int V, int S, int N, int C, float* gw,
                         float* scratch, void* stream);

/* Voxel-count GEMMs of the encoder layer on the tensor cores (tcgen05/TMEM/TMA, csrc/sgc_rows_gemm_tc.cu): every
 * nn.Linear applied to the selected voxel rows -- output_proj and the attention_pooling in/out projections (DCA:815-833;
 * key / value projections per head), the FFN (ENC:335-338) -- and their data gradients:
 *     y[b*batch_y + r*ldy + n] = sum_k x[b*batch_x + r*ldx + k] * W_b[n,k] (+ bias[b*bias_batch + n]),  r < R, n < N, b < B
 * x, y fp32 (16-byte aligned, ld*4 and batch*4 multiples of 16; a batch stride smaller than ld addresses the heads of a
 * [R, H*dh] matrix).  K % 32 == 0, N % 32 == 0.  The weights come packed by sgc_pack_weight_tc / sgc_prepare_weights
 * (kind 1) as a [pack_rows, K] matrix: batch b uses rows [b*pack_batch_rows, +N) of the matrix at
 * wpack + b*pack_batch_elems (bf16 elements; 0 = one matrix shared by all batches).  Both operands are split into bf16
 * hi/lo in shared memory / at pack time and accumulate in fp32 (relative error ~1e-5).  n_cta = output columns per CTA
 * (32/64/128/256, 0 = sgc_rows_gemm_tc_auto_ncta).  No cluster launch: starts as soon as one SM is free. */
/* Debug aid: CTA 0 of every following sgc_rows_gemm_tc launch stores clock64() at 13 points of its life into `stamps` (device,
 * 16 x int64; NULL switches it off).  See tools/rows_gemm_timeline.py. */
int sgc_rows_gemm_tc_set_debug(long long* stamps);
int sgc_rows_gemm_tc_auto_ncta(int R, int N, int B);
int sgc_rows_gemm_tc(const float* x, long long ldx, long long batch_x, int R, int K, int B, const void* wpack,
                     int pack_rows, long long pack_batch_elems, int pack_batch_rows, const float* bias, int bias_batch,
                     int N, float* y, long long ldy, long long batch_y, int n_cta, void* stream);
/* The same with two more operand modes for attention heads narrower than a 32-column k-slab (the 16-wide heads of the
 * C = 128 configs).  a_mode 1: the B batches share ONE x [R, K] (batch_x ignored): y[h] = x W_h^T with W_h the head's
 * weights zero-extended over all K columns.  a_mode 2 (B == 1): y = sum_a x[a] W_a^T over a_batches matrices
 * x[a] [R, K / a_batches] (batch stride batch_x), wpack = the packed [N, K] = [W_0 | W_1 | ...]. */
int sgc_rows_gemm_tc_ex(const float* x, long long ldx, long long batch_x, int R, int K, int B, const void* wpack,
                        int pack_rows, long long pack_batch_elems, int pack_batch_rows, const float* bias, int bias_batch,
                        int N, float* y, long long ldy, long long batch_y, int n_cta, int a_mode, int a_batches, void* stream);

/* Folded projection weights of MSDeformableAttention3D_DFA3D (DCA:417-436): Wcat [C + 4MP, C] = value_proj.weight rows
 * followed by the offset / depth-offset / attention-weight rows permuted to [head*P + point][off_x, off_y, off_d, logit];
 * gbias [4MP] the same permutation of the three small biases.  unfold: the gradients of Wcat / gbias scattered back into
 * seven contiguous parameter gradients (every element written). */
int sgc_fold_wcat(const float* value_w, const float* off_w, const float* dep_w, const float* att_w, const float* off_b,
                  const float* dep_b, const float* att_b, int C, int MP, float* wcat, float* gbias, void* stream);
int sgc_unfold_wcat_grad(const float* gwcat, const float* ggbias, int C, int MP, float* g_value_w, float* g_off_w,
                         float* g_dep_w, float* g_att_w, float* g_off_b, float* g_dep_b, float* g_att_b, void* stream);

/* Weight gradients of the same layers on the tensor cores (reduction over the voxel rows, csrc/sgc_rows_gemm_tc.cu):
 *     out[b*out_b + m*out_m + n*out_n] = scale * sum_r a[b*batch_a + r*lda + m] * b[b*batch_b + r*ldb + n],  m < M, n < N
 * For y = x W^T + bias with upstream gradient g: gW = g^T x (a = g, b = x).  bias_from = 1: bias_out[b*M + m] = sum_r a
 * (the bias gradient), 2: bias_out[b*N + n] = sum_r b, 0: none.  M % 128 == 0, N % 16 == 0; a batch stride smaller than
 * the leading dimension addresses the heads of a [R, H*dh] matrix (per-head key / value weights: a = t[h] or grad_qt[h],
 * b = the head's columns, and the output strides write the transposed result into in_proj_weight's gradient).
 * Both operands are split to bf16 hi/lo in shared memory; split-K partials in `scratch`
 * (sgc_rows_wgrad_tc_scratch_floats floats) are summed in a fixed order (deterministic). */
int sgc_rows_wgrad_tc_scratch_floats(int M, int N, int R, int B);
int sgc_rows_wgrad_tc(const float* a, long long lda, long long batch_a, int M, const float* b, long long ldb,
                      long long batch_b, int N, int R, int B, float* out, long long out_b, long long out_m,
                      long long out_n, float scale, float* bias_out, int bias_from, float* scratch, void* stream);
/* The same for a TABLE of up to 8 products over the same R voxel rows in ONE launch (+ one reduce launch): all weight
 * gradients of an encoder layer (output_proj, query / key / value in-projections, out_proj, both FFN layers).  Every job
 * is one product of the call above (same operand / output conventions; N % 16 == 0 here, so the 16-wide heads of the
 * C = 128 configs qualify).  scratch: sgc_rows_wgrad_group_scratch_floats(jobs, njobs, R) floats. */
typedef struct sgc_wgrad_job {
  const float* a; long long lda, batch_a; int M;
  const float* b; long long ldb, batch_b; int N;
  int B;
  float* out; long long out_b, out_m, out_n;
  float scale;
  float* bias_out; int bias_from;
} sgc_wgrad_job;
long long sgc_rows_wgrad_group_scratch_floats(const sgc_wgrad_job* jobs, int njobs, int R);
int sgc_rows_wgrad_group_tc(const sgc_wgrad_job* jobs, int njobs, int R, float* scratch, void* stream);

/* out[c] = sum_r x[r,c] for a row-major [R,C] matrix (bias gradients), deterministic.  scratch:
 * sgc_colsum_scratch_floats(R,C) floats; counter: one uint32 that is zero on entry (reset to zero on exit). */
int sgc_colsum_scratch_floats(int R, int C);
int sgc_colsum(const float* x, int R, int C, float* out, float* scratch, unsigned int* counter, void* stream);

/* One pass over g [R,C]: the rows-split [3R,C] bf16 (slots per `pattern`) for the weight-gradient GEMM g^T x, and
 * sums[c] = sum_r g[r,c] (the bias gradient).  scratch / counter as for sgc_colsum. */
int sgc_split_rows_colsum(const float* x, int R, int C, int pattern, void* out, float* sums, float* scratch,
                          unsigned int* counter, void* stream);

/* Lift: reference-point sample (DCA:67-116) + offset/weight heads + softmax (DCA:423-436) + sampling
 * locations (DCA:445-461) + 8-head 4-point DFA3D (F3D:277-302) for every visible pair.
 * value [V,S,ldv] (no bias), G [V,S,ldg] (128 ch, [m][p][ox,oy,od,logit]), dist [V,S,D], vbias [C], gbias [128],
 * n_pairs = device pointer to the pair count (view_offsets + V).  Writes samp [cap,32,4], slots [cap,C]. */
int sgc_lift_fwd(const float* value, int ldv, const float* G, int ldg, const float* dist, const float* vbias,
                 const float* gbias, const int* pair_vq, const int* n_pairs, int cap_pairs, const float* ref_cam,
                 int S, int H, int W, int D, int Q, int C, float* samp, float* slots, void* stream);
/* Backward of the above (F3D:303-351 + autograd of the Linear heads).  All grads ACCUMULATE (caller zeroes).
 * scratch: sgc_lift_bwd_scratch_floats(cap_pairs, C) floats (per-CTA bias partials, reduced deterministically). */
/* The same backward as a GATHER over 4x8 pixel tiles (csrc/sgc_lift_tiles.cu): points are filed under the tiles their 2x2
 * corner blocks touch; a CTA stages its tile's value rows in shared memory (bulk async copies), dots / accumulates the
 * records of the tile there and writes every grad_value / grad_G row ONCE -- no zero fill of the two (grad_value / grad_G are
 * fully overwritten), no reductions into them.  grad_dist is zero-filled inside; grad_vbias / grad_gbias are ASSIGNED.
 * S == H*W.  workspace: sgc_lift_bwd_tiles_workspace_bytes() bytes, 256-byte aligned, private to the call. */
long long sgc_lift_bwd_tiles_workspace_bytes(int cap_pairs, int V, int H, int W, int C);
int sgc_lift_bwd_tiles(const float* value, int ldv, const float* G, int ldg, const float* dist, const float* vbias,
                       const int* pair_vq, const int* n_pairs, int cap_pairs, const float* ref_cam, const float* samp,
                       const float* grad_slots, int V, int S, int H, int W, int D, int Q, int C, float* grad_value,
                       float* grad_G, float* grad_dist, float* grad_vbias, float* grad_gbias, void* workspace, void* stream);
int sgc_lift_bwd_scratch_floats(int cap_pairs, int C);
int sgc_lift_bwd(const float* value, int ldv, const float* G, int ldg, const float* dist, const float* vbias,
                 const int* pair_vq, const int* n_pairs, int cap_pairs, const float* ref_cam, const float* samp,
                 const float* grad_slots, int S, int H, int W, int D, int Q, int C, float* grad_value,
                 float* grad_G, float* grad_dist, float* grad_vbias, float* grad_gbias, float* scratch,
                 void* stream);

/* Cross-view fusion (DCA:815-833).  mean [Q,C]: masked mean over views (zeros when no view sees q).
 * attn: qt [8,Q,C] (scaled, key-projected query), t_out [8,Q,C] = sum_v softmax_v(qt.s_v) s_v, alpha [cap,8]. */
int sgc_crossview_mean_fwd(const float* slots, const int* pair_index, int V, int Q, int C, float* mean, void* stream);
int sgc_crossview_attn_fwd(const float* qt, const float* slots, const int* pair_index, int V, int Q, int C,
                           float* t_out, float* alpha, void* stream);
/* Backward in two steps, because grad_mean depends on grad_qt through the host GEMMs:
 *   _qt    : gscore [cap,8] (softmax backward of the view scores) and grad_qt [8,Q,C], both fully written;
 *   _slots : grad_slots [#pairs,C] = grad_mean/n + sum_h (alpha grad_t[h] + gscore qt[h]), fully written. */
int sgc_crossview_attn_bwd_qt(const float* slots, const float* alpha, const int* pair_index, int V, int Q, int C,
                              const float* grad_t, float* gscore, float* grad_qt, void* stream);
/* The same kernels additionally emitting the bf16x3 operand image (sgc_split_bf16x3 pattern 0) of their dense output
 * for the tensor-core GEMM that follows: mean [Q,3C], t [8*Q,3C], grad_qt [8*Q,3C]. */
int sgc_crossview_mean_fwd_split(const float* slots, const int* pair_index, int V, int Q, int C, float* mean, void* split,
                                 void* stream);
int sgc_crossview_attn_fwd_split(const float* qt, const float* slots, const int* pair_index, int V, int Q, int C,
                                 float* t_out, float* alpha, void* split, void* stream);
int sgc_crossview_attn_bwd_qt_split(const float* slots, const float* alpha, const int* pair_index, int V, int Q, int C,
                                    const float* grad_t, float* gscore, float* grad_qt, void* split, void* stream);
int sgc_crossview_attn_bwd_slots(const float* qt, const float* alpha, const float* gscore, const int* pair_index,
                                 int V, int Q, int C, const float* grad_t, const float* grad_mean,
                                 float* grad_slots, void* stream);

/* View-sharded cross-view fusion (config 5): the views are split over shards (GPUs) and every softmax
 * statistic over views is (local partial) -> all-reduce on the host side (NCCL) -> (local finish):
 *   sum_fwd   : sum_v slots (not divided)                       -> [SUM sum, count]  -> mean
 *   scores    : scores [cap,8], local max m_loc [Q,8]           -> [MAX m]
 *   accum     : e = exp(score - m) [cap,8], s_loc [Q,8], o_loc [8,Q,C] -> [SUM s, o] -> t = o / s
 *   bwd_dot   : alpha = e/s, g_alpha [cap,8], D_loc [Q,8] = sum_v alpha g_alpha     -> [SUM D]
 *   bwd_qt    : gscore = alpha (g_alpha - D) [cap,8], gqt_loc [8,Q,C]               -> [SUM gqt]
 *   bwd_slots : as sgc_crossview_attn_bwd_slots with the GLOBAL view count in the mean term. */
int sgc_crossview_sum_fwd(const float* slots, const int* pair_index, int V, int Q, int C, float* sum, void* stream);
int sgc_cvs_scores(const float* qt, const float* slots, const int* pair_index, int V, int Q, int C, float* scores,
                   float* m_loc, void* stream);
int sgc_cvs_accum(const float* scores, const float* m_glob, const float* slots, const int* pair_index, int V, int Q,
                  int C, float* e_out, float* s_loc, float* o_loc, void* stream);
int sgc_cvs_bwd_dot(const float* slots, const float* e_in, const float* s_glob, const int* pair_index, int V, int Q,
                    int C, const float* grad_t, float* alpha, float* galpha, float* d_loc, void* stream);
int sgc_cvs_bwd_qt(const float* slots, const float* alpha, const float* galpha, const float* d_glob,
                   const int* pair_index, int V, int Q, int C, float* gscore, float* gqt_loc, void* stream);
int sgc_cvs_bwd_slots(const float* qt, const float* alpha, const float* gscore, const int* pair_index, int V, int Q,
                      int C, const float* grad_t, const float* grad_mean, const int* count_glob, float* grad_slots,
                      void* stream);
/* Finishing step after an exchange: y[r,c] = x[r,c] / max(s[r, c / (C/heads)], smin) (+ bias[c]); heads == 1 may also emit
 * cnt[r] = (int)s[r] (the global view count).  x / y / bias 16-byte aligned, (C / heads) % 4 == 0. */
int sgc_rows_headscale(const float* x, const float* s, int heads, float smin, const float* bias, int R, int C, float* y,
                       int* cnt, void* stream);

/* nn.LayerNorm backward over voxel rows [R,C] (the two norms of VoxFormerLayer, ENC:262-340), C in {128, 256}:
 * gx fully written; `partial` (sgc_layernorm_bwd_scratch_floats(R,C) floats) receives per-CTA sums that
 * sgc_layernorm_bwd_params reduces in a fixed order into ggamma / gbeta (may run on another stream, after an event).
 * mean / rstd are the [R] statistics of the forward (torch.native_layer_norm). */
int sgc_layernorm_bwd_scratch_floats(int R, int C);
int sgc_layernorm_bwd(const float* x, const float* gy, const float* mean, const float* rstd, const float* gamma,
                      int R, int C, float* gx, float* partial, void* stream);
int sgc_layernorm_bwd_params(const float* partial, int R, int C, float* ggamma, float* gbeta, void* stream);

/* Fused row epilogue / prologue around the tensor-core GEMMs of the encoder layer (output_proj + attention_pooling of
 * DCA:815-837, the norms and the FFN of ENC:262-340): one warp per voxel row of N in {128,256,512} channels.
 *   forward : v = x + bias; relu; v *= mask*mscale; v *= rowscale[r]; v += residual; [pre = v; v = LayerNorm(v)]
 *             -> y (fp32) and ysplit (sgc_split_bf16x3 pattern-0 image of y: the operand of the next GEMM)
 *   backward: v = g (+ g2); [LayerNorm backward; per-CTA (g*xhat, g) sums into partial, reduced by
 *             sgc_layernorm_bwd_params]; gpre = v; v *= mask*mscale; v = gate > 0 ? v*gscale : 0; v *= rowscale[r]
 *             -> gx (fp32), gxsplit (bf16x3 image)
 * Every pointer except x / g may be NULL (= that stage is skipped).  in_heads = H: the input is head-major [H,R,N/H];
 * split_heads = H: the bf16x3 image is per (row, head): [(r*H+h)][slot*dh + d] (operand of a per-head GEMM).
 * partial needs sgc_layernorm_bwd_scratch_floats(R, N) floats. */
typedef struct sgc_rowop_fwd_args {
  const float* x;
  const float* bias;
  const unsigned char* mask;
  const float* rowscale;
  const float* residual;
  const float* gamma;
  const float* beta;
  float* y;
  void* ysplit;
  float* pre;
  float* mean;
  float* rstd;
  float mscale, eps;
  int R, N, relu, in_heads, split_heads;
  const int* rowcount; /* optional [R] view counts: v *= (rowcount[r] > 0), applied where rowscale is */
} sgc_rowop_fwd_args;
typedef struct sgc_rowop_bwd_args {
  const float* g;
  const float* g2;
  const float* pre;
  const float* mean;
  const float* rstd;
  const float* gamma;
  const unsigned char* mask;
  const float* gate;
  const float* rowscale;
  float* partial;
  float* gpre;
  float* gx;
  void* gxsplit;
  float mscale, gscale;
  int R, N, in_heads, split_heads;
  const int* rowcount; /* as in sgc_rowop_fwd_args */
} sgc_rowop_bwd_args;
int sgc_rowop_fwd(const sgc_rowop_fwd_args* args, void* stream);

int sgc_rowop_bwd(const sgc_rowop_bwd_args* args, void* stream);

/* The keep-masks of all dropouts of a step (every level, every layer) in ONE launch: out[j][i] = 1 with probability keep_j,
 * else 0 -- nn.Dropout's mask (FFN / attention dropouts of the VoxFormerLayer, encoder.py:262-340; the reference draws them
 * with ATen's bernoulli_, two launches per layer); applied by sgc_rowop_fwd / _bwd (mask, mscale = 1/keep).
 * Philox4x32-10 keyed by `seed`, counter = (position, job, step).  state = two int64 on the device, zero-initialised once,
 * private to the call site: state[0] is the step number, advanced by the kernel itself, so a CUDA-graph replay draws
 * fresh masks.  out pointers 16-byte aligned, njobs <= 12. */
typedef struct sgc_mask_job {
  unsigned char* out;
  long long n;
  float keep;
} sgc_mask_job;
int sgc_dropout_masks(const sgc_mask_job* jobs, int njobs, long long seed, long long* state, void* stream);

/* Sparse volume construction on channel-last volumes [X,Y,Z,C].
 * upsample: F.interpolate(x2, trilinear, align_corners=False) (ASH:64-69) fused with the occupancy head
 * Linear(C,1)+Sigmoid (ASH:37-39,71).  bwd: grad_in written; grad_w [C], grad_b [1] accumulated. */
int sgc_upsample2x_occ_fwd(const float* vol_in, int X, int Y, int Z, int C, const float* w_occ, const float* b_occ,
                           float* vol_out, float* occ, void* stream);
int sgc_upsample2x_occ_bwd(const float* vol_in, int X, int Y, int Z, int C, const float* w_occ, const float* occ,
                           const float* grad_up, const float* grad_occ, float* gpre_scratch, float* grad_in,
                           float* grad_w, float* grad_b, float* scratch, void* stream);
/* floats of `scratch` above: the gradient is then evaluated axis by axis (three streaming passes over whole rows);
 * scratch = NULL selects the one-launch form (one warp per input voxel gathering its <= 64 output rows). */
long long sgc_upsample2x_occ_bwd_scratch_floats(int X, int Y, int Z, int C);
/* The weight-gradient part of the call above on its own (grad_w accumulated); pass grad_w = NULL above to skip it
 * there and issue it on another stream. */
int sgc_upsample2x_occ_gradw(const float* vol_in, int X, int Y, int Z, int C, const float* gpre, float* grad_w,
                             void* stream);
/* topk_wo_grad (ASH:9-13) + nonzero compaction (DH:66): k largest, ties -> lower index; sel ascending. */
int sgc_topk_select(const float* occ, int N, int k, int* sel, uint8_t* mask, void* stream);
/* The same selection (same deterministic contract) spread over many CTAs for large N (the 204 800-voxel level of the "-L"
 * configs): chunk histograms merged with integer atomics, one pick launch per radix round, then count + ordered compaction.
 * scratch: sgc_topk_scratch_ints(N) ints, private to the call.  k >= 1. */
int sgc_topk_scratch_ints(int N);
int sgc_topk_select_mc(const float* occ, int N, int k, int* sel, uint8_t* mask, int* scratch, void* stream);
/* The same selection as ONE launch for every level size up to sgc_topk_grid_max_n() (229 376) scores: a few CTAs, keys in
 * registers, the threshold found 4 bits per step with the per-step counts merged through packed 64-bit atomics.
 * scratch: sgc_topk_grid_scratch_bytes() bytes, zero-filled ONCE when allocated, then reused by every call issued on the
 * same stream (calls on different streams need their own).  k >= 1. */
int sgc_topk_grid_scratch_bytes(void);
int sgc_topk_grid_max_n(void);
int sgc_topk_select_grid(const float* occ, int N, int k, int* sel, uint8_t* mask, void* scratch, void* stream);
/* AdaptiveSparseHead.occ_loss (ASH:100-103): loss[0] = 0.5 * mean(BCELoss(p, t)) (logs clamped at -100 like torch);
 * bwd: grad_p[i] = g[0] * 0.5/N * (p-t)/max(p(1-p),1e-12), g = the upstream gradient (one float on the device). */
int sgc_occ_loss_fwd(const float* p, const float* t, int N, float* loss, void* stream);
int sgc_occ_loss_bwd(const float* p, const float* t, const float* g, int N, float* grad_p, void* stream);
/* vol[sel[i],:] += y[i,:] (DH:80-81 + ASH:77) and y[i,:] = vol[sel[i],:] (its backward). */
int sgc_scatter_add_rows(float* vol, const int* sel, const float* y, int k, int C, void* stream);
int sgc_gather_rows(const float* vol, const int* sel, float* y, int k, int C, void* stream);

/* The head's per-level validity masks: nn.Upsample(size, mode='trilinear')(valid).round().bool() of
 * dense_heads/imvoxel_head_v2.py:121-123,256-258 for levels of 1, 1/2 and 1/4 of the volume, bit-exact, one launch.  valid
 * [X,Y,Z] int64; v0 / v1 / v2 uint8 0/1 outputs (NULL = not wanted); X, Y, Z multiples of 4 (of 2 without v2). */
int sgc_valid_pyramid(const long long* valid, int X, int Y, int Z, unsigned char* v0, unsigned char* v1, unsigned char* v2,
                      void* stream);

/* ---- the depth-distribution producer in front of the path (csrc/sgc_depth.cu) ----------------
 * Plane-sweep cost volume of DepthNet_Fusion (mmdet3d_plugin/models/im2voxel/depth_utils/depth_est_fusion.py:85-126 homo_warping,
 * :209-232 loop over the neighbour frames), fused: corr[v,d,y,x] = 1/(K sqrt(C)) sum_k sum_c f[v,c,y,x] *
 * grid_sample(f[nbr[v,k]], H_{v,k,d}(x,y))[c].  feat_cl = channel-last [V, H*W, C] (C % 4 == 0, C <= 512, 16-byte aligned);
 * nbr [V,K] int32 neighbour view ids; rt [V,K,12] = rows of (src_proj @ inverse(ref_proj))[:3,:3] then [:3,3]; depth [D]
 * (D <= 32); corr [V,D,H,W].  The backward zero-fills grad_feat_cl ([V, H*W, C]) and accumulates both roles of a feature row. */
int sgc_plane_sweep_fwd(const float* feat_cl, const int* nbr, const float* rt, const float* depth, int V, int K, int D, int H,
                        int W, int C, float* corr, void* stream);
int sgc_plane_sweep_bwd(const float* feat_cl, const int* nbr, const float* rt, const float* depth, const float* grad_corr, int V,
                        int K, int D, int H, int W, int C, float* grad_feat_cl, void* stream);
/* [B,C,S] <-> [B,S,C] tiled transposes (the FPN's NCHW maps <-> the channel-last gather layout). */
int sgc_nchw_to_nhwc(const float* in, int B, int C, int S, float* out, void* stream);
int sgc_nhwc_to_nchw(const float* in, int B, int C, int S, float* out, void* stream);
/* softmax over the depth bins (depth_est_fusion.py:241) + the nearest x1/2, x1/4 pyramid of SGCDet.build_volume
 * (detectors/SGCDet.py:83-85) + the cropped channel-last layout [V, h_l*w_l, D] the lift kernels read.  logits [V,D,H,W];
 * prob [V,D,H,W] or NULL; cl_l NULL = level not wanted; (h_l, w_l) <= ceil((H, W) / 2^l).  Backward: grad_prob / gcl_l may be
 * NULL. */
int sgc_depth_pyramid_fwd(const float* logits, int V, int D, int H, int W, float* prob, float* cl0, int h0, int w0, float* cl1,
                          int h1, int w1, float* cl2, int h2, int w2, void* stream);
int sgc_depth_pyramid_bwd(const float* logits, int V, int D, int H, int W, const float* grad_prob, const float* gcl0, int h0,
                          int w0, const float* gcl1, int h1, int w1, const float* gcl2, int h2, int w2, float* grad_logits,
                          void* stream);

/* Peer memory over NVLink (csrc/sgc_peer.cu) -- SURVEY.md section 8e: the collectives of view sharding (partial sums /
 * counts, score maxima, partial-softmax sums, the backward's normaliser dot and query gradient) and the weight-gradient
 * average of scene-batch data parallelism, as ONE kernel launch each that a CUDA graph can hold.
 *
 * sgc_peer_alloc: a symmetric allocation of `bytes` (zero-filled; synchronises the device) and its 64-byte CUDA IPC handle.
 * The host exchanges the handles (torch.distributed) and maps the peers' allocations with sgc_peer_open.  The first
 * sgc_peer_sig_bytes() of an allocation are used as the signal pad of one collective "channel"; collectives on the same
 * channel must be issued in the same order on every rank, on one stream. */
int sgc_peer_sig_bytes(void);
int sgc_peer_status_offset(void);   /* uint32 in the pad: 1 after a barrier gave up waiting (~4 s) for a peer */
int sgc_peer_alloc(long long bytes, void** ptr, void* handle64);
int sgc_peer_open(const void* handle64, void** ptr);
int sgc_peer_close(void* ptr);
int sgc_peer_free(void* ptr);
/* One-shot all-reduce: bufs / sigs are HOST arrays of `world` device pointers -- rank r's data buffer and signal pad as mapped
 * into this process.  out[i] = scale * sum (op 0) or max (op 1, scale ignored) over the ranks of bufs[r][i], i < n, reduced
 * in rank order on every rank (bit-identical results everywhere).  Every rank calls it with the same n; the cross-GPU
 * barriers are flags in the signal pads (stateless: the launch can be replayed from a CUDA graph); world <= 8.  max_blocks
 * (0 = 128, the same on every rank) and block_threads (0 = 512, or 128) bound the footprint of a collective that runs beside
 * other kernels. */
int sgc_peer_allreduce(const void* const* bufs, void* const* sigs, int rank, int world, long long n, int op, float scale,
                       float* out, int max_blocks, int block_threads, void* stream);
/* dst[k][0:n[k]] = src[k][0:n[k]] for `count` segments (HOST arrays of device pointers / lengths) as one launch of 128-thread
 * CTAs per 64 segments: the gather of gradient tensors into the symmetric buffer and the scatter of the averages back. */
int sgc_peer_copy_segments(const void* const* src, void* const* dst, const long long* n, int count, int max_blocks, void* stream);
/* tensors[k][0:n[k]] <- scale * sum over the ranks, in place, for count <= 64 tensors as ONE launch (gather into the symmetric
 * buffers, meet, reduce in rank order, scatter): the tail of the gradient average. */
int sgc_peer_allreduce_tensors(const void* const* bufs, void* const* sigs, int rank, int world, void* const* tensors,
                               const long long* n, int count, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SGCDET_B200_H */
