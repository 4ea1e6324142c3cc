/*
 * wf_b200.h -- C ABI of libwf_b200.so, the B200 (sm_100a) implementation of the hot path of
 * cansdev/wireframe-3d-prediction: PointCloudToWireframe forward/backward, WireframeLoss and the
 * Hungarian matchers.
 *
 * The reference is pure Python/PyTorch and has no FFI of its own; its boundary is
 * the Python module API.  The entry points below are what a ctypes binding for that path binds:
 * each one names the reference lines it replaces (paths relative to the reference repo).  The
 * host-side mirror of the reference classes (wireframe-3d-prediction_b200/models, /losses) calls
 * ONLY these functions for compute; INTEGRATION.md shows the stub a maintainer would add.
 *
 * Conventions
 *   - plain pointers + sizes; every pointer is DEVICE memory unless the name ends in _host;
 *   - row-major, innermost dimension contiguous, explicit leading dimensions where given;
 *   - the caller owns every buffer (inputs, outputs, workspaces); nothing is allocated or freed;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant, and
 *     keeps no mutable global state beyond one-time function-attribute setup;
 *   - return 0 on success; otherwise a WF_E* code and wf_last_error() (thread-local) explains;
 *   - no C++ exception crosses the boundary.
 */
#ifndef WF_B200_H
#define WF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* wf_stream_t;

enum { WF_OK = 0, WF_EINVAL = 1, WF_ECUDA = 2, WF_EUNSUPPORTED = 3, WF_ETOOBIG = 4 };
enum { WF_F32 = 0, WF_BF16 = 1 };
enum { WF_ACT_NONE = 0, WF_ACT_RELU = 1, WF_ACT_GELU = 2 };
/* per-problem LSAP status, mirrors scipy's errors (losses/WireframeLoss.py:236 raises them) */
enum { WF_LSAP_OK = 0, WF_LSAP_INFEASIBLE = 1, WF_LSAP_INVALID = 2 };

int wf_version(void);
const char* wf_last_error(void);
/* sm count and compute capability of the current device */
int wf_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------
 * Matching
 * ---------------------------------------------------------------------------------------- */

/* losses/WireframeLoss.py:129-244 (_hungarian_matching, the live `final_cost_matrix`):
 * per sample build the V x V cost (L1 cdist + |e-1| for the first counts[b] columns, e for the
 * dummy columns) in shared memory and solve it with the Crouse/Jonker-Volgenant shortest
 * augmenting path in fp64 -- one warp per sample, no host round trip.
 * col_of_row[b*V+i] = assigned column of prediction i (columns >= counts[b] are dummies).
 * status[b] = WF_LSAP_*; counts[b] > V gives WF_LSAP_INFEASIBLE (the reference appends inf rows
 * and scipy raises).  cost_dump (optional, B*V*V floats) receives the matrices. */
int wf_loss_match(const float* pred_v, const float* pred_e, const float* tgt_v,
                  const int64_t* counts, int B, int V, int Vt, int32_t* col_of_row,
                  int32_t* status, float* cost_dump, wf_stream_t stream);

/* scipy.optimize.linear_sum_assignment on B independent float32 matrices
 * (models/WireframeHungarianMatcher.py:70-71, models/HungarianMatcher.py:126-127).
 * Matrix b is nr[b] x nc[b], stored at cost + b*batch_stride with leading dimension ld.
 * col_of_row[b*max_nr+i] = column assigned to row i, or -1 (rows > cols leaves rows free).
 * Identical index choice to scipy, ties included. */
int wf_lsap_batched(const float* cost, int64_t batch_stride, int ld, const int32_t* nr,
                    const int32_t* nc, int B, int max_nr, int max_nc, int32_t* col_of_row,
                    int32_t* status, wf_stream_t stream);

/* models/WireframeHungarianMatcher.py:52-67: C = wv*cdist_L1(pred, tgt) + we*|e_pred - e_tgt|.
 * Targets are concatenated (tgt_off[B+1] prefix offsets); only the block-diagonal part the
 * reference keeps after `C.split(sizes,-1)` is produced: cost[b] is V x (tgt_off[b+1]-tgt_off[b])
 * at cost + b*V*ld. */
int wf_wireframe_matcher_cost(const float* pred_v, const float* pred_e, const float* tgt_v,
                              const float* tgt_e, const int32_t* tgt_off, int B, int V, float wv,
                              float we, float* cost, int ld, wf_stream_t stream);

/* models/HungarianMatcher.py:101-123: -softmax(logits)[label]*wc + L1(boxes)*wb - GIoU*wg. */
int wf_detr_matcher_cost(const float* logits, const float* boxes, const int64_t* tgt_labels,
                         const float* tgt_boxes, const int32_t* tgt_off, int B, int Q, int K,
                         float wc, float wb, float wg, float* cost, int ld, wf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Dense row-MLP primitives, fp32 (heads: SURVEY K6-K10, K14; and the fp32 parity mode of K2-K4)
 * ---------------------------------------------------------------------------------------- */

/* C = alpha*op(A)*op(B) + beta*C (+ bias[n] broadcast over rows).  op(A) is M x K, op(B) is K x N.
 * nn.Linear forward is (transA=0, transB=1); dX is (0,0); dW is (1,0) with beta=1 to accumulate. */
int wf_gemm_f32(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda,
                const float* B, int ldb, float beta, float* C, int ldc, const float* bias,
                wf_stream_t stream);

/* Small-batch products of the 64-row heads (models/PointNetEncoder.py:57-65,115-116; models/VertexPredictor.py:94-117), TF32
 * operands / fp32 accumulate, shaped for weight streaming (csrc/rowmlp.cu).  All dimensions and leading dimensions multiples of
 * 4 floats, pointers 16-byte aligned.
 *   wf_rowmlp_linear: Y[M,N] = X[M,K] * Wop^T (+ bias[N]).  trans_w = 0: W is [N,K] (nn.Linear forward);  trans_w = 1: W is
 *     [K,N] read in place (dX = dZ * W of an nn.Linear whose weight is W).  A cluster of 8 CTAs per 64-column slab splits K and
 *     adds the partial tiles over distributed shared memory in rank order: deterministic, no workspace.
 *   wf_rowmlp_dw: dW[Nr,Kc] = dZ[Mb,Nr]^T * X[Mb,Kc] for Mb <= 128 batch rows (the reduction dimension), every entry written
 *     once (no atomics, no zero fill); db[Nr] = column sums of dZ (exact fp32) when db != NULL. */
int wf_rowmlp_linear(const float* X, int ldx, const float* W, int ldw, int trans_w, const float* bias, int M, int N, int K,
                     float* Y, int ldy, wf_stream_t stream);
int wf_rowmlp_dw(const float* dZ, int ldz, const float* X, int ldx, int Mb, int Nr, int Kc, float* dW, int ldw, float* db,
                 wf_stream_t stream);

/* out = dropout(act(LayerNorm(z))) + residual  -- one fused pass per row.
 * models/PointNetEncoder.py:37-40,58-63; models/VertexPredictor.py:28-54,110,114;
 * models/EdgePredictor.py:31-38,57-66.  gamma==NULL skips the normalisation (pure activation,
 * EdgePredictor.py:65-66).  z/out dtype WF_F32 or WF_BF16.  If stats_in != 0, mean/rstd are
 * inputs (already reduced by the tensor-core epilogue), else they are outputs (may be NULL).
 * keep (optional, M*C bytes) is a dropout keep-mask applied after the activation with keep_scale. */
int wf_ln_act_fwd(const void* z, int z_dtype, const float* gamma, const float* beta, int act,
                  const void* residual, const uint8_t* keep, float keep_scale, void* out,
                  int out_dtype, float* mean, float* rstd, int stats_in, int M, int C, float eps,
                  wf_stream_t stream);

/* Backward of the above w.r.t. z, gamma, beta.  dgamma/dbeta/dcolsum (each C floats, may be NULL)
 * are ACCUMULATED into (caller zeroes); dcolsum receives sum_m dz[m,:] = the bias gradient of the
 * Linear that produced z. */
int wf_ln_act_bwd(const void* dout, int dout_dtype, const void* z, int z_dtype, const float* gamma,
                  const float* beta, const float* mean, const float* rstd, int act,
                  const uint8_t* keep, float keep_scale, void* dz, int dz_dtype, float* dgamma,
                  float* dbeta, float* dcolsum, int M, int C, wf_stream_t stream);

/* column sums: out[c] += sum_m x[m,c]  (bias gradients of Linears without a LayerNorm behind) */
int wf_colsum(const void* x, int dtype, int M, int C, int ld, float* out, wf_stream_t stream);

/* models/VertexPredictor.py:117-127: split (B, 4V) logits into coords (B,V,3), sigmoid(existence)
 * (B,V) and the >0.5 count (B,) int64.  Backward: d_vf = [d_coords, d_prob * p(1-p)]. */
int wf_vertex_split_fwd(const float* vf, int B, int V, float* coords, float* prob, int64_t* count,
                        wf_stream_t stream);
int wf_vertex_split_bwd(const float* d_coords, const float* d_prob, const float* prob, int B, int V,
                        float* d_vf, wf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Encoder
 * ---------------------------------------------------------------------------------------- */

/* models/PointNetEncoder.py:85-86: mask[m] = sum|x[m,:]| > 1e-9, valid[b] = max(1, #mask). */
int wf_point_mask(const float* x, int B, int N, int D, uint8_t* mask, float* valid,
                  wf_stream_t stream);

/* First encoder layer, fused Linear(D->C)+LayerNorm+ReLU, fp32 math (intensity is un-normalised,
 * SURVEY D6), memory-bound: models/PointNetEncoder.py:37-40 (i=0).  h dtype WF_F32 or WF_BF16. */
int wf_enc_l1_fwd(const float* x, const float* W, const float* b, const float* gamma,
                  const float* beta, void* h, int h_dtype, int M, int D, int C, float eps,
                  wf_stream_t stream);
/* Backward: recomputes the layer from x (no saved activations); accumulates dW[C,D], db, dgamma,
 * dbeta (caller zeroes); dx optional (M*D floats, written). */
int wf_enc_l1_bwd(const float* x, const float* W, const float* b, const float* gamma,
                  const float* beta, const void* dh, int dh_dtype, float* dW, float* db,
                  float* dgamma, float* dbeta, float* dx, int M, int D, int C, float eps,
                  wf_stream_t stream);

/* Wide per-point layers on the 5th-gen tensor cores (tcgen05.mma, TMA-fed, TMEM accumulators):
 *   D[M,N] (+)= A * B^T (+ bias)
 * a_kmajor/b_kmajor = 1: operand stored [rows, K] with K contiguous (activations, weights);
 * = 0: operand stored [K, rows] with rows contiguous (the dW = dZ^T * H reduction over points).
 * out_dtype WF_BF16 (store) or WF_F32 (store, or atomic accumulate when accumulate != 0, used with
 * split_k > 1).  rowstats (optional, wf_gemm_rowstats_parts(N) * M * 2 floats, WRITTEN): per N tile and
 * row, the sum and sum of squares of the fp32 result incl. bias -- the LayerNorm statistics of
 * models/PointNetEncoder.py:38, added in tile order by wf_stats_finalize (deterministic).
 * Constraints: lda/ldb % 8 == 0 (16-byte TMA strides), pointers 16-byte aligned. */
int wf_gemm_bf16(const void* A, int lda, int a_kmajor, const void* B, int ldb, int b_kmajor, int M,
                 int N, int K, const float* bias, void* D, int ldd, int out_dtype, int accumulate,
                 int split_k, float* rowstats, wf_stream_t stream);

/* Same kernel on fp32 storage, multiplied as TF32 (kind::tf32, fp32 accumulate): the heads' Linear layers and their
 * dX / dW products in production precision.  Mixed operand majorness is allowed (dX = dY * W reads W as stored).
 * Constraints: lda/ldb/ldd % 4 == 0, pointers 16-byte aligned. */
int wf_gemm_tf32(const float* A, int lda, int a_kmajor, const float* B, int ldb, int b_kmajor, int M,
                 int N, int K, const float* bias, float* D, int ldd, int accumulate, int split_k,
                 wf_stream_t stream);

/* wf_gemm_tf32 with a DETERMINISTIC split-K: each K range stores its partial tile into its own [M,N] fp32 slice of `work`
 * (work_floats >= split_k * M * N), a second kernel adds the slices in order (+ bias, + D when accumulate != 0).  Used for
 * the forward and dX products of the 64-row heads, which have too few output tiles to fill the GPU but must stay
 * bit-reproducible (only weight gradients use the atomic split-K).  N % 4 == 0. */
int wf_gemm_tf32_splitk(const float* A, int lda, int a_kmajor, const float* B, int ldb, int b_kmajor,
                        int M, int N, int K, const float* bias, float* D, int ldd, int accumulate,
                        int split_k, float* work, int64_t work_floats, wf_stream_t stream);

/* bf16 LayerNorm+ReLU passes between the tensor-core layers (models/PointNetEncoder.py:38-39), 16-byte vectorised,
 * HBM-bound.  fwd: h = relu(LN(z)) with the row statistics from the GEMM epilogue.  bwd: dz from dh and z in ONE pass;
 * dgamma/dbeta/dcolsum (the Linear's bias gradient) are accumulated (caller zeroes).  C in {512,1024,2048}. */
int wf_ln_relu_bf16_fwd(const void* z, const float* mean, const float* rstd, const float* gamma,
                        const float* beta, void* h, int M, int C, wf_stream_t stream);
int wf_ln_relu_bf16_bwd(const void* dh, const void* z, const float* mean, const float* rstd,
                        const float* gamma, const float* beta, void* dz, float* dgamma, float* dbeta,
                        float* dcolsum, int M, int C, wf_stream_t stream);

/* rowstats[parts][M] (sum, sumsq) -> (mean, rstd) per row over C columns */
int wf_gemm_rowstats_parts(int N);
int wf_stats_finalize(const float* rowstats, int M, int C, int parts, float eps, float* mean,
                      float* rstd, wf_stream_t stream);

/* fp32 -> bf16 cast, optionally transposed ([R,C] -> [C,R]) -- weight staging for the GEMMs */
int wf_cast_bf16(const float* src, int R, int C, void* dst, int transpose, wf_stream_t stream);

/* models/PointNetEncoder.py:103-111 + models/VertexPredictor.py:86-87: the four reductions of the
 * (B,N,C) point-feature tensor -- masked max(+first argmax)/mean and unmasked max(+argmax)/mean. */
int wf_pool_fwd(const float* pf, const uint8_t* mask, const float* valid, int B, int N, int C,
                float* max_m, int32_t* arg_m, float* avg_m, float* max_u, int32_t* arg_u,
                float* mean_u, wf_stream_t stream);
/* Backward: d_pf[b,n,c] = mask*g_avg/valid + g_mean/N + [n==arg_m]*g_max_m*finite + [n==arg_u]*g_max_u.
 * dbias (optional, C floats, accumulated; bf16 output only) = sum over all points of d_pf in closed form: the bias
 * gradient of the Linear that produced the point features. */
int wf_pool_bwd(const float* g_max_m, const float* g_avg_m, const float* g_max_u,
                const float* g_mean_u, const int32_t* arg_m, const int32_t* arg_u,
                const uint8_t* mask, const float* valid, int B, int N, int C, void* d_pf,
                int d_dtype, float* dbias, wf_stream_t stream);

/* Pooled reductions WITHOUT the (B,N,C) point-feature tensor (production path; north_star "the NxC activation
 * tensor never reaches HBM").
 * wf_gemm_bf16_pool: the final Linear of the per-point MLP (models/PointNetEncoder.py:94) as wf_gemm_bf16 (K-major
 *   bf16 operands) whose epilogue does not store D: for every (cloud, channel) it atomically maximises a packed
 *   64-bit word (order-preserving float bits << 32 | ~row-in-cloud) over all rows (max_u) and over rows with
 *   mask != 0 (max_m) -> max and FIRST argmax, bit-identical to a max over the stored tensor
 *   (models/PointNetEncoder.py:108-110, models/VertexPredictor.py:87).  Both outputs are [clouds, N] words the caller
 *   zeroes; a cloud = points_per_cloud (>= 32) consecutive rows, row_offset = global index of row 0 (chunked calls),
 *   index_offset is added to the stored point index (clouds sharded by points across ranks: SURVEY 8e config 4; the
 *   packed words of all ranks then combine with an integer MAX all-reduce).
 * wf_ln_relu_bf16_fwd_colsum: wf_ln_relu_bf16_fwd that also writes per-row-block column sums of h (all rows / valid
 *   rows) to `part` (wf_seg_part_floats(total_rows, C) floats); wf_seg_mean adds them in block order into
 *   hbar[2][B][C] = mean over all rows, mean over valid rows (sum / valid[b]).  The mean pools of the point features
 *   are then Linear(hbar) (models/PointNetEncoder.py:103-105, models/VertexPredictor.py:86: an affine map commutes
 *   with the mean).  points_per_cloud >= 128, row_offset % 128 == 0.
 * wf_pool_finalize: decodes the packed maxima, adds the bias: lin[2][B][C] = hbar W^T (no bias). */
int wf_gemm_bf16_pool(const void* A, int lda, const void* B, int ldb, int M, int N, int K,
                      const float* bias, int points_per_cloud, int row_offset, int index_offset,
                      const uint8_t* mask, uint64_t* max_u, uint64_t* max_m, wf_stream_t stream);
int wf_ln_relu_bf16_fwd_colsum(const void* z, const float* mean, const float* rstd,
                               const float* gamma, const float* beta, void* h, const uint8_t* mask,
                               int M, int C, int points_per_cloud, int row_offset, float* part,
                               wf_stream_t stream);
int wf_seg_part_floats(int total_rows, int C);
int wf_seg_mean(const float* part, const float* valid, int B, int points_per_cloud, int C,
                float* hbar, wf_stream_t stream);
int wf_pool_finalize(const uint64_t* packed_u, const uint64_t* packed_m, const float* lin,
                     const float* bias, int B, int C, float* max_m, int32_t* arg_m, float* avg_m,
                     float* max_u, int32_t* arg_u, float* mean_u, wf_stream_t stream);

/* Side jobs: HBM-bound row passes executed by 128 spare threads INSIDE the persistent tensor-core GEMM kernel.
 * The wide GEMMs of the per-point MLP (models/PointNetEncoder.py:37-45) are bound by the tensor pipe and leave HBM almost
 * idle; the LayerNorm+ReLU passes between them (:38-39, and their backward) are bound by HBM and leave the tensor pipe
 * idle.  A GEMM launch on one row chunk therefore carries the LayerNorm pass of ANOTHER row chunk (already produced by an
 * earlier launch): both finish in about the time of the longer one.  A segment is rows [0, rows) of its own tensors (the
 * caller offsets the pointers); the rows of every segment are dealt across the CTAs of the launch.
 *   WF_SIDE_LN_FWD         h = relu(LN(z))                         == wf_ln_relu_bf16_fwd         (bit-identical)
 *   WF_SIDE_LN_FWD_COLSUM  ... + per-row-block column sums of h    == wf_ln_relu_bf16_fwd_colsum  (bit-identical)
 *   WF_SIDE_LN_BWD         dz from dh and z, dgamma/dbeta/dbias    == wf_ln_relu_bf16_bwd         (same math; the row sums are
 *                          accumulated atomically                     added in a different order: ~1e-7 relative before rounding)
 * C in {1024, 2048}; up to WF_SIDE_MAX segments per launch, executed in order. */
#define WF_SIDE_LN_FWD 1
#define WF_SIDE_LN_FWD_COLSUM 2
#define WF_SIDE_LN_BWD 3
#define WF_SIDE_MAX 4
typedef struct wf_side_seg {
    int32_t kind;            /* WF_SIDE_* */
    int32_t C;               /* row width (channels) */
    int64_t rows;            /* rows of this segment */
    const void* x0;          /* fwd: z [rows, C] bf16            bwd: dh [rows, C] bf16 */
    const void* x1;          /* fwd: unused                      bwd: z  [rows, C] bf16 */
    const float* mean;       /* [rows] */
    const float* rstd;       /* [rows] */
    const float* gamma;      /* [C] */
    const float* beta;       /* [C] */
    void* out;               /* fwd: h [rows, C] bf16            bwd: dz [rows, C] bf16 */
    float* acc0;             /* bwd: dgamma [C] (accumulated) */
    float* acc1;             /* bwd: dbeta  [C] (accumulated) */
    float* acc2;             /* bwd: dbias  [C] (accumulated) */
    const uint8_t* mask;     /* colsum: [rows] validity, or NULL */
    float* part;             /* colsum: the whole `part` buffer (indexed by global row block) */
    int32_t pool_n;          /* colsum: points per cloud */
    int32_t row_off;         /* colsum: global row index of the segment's row 0 (multiple of 128) */
} wf_side_seg;

/* wf_gemm_bf16 / wf_gemm_bf16_pool (pool_n > 0: the pooling epilogue, D unused) carrying side-job segments.  Needs the
 * 2-SM kernel form (M >= 256 rows); the GEMM results are bit-identical to the plain entry points. */
int wf_gemm_bf16_side(const void* A, int lda, int a_kmajor, const void* B, int ldb, int b_kmajor, int M,
                      int N, int K, const float* bias, void* D, int ldd, int out_dtype, int accumulate,
                      int split_k, float* rowstats, int pool_n, int pool_row_offset, int pool_index_offset,
                      const uint8_t* pool_mask, uint64_t* pool_max_u, uint64_t* pool_max_m,
                      const wf_side_seg* segs, int n_segs, wf_stream_t stream);

/* Linear + LayerNorm + ReLU of models/PointNetEncoder.py:37-40 in ONE launch: Z = A * B^T + bias (bf16, stored: the backward
 * needs it) exactly as wf_gemm_bf16 with row statistics, and H = relu(LN(Z)) produced by the side warps of the same kernel
 * from the 256-row units of Z that have just been completed by all their N tiles -- read back from L2, not from DRAM.
 * mean / rstd [M] are finalised in-kernel (same expression as wf_stats_finalize) and written for the backward.
 * `done`: ceil(M / 256) int32 counters, zeroed by the caller before every launch.  N in {1024, 2048}, M >= 256, K-major
 * operands.  Z, H, mean, rstd are bit-identical to wf_gemm_bf16 + wf_stats_finalize + wf_ln_relu_bf16_fwd.  May carry
 * further side segments (other rows' passes), executed after the own-output work. */
int wf_gemm_bf16_ownln(const void* A, int lda, const void* B, int ldb, int M, int N, int K,
                       const float* bias, void* Z, float* rowstats, const float* gamma, const float* beta,
                       void* H, float* mean, float* rstd, float eps, int32_t* done,
                       const wf_side_seg* segs, int n_segs, wf_stream_t stream);
/* Backward of the four pools THROUGH the final Linear (W [C,K] fp32, h [B*N,K] bf16 its input) without dense GEMMs:
 * the pooled gradients are per-cloud constants plus <= 2C single entries at the argmax rows (train.py:140 autograd
 * reaches the same numbers through a dense (B*N,C) gradient).  dbar[2][B][K] = [g_mean_u; g_avg_m] W (caller).
 *   dh[B*N,K] bf16 (written) = dbar_u/N + mask*dbar_m/valid + sum_{c: argmax(b,c)=n} g(b,c) W[c,:]
 *   dW[C,K] += sum_b g(b,c) h[b,argmax(b,c),:] (caller pre-fills it with G^T hbar);  db[C] written.
 * work: wf_pool_fused_bwd_work_ints(B, C) int32 of scratch.  Deterministic (no floating-point atomics). */
int wf_pool_fused_bwd(const float* g_max_m, const float* g_avg_m, const float* g_max_u,
                      const float* g_mean_u, const int32_t* arg_m, const int32_t* arg_u,
                      const uint8_t* mask, const float* valid, const float* dbar, const float* W,
                      const void* h, int B, int N, int C, int K, int32_t* work, void* dh, float* dW,
                      float* db, wf_stream_t stream);
int wf_pool_fused_bwd_work_ints(int B, int C);

/* ------------------------------------------------------------------------------------------
 * Target preparation (train.py:48-88,112-115): ragged ground truth -> the four target tensors of
 * WireframeLoss in ONE launch.  verts: all samples' vertices concatenated [T,3]; v_off[B+1]; edges: all samples' edges
 * concatenated [total_edges,2] as float32 vertex indices (the dataset's dtype, datasets/building3d.py:180-183); e_off[B+1].
 * Writes tgt_vertices [B,V,3] (zero padded), tgt_existence [B,V], counts [B] (int64, = len(vertices) per sample) and
 * edge_labels [B,max_e]: label 1 at the row-major index of pair (min,max) among the pairs i<j<count
 * (models/EdgePredictor.py:84-89 order), zero elsewhere; edges touching a vertex >= count, self loops and duplicates
 * behave as in the reference's set lookup.  max_e = max_b count_b*(count_b-1)/2 (host-known). */
int wf_pack_targets(const float* verts, const int32_t* v_off, const float* edges, const int32_t* e_off,
                    int B, int V, int max_e, int total_edges, float* tgt_vertices, float* tgt_existence,
                    int64_t* counts, float* edge_labels, wf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Edge head: ragged batch, vertices of all samples concatenated,
 * v_off[B+1] vertex prefix offsets, e_off[B+1] edge prefix offsets (int64), T = v_off[B]
 * (the host mirror knows T: counts are an input in training and one D2H read in inference).
 * ---------------------------------------------------------------------------------------- */

/* predicted_vertices[b, :count_b] -> packed [T,3] (models/PointCloudToWireframe.py:81,91) */
int wf_gather_prefix(const float* verts, int B, int V, const int32_t* v_off, int T, float* packed,
                     wf_stream_t stream);
int wf_scatter_prefix_add(const float* d_packed, int B, int V, const int32_t* v_off, int T,
                          float* d_verts, wf_stream_t stream);

/* 8-head self-attention core of nn.MultiheadAttention (models/EdgePredictor.py:109-111):
 * qkv [T,3*E] (in_proj output), one CTA per (sample, head); out [T,E] is the concatenated heads
 * (input of out_proj).  probs (optional) saves softmax for the backward at p_off[b] + h*c*c.
 * keep (optional) is the attention-dropout keep mask in the same layout. */
int wf_attn_fwd(const float* qkv, const int32_t* v_off, const int64_t* p_off, int B, int heads,
                int head_dim, int max_c, float* out, float* probs, const uint8_t* keep,
                float keep_scale, wf_stream_t stream);
int wf_attn_bwd(const float* d_out, const float* qkv, const float* probs, const int32_t* v_off,
                const int64_t* p_off, int B, int heads, int head_dim, int max_c, float* d_qkv,
                const uint8_t* keep, float keep_scale, wf_stream_t stream);

/* All-pairs first edge layer without materialising the (E,1031) concat
 * (models/EdgePredictor.py:117-134 + edge_mlp.0): z1[e] = P[i] + Q[j] + wd*|v_i - v_j|_2 + b,
 * pairs (i<j) in the reference's row-major order.  P,Q are the per-vertex halves of the layer. */
int wf_edge_pair_fwd(const float* P, const float* Q, const float* verts, const float* wd,
                     const float* bias, const int32_t* v_off, const int64_t* e_off, int B, int T,
                     int C, int ld, float* z1, float* dist, wf_stream_t stream);
/* dP,dQ [T,C] written; d_verts [T,3] and d_wd[C] accumulated (caller zeroes).  The bias gradient
 * is the column sum of dP (every pair contributes once to exactly one dP row): use wf_colsum.
 * ld: row stride in floats of P, Q (forward) and dP, dQ (backward) -- 2C when they are the two halves of ONE
 * stacked [T, 2C] product, so that neither the halves nor their gradients are ever copied apart. */
int wf_edge_pair_bwd(const float* dz1, const float* dist, const float* verts, const float* wd,
                     const int32_t* v_off, const int64_t* e_off, int B, int T, int C, int ld, float* dP,
                     float* dQ, float* d_verts, float* d_wd, wf_stream_t stream);

/* edge_mlp.10 + sigmoid + zero-padding to (B, max_e): models/EdgePredictor.py:137-138,
 * models/PointCloudToWireframe.py:103-112. */
int wf_edge_out_fwd(const float* h, const float* w, const float* bias, const int64_t* e_off, int B,
                    int K, int max_e, float* probs, wf_stream_t stream);
int wf_edge_out_bwd(const float* d_probs, const float* probs, const float* h, const float* w,
                    const int64_t* e_off, int B, int K, int max_e, float* dh, float* dw, float* db,
                    wf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Loss: losses/WireframeLoss.py:38-104,248-283
 * ---------------------------------------------------------------------------------------- */

/* out[0..3] = total, vertex (matched SmoothL1), existence BCE, edge BCE; out[4] = match count;
 * out must hold wf_loss_out_floats() floats (per-CTA partial sums follow, summed in fixed order).
 * col_of_row is wf_loss_match's output.  min_e = min(Ep, El) columns enter the edge term. */
int wf_loss_out_floats(void);
int wf_loss_fwd(const float* pred_v, const float* pred_e, const float* edge_p, const float* tgt_v,
                const float* tgt_e, const float* edge_l, const int32_t* col_of_row,
                const int64_t* counts, int B, int V, int Vt, int Ep, int El, float w_vertex,
                float w_edge, float w_exist, float* out, wf_stream_t stream);
/* g_out[4] (device): upstream gradients of (total, vertex, existence, edge). */
int wf_loss_bwd(const float* g_out, const float* pred_v, const float* pred_e, const float* edge_p,
                const float* tgt_v, const float* tgt_e, const float* edge_l,
                const int32_t* col_of_row, const int64_t* counts, const float* fwd_out, int B, int V,
                int Vt, int Ep, int El, float w_vertex, float w_edge, float w_exist, float* d_pred_v,
                float* d_pred_e, float* d_edge_p, wf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Evaluation post-processing (eval/ap_calculator.py, called from evaluate.py:110).
 * fp64 like numpy/scipy, ragged batches: sample b owns rows off[b]..off[b+1] of each input.
 * ---------------------------------------------------------------------------------------- */

/* eval/ap_calculator.py:8-36 hausdorff_distance_line.  Lines are (L,2,3) doubles holding
 * [start, end-start] (the difference is formed by the host in the segments' own dtype, as numpy does
 * at :24-25); weights[samples] = np.linspace(0,1,samples).  Block b of `out` (at out_off[b], row-major
 * n_pred_b x n_tgt_b) = max(h(pred->tgt), h(tgt->pred)) over the sampled points; bit-equal to the
 * reference's cdist/min/max chain.  max_p = max_b n_pred_b. */
int wf_hausdorff_lines(const double* p_lines, const int64_t* p_off, const double* t_lines,
                       const int64_t* t_off, const int64_t* out_off, int B, int max_p,
                       const double* weights, int samples, double* out, wf_stream_t stream);

/* scipy.spatial.distance.cdist(a_b, b_b) ('euclidean', double) per sample
 * (eval/ap_calculator.py:45,192,225,249); max_block = max_b n_a*n_b. */
int wf_cdist_f64(const double* a, const int64_t* a_off, const double* b, const int64_t* b_off,
                 const int64_t* out_off, int B, int64_t max_block, int dim, double* out,
                 wf_stream_t stream);

/* scipy.optimize.linear_sum_assignment on B fp64 matrices of any shape (eval/ap_calculator.py:161,
 * 193,250), one CTA per matrix: matrix b is nr[b] x nc[b] row-major at cost + c_off[b]; `work` is a
 * scratch buffer of the same size and offsets (holds the transpose of tall matrices).
 * col_of_row[r_off[b]+i] = column of row i or -1; matched_cost (optional, same indexing) = the cost
 * of that pair (what the reference reads back as cost[row_ind, col_ind] at :163,195,251), so the
 * matrices themselves never leave the device; status[b] = WF_LSAP_*.  WF_ETOOBIG when
 * max(max_nr,max_nc) needs more than 200 KB of shared memory (~ 9 000). */
int wf_lsap_f64(const double* cost, const int64_t* c_off, const int32_t* nr, const int32_t* nc,
                const int64_t* r_off, int B, int max_nr, int max_nc, double* work,
                int32_t* col_of_row, double* matched_cost, int32_t* status, wf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* WF_B200_H */
