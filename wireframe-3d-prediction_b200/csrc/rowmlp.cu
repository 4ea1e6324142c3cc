// libwf_b200 -- small-batch row-MLP products (the 64-row feature-fusion and vertex heads,
// models/PointNetEncoder.py:57-65,115-116 and models/VertexPredictor.py:94-117).
//
// With B <= 128 rows these layers are weight streaming: 23.7 M parameters (95 MB fp32) are read once by the forward, once
// by dX, and written once as dW, against 3 GFLOP per pass.  The generic tensor-core GEMM (gemm_tc.cu, kind::tf32) treats them
// as M x N x K problems with 1..16 output tiles and needs split-K with a workspace + reduce launch (forward, dX) or atomics on
// a zero-filled output (dW, whose REDUCTION dimension is the 64 batch rows: two k-blocks per tile, all prologue and epilogue).
// The kernels here are shaped for the streaming instead:
//
//   wf_rowmlp_dw      dW[N][K] = dZ^T X (+ db = column sums of dZ): one CTA per 64 x 256 output tile, both operands (64 rows
//                     each) staged once in shared memory, every output written exactly once (no atomics, no zero fill).
//   wf_rowmlp_linear  Y[M][N] = X W^T (+ bias) (forward) or dX[M][K] = dZ W (W read in place, transposed while staging):
//                     a cluster of 8 CTAs per 64-column output slab, each CTA reduces one eighth of the K range, the eight
//                     partial tiles are added over distributed shared memory in rank order (deterministic), so 8 x more CTAs
//                     stream the weights than there are output slabs and the activations are read once per slab.
//
// Arithmetic: TF32 operands (cvt.rna once, while staging), fp32 accumulation, mma.sync.m16n8k8 -- the same precision class as
// the kind::tf32 path they replace in the production ("bf16") mode; the fp32 parity mode keeps wf_gemm_f32.
#include "wf_common.cuh"

#include <cooperative_groups.h>

namespace wf {
namespace rowmlp {

namespace cg = cooperative_groups;

__device__ __forceinline__ float tf32_rn(float v) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v)); return __uint_as_float(r); }
__device__ __forceinline__ float4 tf32_rn4(float4 v) { return make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w)); }
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t fu(float v) { return __float_as_uint(v); }

// ------------------------------------------------------------------------------------------
// dW[n][k] = sum_b A[b][n] * B[b][k]   (A = dZ [Mb, Nr], B = X [Mb, Kc], Mb <= 128 rows, zero padded to a multiple of 8)
// CTA: 256 threads, output tile 64 (n) x 256 (k); warp w owns columns [32 w, 32 w + 32) as 4 m-tiles x 4 n-tiles.
// MMA roles: M = n (rows of dW), N = k, K = b.  Fragments come from 16-byte shared loads by permuting the row / column
// assignment of the tiles: m-tile q holds rows n0 + 4 r + q (r = fragment row 0..15), n-tile p columns k0 + 4 c + p
// (c = fragment column 0..7) -- so one float4 at [b][4 r ..] feeds the four m-tiles, one at [b][4 c ..] the four n-tiles,
// and a thread's four n-tiles of an accumulator row are 4 consecutive floats of dW (16-byte stores, 128 B per row group).
// Row strides = 8 mod 32 floats: the four k rows x two 16-byte column groups of a quarter warp fall on 8 distinct bank groups.
// ------------------------------------------------------------------------------------------
constexpr int DW_TN = 64, DW_TK = 256, DW_SA = DW_TN + 8, DW_SB = DW_TK + 8, DW_MB = 128;

__global__ void __launch_bounds__(256)
dw_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, int Mb, int Nr, int Kc,
          float* __restrict__ C, int ldc, float* __restrict__ colsum) {
    extern __shared__ __align__(16) float smem[];
    const int mbp = (Mb + 7) & ~7;
    float* As = smem;                              // [mbp][DW_SA]
    float* Bs = smem + (size_t)mbp * DW_SA;        // [mbp][DW_SB]
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, gid = lane >> 2, tig = lane & 3;
    const int n0 = blockIdx.x * DW_TN, k0 = blockIdx.y * DW_TK;
    // ---- stage (TF32-rounded); columns beyond Nr / Kc and rows beyond Mb are zero
    for (int i = t; i < mbp * (DW_TN / 4); i += 256) {
        const int b = i / (DW_TN / 4), c4 = (i - b * (DW_TN / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b < Mb && n0 + c4 < Nr) v = *reinterpret_cast<const float4*>(A + (size_t)b * lda + n0 + c4);     // Nr % 4 == 0
        *reinterpret_cast<float4*>(As + (size_t)b * DW_SA + c4) = tf32_rn4(v);
    }
    for (int i = t; i < mbp * (DW_TK / 4); i += 256) {
        const int b = i / (DW_TK / 4), c4 = (i - b * (DW_TK / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b < Mb && k0 + c4 < Kc) v = *reinterpret_cast<const float4*>(B + (size_t)b * ldb + k0 + c4);     // Kc % 4 == 0
        *reinterpret_cast<float4*>(Bs + (size_t)b * DW_SB + c4) = tf32_rn4(v);
    }
    // db: exact fp32 column sums of dZ, by the CTAs of the first k tile
    if (colsum != nullptr && blockIdx.y == 0 && t < DW_TN && n0 + t < Nr) {
        float s = 0.f;
        for (int b = 0; b < Mb; ++b) s += A[(size_t)b * lda + n0 + t];
        colsum[n0 + t] = s;
    }
    __syncthreads();
    float acc[4][4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[q][p][e] = 0.f;
    const int wc = warp * 32;
    for (int b0 = 0; b0 < mbp; b0 += 8) {
        const float* ar0 = As + (size_t)(b0 + tig) * DW_SA + 4 * gid;
        const float* ar1 = ar0 + 4 * DW_SA;
        const float4 a00 = *reinterpret_cast<const float4*>(ar0), a01 = *reinterpret_cast<const float4*>(ar0 + 32);
        const float4 a10 = *reinterpret_cast<const float4*>(ar1), a11 = *reinterpret_cast<const float4*>(ar1 + 32);
        const float4 bb0 = *reinterpret_cast<const float4*>(Bs + (size_t)(b0 + tig) * DW_SB + wc + 4 * gid);
        const float4 bb1 = *reinterpret_cast<const float4*>(Bs + (size_t)(b0 + tig + 4) * DW_SB + wc + 4 * gid);
        const float x00[4] = {a00.x, a00.y, a00.z, a00.w}, x01[4] = {a01.x, a01.y, a01.z, a01.w};
        const float x10[4] = {a10.x, a10.y, a10.z, a10.w}, x11[4] = {a11.x, a11.y, a11.z, a11.w};
        const float y0[4] = {bb0.x, bb0.y, bb0.z, bb0.w}, y1[4] = {bb1.x, bb1.y, bb1.z, bb1.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int p = 0; p < 4; ++p)
                // a0: (row g, k t), a1: (row g + 8, k t), a2: (row g, k t + 4), a3: (row g + 8, k t + 4);  b0: (k t, col g), b1: (k t + 4, col g)
                mma_tf32(acc[q][p], fu(x00[q]), fu(x01[q]), fu(x10[q]), fu(x11[q]), fu(y0[p]), fu(y1[p]));
    }
    // ---- store: accumulator (q, p): c0 (row g, col 2t), c1 (row g, col 2t + 1), c2 / c3 rows g + 8
    //      actual row n0 + 4 r + q, actual column k0 + wc + 4 c + p  ->  the four p of one (row, c) are consecutive
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int hr = 0; hr < 2; ++hr) {
            const int n = n0 + 4 * (gid + 8 * hr) + q;
            if (n >= Nr) continue;
#pragma unroll
            for (int hc = 0; hc < 2; ++hc) {
                const int k = k0 + wc + 4 * (2 * tig + hc);
                if (k >= Kc) continue;
                const int e = 2 * hr + hc;
                *reinterpret_cast<float4*>(C + (size_t)n * ldc + k) = make_float4(acc[q][0][e], acc[q][1][e], acc[q][2][e], acc[q][3][e]);
            }
        }
}

// ------------------------------------------------------------------------------------------
// Y[m][n] = sum_k X[m][k] * Wop[n][k]  (+ bias[n]);  Wop[n][k] = W[n * ldw + k] (forward: W is [N][K]) or W[k * ldw + n]
// (TRANS: dX = dZ W with W [K = reduction][N = outputs] read in place, transposed while staging).
// Cluster of 8 CTAs per 64-column slab of Y; CTA rank r reduces k in [r K/8, (r+1) K/8) in chunks of 64 through a two-stage
// shared-memory ring; 8 warps = 4 m-tiles (16 rows each: M <= 64 per launch row block) x 2 column halves (4 n-tiles).
// Both staged tiles are [row][k] with k contiguous (stride 80 floats: conflict-free 16-byte fragment loads); a float4 at
// [row][kc + 4 t ..] supplies (k t, k t + 4) of two consecutive MMAs -- the k index is summed, so the same permutation on
// both operands leaves the product unchanged.  The partial 64 x 64 tiles are exchanged over DSMEM: rank r adds rows
// [8 r, 8 r + 8) of all eight partials in rank order and writes them with the bias.
// ------------------------------------------------------------------------------------------
constexpr int LN_BN = 64, LN_KC = 64, LN_S = LN_KC + 16, LN_ST = LN_BN + 2, LN_CL = 8, LN_PS = LN_BN + 4;
constexpr int LN_WBUF = 64 * LN_S;                 // floats per W stage (>= 64 * LN_ST)

template <bool TRANS>
__global__ void __launch_bounds__(256)
linear_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw, const float* __restrict__ bias,
              int M, int N, int K, float* __restrict__ Y, int ldy) {
    extern __shared__ __align__(16) float smem[];
    float* Xs = smem;                                  // [2][64][LN_S]
    float* Ws = smem + 2 * 64 * LN_S;                  // [2][64][LN_S] (rows = outputs, k contiguous) or [2][64][LN_ST] (TRANS: rows = k)
    float* Ps = smem + 2 * 64 * LN_S + 2 * LN_WBUF;    // [64][LN_PS] partial tile of this CTA
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, gid = lane >> 2, tig = lane & 3;
    const int n0 = (blockIdx.x / LN_CL) * LN_BN, m0 = blockIdx.y * 64;
    const int kper = ((K + LN_CL * 8 - 1) / (LN_CL * 8)) * 8;          // k range per rank, a multiple of 8 (K % 4 == 0 -> ke % 4 == 0)
    const int kb = rank * kper, ke = min(K, kb + kper);
    const int mt = warp & 3, nh = warp >> 2;                           // m-tile, column half
    float acc[4][4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[p][e] = 0.f;

    // a chunk [kc, kc + 64) of this rank's range is fetched into registers (16-byte loads) TWO chunks ahead of its use -- a
    // chunk's MMAs (~400 clk per warp) do not cover a global-load latency, two do --, rounded to TF32 and stored one chunk
    // ahead; rows / columns outside the operands arrive as zeros
    struct Regs { float4 x[4], w[4]; };
    auto fetch = [&](Regs& rg, int kc) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int id = t + i * 256, r = id >> 4, c4 = (id & 15) * 4;
            rg.x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + r < M && kc + c4 < ke) rg.x[i] = *reinterpret_cast<const float4*>(X + (size_t)(m0 + r) * ldx + kc + c4);
            rg.w[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!TRANS) { if (n0 + r < N && kc + c4 < ke) rg.w[i] = *reinterpret_cast<const float4*>(W + (size_t)(n0 + r) * ldw + kc + c4); }
            else        { if (kc + r < ke && n0 + c4 < N) rg.w[i] = *reinterpret_cast<const float4*>(W + (size_t)(kc + r) * ldw + n0 + c4); }
        }
    };
    auto store = [&](const Regs& rg, int buf) {
        float* xs = Xs + buf * 64 * LN_S;
        float* ws = Ws + buf * LN_WBUF;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int id = t + i * 256, r = id >> 4, c4 = (id & 15) * 4;
            *reinterpret_cast<float4*>(xs + r * LN_S + c4) = tf32_rn4(rg.x[i]);
            const float4 v = tf32_rn4(rg.w[i]);
            if (!TRANS) *reinterpret_cast<float4*>(ws + r * LN_S + c4) = v;
            else {                                                     // [k][n], stride 66: 8-byte stores, conflict-free scalar fragment loads
                *reinterpret_cast<float2*>(ws + r * LN_ST + c4) = make_float2(v.x, v.y);
                *reinterpret_cast<float2*>(ws + r * LN_ST + c4 + 2) = make_float2(v.z, v.w);
            }
        }
    };
    auto mma_chunk = [&](int buf) {
        const float* xs = Xs + buf * 64 * LN_S + (mt * 16 + gid) * LN_S + 4 * tig;
        const float* wsb = Ws + buf * LN_WBUF;
#pragma unroll
        for (int k16 = 0; k16 < LN_KC; k16 += 16) {
            // X float4 at [row][k16 + 4 t ..]: elements (x, y) are the (k t, k t + 4) slots of the first MMA, (z, w) of the second --
            // the k index is summed over, the same assignment is used for W
            const float4 xa = *reinterpret_cast<const float4*>(xs + k16), xb = *reinterpret_cast<const float4*>(xs + 8 * LN_S + k16);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                float4 wv;
                if (!TRANS) wv = *reinterpret_cast<const float4*>(wsb + (nh * 32 + p * 8 + gid) * LN_S + 4 * tig + k16);
                else {
                    const float* wp = wsb + (k16 + 4 * tig) * LN_ST + nh * 32 + p * 8 + gid;
                    wv = make_float4(wp[0], wp[LN_ST], wp[2 * LN_ST], wp[3 * LN_ST]);
                }
                mma_tf32(acc[p], fu(xa.x), fu(xb.x), fu(xa.y), fu(xb.y), fu(wv.x), fu(wv.y));
                mma_tf32(acc[p], fu(xa.z), fu(xb.z), fu(xa.w), fu(xb.w), fu(wv.z), fu(wv.w));
            }
        }
    };

    // chunk i lives in shared buffer i & 1 and came through register set i & 1
    Regs r0, r1;
    const int nch = kb < ke ? (ke - kb + LN_KC - 1) / LN_KC : 0;
    if (nch > 0) fetch(r0, kb);
    if (nch > 1) fetch(r1, kb + LN_KC);
    if (nch > 0) store(r0, 0);
    __syncthreads();
    for (int i = 0; i < nch; i += 2) {
        // even chunk i (buffer 0): set 0 is free again -> chunk i + 2; then chunk i + 1 (set 1) goes to buffer 1
        if (i + 2 < nch) fetch(r0, kb + (i + 2) * LN_KC);
        mma_chunk(0);
        if (i + 1 < nch) store(r1, 1);
        __syncthreads();
        if (i + 1 >= nch) break;
        // odd chunk i + 1 (buffer 1): set 1 is free -> chunk i + 3; chunk i + 2 (set 0) goes to buffer 0
        if (i + 3 < nch) fetch(r1, kb + (i + 3) * LN_KC);
        mma_chunk(1);
        if (i + 2 < nch) store(r0, 0);
        __syncthreads();
    }
    // ---- partial tile -> own shared memory: c0 (row g, col 2t), c1 (row g, col 2t+1), c2 / c3 rows g + 8
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int col = nh * 32 + p * 8 + 2 * tig, row = mt * 16 + gid;
        *reinterpret_cast<float2*>(Ps + row * LN_PS + col) = make_float2(acc[p][0], acc[p][1]);
        *reinterpret_cast<float2*>(Ps + (row + 8) * LN_PS + col) = make_float2(acc[p][2], acc[p][3]);
    }
    cluster.sync();
    // ---- rank r finishes rows [8 r, 8 r + 8): 8 rows x 64 columns = 128 float4, one per thread t < 128
    if (t < 128) {
        const int row = rank * 8 + (t >> 4), c4 = (t & 15) * 4;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < LN_CL; ++r) {                              // rank order: deterministic
            const float* rp = cluster.map_shared_rank(Ps, r);
            const float4 v = *reinterpret_cast<const float4*>(rp + row * LN_PS + c4);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        if (m0 + row < M && n0 + c4 < N) {                             // N % 4 == 0
            if (bias != nullptr) { const float4 bv = *reinterpret_cast<const float4*>(bias + n0 + c4); s.x += bv.x; s.y += bv.y; s.z += bv.z; s.w += bv.w; }
            *reinterpret_cast<float4*>(Y + (size_t)(m0 + row) * ldy + n0 + c4) = s;
        }
    }
    cluster.sync();                                                    // nobody leaves while its partial tile may still be read
}

}  // namespace rowmlp
}  // namespace wf

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int wf_rowmlp_dw(const float* dZ, int ldz, const float* X, int ldx, int Mb, int Nr, int Kc, float* dW, int ldw,
                            float* db, wf_stream_t stream) {
    using namespace wf;
    if (Nr <= 0 || Kc <= 0) return WF_OK;
    WF_CHECK_ARG(Mb >= 1 && Mb <= rowmlp::DW_MB, "wf_rowmlp_dw: 1 <= rows <= %d (got %d); use wf_gemm_tf32 / wf_gemm_f32", rowmlp::DW_MB, Mb);
    WF_CHECK_ARG(Nr % 4 == 0 && Kc % 4 == 0 && ldz % 4 == 0 && ldx % 4 == 0 && ldw % 4 == 0 && al16(dZ) && al16(X) && al16(dW),
                 "wf_rowmlp_dw: dimensions / leading dimensions must be multiples of 4 floats and pointers 16-byte aligned");
    const int mbp = (Mb + 7) & ~7;
    const size_t smem = (size_t)mbp * (rowmlp::DW_SA + rowmlp::DW_SB) * sizeof(float);
    WF_CUDA(cudaFuncSetAttribute(rowmlp::dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    dim3 grid(cdiv(Nr, rowmlp::DW_TN), cdiv(Kc, rowmlp::DW_TK));
    rowmlp::dw_kernel<<<grid, 256, smem, as_stream(stream)>>>(dZ, ldz, X, ldx, Mb, Nr, Kc, dW, ldw, db);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_rowmlp_linear(const float* X, int ldx, const float* W, int ldw, int trans_w, const float* bias, int M, int N,
                                int K, float* Y, int ldy, wf_stream_t stream) {
    using namespace wf;
    if (M <= 0 || N <= 0) return WF_OK;
    WF_CHECK_ARG(K >= 1 && M <= 65535 * 64, "wf_rowmlp_linear: bad dims");
    WF_CHECK_ARG(N % 4 == 0 && K % 4 == 0 && ldx % 4 == 0 && ldw % 4 == 0 && ldy % 4 == 0 && al16(X) && al16(W) && al16(Y) &&
                     (bias == nullptr || al16(bias)),
                 "wf_rowmlp_linear: dimensions / leading dimensions must be multiples of 4 floats and pointers 16-byte aligned");
    const size_t smem = (size_t)(2 * 64 * rowmlp::LN_S + 2 * rowmlp::LN_WBUF + 64 * rowmlp::LN_PS) * sizeof(float);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cdiv(N, rowmlp::LN_BN) * rowmlp::LN_CL, cdiv(M, 64));
    cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = as_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = rowmlp::LN_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (trans_w) {
        WF_CUDA(cudaFuncSetAttribute(rowmlp::linear_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        WF_CUDA(cudaLaunchKernelEx(&cfg, rowmlp::linear_kernel<true>, X, ldx, W, ldw, bias, M, N, K, Y, ldy));
    } else {
        WF_CUDA(cudaFuncSetAttribute(rowmlp::linear_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        WF_CUDA(cudaLaunchKernelEx(&cfg, rowmlp::linear_kernel<false>, X, ldx, W, ldw, bias, M, N, K, Y, ldy));
    }
    return WF_OK;
}
