// Evaluation post-processing that follows the hot path: the arithmetic of the
// reference's eval/ap_calculator.py on the GPU, fp64 like numpy/scipy, over RAGGED batches of samples.
//
//   hausdorff_kernel  eval/ap_calculator.py:8-36   symmetric Hausdorff distance between sampled segments
//   cdist_kernel      eval/ap_calculator.py:45,192,225,249  scipy cdist(..., 'euclidean')
//   lsap_cta_kernel   eval/ap_calculator.py:161,193,250     scipy linear_sum_assignment, one CTA per matrix
//
// Bit-exactness: every product and sum is rounded separately (__dmul_rn/__dadd_rn: no FMA contraction),
// in scipy's order (d0*d0 + d1*d1) + d2*d2, then an IEEE sqrt.  sqrt_rn is monotone, so the min/max
// reductions of the Hausdorff distance run on SQUARED distances and one sqrt is taken at the end --
// same bits as reducing the square-rooted matrix.  The LSAP kernel is the CTA-wide form of lsap.cu's
// solver (same closed-form restatement of scipy's tie rules) for matrices that do not fit a warp's
// shared memory: e.g. 2016 predicted edges x 90 label edges.  The matrix stays in global memory
// (L2-resident, <= 1.5 MB), the per-column state in shared memory.
#include "wf_common.cuh"

#include <math_constants.h>

namespace wf {
namespace evalpost {

constexpr int MAX_SAMPLES = 64;

// lines: (L, 2, 3) doubles = [start xyz, (end - start) xyz]  (the host forms end-start in the segments'
// own dtype, as numpy does at :24-25).  One CTA per predicted line, threads over the sample's target lines.
template <int S>
__global__ void hausdorff_kernel(const double* __restrict__ p_lines, const long long* __restrict__ p_off,
                                 const double* __restrict__ t_lines, const long long* __restrict__ t_off,
                                 const long long* __restrict__ out_off, const double* __restrict__ weights,
                                 int samples_rt, double* __restrict__ out) {
    const int samples = S > 0 ? S : samples_rt;
    const int b = blockIdx.y;
    const long long p0 = p_off[b], n_p = p_off[b + 1] - p0;
    const long long t0 = t_off[b], n_t = t_off[b + 1] - t0;
    __shared__ double pp[MAX_SAMPLES][3];
    for (long long n = blockIdx.x; n < n_p; n += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < samples * 3) {
            const int i = threadIdx.x / 3, k = threadIdx.x - 3 * i;
            const double* ln = p_lines + (p0 + n) * 6;
            pp[i][k] = __dadd_rn(ln[k], __dmul_rn(weights[i], ln[3 + k]));
        }
        __syncthreads();
        for (long long m = threadIdx.x; m < n_t; m += blockDim.x) {
            const double* ln = t_lines + (t0 + m) * 6;
            const double sx = ln[0], sy = ln[1], sz = ln[2], dx = ln[3], dy = ln[4], dz = ln[5];
            double rowmin[S > 0 ? S : MAX_SAMPLES];
#pragma unroll (S > 0 ? S : 1)
            for (int i = 0; i < (S > 0 ? S : MAX_SAMPLES); ++i) rowmin[i] = CUDART_INF;
            double h_tp = 0.0;                                 // max over target points of min over pred points
#pragma unroll 1
            for (int j = 0; j < samples; ++j) {
                const double w = weights[j];
                const double tx = __dadd_rn(sx, __dmul_rn(w, dx));
                const double ty = __dadd_rn(sy, __dmul_rn(w, dy));
                const double tz = __dadd_rn(sz, __dmul_rn(w, dz));
                double colmin = CUDART_INF;
#pragma unroll (S > 0 ? S : 1)
                for (int i = 0; i < (S > 0 ? S : MAX_SAMPLES); ++i) {
                    if (S == 0 && i >= samples) break;
                    const double ex = __dsub_rn(pp[i][0], tx), ey = __dsub_rn(pp[i][1], ty), ez = __dsub_rn(pp[i][2], tz);
                    const double s = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
                    colmin = s < colmin ? s : colmin;
                    rowmin[i] = s < rowmin[i] ? s : rowmin[i];
                }
                h_tp = colmin > h_tp ? colmin : h_tp;
            }
            double h_pt = 0.0;
#pragma unroll (S > 0 ? S : 1)
            for (int i = 0; i < (S > 0 ? S : MAX_SAMPLES); ++i) {
                if (S == 0 && i >= samples) break;
                h_pt = rowmin[i] > h_pt ? rowmin[i] : h_pt;
            }
            out[out_off[b] + n * n_t + m] = __dsqrt_rn(h_pt > h_tp ? h_pt : h_tp);
        }
    }
}

// a: rows a_off[b]..a_off[b+1] of (.., dim), b likewise; out block b at out_off[b], row-major (n_a, n_b).
__global__ void cdist_kernel(const double* __restrict__ a, const long long* __restrict__ a_off,
                             const double* __restrict__ bmat, const long long* __restrict__ b_off,
                             const long long* __restrict__ out_off, int dim, double* __restrict__ out) {
    const int b = blockIdx.y;
    const long long a0 = a_off[b], na = a_off[b + 1] - a0;
    const long long b0 = b_off[b], nb = b_off[b + 1] - b0;
    const long long total = na * nb;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        const long long i = k / nb, j = k - i * nb;
        const double* x = a + (a0 + i) * dim;
        const double* y = bmat + (b0 + j) * dim;
        double s = 0.0;
        for (int d = 0; d < dim; ++d) {
            const double e = __dsub_rn(x[d], y[d]);
            s = __dadd_rn(s, __dmul_rn(e, e));
        }
        out[out_off[b] + k] = __dsqrt_rn(s);
    }
}

// ---------------------------------------------------------------------------------------------------
// CTA-wide shortest-augmenting-path LSAP on fp64 matrices of any shape (after transposition nr <= nc).
// ---------------------------------------------------------------------------------------------------
constexpr int LSAP_THREADS = 256;
constexpr int LSAP_WARPS = LSAP_THREADS / 32;

struct CtaWork {
    double *u, *v, *dist;
    int *pred, *col_of_row, *row_of_col, *pool;
    uint8_t *row_seen, *col_seen;
};

__host__ __device__ inline size_t cta_work_bytes(int nr, int nc) {
    size_t b = (size_t)(nr + 2 * nc) * sizeof(double);
    b += (size_t)(nr + 3 * nc) * sizeof(int);
    b += (size_t)(nr + nc);
    return (b + 15) & ~(size_t)15;
}

__device__ inline CtaWork cta_carve(uint8_t* base, int nr, int nc) {
    CtaWork w;
    size_t off = 0;
    w.u = reinterpret_cast<double*>(base + off); off += (size_t)nr * sizeof(double);
    w.v = reinterpret_cast<double*>(base + off); off += (size_t)nc * sizeof(double);
    w.dist = reinterpret_cast<double*>(base + off); off += (size_t)nc * sizeof(double);
    w.pred = reinterpret_cast<int*>(base + off); off += (size_t)nc * sizeof(int);
    w.col_of_row = reinterpret_cast<int*>(base + off); off += (size_t)nr * sizeof(int);
    w.row_of_col = reinterpret_cast<int*>(base + off); off += (size_t)nc * sizeof(int);
    w.pool = reinterpret_cast<int*>(base + off); off += (size_t)nc * sizeof(int);
    w.row_seen = base + off; off += (size_t)nr;
    w.col_seen = base + off;
    return w;
}

__device__ __forceinline__ double warp_min_f64(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double y = __shfl_xor_sync(0xffffffffu, x, o);
        x = y < x ? y : x;
    }
    return x;
}

__global__ void __launch_bounds__(LSAP_THREADS)
lsap_cta_kernel(const double* cost, const long long* __restrict__ c_off, const int* __restrict__ nr_arr,
                const int* __restrict__ nc_arr, const long long* __restrict__ r_off, double* work,
                int32_t* __restrict__ col_of_row_out, double* __restrict__ matched_out, int32_t* __restrict__ status) {
    extern __shared__ __align__(16) uint8_t sm[];
    __shared__ double red_d[LSAP_WARPS];
    __shared__ int red_first[LSAP_WARPS], red_ulast[LSAP_WARPS];
    __shared__ int s_bad;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    const int nr0 = nr_arr[b], nc0 = nc_arr[b];
    int32_t* out = col_of_row_out + r_off[b];
    for (int i = tid; i < nr0; i += LSAP_THREADS) out[i] = -1;
    if (nr0 <= 0 || nc0 <= 0) { if (tid == 0) status[b] = WF_LSAP_OK; return; }
    const bool flip = nc0 < nr0;                                    // tall -> solve the transpose (scipy does)
    const int nr = flip ? nc0 : nr0, nc = flip ? nr0 : nc0;
    const double* src = cost + c_off[b];
    const double* C = src;
    if (tid == 0) s_bad = 0;
    __syncthreads();
    int bad = 0;
    if (flip) {
        double* T = work + c_off[b];
        for (long long k = tid; k < (long long)nr0 * nc0; k += LSAP_THREADS) {
            const long long i = k / nc0, j = k - i * nc0;
            const double c = src[k];
            bad |= (c != c) || (c == -CUDART_INF);
            T[j * nc + i] = c;
        }
        C = T;
    } else {
        for (long long k = tid; k < (long long)nr0 * nc0; k += LSAP_THREADS) {
            const double c = src[k];
            bad |= (c != c) || (c == -CUDART_INF);
        }
    }
    if (bad) s_bad = 1;
    CtaWork w = cta_carve(sm, nr, nc);
    for (int i = tid; i < nr; i += LSAP_THREADS) { w.u[i] = 0.0; w.col_of_row[i] = -1; }
    for (int j = tid; j < nc; j += LSAP_THREADS) { w.v[j] = 0.0; w.row_of_col[j] = -1; w.pred[j] = -1; }
    __syncthreads();                                                // also orders the transposed copy (same CTA)
    if (s_bad) { if (tid == 0) status[b] = WF_LSAP_INVALID; return; }

    for (int cur = 0; cur < nr; ++cur) {
        for (int s = tid; s < nc; s += LSAP_THREADS) { w.pool[s] = nc - 1 - s; w.dist[s] = CUDART_INF; w.col_seen[s] = 0; }
        for (int i = tid; i < nr; i += LSAP_THREADS) w.row_seen[i] = 0;
        __syncthreads();
        int live = nc, sink = -1, row = cur;
        double frontier = 0.0;
        while (sink < 0) {
            if (tid == 0) w.row_seen[row] = 1;
            const double u_row = w.u[row];
            const double* crow = C + (size_t)row * nc;
            double lmin = CUDART_INF;
            for (int s = tid; s < live; s += LSAP_THREADS) {
                const int j = w.pool[s];
                const double cand = __dsub_rn(__dsub_rn(__dadd_rn(frontier, crow[j]), u_row), w.v[j]);
                double d = w.dist[j];
                if (cand < d) { d = cand; w.dist[j] = cand; w.pred[j] = row; }
                lmin = d < lmin ? d : lmin;
            }
            lmin = warp_min_f64(lmin);
            if (lane == 0) red_d[warp] = lmin;
            __syncthreads();
            double m = red_d[0];
#pragma unroll
            for (int k = 1; k < LSAP_WARPS; ++k) m = red_d[k] < m ? red_d[k] : m;
            if (m == CUDART_INF) { if (tid == 0) status[b] = WF_LSAP_INFEASIBLE; return; }
            // scipy's scan keeps the first slot at the minimum unless a later slot at the minimum holds an
            // unassigned column, in which case the LAST such slot wins (see lsap.cu)
            int first = 0x7fffffff, ulast = -1;
            for (int s = tid; s < live; s += LSAP_THREADS) {
                const int j = w.pool[s];
                if (w.dist[j] == m) {
                    first = min(first, s);
                    if (w.row_of_col[j] < 0) ulast = max(ulast, s);
                }
            }
            first = __reduce_min_sync(0xffffffffu, first);
            if (lane == 0) red_first[warp] = first;
            __syncthreads();
            int s1 = red_first[0];
#pragma unroll
            for (int k = 1; k < LSAP_WARPS; ++k) s1 = min(s1, red_first[k]);
            ulast = __reduce_max_sync(0xffffffffu, ulast == s1 ? -1 : ulast);
            if (lane == 0) red_ulast[warp] = ulast;
            __syncthreads();
            int ubest = red_ulast[0];
#pragma unroll
            for (int k = 1; k < LSAP_WARPS; ++k) ubest = max(ubest, red_ulast[k]);
            const int slot = ubest >= 0 ? ubest : s1;
            const int j = w.pool[slot];
            const int owner = w.row_of_col[j];
            __syncthreads();
            if (tid == 0) { w.col_seen[j] = 1; w.pool[slot] = w.pool[live - 1]; }
            --live;
            frontier = m;
            if (owner < 0) sink = j; else row = owner;
            __syncthreads();
        }
        for (int i = tid; i < nr; i += LSAP_THREADS) {
            if (i == cur) w.u[i] = __dadd_rn(w.u[i], frontier);
            else if (w.row_seen[i]) w.u[i] = __dadd_rn(w.u[i], __dsub_rn(frontier, w.dist[w.col_of_row[i]]));
        }
        for (int j = tid; j < nc; j += LSAP_THREADS)
            if (w.col_seen[j]) w.v[j] = __dsub_rn(w.v[j], __dsub_rn(frontier, w.dist[j]));
        __syncthreads();
        if (tid == 0) {
            int j = sink;
            while (true) {
                const int i = w.pred[j];
                const int prev = w.col_of_row[i];
                w.row_of_col[j] = i;
                w.col_of_row[i] = j;
                j = prev;
                if (i == cur) break;
            }
        }
        __syncthreads();
    }
    if (tid == 0) status[b] = WF_LSAP_OK;
    double* mo = matched_out ? matched_out + r_off[b] : nullptr;   // cost of each original row's assignment
    for (int i = tid; i < nr0; i += LSAP_THREADS) {
        const int j = flip ? w.row_of_col[i] : w.col_of_row[i];     // original row i = transposed column i
        out[i] = j;
        if (mo) mo[i] = j >= 0 ? src[(size_t)i * nc0 + j] : 0.0;
    }
}

}  // namespace evalpost
}  // namespace wf

extern "C" int wf_hausdorff_lines(const double* p_lines, const int64_t* p_off, const double* t_lines, const int64_t* t_off,
                                  const int64_t* out_off, int B, int max_p, const double* weights, int samples,
                                  double* out, wf_stream_t stream) {
    using namespace wf;
    if (B <= 0 || max_p <= 0) return WF_OK;
    WF_CHECK_ARG(samples >= 1 && samples <= evalpost::MAX_SAMPLES, "wf_hausdorff_lines: samples must be in 1..%d", evalpost::MAX_SAMPLES);
    const int per_sample = max_p < 4096 ? max_p : 4096;
    dim3 grid(per_sample, B);
    const long long* po = reinterpret_cast<const long long*>(p_off);
    const long long* to = reinterpret_cast<const long long*>(t_off);
    const long long* oo = reinterpret_cast<const long long*>(out_off);
    if (samples == 20)
        evalpost::hausdorff_kernel<20><<<grid, 128, 0, as_stream(stream)>>>(p_lines, po, t_lines, to, oo, weights, samples, out);
    else
        evalpost::hausdorff_kernel<0><<<grid, 192, 0, as_stream(stream)>>>(p_lines, po, t_lines, to, oo, weights, samples, out);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_cdist_f64(const double* a, const int64_t* a_off, const double* b, const int64_t* b_off,
                            const int64_t* out_off, int B, int64_t max_block, int dim, double* out, wf_stream_t stream) {
    using namespace wf;
    if (B <= 0 || max_block <= 0) return WF_OK;
    WF_CHECK_ARG(dim >= 1, "wf_cdist_f64: bad dim");
    int gx = cdiv(max_block, 256);
    if (gx > 1024) gx = 1024;
    dim3 grid(gx, B);
    evalpost::cdist_kernel<<<grid, 256, 0, as_stream(stream)>>>(a, reinterpret_cast<const long long*>(a_off), b,
                                                               reinterpret_cast<const long long*>(b_off),
                                                               reinterpret_cast<const long long*>(out_off), dim, out);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_lsap_f64(const double* cost, const int64_t* c_off, const int32_t* nr, const int32_t* nc,
                           const int64_t* r_off, int B, int max_nr, int max_nc, double* work, int32_t* col_of_row,
                           double* matched_cost, int32_t* status, wf_stream_t stream) {
    using namespace wf;
    using namespace wf::evalpost;
    if (B <= 0) return WF_OK;
    WF_CHECK_ARG(max_nr >= 0 && max_nc >= 0, "wf_lsap_f64: bad sizes");
    const int hi = max_nr < max_nc ? max_nc : max_nr;
    // every problem has, after transposition, rows <= cols <= hi: carve bound (hi, hi)
    const size_t smem = cta_work_bytes(hi > 0 ? hi : 1, hi > 0 ? hi : 1);
    if (smem > 200 * 1024) { set_error("wf_lsap_f64: %d x %d does not fit shared memory (%zu bytes)", max_nr, max_nc, smem); return WF_ETOOBIG; }
    if (smem > 48 * 1024) WF_CUDA(cudaFuncSetAttribute((const void*)lsap_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lsap_cta_kernel<<<B, LSAP_THREADS, smem, as_stream(stream)>>>(cost, reinterpret_cast<const long long*>(c_off), nr, nc,
                                                                  reinterpret_cast<const long long*>(r_off), work, col_of_row, matched_cost, status);
    WF_LAUNCH_CHECK();
    return WF_OK;
}
