// Batched linear-sum-assignment on the GPU: one warp per cost matrix, matrix and dual variables in
// shared memory, fp64 arithmetic, no host round trip.
//
// Replaces the per-sample `.cpu().numpy()` + scipy.optimize.linear_sum_assignment calls at
// losses/WireframeLoss.py:235-236, models/WireframeHungarianMatcher.py:68-71 and
// models/HungarianMatcher.py:124-127.  The algorithm is Crouse's shortest-augmenting-path variant of
// Jonker-Volgenant (IEEE TAES 52(4), 2016) -- the one scipy documents -- and the index choice is
// identical to scipy's, ties included (differentially tested against scipy in tests/).
//
// Parallelisation: the 32 lanes split the pool of unscanned columns of the current Dijkstra step
// (slot s -> lane s%32).  The sequential scan's tie rule "a strictly smaller distance wins; at equal
// distance a still-unassigned column replaces the incumbent" is order dependent, so it is restated
// in closed form: with m the minimum, s1 the first pool slot holding m, the winner is the LAST slot
// holding m whose column is unassigned and that is not s1, else s1.  Two warp reductions per step.
#include "wf_common.cuh"

#include <math_constants.h>

namespace wf {
namespace lsap {

struct Work {          // per-warp shared-memory carve-up (all sizes for nr <= nc after transposition)
    float* cost;       // nr * ldc
    double* u;         // nr
    double* v;         // nc
    double* dist;      // nc
    int* pred;         // nc
    int* col_of_row;   // nr
    int* row_of_col;   // nc
    int* pool;         // nc
    uint8_t* row_seen; // nr
    uint8_t* col_seen; // nc
};

__host__ __device__ inline size_t work_bytes(int nr, int nc) {
    size_t b = 0;
    b += (size_t)nr * nc * sizeof(float);
    b = (b + 7) & ~(size_t)7;
    b += (size_t)(nr + 2 * nc) * sizeof(double);
    b += (size_t)(nr + 3 * nc) * sizeof(int);
    b += (size_t)(nr + nc);
    return (b + 15) & ~(size_t)15;
}

__device__ inline Work carve(uint8_t* base, int nr, int nc) {
    Work w;
    size_t off = 0;
    w.cost = reinterpret_cast<float*>(base); off += (size_t)nr * nc * sizeof(float);
    off = (off + 7) & ~(size_t)7;
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
__device__ __forceinline__ int warp_min_i32(int x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = min(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}
__device__ __forceinline__ int warp_max_i32(int x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = max(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}

// Solve the nr x nc (nr <= nc) problem held in w.cost (leading dimension nc).  Whole warp calls.
// Returns WF_LSAP_*; on success w.col_of_row / w.row_of_col hold the assignment.
__device__ int solve_warp(const Work& w, int nr, int nc, int lane) {
    // validity: NaN or -inf anywhere -> invalid (scipy checks before solving)
    int bad = 0;
    for (int k = lane; k < nr * nc; k += 32) {
        const float c = w.cost[k];
        bad |= (c != c) || (c == -CUDART_INF_F);
    }
    if (__any_sync(0xffffffffu, bad)) return WF_LSAP_INVALID;
    for (int i = lane; i < nr; i += 32) { w.u[i] = 0.0; w.col_of_row[i] = -1; }
    for (int j = lane; j < nc; j += 32) { w.v[j] = 0.0; w.row_of_col[j] = -1; w.pred[j] = -1; }
    __syncwarp();

    for (int cur = 0; cur < nr; ++cur) {
        // ---- Dijkstra sweep from row `cur`
        for (int s = lane; s < nc; s += 32) { w.pool[s] = nc - 1 - s; w.dist[s] = CUDART_INF; w.col_seen[s] = 0; }
        for (int i = lane; i < nr; i += 32) w.row_seen[i] = 0;
        __syncwarp();
        int live = nc, sink = -1, row = cur;
        double frontier = 0.0;
        while (sink < 0) {
            if (lane == 0) w.row_seen[row] = 1;
            const double u_row = w.u[row];
            const float* crow = w.cost + (size_t)row * nc;
            double lmin = CUDART_INF;
            for (int s = lane; s < live; s += 32) {
                const int j = w.pool[s];
                const double cand = ((frontier + (double)crow[j]) - u_row) - w.v[j];
                double d = w.dist[j];
                if (cand < d) { d = cand; w.dist[j] = cand; w.pred[j] = row; }
                lmin = d < lmin ? d : lmin;
            }
            const double m = warp_min_f64(lmin);
            if (m == CUDART_INF) return WF_LSAP_INFEASIBLE;
            int first = 0x7fffffff, ulast = -1;
            for (int s = lane; s < live; s += 32) {
                const int j = w.pool[s];
                if (w.dist[j] == m) {
                    first = min(first, s);
                    if (w.row_of_col[j] < 0) ulast = max(ulast, s);
                }
            }
            const int s1 = warp_min_i32(first);
            const int ubest = warp_max_i32(ulast == s1 ? -1 : ulast);
            const int slot = ubest >= 0 ? ubest : s1;
            frontier = m;
            const int j = w.pool[slot];
            const int owner = w.row_of_col[j];
            __syncwarp();
            if (lane == 0) { w.col_seen[j] = 1; w.pool[slot] = w.pool[live - 1]; }
            --live;
            if (owner < 0) sink = j; else row = owner;
            __syncwarp();
        }
        // ---- dual update (reads dist of the columns matched to scanned rows, then rewrites v)
        for (int i = lane; i < nr; i += 32) {
            if (i == cur) w.u[i] = w.u[i] + frontier;
            else if (w.row_seen[i]) w.u[i] = w.u[i] + (frontier - w.dist[w.col_of_row[i]]);
        }
        for (int j = lane; j < nc; j += 32)
            if (w.col_seen[j]) w.v[j] = w.v[j] - (frontier - w.dist[j]);
        __syncwarp();
        // ---- flip the alternating path
        if (lane == 0) {
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
        __syncwarp();
    }
    return WF_LSAP_OK;
}

// ------------------------------------------------------------------------------------------
// Register-resident variant for nc <= 32*CPL: lane l owns columns l, l+32, ... and keeps their dual
// variable v, tentative distance, pool slot and owner row in registers; the Dijkstra step is then one
// shared-memory read of the cost entry, three fp64 adds and six REDUX warp reductions -- no pool array,
// no second pass over shared memory.  Semantics (incl. scipy's tie rules) identical to solve_warp:
// the pool is implicit -- slot[j] is column j's position in scipy's `remaining` array, seeded as
// nc-1-j, and "move the last entry into the freed slot" becomes "the column at slot live-1 takes slot s".
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long f64_key(double d) {        // order-preserving map to u64
    const unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
__device__ __forceinline__ double warp_min_f64_redux(double x) {
    const unsigned long long k = f64_key(x);
    const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    return key_f64(((unsigned long long)mh << 32) | ml);
}

template <int CPL>
__device__ int solve_warp_reg(const Work& w, int nr, int nc, int lane) {
    int bad = 0;
    for (int k = lane; k < nr * nc; k += 32) {
        const float c = w.cost[k];
        bad |= (c != c) || (c == -CUDART_INF_F);
    }
    if (__any_sync(0xffffffffu, bad)) return WF_LSAP_INVALID;
    for (int i = lane; i < nr; i += 32) { w.u[i] = 0.0; w.col_of_row[i] = -1; }
    for (int j = lane; j < nc; j += 32) { w.row_of_col[j] = -1; w.pred[j] = -1; }
    double v[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) v[k] = 0.0;
    __syncwarp();

    for (int cur = 0; cur < nr; ++cur) {
        double dist[CPL];
        int slot[CPL], rowof[CPL];
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const int j = lane + 32 * k;
            dist[k] = CUDART_INF;
            slot[k] = j < nc ? nc - 1 - j : -2;              // -2: column does not exist, -1: scanned
            rowof[k] = j < nc ? w.row_of_col[j] : 0;
        }
        for (int i = lane; i < nr; i += 32) w.row_seen[i] = 0;
        __syncwarp();
        int live = nc, sink = -1, row = cur;
        double frontier = 0.0;
        while (sink < 0) {
            if (lane == 0) w.row_seen[row] = 1;
            const double u_row = w.u[row];
            const float* crow = w.cost + (size_t)row * nc;
            double lmin = CUDART_INF;
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                if (slot[k] >= 0) {
                    const int j = lane + 32 * k;
                    const double cand = ((frontier + (double)crow[j]) - u_row) - v[k];
                    if (cand < dist[k]) { dist[k] = cand; w.pred[j] = row; }
                    lmin = dist[k] < lmin ? dist[k] : lmin;
                }
            }
            const double m = warp_min_f64_redux(lmin);
            if (m == CUDART_INF) return WF_LSAP_INFEASIBLE;
            int first = 0x7fffffff, ulast = -1;
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                if (slot[k] >= 0 && dist[k] == m) {
                    first = min(first, slot[k]);
                    if (rowof[k] < 0) ulast = max(ulast, slot[k]);
                }
            }
            const int s1 = __reduce_min_sync(0xffffffffu, first);
            const int ubest = __reduce_max_sync(0xffffffffu, ulast == s1 ? -1 : ulast);
            const int s = ubest >= 0 ? ubest : s1;
            int mine_j = -1, mine_owner = -1;
#pragma unroll
            for (int k = 0; k < CPL; ++k)
                if (slot[k] == s) { mine_j = lane + 32 * k; mine_owner = rowof[k] + 1; }
            const int jsel = __reduce_max_sync(0xffffffffu, mine_j);
            const int owner = __reduce_max_sync(0xffffffffu, mine_owner) - 1;     // -1 -> unassigned column
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                if (slot[k] == s) slot[k] = -1;                   // scanned
                else if (slot[k] == live - 1) slot[k] = s;        // the pool's last entry moves into the hole
            }
            --live;
            frontier = m;
            if (owner < 0) sink = jsel; else row = owner;
        }
        // publish the distances of scanned columns, then the dual update
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const int j = lane + 32 * k;
            if (j < nc) w.dist[j] = dist[k];
        }
        __syncwarp();
        for (int i = lane; i < nr; i += 32) {
            if (i == cur) w.u[i] = w.u[i] + frontier;
            else if (w.row_seen[i]) w.u[i] = w.u[i] + (frontier - w.dist[w.col_of_row[i]]);
        }
#pragma unroll
        for (int k = 0; k < CPL; ++k)
            if (slot[k] == -1) v[k] = v[k] - (frontier - dist[k]);
        __syncwarp();
        if (lane == 0) {
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
        __syncwarp();
    }
    return WF_LSAP_OK;
}

__device__ __forceinline__ int solve_any(const Work& w, int nr, int nc, int lane) {
    if (nc <= 32) return solve_warp_reg<1>(w, nr, nc, lane);
    if (nc <= 64) return solve_warp_reg<2>(w, nr, nc, lane);
    if (nc <= 128) return solve_warp_reg<4>(w, nr, nc, lane);
    return solve_warp(w, nr, nc, lane);
}

constexpr int WARPS_PER_CTA = 4;          // fewer when a problem's working set is large (launch_cfg); kernels read blockDim

// Generic: raw float32 matrices from global memory.
__global__ void lsap_batched_kernel(const float* __restrict__ cost, long long batch_stride, int ld,
                                    const int* __restrict__ nr_arr, const int* __restrict__ nc_arr, int B, int max_nr,
                                    int32_t* __restrict__ col_of_row, int32_t* __restrict__ status, size_t per_warp) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * (int)(blockDim.x >> 5) + warp;
    if (b >= B) return;
    const int nr0 = nr_arr[b], nc0 = nc_arr[b];
    int32_t* out = col_of_row + (size_t)b * max_nr;
    for (int i = lane; i < max_nr; i += 32) out[i] = -1;
    if (nr0 <= 0 || nc0 <= 0) { if (lane == 0) status[b] = WF_LSAP_OK; return; }
    const bool flip = nc0 < nr0;                       // tall -> solve the transpose
    const int nr = flip ? nc0 : nr0, nc = flip ? nr0 : nc0;
    Work w = carve(sm + per_warp * warp, nr, nc);
    const float* src = cost + (size_t)b * batch_stride;
    for (int k = lane; k < nr0 * nc0; k += 32) {
        const int i = k / nc0, j = k - i * nc0;
        const float c = src[(size_t)i * ld + j];
        if (flip) w.cost[(size_t)j * nc + i] = c; else w.cost[(size_t)i * nc + j] = c;
    }
    __syncwarp();
    const int st = solve_any(w, nr, nc, lane);
    if (lane == 0) status[b] = st;
    if (st != WF_LSAP_OK) return;
    __syncwarp();
    if (!flip) { for (int i = lane; i < nr; i += 32) out[i] = w.col_of_row[i]; }
    else       { for (int i = lane; i < nc; i += 32) out[i] = w.row_of_col[i]; }   // original row i = transposed column i
}

// Loss-style matrix built in place (losses/WireframeLoss.py:142,206-224), always V x V.
__global__ void loss_match_kernel(const float* __restrict__ pred_v, const float* __restrict__ pred_e,
                                  const float* __restrict__ tgt_v, const long long* __restrict__ counts, int B, int V,
                                  int Vt, int32_t* __restrict__ col_of_row, int32_t* __restrict__ status,
                                  float* __restrict__ cost_dump, size_t per_warp) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * (int)(blockDim.x >> 5) + warp;
    if (b >= B) return;
    const long long cnt = counts[b];
    int32_t* out = col_of_row + (size_t)b * V;
    for (int i = lane; i < V; i += 32) out[i] = -1;
    if (cnt > V || cnt > Vt || cnt < 0) { if (lane == 0) status[b] = WF_LSAP_INFEASIBLE; return; }
    Work w = carve(sm + per_warp * warp, V, V);
    const float* pv = pred_v + (size_t)b * V * 3;
    const float* pe = pred_e + (size_t)b * V;
    const float* tv = tgt_v + (size_t)b * Vt * 3;
    for (int k = lane; k < V * V; k += 32) {
        const int i = k / V, j = k - i * V;
        const float e = pe[i];
        float c;
        if (j < cnt) {
            // torch.cdist(p=1): |dx| + |dy| + |dz| added in that order, one rounding each; then + |e-1|
            float d = fabsf(__fsub_rn(pv[i * 3 + 0], tv[j * 3 + 0]));
            d = __fadd_rn(d, fabsf(__fsub_rn(pv[i * 3 + 1], tv[j * 3 + 1])));
            d = __fadd_rn(d, fabsf(__fsub_rn(pv[i * 3 + 2], tv[j * 3 + 2])));
            c = __fadd_rn(d, fabsf(__fsub_rn(e, 1.0f)));
        } else {
            c = e;                                         // dummy "no object" column
        }
        w.cost[k] = c;
        if (cost_dump) cost_dump[(size_t)b * V * V + k] = c;
    }
    __syncwarp();
    const int st = solve_any(w, V, V, lane);
    if (lane == 0) status[b] = st;
    if (st != WF_LSAP_OK) return;
    __syncwarp();
    for (int i = lane; i < V; i += 32) out[i] = w.col_of_row[i];
}

__global__ void wireframe_cost_kernel(const float* __restrict__ pred_v, const float* __restrict__ pred_e,
                                      const float* __restrict__ tgt_v, const float* __restrict__ tgt_e,
                                      const int* __restrict__ tgt_off, int V, float wv, float we,
                                      float* __restrict__ cost, int ld) {
    const int b = blockIdx.y;
    const int t0 = tgt_off[b], T = tgt_off[b + 1] - t0;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= V * T) return;
    const int i = k / T, j = k - i * T;
    const float* p = pred_v + ((size_t)b * V + i) * 3;
    const float* t = tgt_v + (size_t)(t0 + j) * 3;
    float d = fabsf(__fsub_rn(p[0], t[0]));
    d = __fadd_rn(d, fabsf(__fsub_rn(p[1], t[1])));
    d = __fadd_rn(d, fabsf(__fsub_rn(p[2], t[2])));
    const float ce = fabsf(__fsub_rn(pred_e[(size_t)b * V + i], tgt_e[t0 + j]));
    cost[((size_t)b * V + i) * ld + j] = __fadd_rn(__fmul_rn(wv, d), __fmul_rn(we, ce));
}

// models/HungarianMatcher.py:101-123.  One thread per (query, target); softmax row recomputed per thread
// (K is small); arithmetic order follows the reference expression C = wb*L1 + wc*(-p) + wg*(-giou).
__global__ void detr_cost_kernel(const float* __restrict__ logits, const float* __restrict__ boxes,
                                 const long long* __restrict__ labels, const float* __restrict__ tboxes,
                                 const int* __restrict__ tgt_off, int Q, int K, float wc, float wb, float wg,
                                 float* __restrict__ cost, int ld) {
    const int b = blockIdx.y;
    const int t0 = tgt_off[b], T = tgt_off[b + 1] - t0;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Q * T) return;
    const int i = k / T, j = k - i * T;
    const float* lg = logits + ((size_t)b * Q + i) * K;
    float mx = -CUDART_INF_F;
    for (int c = 0; c < K; ++c) mx = fmaxf(mx, lg[c]);
    float den = 0.f;
    for (int c = 0; c < K; ++c) den += expf(lg[c] - mx);
    const float prob = expf(lg[labels[t0 + j]] - mx) / den;
    const float* bq = boxes + ((size_t)b * Q + i) * 4;
    const float* bt = tboxes + (size_t)(t0 + j) * 4;
    float l1 = fabsf(__fsub_rn(bq[0], bt[0]));
    l1 = __fadd_rn(l1, fabsf(__fsub_rn(bq[1], bt[1])));
    l1 = __fadd_rn(l1, fabsf(__fsub_rn(bq[2], bt[2])));
    l1 = __fadd_rn(l1, fabsf(__fsub_rn(bq[3], bt[3])));
    // cxcywh -> xyxy
    const float ax0 = __fsub_rn(bq[0], __fmul_rn(0.5f, bq[2])), ay0 = __fsub_rn(bq[1], __fmul_rn(0.5f, bq[3]));
    const float ax1 = __fadd_rn(bq[0], __fmul_rn(0.5f, bq[2])), ay1 = __fadd_rn(bq[1], __fmul_rn(0.5f, bq[3]));
    const float bx0 = __fsub_rn(bt[0], __fmul_rn(0.5f, bt[2])), by0 = __fsub_rn(bt[1], __fmul_rn(0.5f, bt[3]));
    const float bx1 = __fadd_rn(bt[0], __fmul_rn(0.5f, bt[2])), by1 = __fadd_rn(bt[1], __fmul_rn(0.5f, bt[3]));
    const float area_a = __fmul_rn(__fsub_rn(ax1, ax0), __fsub_rn(ay1, ay0));
    const float area_b = __fmul_rn(__fsub_rn(bx1, bx0), __fsub_rn(by1, by0));
    const float iw = fmaxf(__fsub_rn(fminf(ax1, bx1), fmaxf(ax0, bx0)), 0.f);
    const float ih = fmaxf(__fsub_rn(fminf(ay1, by1), fmaxf(ay0, by0)), 0.f);
    const float inter = __fmul_rn(iw, ih);
    const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    const float iou = __fdiv_rn(inter, uni);
    const float hw = fmaxf(__fsub_rn(fmaxf(ax1, bx1), fminf(ax0, bx0)), 0.f);
    const float hh = fmaxf(__fsub_rn(fmaxf(ay1, by1), fminf(ay0, by0)), 0.f);
    const float hull = __fmul_rn(hw, hh);
    const float giou = __fsub_rn(iou, __fdiv_rn(__fsub_rn(hull, uni), hull));
    float c = __fadd_rn(__fmul_rn(wb, l1), __fmul_rn(wc, -prob));
    c = __fadd_rn(c, __fmul_rn(wg, -giou));
    cost[((size_t)b * Q + i) * ld + j] = c;
}

static int launch_cfg(size_t per_warp, size_t* smem_out, int* warps_out, const void* kernel) {
    int warps = WARPS_PER_CTA;
    while (warps > 1 && per_warp * warps > 227 * 1024) warps >>= 1;       // large matrices: fewer samples per CTA
    const size_t smem = per_warp * warps;
    *warps_out = warps;
    if (smem > 227 * 1024) { set_error("LSAP problem too large for shared memory (%zu bytes per matrix)", smem); return WF_ETOOBIG; }
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return WF_ECUDA; }
    }
    *smem_out = smem;
    return WF_OK;
}

}  // namespace lsap
}  // namespace wf

extern "C" int wf_loss_match(const float* pred_v, const float* pred_e, const float* tgt_v, const int64_t* counts, int B,
                             int V, int Vt, int32_t* col_of_row, int32_t* status, float* cost_dump, wf_stream_t stream) {
    using namespace wf;
    using namespace wf::lsap;
    if (B <= 0 || V <= 0) return WF_OK;
    WF_CHECK_ARG(Vt >= 0, "wf_loss_match: bad Vt");
    size_t smem;
    const size_t per_warp = work_bytes(V, V);
    int warps;
    int rc = launch_cfg(per_warp, &smem, &warps, (const void*)loss_match_kernel);
    if (rc != WF_OK) return rc;
    loss_match_kernel<<<cdiv(B, warps), warps * 32, smem, as_stream(stream)>>>(
        pred_v, pred_e, tgt_v, reinterpret_cast<const long long*>(counts), B, V, Vt, col_of_row, status, cost_dump, per_warp);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_lsap_batched(const float* cost, int64_t batch_stride, int ld, const int32_t* nr, const int32_t* nc, int B,
                               int max_nr, int max_nc, int32_t* col_of_row, int32_t* status, wf_stream_t stream) {
    using namespace wf;
    using namespace wf::lsap;
    if (B <= 0 || max_nr <= 0) return WF_OK;
    const int lo = max_nr < max_nc ? max_nr : max_nc, hi = max_nr < max_nc ? max_nc : max_nr;
    // worst case carve for any (nr<=max_nr, nc<=max_nc) after transposition: rows<=cols
    const size_t per_warp = work_bytes(lo > 0 ? lo : 1, hi > 0 ? hi : 1);
    size_t smem;
    int warps;
    int rc = launch_cfg(per_warp, &smem, &warps, (const void*)lsap_batched_kernel);
    if (rc != WF_OK) return rc;
    lsap_batched_kernel<<<cdiv(B, warps), warps * 32, smem, as_stream(stream)>>>(
        cost, (long long)batch_stride, ld, nr, nc, B, max_nr, col_of_row, status, per_warp);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_wireframe_matcher_cost(const float* pred_v, const float* pred_e, const float* tgt_v, const float* tgt_e,
                                         const int32_t* tgt_off, int B, int V, float wv, float we, float* cost, int ld,
                                         wf_stream_t stream) {
    using namespace wf;
    if (B <= 0 || V <= 0 || ld <= 0) return WF_OK;
    dim3 grid(cdiv((long long)V * ld, 256), B);
    lsap::wireframe_cost_kernel<<<grid, 256, 0, as_stream(stream)>>>(pred_v, pred_e, tgt_v, tgt_e, tgt_off, V, wv, we, cost, ld);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_detr_matcher_cost(const float* logits, const float* boxes, const int64_t* tgt_labels, const float* tgt_boxes,
                                    const int32_t* tgt_off, int B, int Q, int K, float wc, float wb, float wg, float* cost,
                                    int ld, wf_stream_t stream) {
    using namespace wf;
    if (B <= 0 || Q <= 0 || ld <= 0) return WF_OK;
    dim3 grid(cdiv((long long)Q * ld, 256), B);
    lsap::detr_cost_kernel<<<grid, 256, 0, as_stream(stream)>>>(logits, boxes, reinterpret_cast<const long long*>(tgt_labels),
                                                               tgt_boxes, tgt_off, Q, K, wc, wb, wg, cost, ld);
    WF_LAUNCH_CHECK();
    return WF_OK;
}
