// Encoder-side kernels that are NOT tensor-core GEMMs: validity mask, the narrow first layer
// (memory-bound, fp32, fused Linear+LayerNorm+ReLU), LayerNorm statistics finalisation, weight
// staging casts, and the four pooled reductions with their backward.
#include "wf_common.cuh"
#include "ln_side.cuh"

#include <stdlib.h>

#include <math_constants.h>
#include <type_traits>

namespace wf {
namespace enc {

// models/PointNetEncoder.py:85-86.  Grid (slabs of PM_SLAB points, clouds): a single CTA per cloud walked a million-point scan
// alone (1.6 ms for 32 MB).  The per-cloud count is accumulated as a float: sums of 0/1 stay exact below 2^24 points per cloud
// whatever the order of the atomics (larger clouds take one slab per cloud, i.e. the serial order).
constexpr int PM_SLAB = 2048;

__global__ void __launch_bounds__(256)
point_mask_kernel(const float* __restrict__ x, int N, int D, int slab, uint8_t* __restrict__ mask, float* __restrict__ valid) {
    const int b = blockIdx.y;
    const int n0 = blockIdx.x * slab, n1 = min(N, n0 + slab);
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    int local = 0;
    const bool vec = D == 8 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    for (int n = n0 + threadIdx.x; n < n1; n += blockDim.x) {
        const float* p = x + ((size_t)b * N + n) * D;
        float s = 0.f;
        if (vec) {
            const float4 u = reinterpret_cast<const float4*>(p)[0], v = reinterpret_cast<const float4*>(p)[1];
            s += fabsf(u.x); s += fabsf(u.y); s += fabsf(u.z); s += fabsf(u.w);       // same order as the scalar loop
            s += fabsf(v.x); s += fabsf(v.y); s += fabsf(v.z); s += fabsf(v.w);
        } else {
            for (int d = 0; d < D; ++d) s += fabsf(p[d]);
        }
        const int m = s > 1e-9f;
        mask[(size_t)b * N + n] = (uint8_t)m;
        local += m;
    }
    local = __reduce_add_sync(0xffffffffu, local);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(&cnt, local);
    __syncthreads();
    if (threadIdx.x == 0 && cnt) atomicAdd(valid + b, (float)cnt);
}

__global__ void clamp_valid_kernel(float* __restrict__ valid, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) valid[b] = fmaxf(valid[b], 1.0f);
}

// ------------------------------------------------------------------------------------------
// layer 1: one warp per point, lane owns channel pairs (64*i + 2*lane, +1), W^T staged in smem
// ------------------------------------------------------------------------------------------
template <int D, int CP>
struct L1Smem {
    float wt[D][64 * CP];      // W transposed: [k][c]
    float b[64 * CP], g[64 * CP], be[64 * CP];
};

template <int D, int CP>
__device__ __forceinline__ void l1_stage(L1Smem<D, CP>& s, const float* W, const float* b, const float* g, const float* be) {
    constexpr int C = 64 * CP;
    for (int i = threadIdx.x; i < C * D; i += blockDim.x) { const int c = i / D, k = i - c * D; s.wt[k][c] = W[i]; }
    for (int c = threadIdx.x; c < C; c += blockDim.x) { s.b[c] = b[c]; s.g[c] = g[c]; s.be[c] = be[c]; }
    __syncthreads();
}

// z (pre-LN) for this lane's 2*CP channels of point `row`; returns mean and rstd
template <int D, int CP>
__device__ __forceinline__ void l1_point(const L1Smem<D, CP>& s, const float* __restrict__ x, size_t row, int lane,
                                         float (&xv)[D], float (&z)[2 * CP], float& mu, float& rs, float eps) {
    constexpr int C = 64 * CP;
#pragma unroll
    for (int k = 0; k < D; k += 4) {
        const float4 t = *reinterpret_cast<const float4*>(x + row * D + k);
        xv[k] = t.x; xv[k + 1] = t.y; xv[k + 2] = t.z; xv[k + 3] = t.w;
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < CP; ++i) {
        const int c = 64 * i + 2 * lane;
        float a0 = s.b[c], a1 = s.b[c + 1];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const float2 w = *reinterpret_cast<const float2*>(&s.wt[k][c]);
            a0 = fmaf(xv[k], w.x, a0); a1 = fmaf(xv[k], w.y, a1);
        }
        z[2 * i] = a0; z[2 * i + 1] = a1;
        sum += a0 + a1;
    }
    mu = warp_sum(sum) / (float)C;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 2 * CP; ++i) { const float d = z[i] - mu; var = fmaf(d, d, var); }
    rs = rsqrtf(warp_sum(var) / (float)C + eps);
}

template <int D, int CP, int HDT>
__global__ void __launch_bounds__(256)
l1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b,
              const float* __restrict__ g, const float* __restrict__ be, void* __restrict__ h, int M, float eps) {
    __shared__ L1Smem<D, CP> s;
    constexpr int C = 64 * CP;
    l1_stage<D, CP>(s, W, b, g, be);
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (size_t row = (size_t)blockIdx.x * wpb + (threadIdx.x >> 5); row < (size_t)M; row += (size_t)gridDim.x * wpb) {
        float xv[D], z[2 * CP], mu, rs;
        l1_point<D, CP>(s, x, row, lane, xv, z, mu, rs, eps);
#pragma unroll
        for (int i = 0; i < CP; ++i) {
            const int c = 64 * i + 2 * lane;
            const float y0 = fmaxf((z[2 * i] - mu) * rs * s.g[c] + s.be[c], 0.f);
            const float y1 = fmaxf((z[2 * i + 1] - mu) * rs * s.g[c + 1] + s.be[c + 1], 0.f);
            if (HDT == WF_BF16) {
                reinterpret_cast<__nv_bfloat162*>(h)[(row * C + c) >> 1] = __floats2bfloat162_rn(y0, y1);
            } else {
                reinterpret_cast<float2*>(h)[(row * C + c) >> 1] = make_float2(y0, y1);
            }
        }
    }
}

// backward: two warps share a point.  Both recompute the layer (z, LayerNorm statistics and the two
// backward row reductions need all C channels; the arithmetic is trivial next to the 1 KB/point
// gradient read), but each keeps register accumulators for only HALF of its channels, which is what
// keeps dW (2*CP x D per lane otherwise) out of local memory.  Block reduce in smem, then atomics.
template <int D, int CP, int GDT, bool WANT_DX>
__global__ void __launch_bounds__(256)
l1_bwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b,
              const float* __restrict__ g, const float* __restrict__ be, const void* __restrict__ dh,
              float* __restrict__ dW, float* __restrict__ db, float* __restrict__ dg, float* __restrict__ dbe,
              float* __restrict__ dx, int M, float eps) {
    __shared__ L1Smem<D, CP> s;
    constexpr int C = 64 * CP;
    constexpr int HC = CP;                   // accumulated channels per lane (half of 2*CP)
    l1_stage<D, CP>(s, W, b, g, be);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = warp & 1, pair = warp >> 1;
    const int ppb = blockDim.x >> 6;         // point pairs (of warps) per block
    float aW[HC][D], ab[HC], ag[HC], abe[HC];
#pragma unroll
    for (int i = 0; i < HC; ++i) {
        ab[i] = ag[i] = abe[i] = 0.f;
#pragma unroll
        for (int k = 0; k < D; ++k) aW[i][k] = 0.f;
    }
    for (size_t row = (size_t)blockIdx.x * ppb + pair; row < (size_t)M; row += (size_t)gridDim.x * ppb) {
        float xv[D], z[2 * CP], mu, rs;
        l1_point<D, CP>(s, x, row, lane, xv, z, mu, rs, eps);
        float gy[2 * CP];
        float c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int i = 0; i < CP; ++i) {
            const int c = 64 * i + 2 * lane;
            float d0, d1;
            if (GDT == WF_BF16) {
                const float2 t = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(dh)[(row * C + c) >> 1]);
                d0 = t.x; d1 = t.y;
            } else {
                const float2 t = reinterpret_cast<const float2*>(dh)[(row * C + c) >> 1];
                d0 = t.x; d1 = t.y;
            }
            z[2 * i] = (z[2 * i] - mu) * rs; z[2 * i + 1] = (z[2 * i + 1] - mu) * rs;          // z becomes xhat
            const float y0 = z[2 * i] * s.g[c] + s.be[c], y1 = z[2 * i + 1] * s.g[c + 1] + s.be[c + 1];
            gy[2 * i] = y0 > 0.f ? d0 : 0.f; gy[2 * i + 1] = y1 > 0.f ? d1 : 0.f;
            const float h0 = gy[2 * i] * s.g[c], h1 = gy[2 * i + 1] * s.g[c + 1];
            c1 += h0 + h1; c2 = fmaf(h0, z[2 * i], fmaf(h1, z[2 * i + 1], c2));
        }
        c1 = warp_sum(c1) / (float)C; c2 = warp_sum(c2) / (float)C;
        // this warp's half: register index ii <-> channel slot (half ? ii + CP : ii)
#pragma unroll
        for (int ii = 0; ii < HC; ++ii) {
            const float gyv = half ? gy[ii + CP] : gy[ii];
            const float xh = half ? z[ii + CP] : z[ii];
            const int slot = ii + half * CP;
            const int c = 64 * (slot >> 1) + 2 * lane + (slot & 1);
            const float dzv = rs * (gyv * s.g[c] - c1 - xh * c2);
            ab[ii] += dzv; ag[ii] = fmaf(gyv, xh, ag[ii]); abe[ii] += gyv;
#pragma unroll
            for (int k = 0; k < D; ++k) aW[ii][k] = fmaf(dzv, xv[k], aW[ii][k]);
        }
        if (WANT_DX && half == 0) {
            float dxl[D];
#pragma unroll
            for (int k = 0; k < D; ++k) dxl[k] = 0.f;
#pragma unroll
            for (int i = 0; i < 2 * CP; ++i) {
                const int c = 64 * (i >> 1) + 2 * lane + (i & 1);
                const float dzv = rs * (gy[i] * s.g[c] - c1 - z[i] * c2);
#pragma unroll
                for (int k = 0; k < D; ++k) dxl[k] = fmaf(dzv, s.wt[k][c], dxl[k]);
            }
#pragma unroll
            for (int k = 0; k < D; ++k) { const float t = warp_sum(dxl[k]); if (lane == k) dx[row * D + k] = t; }
        }
    }
    // block reduction through the (now idle) weight staging buffer, then one atomic per entry
    __syncthreads();
    float* red = &s.wt[0][0];                 // C*D floats
    float* red2 = s.b;                        // 3*C floats contiguous (b, g, be)
    for (int i = threadIdx.x; i < C * D; i += blockDim.x) red[i] = 0.f;
    for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) red2[i] = 0.f;
    __syncthreads();
    for (int psel = 0; psel < ppb; ++psel) {  // the two halves of a pair touch disjoint channels
        if (pair == psel) {
#pragma unroll
            for (int ii = 0; ii < HC; ++ii) {
                const int slot = ii + half * CP;
                const int c = 64 * (slot >> 1) + 2 * lane + (slot & 1);
#pragma unroll
                for (int k = 0; k < D; ++k) red[c * D + k] += aW[ii][k];
                red2[c] += ab[ii]; red2[C + c] += ag[ii]; red2[2 * C + c] += abe[ii];
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < C * D; i += blockDim.x) atomicAdd(dW + i, red[i]);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        atomicAdd(db + c, red2[c]); atomicAdd(dg + c, red2[C + c]); atomicAdd(dbe + c, red2[2 * C + c]);
    }
}

// ------------------------------------------------------------------------------------------
// layer 1, channel-stationary variant (the production kernels): a CTA of 256 threads, thread t owns channels
// (2t, 2t+1) of all 512 for EVERY point it sees, so the weights (2 x 8), the parameter-gradient accumulators
// (2 x 8 + 6) and LayerNorm gain/shift live in a few registers and no weight is ever re-read from shared memory.
//
// LayerNorm statistics without a reduction over channels: z_c = W_c.x + b_c is affine in the 8-vector x, so
//   z_c - mean_c(z) = Wt_c.x + bt_c           (Wt = W - column mean, bt = b - mean(b): "centred" layer)
//   var_c(z)        = x'^T S x',  x' = [x; 1],  S = (1/C) [Wt | bt]^T [Wt | bt]   (9 x 9, PSD)
//                   = |R x'|^2  with S = R^T R (Cholesky, upper triangular R): a sum of squares, no cancellation
// R is rebuilt per CTA in the prologue (512 x 45 products + a 9 x 9 factorisation in fp64 by one thread); each point's
// rstd then costs 45 FMAs, done by ONE thread per point while the point is staged in shared memory.
// Arithmetic on channel pairs uses the packed fp32 instructions (fma/mul/add.f32x2 -> FFMA2/FMUL2/FADD2).
// ------------------------------------------------------------------------------------------
namespace l1c {

constexpr int C = 512, D = 8, NT = 256, PB = 256;       // channels, inputs, threads per CTA, points per staging block
typedef unsigned long long u64;

__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void up2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

struct __align__(16) Smem {
    u64 xs2[PB][D];          // staged points, every coordinate duplicated into both halves (operand of the f32x2 ops); read as
                             // 16-byte pairs of coordinates (half the shared-memory instructions of 8-byte reads)
    u64 rs2[PB];             // (rstd, rstd) per staged point
    float red[NT / 32][48];  // block reductions of the prologue
    float R[48];             // packed upper-triangular factor: R[i][j], j >= i, row-major
    float cm[12];            // column means of W (8) and mean of b
};

template <int N_>
__device__ __forceinline__ void block_sum(Smem& s, float* v, float* out) {      // out[0..N_) valid in all threads
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < N_; ++i) { const float t = warp_sum(v[i]); if (lane == 0) s.red[warp][i] = t; }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N_; ++i) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) t += s.red[w][i];
        out[i] = t;
    }
    __syncthreads();
}

// Loads this thread's two channels, centres them, builds R.  w2[k] = (Wt[2t][k], Wt[2t+1][k]), b2 = centred bias pair.
__device__ __forceinline__ void prologue(Smem& s, const float* __restrict__ W, const float* __restrict__ b, u64 (&w2)[D], u64& b2) {
    const int t = threadIdx.x;
    float w[2][D], bb[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(W + (size_t)(2 * t + i) * D), c = *reinterpret_cast<const float4*>(W + (size_t)(2 * t + i) * D + 4);
        w[i][0] = a.x; w[i][1] = a.y; w[i][2] = a.z; w[i][3] = a.w; w[i][4] = c.x; w[i][5] = c.y; w[i][6] = c.z; w[i][7] = c.w;
        bb[i] = b[2 * t + i];
    }
    float v[48], o[48];
#pragma unroll
    for (int k = 0; k < D; ++k) v[k] = w[0][k] + w[1][k];
    v[8] = bb[0] + bb[1];
    block_sum<9>(s, v, o);
#pragma unroll
    for (int k = 0; k < D; ++k) { const float m = o[k] * (1.0f / C); w[0][k] -= m; w[1][k] -= m; if (t == 0) s.cm[k] = m; }
    { const float m = o[8] * (1.0f / C); bb[0] -= m; bb[1] -= m; if (t == 0) s.cm[8] = m; }
    int idx = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int j = i; j < 9; ++j) {
            const float ai0 = i < 8 ? w[0][i] : bb[0], aj0 = j < 8 ? w[0][j] : bb[0];
            const float ai1 = i < 8 ? w[1][i] : bb[1], aj1 = j < 8 ? w[1][j] : bb[1];
            v[idx++] = ai0 * aj0 + ai1 * aj1;
        }
    block_sum<45>(s, v, o);
    if (t == 0) {
        // 9 x 9 Cholesky, fully unrolled so that S and R stay in registers (entries of S are sums of 512 products of O(1)
        // numbers; fp32 with rsqrt is accurate to ~1e-6 relative on rstd, far below the bf16 output rounding)
        float S[9][9], Rf[9][9];
        int q = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i)
#pragma unroll
            for (int j = i; j < 9; ++j) S[i][j] = o[q++] * (1.0f / C);
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            float d = S[i][i];
#pragma unroll
            for (int m = 0; m < i; ++m) d = fmaf(-Rf[m][i], Rf[m][i], d);
            const bool okp = d > 1e-7f * S[i][i] && d > 0.f;                      // rank-deficient directions drop out
            const float inv = okp ? rsqrtf(d) : 0.f;
            Rf[i][i] = okp ? d * inv : 0.f;
#pragma unroll
            for (int j = i + 1; j < 9; ++j) {
                float e = S[i][j];
#pragma unroll
                for (int m = 0; m < i; ++m) e = fmaf(-Rf[m][i], Rf[m][j], e);
                Rf[i][j] = e * inv;
            }
        }
        q = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i)
#pragma unroll
            for (int j = i; j < 9; ++j) s.R[q++] = Rf[i][j];
    }
#pragma unroll
    for (int k = 0; k < D; ++k) w2[k] = pk2(w[0][k], w[1][k]);
    b2 = pk2(bb[0], bb[1]);
    __syncthreads();
}

// stage up to PB points starting at p0: thread t owns point p0 + t (coordinates duplicated, rstd from the factor R)
__device__ __forceinline__ void stage(Smem& s, const float4& xa, const float4& xb, bool ok, float eps) {
    const int t = threadIdx.x;
    const float xv[9] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w, 1.0f};
    float var = 0.f;
    int q = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        float a = 0.f;
#pragma unroll
        for (int j = i; j < 9; ++j) a = fmaf(s.R[q++], xv[j], a);
        var = fmaf(a, a, var);
    }
    const float rs = ok ? rsqrtf(var + eps) : 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) s.xs2[t][k] = pk2(xv[k], xv[k]);
    s.rs2[t] = pk2(rs, rs);
}

template <int HDT>
__global__ void __launch_bounds__(NT, 2)
fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ g,
           const float* __restrict__ be, void* __restrict__ h, int M, float eps) {
    __shared__ Smem s;
    const int t = threadIdx.x;
    u64 w2[D], b2;
    prologue(s, W, b, w2, b2);
    const u64 g2 = pk2(g[2 * t], g[2 * t + 1]), be2 = pk2(be[2 * t], be[2 * t + 1]);
    const int nblk = (M + PB - 1) / PB;
    float4 xa = make_float4(0, 0, 0, 0), xb = xa;
    int blk = blockIdx.x;
    if (blk < nblk && blk * PB + t < M) { const float4* src = reinterpret_cast<const float4*>(x + (size_t)(blk * PB + t) * D); xa = src[0]; xb = src[1]; }
    for (; blk < nblk; blk += gridDim.x) {
        const int p0 = blk * PB;
        __syncthreads();                                  // previous block's readers are done with the staging buffers
        stage(s, xa, xb, p0 + t < M, eps);
        const int nb = blk + gridDim.x;                   // prefetch the next block's point while this one is processed
        if (nb < nblk && nb * PB + t < M) { const float4* src = reinterpret_cast<const float4*>(x + (size_t)(nb * PB + t) * D); xa = src[0]; xb = src[1]; }
        __syncthreads();
        const int np = min(PB, M - p0);
#pragma unroll 4
        for (int p = 0; p < np; ++p) {
            u64 z = b2;
#pragma unroll
            for (int k = 0; k < D; ++k) z = fma2(s.xs2[p][k], w2[k], z);
            const u64 y = fma2(mul2(z, s.rs2[p]), g2, be2);
            float y0, y1;
            up2(y, y0, y1);
            y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f);
            const size_t o = ((size_t)(p0 + p) * C + 2 * t) >> 1;
            if (HDT == WF_BF16) reinterpret_cast<__nv_bfloat162*>(h)[o] = __floats2bfloat162_rn(y0, y1);
            else reinterpret_cast<float2*>(h)[o] = make_float2(y0, y1);
        }
    }
}

// backward without dx.  Points are processed 8 at a time: phase 1 recomputes the layer and forms this thread's share of
// the two LayerNorm-backward row sums, one block reduction per 8 points (warp w reduces point w), phase 2 finishes dz and
// accumulates dW (2 x 8 per thread), db, dgamma, dbeta in registers; one atomic per entry per CTA at the end.
constexpr int PG = 8;

template <int GDT>
__global__ void __launch_bounds__(NT, 2)
bwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ g,
           const float* __restrict__ be, const void* __restrict__ dh, float* __restrict__ dW, float* __restrict__ db,
           float* __restrict__ dg, float* __restrict__ dbe, int M, float eps) {
    __shared__ Smem s;
    __shared__ float part[PG][2][NT];                     // per point, per row sum, per thread
    __shared__ float2 csum[PG];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    u64 w2[D], b2;
    prologue(s, W, b, w2, b2);
    const u64 g2 = pk2(g[2 * t], g[2 * t + 1]), be2 = pk2(be[2 * t], be[2 * t + 1]);
    u64 aW[D], ab = 0ull, ag = 0ull, abe = 0ull;          // bit pattern 0 == (0.f, 0.f)
#pragma unroll
    for (int k = 0; k < D; ++k) aW[k] = 0ull;
    const int nblk = (M + PB - 1) / PB;
    // raw gradient words of one 8-point group: (bf16, bf16) in 32 bits, or two floats
    typedef typename std::conditional<GDT == WF_BF16, uint32_t, float2>::type Raw;
    Raw cur[PG], nxt[PG];
    bool have_cur = false;
    auto load_group = [&](Raw (&dst)[PG], int pb0, int q, int npts) {
#pragma unroll
        for (int i = 0; i < PG; ++i) {
            if (q + i < npts) dst[i] = reinterpret_cast<const Raw*>(dh)[((size_t)(pb0 + q + i) * C + 2 * t) >> 1];
            else dst[i] = Raw();
        }
    };
    auto decode = [](const Raw& r, float& d0, float& d1) {
        if constexpr (GDT == WF_BF16) { d0 = __uint_as_float(r << 16); d1 = __uint_as_float(r & 0xFFFF0000u); }
        else { d0 = r.x; d1 = r.y; }
    };
    float4 xa = make_float4(0, 0, 0, 0), xb = xa;
    int blk = blockIdx.x;
    if (blk < nblk && blk * PB + t < M) { const float4* src = reinterpret_cast<const float4*>(x + (size_t)(blk * PB + t) * D); xa = src[0]; xb = src[1]; }
    for (; blk < nblk; blk += gridDim.x) {
        const int p0 = blk * PB;
        __syncthreads();
        stage(s, xa, xb, p0 + t < M, eps);
        const int nb = blk + gridDim.x;
        if (nb < nblk && nb * PB + t < M) { const float4* src = reinterpret_cast<const float4*>(x + (size_t)(nb * PB + t) * D); xa = src[0]; xb = src[1]; }
        __syncthreads();
        const int np = min(PB, M - p0);
        if (!have_cur) { load_group(cur, p0, 0, np); have_cur = true; }
        for (int q0 = 0; q0 < np; q0 += PG) {
            u64 xh[PG], gy[PG];
            // the next group's gradient words are requested now and consumed one group later (also across staging blocks)
            if (q0 + PG < np) load_group(nxt, p0, q0 + PG, np);
            else if (blk + (int)gridDim.x < nblk) load_group(nxt, (blk + (int)gridDim.x) * PB, 0, min(PB, M - (blk + (int)gridDim.x) * PB));
            // ---- phase 1
#pragma unroll
            for (int i = 0; i < PG; ++i) {
                const int p = q0 + i;
                float d0, d1;
                decode(cur[i], d0, d1);
                u64 z = b2;
#pragma unroll
                for (int k = 0; k < D; k += 2) {                                    // p < PB always (staging buffer is PB long)
                    const ulonglong2 xv = *reinterpret_cast<const ulonglong2*>(&s.xs2[p][k]);
                    z = fma2(xv.x, w2[k], z); z = fma2(xv.y, w2[k + 1], z);
                }
                xh[i] = mul2(z, s.rs2[p]);
                float y0, y1;
                up2(fma2(xh[i], g2, be2), y0, y1);
                gy[i] = pk2(y0 > 0.f ? d0 : 0.f, y1 > 0.f ? d1 : 0.f);
                const u64 gh = mul2(gy[i], g2);
                float a0, a1, c0, c1;
                up2(gh, a0, a1);
                up2(mul2(gh, xh[i]), c0, c1);
                part[i][0][t] = a0 + a1;
                part[i][1][t] = c0 + c1;
            }
#pragma unroll
            for (int i = 0; i < PG; ++i) cur[i] = nxt[i];
            __syncthreads();
            {   // warp w reduces point q0 + w
                float a = 0.f, c = 0.f;
#pragma unroll
                for (int j = 0; j < NT / 32; ++j) { a += part[warp][0][lane + 32 * j]; c += part[warp][1][lane + 32 * j]; }
                a = warp_sum(a); c = warp_sum(c);
                if (lane == 0) csum[warp] = make_float2(a * (1.0f / C), c * (1.0f / C));
            }
            __syncthreads();
            // ---- phase 2
#pragma unroll
            for (int i = 0; i < PG; ++i) {
                const int p = q0 + i;
                if (p < np) {                                                       // block-uniform
                    const float2 cs = csum[i];
                    const u64 c1 = pk2(-cs.x, -cs.x), c2 = pk2(-cs.y, -cs.y);
                    // dz = rstd * (g*gamma - c1 - xhat*c2)
                    const u64 dz = mul2(s.rs2[p], add2(fma2(xh[i], c2, mul2(gy[i], g2)), c1));
                    ab = add2(ab, dz);
                    ag = fma2(gy[i], xh[i], ag);
                    abe = add2(abe, gy[i]);
#pragma unroll
                    for (int k = 0; k < D; k += 2) {
                        const ulonglong2 xv = *reinterpret_cast<const ulonglong2*>(&s.xs2[p][k]);
                        aW[k] = fma2(dz, xv.x, aW[k]); aW[k + 1] = fma2(dz, xv.y, aW[k + 1]);
                    }
                }
            }
        }
    }
    float lo, hi;
#pragma unroll
    for (int k = 0; k < D; ++k) { up2(aW[k], lo, hi); atomicAdd(dW + (size_t)(2 * t) * D + k, lo); atomicAdd(dW + (size_t)(2 * t + 1) * D + k, hi); }
    up2(ab, lo, hi); atomicAdd(db + 2 * t, lo); atomicAdd(db + 2 * t + 1, hi);
    up2(ag, lo, hi); atomicAdd(dg + 2 * t, lo); atomicAdd(dg + 2 * t + 1, hi);
    up2(abe, lo, hi); atomicAdd(dbe + 2 * t, lo); atomicAdd(dbe + 2 * t + 1, hi);
}

}  // namespace l1c

// ------------------------------------------------------------------------------------------
// layer 1 forward on the legacy tensor-core path (mma.sync m16n8k8, TF32 operands, fp32 accumulate) with the 3xTF32
// split: x = x_hi + x_lo, W = W_hi + W_lo (each part exactly representable in TF32), z = x_lo W_hi + x_hi W_lo + x_hi W_hi
// -- fp32-grade products (error ~2^-21 of each term), which the un-normalised intensity column needs.
// The channel-stationary SIMT kernel above waits on the FMA pipe (8 packed FMAs per channel pair and point); here a warp
// owns 64 channels (8 n-tiles) for every staged point: weight fragments, bias, gain and shift live in ~80 registers, a
// group of 16 points costs 8 + 2 shared loads, 24 MMAs and a packed epilogue.  LayerNorm statistics as in l1c (centred layer
// + Cholesky factor, rstd per point computed once while staging).
// ------------------------------------------------------------------------------------------
namespace l1m {

constexpr int C = l1c::C, D = l1c::D, NT = l1c::NT, PB = l1c::PB, XS = 12;       // XS: padded row stride of the staged points (bank-conflict free)
typedef unsigned long long u64;

struct Smem {
    l1c::Smem base;          // prologue scratch, Cholesky factor R, column means
    float xh[PB][XS];        // staged points, TF32 "big" parts
    float xl[PB][XS];        // TF32 "small" parts
    float rs[PB];            // rstd per staged point (0 for rows beyond M)
};

__device__ __forceinline__ float tf32_rn(float v) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v)); return __uint_as_float(r); }
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int HDT>
__global__ void __launch_bounds__(NT, 2)
fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ g,
           const float* __restrict__ be, void* __restrict__ h, int M, float eps) {
    __shared__ Smem s;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, gid = lane >> 2, tig = lane & 3;
    {
        u64 w2[D], b2;                                      // the prologue's own per-thread weights are not needed here
        l1c::prologue(s.base, W, b, w2, b2);
    }
    // this thread's fragments for the warp's 8 n-tiles: B[k][n] = Wt[channel n][k], k = tig / tig + 4, n = gid
    uint32_t bh[8][2], bl[8][2];
    u64 bias2[8], g2[8], be2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int chb = warp * 64 + j * 8 + gid;                                   // B fragment column
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int k = tig + 4 * q;
            const float wv = W[(size_t)chb * D + k] - s.base.cm[k];
            const float hi = tf32_rn(wv);
            bh[j][q] = __float_as_uint(hi); bl[j][q] = __float_as_uint(tf32_rn(wv - hi));
        }
        const int chc = warp * 64 + j * 8 + 2 * tig;                               // accumulator columns (chc, chc + 1)
        bias2[j] = l1c::pk2(b[chc] - s.base.cm[8], b[chc + 1] - s.base.cm[8]);
        g2[j] = l1c::pk2(g[chc], g[chc + 1]); be2[j] = l1c::pk2(be[chc], be[chc + 1]);
    }
    const int nblk = (M + PB - 1) / PB;
    float4 xa = make_float4(0, 0, 0, 0), xb = xa;
    int blk = blockIdx.x;
    if (blk < nblk && blk * PB + t < M) { const float4* src = reinterpret_cast<const float4*>(x + (size_t)(blk * PB + t) * D); xa = src[0]; xb = src[1]; }
    for (; blk < nblk; blk += gridDim.x) {
        const int p0 = blk * PB;
        __syncthreads();                                  // previous block's readers are done with the staging buffers
        {   // stage point p0 + t: rstd from the factor R, TF32 split of its 8 features
            const float xv[9] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w, 1.0f};
            float var = 0.f;
            int q = 0;
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                float a = 0.f;
#pragma unroll
                for (int j = i; j < 9; ++j) a = fmaf(s.base.R[q++], xv[j], a);
                var = fmaf(a, a, var);
            }
            s.rs[t] = p0 + t < M ? rsqrtf(var + eps) : 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) { const float hi = tf32_rn(xv[k]); s.xh[t][k] = hi; s.xl[t][k] = tf32_rn(xv[k] - hi); }
        }
        const int nb = blk + gridDim.x;                   // prefetch the next block's point while this one is processed
        if (nb < nblk && nb * PB + t < M) { const float4* src = reinterpret_cast<const float4*>(x + (size_t)(nb * PB + t) * D); xa = src[0]; xb = src[1]; }
        __syncthreads();
        const int np = min(PB, M - p0);
        uint32_t* const otile = reinterpret_cast<uint32_t*>(&s.base.xs2[0][0]) + warp * 512;      // 16 rows x 32 words per warp
        for (int g0 = 0; g0 < np; g0 += 16) {
            const int ra = g0 + gid, rb = ra + 8;         // this thread's two accumulator rows (staged point indices)
            uint32_t ah[4], al[4];
            ah[0] = __float_as_uint(s.xh[ra][tig]); ah[1] = __float_as_uint(s.xh[rb][tig]);
            ah[2] = __float_as_uint(s.xh[ra][tig + 4]); ah[3] = __float_as_uint(s.xh[rb][tig + 4]);
            al[0] = __float_as_uint(s.xl[ra][tig]); al[1] = __float_as_uint(s.xl[rb][tig]);
            al[2] = __float_as_uint(s.xl[ra][tig + 4]); al[3] = __float_as_uint(s.xl[rb][tig + 4]);
            const float rsa = s.rs[ra], rsb = s.rs[rb];
            const u64 rsa2 = l1c::pk2(rsa, rsa), rsb2 = l1c::pk2(rsb, rsb);
            const bool oka = ra < np, okb = rb < np;
            const size_t rowa = (size_t)(p0 + ra) * C, rowb = (size_t)(p0 + rb) * C;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float d[4];
                l1c::up2(bias2[j], d[0], d[1]); d[2] = d[0]; d[3] = d[1];
                mma_tf32(d, al, bh[j][0], bh[j][1]);      // small terms first
                mma_tf32(d, ah, bl[j][0], bl[j][1]);
                mma_tf32(d, ah, bh[j][0], bh[j][1]);
                float y0, y1, y2, y3;
                l1c::up2(l1c::fma2(l1c::mul2(l1c::pk2(d[0], d[1]), rsa2), g2[j], be2[j]), y0, y1);
                l1c::up2(l1c::fma2(l1c::mul2(l1c::pk2(d[2], d[3]), rsb2), g2[j], be2[j]), y2, y3);
                y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); y2 = fmaxf(y2, 0.f); y3 = fmaxf(y3, 0.f);
                const int ch = warp * 64 + j * 8 + 2 * tig;
                if (HDT == WF_BF16) {
                    // the warp's 16 x 64 bf16 tile is collected in shared memory (the l1c staging area, unused here; 4-word
                    // groups XOR-swizzled with the row) and leaves as 128-byte row segments, not as 16-byte pieces
                    const __nv_bfloat162 va = __floats2bfloat162_rn(y0, y1), vb = __floats2bfloat162_rn(y2, y3);
                    const int w = j * 4 + tig;
                    otile[gid * 32 + (w ^ ((gid & 7) << 2))] = *reinterpret_cast<const uint32_t*>(&va);
                    otile[(gid + 8) * 32 + (w ^ ((gid & 7) << 2))] = *reinterpret_cast<const uint32_t*>(&vb);
                } else {
                    if (oka) reinterpret_cast<float2*>(h)[(rowa + ch) >> 1] = make_float2(y0, y1);
                    if (okb) reinterpret_cast<float2*>(h)[(rowb + ch) >> 1] = make_float2(y2, y3);
                }
            }
            if (HDT == WF_BF16) {
                __syncwarp();
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int row = it * 4 + (lane >> 3), c16 = lane & 7;
                    const uint4 v = *reinterpret_cast<const uint4*>(otile + row * 32 + ((c16 ^ (row & 7)) << 2));
                    if (g0 + row < np)
                        *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(h) + (size_t)(p0 + g0 + row) * C + warp * 64 + c16 * 8) = v;
                }
                __syncwarp();
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// layer 1 backward (no dx) on the same tensor-core formulation, bf16 gradient in.  The channel-stationary SIMT kernel
// (l1c::bwd_kernel) spends its time on shared-memory round trips (two staged-point reads per FMA group, row sums through a
// per-thread partial array): 0.67 ms for 0.68 GB, 0.15 of the HBM roofline.  Here the layer is recomputed TRANSPOSED,
//   z^T (16 channels x 8 points) = Wt (16 x 8) . x^T (8 x 8),          3xTF32 as in the forward,
// so that the accumulator fragment of a thread -- channel rows (g, g+8) x points (2t, 2t+1) -- IS the A fragment of the
// weight-gradient product dW (16 channels x 8 inputs) += dz^T (16 x 8 points) . x (8 points x 8 inputs): A wants columns
// (t, t+4), and since the point index is summed over, reading x in the permuted order (2t, 2t+1) makes the two agree
// with no data movement.  A warp owns 64 channels as 4 m-tiles whose rows are assigned so that a thread's 8 channels are
// CONSECUTIVE (channel = 64 warp + 8 g + 2 j + half): its share of a point's gradient row is one 16-byte load, a warp
// reads 4 points x 128 contiguous bytes per instruction, and channel pairs (2j, 2j+1) sit in (c0,c2)/(c1,c3) for the
// packed f32x2 epilogue.  Steps of 16 points: phase 1 forms xhat, the masked gradient (kept as bf16 -- it is dh or 0) and
// the two LayerNorm-backward row sums (in-thread over 8 channels, shuffles over the 8 row groups, one partial per warp
// and point in shared memory, added in warp order: deterministic); ONE __syncthreads per step (the partial buffer is
// double-buffered); phase 2 finishes dz, accumulates db / dgamma / dbeta per thread and dW by 3xTF32 MMAs.
// ------------------------------------------------------------------------------------------
constexpr int BPMAX = 16;                                 // points per step: 8 per n-tile, NTL n-tiles (1 or 2)

struct SmemB {
    l1c::Smem base;          // prologue scratch, Cholesky factor R, column means
    float xh[PB][XS];        // staged points, TF32 "big" parts
    float xl[PB][XS];        // TF32 "small" parts
    float rs[PB];            // rstd per staged point (0 for rows beyond M)
    float2 part[2][BPMAX][NT / 32];                          // [buffer][point of the step][warp] (sum gh, sum gh * xhat)
    uint4 wh[NT / 32][4][32];                             // A fragments of the centred weights, TF32 big parts: [warp][m-tile][lane]
    uint4 wl[NT / 32][4][32];                             // small parts
    u64 bias2[NT / 32][4][8];                             // centred bias pairs: [warp][m-tile][row group]
};

// NTL = 2: 16 points per block-wide barrier, ~170 registers -> one CTA per SM;  NTL = 1: 8 points per barrier, two CTAs per SM
template <int NTL>
__global__ void __launch_bounds__(NT, NTL == 1 ? 2 : 1)
bwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ g,
           const float* __restrict__ be, const __nv_bfloat16* __restrict__ dh, float* __restrict__ dW, float* __restrict__ db,
           float* __restrict__ dg, float* __restrict__ dbe, int M, float eps) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    SmemB& s = *reinterpret_cast<SmemB*>(smem_raw);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, gid = lane >> 2, tig = lane & 3;
    {
        u64 w2[D], b2;
        l1c::prologue(s.base, W, b, w2, b2);
    }
    const int ch8 = warp * 64 + gid * 8;                  // this thread's 8 consecutive channels
    // A fragments of the centred weights for the 4 m-tiles: rows (g, g+8) = channels (ch8 + 2j, ch8 + 2j + 1), columns (t, t+4)
    // (kept in shared memory, fragment-major: two 16-byte loads per m-tile and n-tile instead of 32 registers)
    u64 g2[4], be2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c0 = ch8 + 2 * j, c1 = c0 + 1;
        const float wv[4] = {W[(size_t)c0 * D + tig] - s.base.cm[tig], W[(size_t)c1 * D + tig] - s.base.cm[tig],
                             W[(size_t)c0 * D + tig + 4] - s.base.cm[tig + 4], W[(size_t)c1 * D + tig + 4] - s.base.cm[tig + 4]};
        uint32_t ah[4], al[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float hi = tf32_rn(wv[q]);
            ah[q] = __float_as_uint(hi); al[q] = __float_as_uint(tf32_rn(wv[q] - hi));
        }
        s.wh[warp][j][lane] = make_uint4(ah[0], ah[1], ah[2], ah[3]);
        s.wl[warp][j][lane] = make_uint4(al[0], al[1], al[2], al[3]);
        if (tig == 0) s.bias2[warp][j][gid] = l1c::pk2(b[c0] - s.base.cm[8], b[c1] - s.base.cm[8]);
        g2[j] = l1c::pk2(g[c0], g[c1]); be2[j] = l1c::pk2(be[c0], be[c1]);
    }
    float aW[4][4];                                       // dW accumulators: channels (c0, c0, c1, c1) x inputs (2t, 2t+1, 2t, 2t+1)
    u64 adb[4], adg[4], adbe[4];                          // per channel pair, over this thread's points only
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        adb[j] = adg[j] = adbe[j] = 0ull;
#pragma unroll
        for (int q = 0; q < 4; ++q) aW[j][q] = 0.f;
    }
    constexpr int BP = 8 * NTL;
    const int nblk = (M + PB - 1) / PB;
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    // gradient words of one step: [n-tile][point 2t / 2t+1] -> 8 bf16 of this thread's channels
    uint4 cur[NTL][2], nxt[NTL][2];
    auto load_step = [&](uint4 (&dst)[NTL][2], int pbase, int npts_left) {       // pbase: global index of the step's first point
#pragma unroll
        for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int q = nt * 8 + 2 * tig + e;
                dst[nt][e] = q < npts_left ? *reinterpret_cast<const uint4*>(dh + (size_t)(pbase + q) * C + ch8) : zero4;
            }
    };
    float4 xa = make_float4(0, 0, 0, 0), xb = xa;
    int blk = blockIdx.x;
    if (blk < nblk) {
        if (blk * PB + t < M) { const float4* src = reinterpret_cast<const float4*>(x + (size_t)(blk * PB + t) * D); xa = src[0]; xb = src[1]; }
        load_step(cur, blk * PB, min(PB, M - blk * PB));
    }
    int buf = 0;
    for (; blk < nblk; blk += gridDim.x) {
        const int p0 = blk * PB;
        __syncthreads();                                  // previous block's readers are done with the staging buffers
        {
            const float xv[9] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w, 1.0f};
            float var = 0.f;
            int q = 0;
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                float a = 0.f;
#pragma unroll
                for (int j = i; j < 9; ++j) a = fmaf(s.base.R[q++], xv[j], a);
                var = fmaf(a, a, var);
            }
            s.rs[t] = p0 + t < M ? rsqrtf(var + eps) : 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) { const float hi = tf32_rn(xv[k]); s.xh[t][k] = hi; s.xl[t][k] = tf32_rn(xv[k] - hi); }
        }
        const int nb = blk + gridDim.x;
        if (nb < nblk && nb * PB + t < M) { const float4* src = reinterpret_cast<const float4*>(x + (size_t)(nb * PB + t) * D); xa = src[0]; xb = src[1]; }
        __syncthreads();
        const int np = min(PB, M - p0);
        for (int q0 = 0; q0 < np; q0 += BP) {
            // the next step's gradient words are requested now and consumed one step later (also across staging blocks)
            if (q0 + BP < np) load_step(nxt, p0 + q0 + BP, np - q0 - BP);
            else if (nb < nblk) load_step(nxt, nb * PB, min(PB, M - nb * PB));
            u64 xhat[NTL][4][2];                            // [n-tile][m-tile][point e]: (channel c0, channel c1)
            // ---- phase 1: recompute, mask, row sums
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) {
                const int pr = q0 + nt * 8 + gid;         // B fragment: column n = g -> staged point, rows k = t, t+4
                const uint32_t bh0 = __float_as_uint(s.xh[pr][tig]), bh1 = __float_as_uint(s.xh[pr][tig + 4]);
                const uint32_t bl0 = __float_as_uint(s.xl[pr][tig]), bl1 = __float_as_uint(s.xl[pr][tig + 4]);
                const int pa = q0 + nt * 8 + 2 * tig;     // accumulator columns: staged points pa, pa + 1
                const float rsa = s.rs[pa], rsb = s.rs[pa + 1];
                const u64 rsa2 = l1c::pk2(rsa, rsa), rsb2 = l1c::pk2(rsb, rsb);
                u64 s1a = 0ull, s2a = 0ull, s1b = 0ull, s2b = 0ull;
                uint32_t wa[4] = {cur[nt][0].x, cur[nt][0].y, cur[nt][0].z, cur[nt][0].w};
                uint32_t wb[4] = {cur[nt][1].x, cur[nt][1].y, cur[nt][1].z, cur[nt][1].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float d[4];
                    l1c::up2(s.bias2[warp][j][gid], d[0], d[2]); d[1] = d[0]; d[3] = d[2];
                    {
                        const uint4 fh = s.wh[warp][j][lane], fl = s.wl[warp][j][lane];
                        const uint32_t ah[4] = {fh.x, fh.y, fh.z, fh.w}, al[4] = {fl.x, fl.y, fl.z, fl.w};
                        mma_tf32(d, al, bh0, bh1);
                        mma_tf32(d, ah, bl0, bl1);
                        mma_tf32(d, ah, bh0, bh1);
                    }
                    const u64 xa2 = l1c::mul2(l1c::pk2(d[0], d[2]), rsa2), xb2 = l1c::mul2(l1c::pk2(d[1], d[3]), rsb2);
                    float y0, y1, y2, y3;
                    l1c::up2(l1c::fma2(xa2, g2[j], be2[j]), y0, y1);
                    l1c::up2(l1c::fma2(xb2, g2[j], be2[j]), y2, y3);
                    wa[j] &= (y0 > 0.f ? 0x0000FFFFu : 0u) | (y1 > 0.f ? 0xFFFF0000u : 0u);        // g = dh * [y > 0]
                    wb[j] &= (y2 > 0.f ? 0x0000FFFFu : 0u) | (y3 > 0.f ? 0xFFFF0000u : 0u);
                    const u64 gha = l1c::mul2(lnb::bf2(wa[j]), g2[j]), ghb = l1c::mul2(lnb::bf2(wb[j]), g2[j]);
                    s1a = l1c::add2(s1a, gha); s2a = l1c::fma2(gha, xa2, s2a);
                    s1b = l1c::add2(s1b, ghb); s2b = l1c::fma2(ghb, xb2, s2b);
                    xhat[nt][j][0] = xa2; xhat[nt][j][1] = xb2;
                }
                cur[nt][0] = make_uint4(wa[0], wa[1], wa[2], wa[3]);
                cur[nt][1] = make_uint4(wb[0], wb[1], wb[2], wb[3]);
                float v[4], lo, hi;
                l1c::up2(s1a, lo, hi); v[0] = lo + hi; l1c::up2(s2a, lo, hi); v[1] = lo + hi;
                l1c::up2(s1b, lo, hi); v[2] = lo + hi; l1c::up2(s2b, lo, hi); v[3] = lo + hi;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    v[q] += __shfl_xor_sync(0xffffffffu, v[q], 4);
                    v[q] += __shfl_xor_sync(0xffffffffu, v[q], 8);
                    v[q] += __shfl_xor_sync(0xffffffffu, v[q], 16);
                }
                if (gid == 0) {
                    s.part[buf][nt * 8 + 2 * tig][warp] = make_float2(v[0], v[1]);
                    s.part[buf][nt * 8 + 2 * tig + 1][warp] = make_float2(v[2], v[3]);
                }
            }
            __syncthreads();
            // ---- phase 2: dz, parameter gradients
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) {
                const int pa = q0 + nt * 8 + 2 * tig;
                float c1a = 0.f, c2a = 0.f, c1b = 0.f, c2b = 0.f;
                {
                    const float4* pp = reinterpret_cast<const float4*>(&s.part[buf][nt * 8 + 2 * tig][0]);      // 2 points x 8 warps x (s1, s2)
#pragma unroll
                    for (int w4 = 0; w4 < 4; ++w4) { const float4 u = pp[w4]; c1a += u.x; c2a += u.y; c1a += u.z; c2a += u.w; }
#pragma unroll
                    for (int w4 = 4; w4 < 8; ++w4) { const float4 u = pp[w4]; c1b += u.x; c2b += u.y; c1b += u.z; c2b += u.w; }
                }
                const float rsa = s.rs[pa], rsb = s.rs[pa + 1];
                const u64 rsa2 = l1c::pk2(rsa, rsa), rsb2 = l1c::pk2(rsb, rsb);
                const u64 n1a = l1c::pk2(-c1a * (1.0f / C), -c1a * (1.0f / C)), n2a = l1c::pk2(-c2a * (1.0f / C), -c2a * (1.0f / C));
                const u64 n1b = l1c::pk2(-c1b * (1.0f / C), -c1b * (1.0f / C)), n2b = l1c::pk2(-c2b * (1.0f / C), -c2b * (1.0f / C));
                // B fragment of the weight-gradient product: rows k = (t, t+4) <-> points (pa, pa+1), column n = g = input feature
                const uint32_t xh0 = __float_as_uint(s.xh[pa][gid]), xh1 = __float_as_uint(s.xh[pa + 1][gid]);
                const uint32_t xl0 = __float_as_uint(s.xl[pa][gid]), xl1 = __float_as_uint(s.xl[pa + 1][gid]);
                const uint32_t wa[4] = {cur[nt][0].x, cur[nt][0].y, cur[nt][0].z, cur[nt][0].w};
                const uint32_t wb[4] = {cur[nt][1].x, cur[nt][1].y, cur[nt][1].z, cur[nt][1].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const u64 gya = lnb::bf2(wa[j]), gyb = lnb::bf2(wb[j]);
                    const u64 xa2 = xhat[nt][j][0], xb2 = xhat[nt][j][1];
                    // dz = rstd * (g * gamma - c1 - xhat * c2)
                    const u64 dza = l1c::mul2(rsa2, l1c::add2(l1c::fma2(xa2, n2a, l1c::mul2(gya, g2[j])), n1a));
                    const u64 dzb = l1c::mul2(rsb2, l1c::add2(l1c::fma2(xb2, n2b, l1c::mul2(gyb, g2[j])), n1b));
                    adb[j] = l1c::add2(adb[j], l1c::add2(dza, dzb));
                    adg[j] = l1c::fma2(gya, xa2, l1c::fma2(gyb, xb2, adg[j]));
                    adbe[j] = l1c::add2(adbe[j], l1c::add2(gya, gyb));
                    float e[4];                            // A fragment order: (c0, pa), (c1, pa), (c0, pa+1), (c1, pa+1)
                    l1c::up2(dza, e[0], e[1]); l1c::up2(dzb, e[2], e[3]);
                    uint32_t eh[4], el[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float hi = tf32_rn(e[q]);
                        eh[q] = __float_as_uint(hi); el[q] = __float_as_uint(tf32_rn(e[q] - hi));
                    }
                    mma_tf32(aW[j], el, xh0, xh1);
                    mma_tf32(aW[j], eh, xl0, xl1);
                    mma_tf32(aW[j], eh, xh0, xh1);
                }
            }
            buf ^= 1;
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) { cur[nt][0] = nxt[nt][0]; cur[nt][1] = nxt[nt][1]; }
        }
    }
    // dW: the MMA has already summed over the step's points; one atomic per entry per warp-owner thread
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c0 = ch8 + 2 * j, c1 = c0 + 1;
        atomicAdd(dW + (size_t)c0 * D + 2 * tig, aW[j][0]); atomicAdd(dW + (size_t)c0 * D + 2 * tig + 1, aW[j][1]);
        atomicAdd(dW + (size_t)c1 * D + 2 * tig, aW[j][2]); atomicAdd(dW + (size_t)c1 * D + 2 * tig + 1, aW[j][3]);
        float v[6];
        l1c::up2(adb[j], v[0], v[1]); l1c::up2(adg[j], v[2], v[3]); l1c::up2(adbe[j], v[4], v[5]);
#pragma unroll
        for (int q = 0; q < 6; ++q) {                     // the four point-column lanes of a row group
            v[q] += __shfl_xor_sync(0xffffffffu, v[q], 1);
            v[q] += __shfl_xor_sync(0xffffffffu, v[q], 2);
        }
        if (tig == 0) {
            atomicAdd(db + c0, v[0]); atomicAdd(db + c1, v[1]);
            atomicAdd(dg + c0, v[2]); atomicAdd(dg + c1, v[3]);
            atomicAdd(dbe + c0, v[4]); atomicAdd(dbe + c1, v[5]);
        }
    }
}

}  // namespace l1m

__global__ void stats_finalize_kernel(const float2* __restrict__ st, int M, int parts, float invC, float eps,
                                      float* __restrict__ mean, float* __restrict__ rstd) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    float mu, rs;
    lnb::stats_from_parts(st + i, (size_t)M, parts, invC, eps, mu, rs);
    mean[i] = mu; rstd[i] = rs;
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __float2bfloat16_rn(src[i]);
}
__global__ void cast_bf16_t_kernel(const float* __restrict__ src, int R, int C, __nv_bfloat16* __restrict__ dst) {
    __shared__ float t[32][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int r = blockIdx.y * 32 + j;
        t[j][threadIdx.x] = (r < R && c < C) ? src[(size_t)r * C + c] : 0.f;
    }
    __syncthreads();
    const int r2 = blockIdx.y * 32 + threadIdx.x;          // output column = source row
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int c2 = blockIdx.x * 32 + j;                // output row = source column
        if (c2 < C && r2 < R) dst[(size_t)c2 * R + r2] = __float2bfloat16_rn(t[threadIdx.x][j]);
    }
}

// ------------------------------------------------------------------------------------------
// pooled reductions over points: block = 32 channels x 8 point lanes
// ------------------------------------------------------------------------------------------
__global__ void pool_fwd_kernel(const float* __restrict__ pf, const uint8_t* __restrict__ mask,
                                const float* __restrict__ valid, int N, int C, float* __restrict__ max_m,
                                int* __restrict__ arg_m, float* __restrict__ avg_m, float* __restrict__ max_u,
                                int* __restrict__ arg_u, float* __restrict__ mean_u) {
    const int b = blockIdx.y, c = blockIdx.x * 32 + threadIdx.x, ty = threadIdx.y;
    float mm = -CUDART_INF_F, mu_ = -CUDART_INF_F, sm_ = 0.f, su = 0.f;
    int am = -1, au = -1;
    if (c < C) {
        for (int n = ty; n < N; n += 8) {
            const float v = pf[((size_t)b * N + n) * C + c];
            const bool ok = mask[(size_t)b * N + n] != 0;
            su += v;
            if (v > mu_ || au < 0) { mu_ = v; au = n; }
            if (ok) { sm_ += v; if (v > mm || am < 0) { mm = v; am = n; } }
        }
    }
    __shared__ float s_mm[8][33], s_mu[8][33], s_sm[8][33], s_su[8][33];
    __shared__ int s_am[8][33], s_au[8][33];
    s_mm[ty][threadIdx.x] = mm; s_mu[ty][threadIdx.x] = mu_; s_sm[ty][threadIdx.x] = sm_; s_su[ty][threadIdx.x] = su;
    s_am[ty][threadIdx.x] = am; s_au[ty][threadIdx.x] = au;
    __syncthreads();
    if (ty == 0 && c < C) {
        for (int k = 1; k < 8; ++k) {
            const float v1 = s_mm[k][threadIdx.x]; const int a1 = s_am[k][threadIdx.x];
            if (a1 >= 0 && (am < 0 || v1 > mm || (v1 == mm && a1 < am))) { mm = v1; am = a1; }
            const float v2 = s_mu[k][threadIdx.x]; const int a2 = s_au[k][threadIdx.x];
            if (a2 >= 0 && (au < 0 || v2 > mu_ || (v2 == mu_ && a2 < au))) { mu_ = v2; au = a2; }
            sm_ += s_sm[k][threadIdx.x]; su += s_su[k][threadIdx.x];
        }
        const size_t o = (size_t)b * C + c;
        const bool fin = am >= 0 && isfinite(mm);          // models/PointNetEncoder.py:111
        max_m[o] = fin ? mm : 0.f; arg_m[o] = fin ? am : -1;
        avg_m[o] = sm_ / valid[b];
        max_u[o] = mu_; arg_u[o] = au;
        mean_u[o] = su / (float)N;
    }
}

template <int DDT>
__global__ void pool_bwd_kernel(const float* __restrict__ g_max_m, const float* __restrict__ g_avg_m,
                                const float* __restrict__ g_max_u, const float* __restrict__ g_mean_u,
                                const int* __restrict__ arg_m, const int* __restrict__ arg_u,
                                const uint8_t* __restrict__ mask, const float* __restrict__ valid, int N, int C,
                                void* __restrict__ d_pf) {
    const int b = blockIdx.z;
    const float invN = 1.0f / (float)N;
    const float inv_valid = 1.0f / valid[b];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
        const size_t o = (size_t)b * C + c;
        const float ga = g_avg_m ? g_avg_m[o] * inv_valid : 0.f;
        const float gu = g_mean_u ? g_mean_u[o] * invN : 0.f;
        const float gm = g_max_m ? g_max_m[o] : 0.f, gx = g_max_u ? g_max_u[o] : 0.f;
        const int am = g_max_m ? arg_m[o] : -1, au = g_max_u ? arg_u[o] : -1;
        for (int n = blockIdx.y; n < N; n += gridDim.y) {
            float v = gu;
            if (mask[(size_t)b * N + n]) v += ga;
            if (am == n) v += gm;
            if (au == n) v += gx;
            elem<DDT>::st(d_pf, ((size_t)b * N + n) * C + c, v);
        }
    }
}

// bf16 variant: thread owns 8 consecutive channels (one 16-byte store per point), per-channel terms live in registers
__global__ void __launch_bounds__(256)
pool_bwd_bf16_kernel(const float* __restrict__ g_max_m, const float* __restrict__ g_avg_m, const float* __restrict__ g_max_u,
                     const float* __restrict__ g_mean_u, const int* __restrict__ arg_m, const int* __restrict__ arg_u,
                     const uint8_t* __restrict__ mask, const float* __restrict__ valid, int N, int C, uint4* __restrict__ d_pf,
                     float* __restrict__ dbias) {
    const int b = blockIdx.z;
    const int c0 = threadIdx.x * 8;
    const float invN = 1.0f / (float)N, inv_valid = 1.0f / valid[b];
    float ga[8], gu[8], gm[8], gx[8];
    int am[8], au[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const size_t o = (size_t)b * C + c0 + i;
        ga[i] = g_avg_m ? g_avg_m[o] * inv_valid : 0.f;
        gu[i] = g_mean_u ? g_mean_u[o] * invN : 0.f;
        gm[i] = g_max_m ? g_max_m[o] : 0.f; gx[i] = g_max_u ? g_max_u[o] : 0.f;
        am[i] = g_max_m ? arg_m[o] : -1; au[i] = g_max_u ? arg_u[o] : -1;
    }
    // bias gradient of the final Linear = sum over points of d_pf, in closed form from the pooled gradients
    // (mean terms sum back to the pooled gradient, each max term lands on exactly one point)
    if (dbias != nullptr && blockIdx.y == 0 && threadIdx.y == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float t = gu[i] * (float)N;
            if (am[i] >= 0) t += ga[i] * valid[b] + gm[i];      // am < 0 <=> no valid point in this cloud: masked terms vanish
            if (au[i] >= 0) t += gx[i];
            atomicAdd(dbias + c0 + i, t);
        }
    }
    const int C8 = C >> 3;
    for (int n = blockIdx.y * blockDim.y + threadIdx.y; n < N; n += gridDim.y * blockDim.y) {
        const bool mk = mask[(size_t)b * N + n] != 0;
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            float v0 = gu[i] + (mk ? ga[i] : 0.f), v1 = gu[i + 1] + (mk ? ga[i + 1] : 0.f);
            if (am[i] == n) v0 += gm[i];
            if (au[i] == n) v0 += gx[i];
            if (am[i + 1] == n) v1 += gm[i + 1];
            if (au[i + 1] == n) v1 += gx[i + 1];
            h[i >> 1] = __floats2bfloat162_rn(v0, v1);
        }
        d_pf[((size_t)b * N + n) * C8 + threadIdx.x] = u;
    }
}

}  // namespace enc
}  // namespace wf

extern "C" int wf_point_mask(const float* x, int B, int N, int D, uint8_t* mask, float* valid, wf_stream_t stream) {
    using namespace wf;
    if (B <= 0) return WF_OK;
    const int slab = N < (1 << 24) ? enc::PM_SLAB : N;
    WF_CUDA(cudaMemsetAsync(valid, 0, (size_t)B * sizeof(float), as_stream(stream)));
    if (N > 0) {
        enc::point_mask_kernel<<<dim3(cdiv(N, slab), B), 256, 0, as_stream(stream)>>>(x, N, D, slab, mask, valid);
        WF_LAUNCH_CHECK();
    }
    enc::clamp_valid_kernel<<<cdiv(B, 256), 256, 0, as_stream(stream)>>>(valid, B);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_enc_l1_fwd(const float* x, const float* W, const float* b, const float* gamma, const float* beta, void* h,
                             int h_dtype, int M, int D, int C, float eps, wf_stream_t stream) {
    using namespace wf;
    if (M <= 0) return WF_OK;
    WF_CHECK_ARG(D == 8 && C == 512, "wf_enc_l1_fwd: built for D=8, C=512 (got D=%d C=%d); use wf_gemm_f32 + wf_ln_act_fwd", D, C);
    WF_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0, "wf_enc_l1_fwd: x must be 16-byte aligned");
    WF_CHECK_ARG((reinterpret_cast<uintptr_t>(W) & 15) == 0, "wf_enc_l1_fwd: W must be 16-byte aligned");
    const int grid = min(cdiv(M, enc::l1c::PB), sm_count() * 2);
    // WF_B200_L1_MMA=0 selects the channel-stationary SIMT forward instead of the 3xTF32 tensor-core one
    const char* e_mma = getenv("WF_B200_L1_MMA");           // read per call: the tests flip it
    const bool l1_mma = !(e_mma && e_mma[0] == '0');
    if (l1_mma && (h_dtype == WF_BF16 || h_dtype == WF_F32)) {
        if (h_dtype == WF_BF16) enc::l1m::fwd_kernel<WF_BF16><<<grid, enc::l1m::NT, 0, as_stream(stream)>>>(x, W, b, gamma, beta, h, M, eps);
        else enc::l1m::fwd_kernel<WF_F32><<<grid, enc::l1m::NT, 0, as_stream(stream)>>>(x, W, b, gamma, beta, h, M, eps);
        WF_LAUNCH_CHECK();
        return WF_OK;
    }
    if (h_dtype == WF_BF16) enc::l1c::fwd_kernel<WF_BF16><<<grid, enc::l1c::NT, 0, as_stream(stream)>>>(x, W, b, gamma, beta, h, M, eps);
    else if (h_dtype == WF_F32) enc::l1c::fwd_kernel<WF_F32><<<grid, enc::l1c::NT, 0, as_stream(stream)>>>(x, W, b, gamma, beta, h, M, eps);
    else { set_error("wf_enc_l1_fwd: bad dtype"); return WF_EINVAL; }
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_enc_l1_bwd(const float* x, const float* W, const float* b, const float* gamma, const float* beta,
                             const void* dh, int dh_dtype, float* dW, float* db, float* dgamma, float* dbeta, float* dx, int M,
                             int D, int C, float eps, wf_stream_t stream) {
    using namespace wf;
    if (M <= 0) return WF_OK;
    WF_CHECK_ARG(D == 8 && C == 512, "wf_enc_l1_bwd: built for D=8, C=512 (got D=%d C=%d)", D, C);
    WF_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(W)) & 15) == 0, "wf_enc_l1_bwd: x, W must be 16-byte aligned");
    if (dx == nullptr) {                                   // training: no input gradient
        const int g2 = min(cdiv(M, enc::l1c::PB), sm_count() * 2);
        // WF_B200_L1_BWD=mma1|mma2 selects the 3xTF32 mma.sync kernel (8 / 16 points per barrier) for a bf16 gradient.  Measured
        // on a B200 at 640,000 points: 0.97 / 0.76 ms against 0.67 ms for the channel-stationary SIMT kernel -- legacy
        // mma.sync.m16n8k8.tf32 retires about one instruction per 6 clocks per SM here (the forward kernel shows the same
        // rate), and the backward needs 6 of them per 16 x 8 tile -- so the SIMT kernel stays the default (read per call: the
        // tests flip it)
        const char* e_bwd = getenv("WF_B200_L1_BWD");
        const bool l1_mma = e_bwd && e_bwd[0] == 'm';
        if (l1_mma && dh_dtype == WF_BF16 && (reinterpret_cast<uintptr_t>(dh) & 15) == 0) {
            const int ntl = e_bwd[3] == '2' ? 2 : 1;
            const auto* dhb = static_cast<const __nv_bfloat16*>(dh);
            const int smem = (int)sizeof(enc::l1m::SmemB);
            if (ntl == 2) {
                WF_CUDA(cudaFuncSetAttribute(enc::l1m::bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                enc::l1m::bwd_kernel<2><<<min(cdiv(M, enc::l1c::PB), sm_count()), enc::l1m::NT, smem, as_stream(stream)>>>(
                    x, W, b, gamma, beta, dhb, dW, db, dgamma, dbeta, M, eps);
            } else {
                WF_CUDA(cudaFuncSetAttribute(enc::l1m::bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                enc::l1m::bwd_kernel<1><<<g2, enc::l1m::NT, smem, as_stream(stream)>>>(x, W, b, gamma, beta, dhb, dW, db, dgamma, dbeta, M, eps);
            }
            WF_LAUNCH_CHECK();
            return WF_OK;
        }
        if (dh_dtype == WF_BF16) enc::l1c::bwd_kernel<WF_BF16><<<g2, enc::l1c::NT, 0, as_stream(stream)>>>(x, W, b, gamma, beta, dh, dW, db, dgamma, dbeta, M, eps);
        else if (dh_dtype == WF_F32) enc::l1c::bwd_kernel<WF_F32><<<g2, enc::l1c::NT, 0, as_stream(stream)>>>(x, W, b, gamma, beta, dh, dW, db, dgamma, dbeta, M, eps);
        else { set_error("wf_enc_l1_bwd: bad dtype"); return WF_EINVAL; }
        WF_LAUNCH_CHECK();
        return WF_OK;
    }
    const int grid = min(cdiv(M, 4), sm_count() * 4);
#define WF_L1B(DT, DX) enc::l1_bwd_kernel<8, 8, DT, DX><<<grid, 256, 0, as_stream(stream)>>>(x, W, b, gamma, beta, dh, dW, db, dgamma, dbeta, dx, M, eps)
    if (dh_dtype == WF_BF16) { if (dx) WF_L1B(WF_BF16, true); else WF_L1B(WF_BF16, false); }
    else if (dh_dtype == WF_F32) { if (dx) WF_L1B(WF_F32, true); else WF_L1B(WF_F32, false); }
#undef WF_L1B
    else { set_error("wf_enc_l1_bwd: bad dtype"); return WF_EINVAL; }
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_stats_finalize(const float* rowstats, int M, int C, int parts, float eps, float* mean, float* rstd,
                                 wf_stream_t stream) {
    using namespace wf;
    if (M <= 0) return WF_OK;
    WF_CHECK_ARG(parts >= 1 && (reinterpret_cast<uintptr_t>(rowstats) & 7) == 0, "wf_stats_finalize: parts >= 1, 8-byte aligned rowstats");
    enc::stats_finalize_kernel<<<cdiv(M, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float2*>(rowstats), M, parts,
                                                                            1.0f / (float)C, eps, mean, rstd);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_cast_bf16(const float* src, int R, int C, void* dst, int transpose, wf_stream_t stream) {
    using namespace wf;
    if (R <= 0 || C <= 0) return WF_OK;
    if (!transpose) {
        const size_t n = (size_t)R * C;
        enc::cast_bf16_kernel<<<cdiv((long long)n, 256), 256, 0, as_stream(stream)>>>(src, static_cast<__nv_bfloat16*>(dst), n);
    } else {
        dim3 grid(cdiv(C, 32), cdiv(R, 32)), block(32, 8);
        enc::cast_bf16_t_kernel<<<grid, block, 0, as_stream(stream)>>>(src, R, C, static_cast<__nv_bfloat16*>(dst));
    }
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_pool_fwd(const float* pf, const uint8_t* mask, const float* valid, int B, int N, int C, float* max_m,
                           int32_t* arg_m, float* avg_m, float* max_u, int32_t* arg_u, float* mean_u, wf_stream_t stream) {
    using namespace wf;
    if (B <= 0 || C <= 0) return WF_OK;
    WF_CHECK_ARG(N > 0, "wf_pool_fwd: N must be positive");
    dim3 grid(cdiv(C, 32), B), block(32, 8);
    enc::pool_fwd_kernel<<<grid, block, 0, as_stream(stream)>>>(pf, mask, valid, N, C, max_m, arg_m, avg_m, max_u, arg_u, mean_u);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_pool_bwd(const float* g_max_m, const float* g_avg_m, const float* g_max_u, const float* g_mean_u,
                           const int32_t* arg_m, const int32_t* arg_u, const uint8_t* mask, const float* valid, int B, int N,
                           int C, void* d_pf, int d_dtype, float* dbias, wf_stream_t stream) {
    using namespace wf;
    if (B <= 0 || N <= 0 || C <= 0) return WF_OK;
    WF_CHECK_ARG(B <= 65535, "wf_pool_bwd: B > 65535");
    if (d_dtype == WF_BF16 && C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0 && (reinterpret_cast<uintptr_t>(d_pf) & 15) == 0) {
        dim3 block(C / 8, 256 / (C / 8));
        const int want = cdiv(N, (int)block.y);
        dim3 g2(1, want < 64 ? want : 64, B);
        enc::pool_bwd_bf16_kernel<<<g2, block, 0, as_stream(stream)>>>(g_max_m, g_avg_m, g_max_u, g_mean_u, arg_m, arg_u, mask, valid,
                                                                   N, C, static_cast<uint4*>(d_pf), dbias);
        WF_LAUNCH_CHECK();
        return WF_OK;
    }
    WF_CHECK_ARG(dbias == nullptr, "wf_pool_bwd: dbias is only produced by the bf16 vector kernel (C %% 8 == 0, C <= 2048)");
    dim3 grid(cdiv(C, 256), N < 16384 ? N : 16384, B);
    if (d_dtype == WF_F32) enc::pool_bwd_kernel<WF_F32><<<grid, 256, 0, as_stream(stream)>>>(g_max_m, g_avg_m, g_max_u, g_mean_u, arg_m, arg_u, mask, valid, N, C, d_pf);
    else if (d_dtype == WF_BF16) enc::pool_bwd_kernel<WF_BF16><<<grid, 256, 0, as_stream(stream)>>>(g_max_m, g_avg_m, g_max_u, g_mean_u, arg_m, arg_u, mask, valid, N, C, d_pf);
    else { set_error("wf_pool_bwd: bad dtype"); return WF_EINVAL; }
    WF_LAUNCH_CHECK();
    return WF_OK;
}
