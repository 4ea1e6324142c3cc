// fp32 row-MLP primitives: strided SIMT GEMM, fused LayerNorm+activation(+dropout)(+residual)
// forward/backward, column sums, vertex-head split.  These carry the small-M heads
// (feature_fusion, VertexPredictor, EdgePredictor) and the fp32 parity mode of the encoder;
// the wide encoder layers in production precision run on gemm_tc.cu instead.
#include "wf_common.cuh"

namespace wf {
namespace dense {

// ------------------------------------------------------------------------------------------
// C = alpha * op(A) * op(B) + beta * C + bias      64x64 CTA tile, 16-deep k-slab, 4x4 per thread
// ------------------------------------------------------------------------------------------
constexpr int TK = 16;

// R: register tile R x R per thread, CTA tile 16R x 16R (R = 4: 64 x 64; R = 2: 32 x 32 for products whose output has too few
// 64 x 64 tiles to occupy the GPU -- the 128-row pooled products ran on 16 CTAs)
template <bool TA, bool TB, int R>
__global__ void __launch_bounds__(256)
gemm_f32_kernel(int M, int N, int K, float alpha, const float* __restrict__ A, int lda,
                const float* __restrict__ B, int ldb, float beta, float* __restrict__ C, int ldc,
                const float* __restrict__ bias, int k_chunk) {
    constexpr int TM = 16 * R, TN = 16 * R;
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int tx = tid & 15, ty = tid >> 4;          // 16 x 16 threads, each R x R outputs
    float acc[R][R];
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) acc[i][j] = 0.f;

    // split-K: blockIdx.z owns [kz0, kz1) and accumulates atomically (host pre-zeroes C when beta == 0)
    const int kz0 = blockIdx.z * k_chunk;
    K = min(K, kz0 + k_chunk);
    for (int k0 = kz0; k0 < K; k0 += TK) {
        // ---- stage A (op(A) is M x K)
#pragma unroll
        for (int i = 0; i < R; ++i) {
            int m, k;
            if (!TA) { k = tid & 15; m = (tid >> 4) + 16 * i; }                 // k contiguous in memory
            else     { m = tid & (TM - 1); k = tid / TM + (256 / TM) * i; }     // m contiguous in memory
            const int gm = m0 + m, gk = k0 + k;
            float v = 0.f;
            if (gm < M && gk < K) v = TA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk];
            As[k][m] = v;
        }
        // ---- stage B (op(B) is K x N)
#pragma unroll
        for (int i = 0; i < R; ++i) {
            int n, k;
            if (!TB) { n = tid & (TN - 1); k = tid / TN + (256 / TN) * i; }     // n contiguous
            else     { k = tid & 15; n = (tid >> 4) + 16 * i; }                 // k contiguous
            const int gn = n0 + n, gk = k0 + k;
            float v = 0.f;
            if (gn < N && gk < K) v = TB ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn];
            Bs[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            float a[R], b[R];
#pragma unroll
            for (int i = 0; i < R; ++i) a[i] = As[k][ty * R + i];
#pragma unroll
            for (int j = 0; j < R; ++j) b[j] = Bs[k][tx * R + j];
#pragma unroll
            for (int i = 0; i < R; ++i)
#pragma unroll
                for (int j = 0; j < R; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const int gm = m0 + ty * R + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int gn = n0 + tx * R + j;
            if (gn >= N) continue;
            float v = alpha * acc[i][j];
            float* dst = C + (size_t)gm * ldc + gn;
            if (gridDim.z > 1) {
                if (bias && blockIdx.z == 0) v += bias[gn];
                atomicAdd(dst, v);
            } else {
                if (bias) v += bias[gn];
                if (beta != 0.f) v += beta * (*dst);
                *dst = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// LayerNorm + activation forward: one warp per row
// ------------------------------------------------------------------------------------------
template <int ZDT, int ODT>
__global__ void __launch_bounds__(256)
ln_act_fwd_kernel(const void* __restrict__ z, const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                  const void* __restrict__ residual, const uint8_t* __restrict__ keep, float keep_scale,
                  void* __restrict__ out, float* __restrict__ mean, float* __restrict__ rstd, int stats_in, int M, int C,
                  float eps) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const size_t base = (size_t)row * C;
    float mu = 0.f, rs = 1.f;
    if (gamma != nullptr) {
        if (stats_in) { mu = mean[row]; rs = rstd[row]; }
        else {
            float s = 0.f;
            for (int c = lane; c < C; c += 32) s += elem<ZDT>::ld(z, base + c);
            mu = warp_sum(s) / (float)C;
            float v = 0.f;
            for (int c = lane; c < C; c += 32) { const float d = elem<ZDT>::ld(z, base + c) - mu; v = fmaf(d, d, v); }
            rs = rsqrtf(warp_sum(v) / (float)C + eps);
            if (lane == 0) { if (mean) mean[row] = mu; if (rstd) rstd[row] = rs; }
        }
    }
    for (int c = lane; c < C; c += 32) {
        float y = elem<ZDT>::ld(z, base + c);
        if (gamma != nullptr) y = (y - mu) * rs * gamma[c] + beta[c];
        y = act_f(act, y);
        if (keep != nullptr) y = keep[base + c] ? y * keep_scale : 0.f;
        if (residual != nullptr) y += elem<ODT>::ld(residual, base + c);
        elem<ODT>::st(out, base + c, y);
    }
}

// ------------------------------------------------------------------------------------------
// LayerNorm + activation backward.  Block = 256 threads, ROWS rows.
//   phase 1 (warp per row): the two row reductions  c1 = mean(g*gamma), c2 = mean(g*gamma*xhat)
//   phase 2 (thread per column set): dz = rstd*(g*gamma - c1 - xhat*c2); per-column partial sums
//            of g (dbeta), g*xhat (dgamma) and dz (bias grad) stay in registers, one atomic each.
// ------------------------------------------------------------------------------------------
constexpr int LNB_ROWS = 32;

template <int GDT, int ZDT, int DDT, int CPT>
__global__ void __launch_bounds__(256)
ln_act_bwd_kernel(const void* __restrict__ dout, const void* __restrict__ z, const float* __restrict__ gamma,
                  const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ rstd, int act,
                  const uint8_t* __restrict__ keep, float keep_scale, void* __restrict__ dz, float* __restrict__ dgamma,
                  float* __restrict__ dbeta, float* __restrict__ dcolsum, int M, int C) {
    __shared__ float c1s[LNB_ROWS], c2s[LNB_ROWS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = blockIdx.x * LNB_ROWS;
    const bool has_ln = gamma != nullptr;
    if (has_ln) {
        for (int rr = warp; rr < LNB_ROWS; rr += 8) {
            const int row = r0 + rr;
            if (row >= M) break;
            const size_t base = (size_t)row * C;
            const float mu = mean[row], rs = rstd[row];
            float a = 0.f, b = 0.f;
            for (int c = lane; c < C; c += 32) {
                const float xh = (elem<ZDT>::ld(z, base + c) - mu) * rs;
                const float y = xh * gamma[c] + beta[c];
                float g = elem<GDT>::ld(dout, base + c);
                if (keep != nullptr) g = keep[base + c] ? g * keep_scale : 0.f;
                g *= act_grad_f(act, y) * gamma[c];
                a += g; b = fmaf(g, xh, b);
            }
            a = warp_sum(a); b = warp_sum(b);
            if (lane == 0) { c1s[rr] = a / (float)C; c2s[rr] = b / (float)C; }
        }
    }
    __syncthreads();
    float acc_g[CPT], acc_gx[CPT], acc_dz[CPT], gam[CPT], bet[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        acc_g[i] = acc_gx[i] = acc_dz[i] = 0.f;
        const int c = tid + 256 * i;
        gam[i] = (has_ln && c < C) ? gamma[c] : 1.f;
        bet[i] = (has_ln && c < C) ? beta[c] : 0.f;
    }
    for (int rr = 0; rr < LNB_ROWS; ++rr) {
        const int row = r0 + rr;
        if (row >= M) break;
        const size_t base = (size_t)row * C;
        const float mu = has_ln ? mean[row] : 0.f, rs = has_ln ? rstd[row] : 1.f;
        const float c1 = has_ln ? c1s[rr] : 0.f, c2 = has_ln ? c2s[rr] : 0.f;
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
            const int c = tid + 256 * i;
            if (c >= C) continue;
            const float zv = elem<ZDT>::ld(z, base + c);
            const float xh = has_ln ? (zv - mu) * rs : zv;
            const float y = has_ln ? xh * gam[i] + bet[i] : zv;
            float g = elem<GDT>::ld(dout, base + c);
            if (keep != nullptr) g = keep[base + c] ? g * keep_scale : 0.f;
            g *= act_grad_f(act, y);
            float d;
            if (has_ln) {
                d = rs * (g * gam[i] - c1 - xh * c2);
                acc_g[i] += g; acc_gx[i] = fmaf(g, xh, acc_gx[i]);
            } else {
                d = g;
            }
            acc_dz[i] += d;
            elem<DDT>::st(dz, base + c, d);
        }
    }
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        const int c = tid + 256 * i;
        if (c >= C) continue;
        if (has_ln && dgamma) atomicAdd(dgamma + c, acc_gx[i]);
        if (has_ln && dbeta) atomicAdd(dbeta + c, acc_g[i]);
        if (dcolsum) atomicAdd(dcolsum + c, acc_dz[i]);
    }
}

// ------------------------------------------------------------------------------------------
// Vectorised fp32 LayerNorm+activation kernels (C % 4 == 0): a row is owned by NT threads -- one warp (C <= 1024: the
// edge head's E x {128,256,512} tensors, reductions by shuffles only) or a whole CTA of 256 threads (C up to 8192: the
// 64-row heads, where a row per CTA is what gives the launch any parallelism).  Thread t holds float4 columns
// t + NT*i, i < VPT, of its row in registers: one pass over memory forward, one pass backward.
// ------------------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ float row_sum(float v, float* red) {          // red: NT/32 floats of shared memory (NT > 32)
    v = warp_sum(v);
    if (NT == 32) return v;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) t += red[w];
    return t;
}

template <int NT, int VPT>
__global__ void __launch_bounds__(256)
ln_act_fwd_v4_kernel(const float* __restrict__ z, const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                     const float* __restrict__ residual, const uint8_t* __restrict__ keep, float keep_scale,
                     float* __restrict__ out, float* __restrict__ mean, float* __restrict__ rstd, int M, int C, float eps) {
    __shared__ float red[8];
    const int t = NT == 32 ? (threadIdx.x & 31) : threadIdx.x;
    const int rows_per_cta = 256 / NT;
    const int C4 = C >> 2;
    float4 gm[VPT], bt[VPT];
    if (gamma != nullptr) {
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int j = t + NT * i;
            gm[i] = j < C4 ? reinterpret_cast<const float4*>(gamma)[j] : make_float4(0, 0, 0, 0);
            bt[i] = j < C4 ? reinterpret_cast<const float4*>(beta)[j] : make_float4(0, 0, 0, 0);
        }
    }
    for (int row = blockIdx.x * rows_per_cta + (NT == 32 ? (threadIdx.x >> 5) : 0); row < M; row += gridDim.x * rows_per_cta) {
        const size_t base4 = (size_t)row * C4;
        float4 v[VPT];
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int j = t + NT * i;
            v[i] = j < C4 ? reinterpret_cast<const float4*>(z)[base4 + j] : make_float4(0, 0, 0, 0);
        }
        float mu = 0.f, rs = 1.f;
        if (gamma != nullptr) {
            float sm = 0.f;
#pragma unroll
            for (int i = 0; i < VPT; ++i) sm += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            mu = row_sum<NT>(sm, red) / (float)C;
            float var = 0.f;
#pragma unroll
            for (int i = 0; i < VPT; ++i) {
                if (t + NT * i < C4) {
                    const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
                    var += a * a + b * b + c * c + d * d;
                }
            }
            rs = rsqrtf(row_sum<NT>(var, red) / (float)C + eps);
            if (t == 0) { if (mean) mean[row] = mu; if (rstd) rstd[row] = rs; }
        }
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int j = t + NT * i;
            if (j >= C4) continue;
            float y[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
            if (gamma != nullptr) {
                const float g4[4] = {gm[i].x, gm[i].y, gm[i].z, gm[i].w}, b4[4] = {bt[i].x, bt[i].y, bt[i].z, bt[i].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) y[e] = (y[e] - mu) * rs * g4[e] + b4[e];
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) y[e] = act_f(act, y[e]);
            if (keep != nullptr) {
                const uchar4 k = reinterpret_cast<const uchar4*>(keep)[base4 + j];
                y[0] = k.x ? y[0] * keep_scale : 0.f; y[1] = k.y ? y[1] * keep_scale : 0.f;
                y[2] = k.z ? y[2] * keep_scale : 0.f; y[3] = k.w ? y[3] * keep_scale : 0.f;
            }
            if (residual != nullptr) {
                const float4 r = reinterpret_cast<const float4*>(residual)[base4 + j];
                y[0] += r.x; y[1] += r.y; y[2] += r.z; y[3] += r.w;
            }
            reinterpret_cast<float4*>(out)[base4 + j] = make_float4(y[0], y[1], y[2], y[3]);
        }
    }
}

// backward: column partial sums (dgamma, dbeta, bias gradient) stay in 12*VPT registers over all rows a thread sees;
// NT == 32: the CTA's 8 warps are combined through shared memory first, then one atomic per column per CTA.
template <int NT, int VPT>
__global__ void __launch_bounds__(256, (NT == 32 && VPT == 4) ? 2 : 1)     // 512-wide rows per warp: two CTAs per SM (<= 128 registers)
ln_act_bwd_v4_kernel(const float* __restrict__ dout, const float* __restrict__ z, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ rstd, int act,
                     const uint8_t* __restrict__ keep, float keep_scale, float* __restrict__ dz, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, float* __restrict__ dcolsum, int M, int C, int rows_per_owner) {
    __shared__ float red[8];
    __shared__ float comb[NT == 32 ? 8 : 1][NT == 32 ? 3 * 4 * VPT * 32 : 1];
    const int t = NT == 32 ? (threadIdx.x & 31) : threadIdx.x;
    const int owner = NT == 32 ? blockIdx.x * 8 + (threadIdx.x >> 5) : blockIdx.x;      // a warp or a CTA
    const int C4 = C >> 2;
    const bool has_ln = gamma != nullptr;
    float4 gm[VPT], bt[VPT];
    float acc_g[VPT][4], acc_gx[VPT][4], acc_dz[VPT][4];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int j = t + NT * i;
        gm[i] = (has_ln && j < C4) ? reinterpret_cast<const float4*>(gamma)[j] : make_float4(1, 1, 1, 1);
        bt[i] = (has_ln && j < C4) ? reinterpret_cast<const float4*>(beta)[j] : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int e = 0; e < 4; ++e) acc_g[i][e] = acc_gx[i][e] = acc_dz[i][e] = 0.f;
    }
    const int r_begin = owner * rows_per_owner, r_end = min(M, r_begin + rows_per_owner);
    for (int row = r_begin; row < r_end; ++row) {
        const size_t base4 = (size_t)row * C4;
        const float mu = has_ln ? mean[row] : 0.f, rs = has_ln ? rstd[row] : 1.f;
        float xh[VPT][4], g[VPT][4];
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int j = t + NT * i;
            if (j < C4) {
                const float4 zv = reinterpret_cast<const float4*>(z)[base4 + j];
                const float4 dv = reinterpret_cast<const float4*>(dout)[base4 + j];
                const float z4[4] = {zv.x, zv.y, zv.z, zv.w};
                float d4[4] = {dv.x, dv.y, dv.z, dv.w};
                if (keep != nullptr) {
                    const uchar4 k = reinterpret_cast<const uchar4*>(keep)[base4 + j];
                    d4[0] = k.x ? d4[0] * keep_scale : 0.f; d4[1] = k.y ? d4[1] * keep_scale : 0.f;
                    d4[2] = k.z ? d4[2] * keep_scale : 0.f; d4[3] = k.w ? d4[3] * keep_scale : 0.f;
                }
                const float g4[4] = {gm[i].x, gm[i].y, gm[i].z, gm[i].w}, b4[4] = {bt[i].x, bt[i].y, bt[i].z, bt[i].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    xh[i][e] = has_ln ? (z4[e] - mu) * rs : z4[e];
                    const float y = has_ln ? xh[i][e] * g4[e] + b4[e] : z4[e];
                    g[i][e] = d4[e] * act_grad_f(act, y);
                    const float gh = g[i][e] * g4[e];
                    a += gh; b = fmaf(gh, xh[i][e], b);
                }
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) xh[i][e] = g[i][e] = 0.f;
            }
        }
        float c1 = 0.f, c2 = 0.f;
        if (has_ln) { c1 = row_sum<NT>(a, red) / (float)C; c2 = row_sum<NT>(b, red) / (float)C; }
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int j = t + NT * i;
            if (j >= C4) continue;
            const float g4[4] = {gm[i].x, gm[i].y, gm[i].z, gm[i].w};
            float d[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (has_ln) {
                    d[e] = rs * (g[i][e] * g4[e] - c1 - xh[i][e] * c2);
                    acc_g[i][e] += g[i][e]; acc_gx[i][e] = fmaf(g[i][e], xh[i][e], acc_gx[i][e]);
                } else {
                    d[e] = g[i][e];
                }
                acc_dz[i][e] += d[e];
            }
            reinterpret_cast<float4*>(dz)[base4 + j] = make_float4(d[0], d[1], d[2], d[3]);
        }
    }
    if (NT == 32) {
        // combine the 8 warps of the CTA, then one atomic per column
        const int warp = threadIdx.x >> 5;
#pragma unroll
        for (int i = 0; i < VPT; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int slot = (i * 4 + e) * 32 + t;
                comb[warp][slot] = acc_gx[i][e]; comb[warp][4 * VPT * 32 + slot] = acc_g[i][e]; comb[warp][8 * VPT * 32 + slot] = acc_dz[i][e];
            }
        __syncthreads();
        for (int s = threadIdx.x; s < 3 * 4 * VPT * 32; s += 256) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) v += comb[w][s];
            const int kind = s / (4 * VPT * 32), r = s - kind * (4 * VPT * 32);
            const int ie = r >> 5, tt = r & 31, i = ie >> 2, e = ie & 3;
            const int c = 4 * (tt + 32 * i) + e;
            if (c < C) {
                if (kind == 0) { if (has_ln && dgamma) atomicAdd(dgamma + c, v); }
                else if (kind == 1) { if (has_ln && dbeta) atomicAdd(dbeta + c, v); }
                else if (dcolsum) atomicAdd(dcolsum + c, v);
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int j = t + NT * i;
            if (j >= C4) continue;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = 4 * j + e;
                if (has_ln && dgamma) atomicAdd(dgamma + c, acc_gx[i][e]);
                if (has_ln && dbeta) atomicAdd(dbeta + c, acc_g[i][e]);
                if (dcolsum) atomicAdd(dcolsum + c, acc_dz[i][e]);
            }
        }
    }
}

// column sums: block (32 columns x 8 row lanes), chunk of rows per block, one atomic per column
template <int DT>
__global__ void colsum_kernel(const void* __restrict__ x, int M, int C, int ld, float* __restrict__ out, int rows_per_block) {
    __shared__ float part[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
    float s = 0.f;
    if (c < C)
        for (int r = r0 + threadIdx.y; r < r1; r += 8) s += elem<DT>::ld(x, (size_t)r * ld + c);
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x];
        atomicAdd(out + c, t);
    }
}

__global__ void vertex_split_fwd_kernel(const float* __restrict__ vf, int B, int V, float* __restrict__ coords,
                                        float* __restrict__ prob, long long* __restrict__ count) {
    const int b = blockIdx.x;
    int local = 0;
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
        const float* s = vf + ((size_t)b * V + v) * 4;
        float* d = coords + ((size_t)b * V + v) * 3;
        d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
        const float p = 1.0f / (1.0f + expf(-s[3]));
        prob[(size_t)b * V + v] = p;
        local += p > 0.5f ? 1 : 0;
    }
    __shared__ int tot;
    if (threadIdx.x == 0) tot = 0;
    __syncthreads();
    atomicAdd(&tot, local);
    __syncthreads();
    if (threadIdx.x == 0) count[b] = tot;
}

__global__ void vertex_split_bwd_kernel(const float* __restrict__ d_coords, const float* __restrict__ d_prob,
                                        const float* __restrict__ prob, int n, float* __restrict__ d_vf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float* d = d_vf + (size_t)i * 4;
    d[0] = d_coords ? d_coords[(size_t)i * 3 + 0] : 0.f;
    d[1] = d_coords ? d_coords[(size_t)i * 3 + 1] : 0.f;
    d[2] = d_coords ? d_coords[(size_t)i * 3 + 2] : 0.f;
    const float p = prob[i];
    d[3] = d_prob ? d_prob[i] * p * (1.0f - p) : 0.f;
}

}  // namespace dense
}  // namespace wf

extern "C" int wf_gemm_f32(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda, const float* B,
                           int ldb, float beta, float* C, int ldc, const float* bias, wf_stream_t stream) {
    using namespace wf;
    using namespace wf::dense;
    if (M <= 0 || N <= 0) return WF_OK;
    WF_CHECK_ARG(K >= 0 && lda > 0 && ldb > 0 && ldc >= N, "wf_gemm_f32: bad dims");
    // few 64 x 64 tiles: 32 x 32 tiles put 4x the CTAs on the product (same summation order per output: results are identical)
    const bool small = (long long)cdiv(N, 64) * cdiv(M, 64) * 2 <= sm_count();
    const int T = small ? 32 : 64;
    dim3 grid(cdiv(N, T), cdiv(M, T));
    cudaStream_t s = as_stream(stream);
    // split-K when the output is small and the reduction long (weight gradients of the edge head: K = #edges)
    int split = 1;
    const long long ctas = (long long)grid.x * grid.y;
    // (transA only: forward / dX products keep a fixed summation order, so duplicate input rows stay bit-identical)
    if (transA && ctas < sm_count() && K >= 1024 && (beta == 0.f || beta == 1.f)) {
        split = (int)((2LL * sm_count() + ctas - 1) / ctas);
        if (split > K / 256) split = K / 256;
        if (split < 1) split = 1;
    }
    int k_chunk = K;
    if (split > 1) {
        k_chunk = ((K + split - 1) / split + TK - 1) / TK * TK;
        split = (K + k_chunk - 1) / k_chunk;
        if (beta == 0.f) WF_CUDA(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), M, s));
    }
    grid.z = split;
#define WF_GF(TA_, TB_) do { if (small) gemm_f32_kernel<TA_, TB_, 2><<<grid, 256, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, k_chunk); \
                             else gemm_f32_kernel<TA_, TB_, 4><<<grid, 256, 0, s>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias, k_chunk); } while (0)
    if (!transA && !transB) WF_GF(false, false);
    else if (!transA && transB) WF_GF(false, true);
    else if (transA && !transB) WF_GF(true, false);
    else WF_GF(true, true);
#undef WF_GF
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_ln_act_fwd(const void* z, int z_dtype, const float* gamma, const float* beta, int act, const void* residual,
                             const uint8_t* keep, float keep_scale, void* out, int out_dtype, float* mean, float* rstd,
                             int stats_in, int M, int C, float eps, wf_stream_t stream) {
    using namespace wf;
    using namespace wf::dense;
    if (M <= 0 || C <= 0) return WF_OK;
    WF_CHECK_ARG(!(gamma && !beta), "wf_ln_act_fwd: gamma without beta");
    WF_CHECK_ARG(!(stats_in && (!mean || !rstd)), "wf_ln_act_fwd: stats_in needs mean/rstd");
    cudaStream_t s = as_stream(stream);
    const bool al16 = ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(gamma) |
                        reinterpret_cast<uintptr_t>(beta) | reinterpret_cast<uintptr_t>(residual)) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(keep) & 3) == 0;
    if (z_dtype == WF_F32 && out_dtype == WF_F32 && !stats_in && C % 4 == 0 && C <= 8192 && al16) {
        const float* zf = static_cast<const float*>(z); const float* rf = static_cast<const float*>(residual);
        float* of = static_cast<float*>(out);
#define WF_LNV(NT, VPT, G) ln_act_fwd_v4_kernel<NT, VPT><<<G, 256, 0, s>>>(zf, gamma, beta, act, rf, keep, keep_scale, of, mean, rstd, M, C, eps)
        if (C <= 512 || (C <= 1024 && M >= 4096)) {          // warp per row
            const int g = min(cdiv(M, 8), sm_count() * 16);
            if (C <= 128) WF_LNV(32, 1, g); else if (C <= 256) WF_LNV(32, 2, g); else if (C <= 512) WF_LNV(32, 4, g); else WF_LNV(32, 8, g);
        } else {                                              // CTA per row
            const int g = min(M, sm_count() * 16);
            if (C <= 1024) WF_LNV(256, 1, g); else if (C <= 2048) WF_LNV(256, 2, g); else if (C <= 4096) WF_LNV(256, 4, g); else WF_LNV(256, 8, g);
        }
#undef WF_LNV
        WF_LAUNCH_CHECK();
        return WF_OK;
    }
    dim3 grid(cdiv(M, 8));
#define WF_LNF(ZD, OD) ln_act_fwd_kernel<ZD, OD><<<grid, 256, 0, s>>>(z, gamma, beta, act, residual, keep, keep_scale, out, mean, rstd, stats_in, M, C, eps)
    if (z_dtype == WF_F32 && out_dtype == WF_F32) WF_LNF(WF_F32, WF_F32);
    else if (z_dtype == WF_BF16 && out_dtype == WF_BF16) WF_LNF(WF_BF16, WF_BF16);
    else if (z_dtype == WF_F32 && out_dtype == WF_BF16) WF_LNF(WF_F32, WF_BF16);
    else if (z_dtype == WF_BF16 && out_dtype == WF_F32) WF_LNF(WF_BF16, WF_F32);
    else { set_error("wf_ln_act_fwd: bad dtypes"); return WF_EINVAL; }
#undef WF_LNF
    WF_LAUNCH_CHECK();
    return WF_OK;
}

namespace wf { namespace dense {
template <int GD, int ZD, int DD>
static int launch_ln_bwd(const void* dout, const void* z, const float* gamma, const float* beta, const float* mean,
                         const float* rstd, int act, const uint8_t* keep, float keep_scale, void* dz, float* dgamma,
                         float* dbeta, float* dcolsum, int M, int C, cudaStream_t s) {
    dim3 grid(cdiv(M, LNB_ROWS));
#define WF_LNB(CPT) ln_act_bwd_kernel<GD, ZD, DD, CPT><<<grid, 256, 0, s>>>(dout, z, gamma, beta, mean, rstd, act, keep, keep_scale, dz, dgamma, dbeta, dcolsum, M, C)
    if (C <= 256) WF_LNB(1);
    else if (C <= 512) WF_LNB(2);
    else if (C <= 1024) WF_LNB(4);
    else if (C <= 2048) WF_LNB(8);
    else if (C <= 4096) WF_LNB(16);
    else { set_error("wf_ln_act_bwd: C=%d > 4096 not built", C); return WF_EUNSUPPORTED; }
#undef WF_LNB
    return WF_OK;
}
}}  // namespace wf::dense

extern "C" int wf_ln_act_bwd(const void* dout, int dout_dtype, const void* z, int z_dtype, const float* gamma,
                             const float* beta, const float* mean, const float* rstd, int act, const uint8_t* keep,
                             float keep_scale, void* dz, int dz_dtype, float* dgamma, float* dbeta, float* dcolsum, int M,
                             int C, wf_stream_t stream) {
    using namespace wf;
    using namespace wf::dense;
    if (M <= 0 || C <= 0) return WF_OK;
    WF_CHECK_ARG(!(gamma && (!beta || !mean || !rstd)), "wf_ln_act_bwd: LayerNorm backward needs beta, mean, rstd");
    cudaStream_t s = as_stream(stream);
    int rc;
    const bool al16 = ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dz) |
                        reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(keep) & 3) == 0;
    if (dout_dtype == WF_F32 && z_dtype == WF_F32 && dz_dtype == WF_F32 && C % 4 == 0 && C <= 8192 && al16) {
        const float* df = static_cast<const float*>(dout); const float* zf = static_cast<const float*>(z);
        float* of = static_cast<float*>(dz);
#define WF_LNV(NT, VPT, G, RPO) ln_act_bwd_v4_kernel<NT, VPT><<<G, 256, 0, s>>>(df, zf, gamma, beta, mean, rstd, act, keep, keep_scale, of, dgamma, dbeta, dcolsum, M, C, RPO)
        if (C <= 512) {                                       // warp per row; a warp owns rpo consecutive rows
            int rpo = cdiv(M, sm_count() * 8 * 8);            // ~8 CTAs of 8 warps per SM
            rpo = rpo < 1 ? 1 : (rpo > 16 ? 16 : rpo);
            const int g = cdiv(M, 8 * rpo);
            if (C <= 128) WF_LNV(32, 1, g, rpo); else if (C <= 256) WF_LNV(32, 2, g, rpo); else WF_LNV(32, 4, g, rpo);
        } else {                                              // CTA per row(s)
            int rpo = cdiv(M, sm_count() * 8);
            rpo = rpo < 1 ? 1 : rpo;
            const int g = cdiv(M, rpo);
            if (C <= 1024) WF_LNV(256, 1, g, rpo); else if (C <= 2048) WF_LNV(256, 2, g, rpo); else if (C <= 4096) WF_LNV(256, 4, g, rpo); else WF_LNV(256, 8, g, rpo);
        }
#undef WF_LNV
        WF_LAUNCH_CHECK();
        return WF_OK;
    }
    if (dout_dtype == WF_F32 && z_dtype == WF_F32 && dz_dtype == WF_F32)
        rc = launch_ln_bwd<WF_F32, WF_F32, WF_F32>(dout, z, gamma, beta, mean, rstd, act, keep, keep_scale, dz, dgamma, dbeta, dcolsum, M, C, s);
    else if (dout_dtype == WF_BF16 && z_dtype == WF_BF16 && dz_dtype == WF_BF16)
        rc = launch_ln_bwd<WF_BF16, WF_BF16, WF_BF16>(dout, z, gamma, beta, mean, rstd, act, keep, keep_scale, dz, dgamma, dbeta, dcolsum, M, C, s);
    else { set_error("wf_ln_act_bwd: dtype combination not built"); return WF_EUNSUPPORTED; }
    if (rc != WF_OK) return rc;
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_colsum(const void* x, int dtype, int M, int C, int ld, float* out, wf_stream_t stream) {
    using namespace wf;
    if (M <= 0 || C <= 0) return WF_OK;
    const int rows_per_block = 512;
    dim3 grid(cdiv(C, 32), cdiv(M, rows_per_block)), block(32, 8);
    if (dtype == WF_F32) dense::colsum_kernel<WF_F32><<<grid, block, 0, as_stream(stream)>>>(x, M, C, ld, out, rows_per_block);
    else if (dtype == WF_BF16) dense::colsum_kernel<WF_BF16><<<grid, block, 0, as_stream(stream)>>>(x, M, C, ld, out, rows_per_block);
    else { set_error("wf_colsum: bad dtype"); return WF_EINVAL; }
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_vertex_split_fwd(const float* vf, int B, int V, float* coords, float* prob, int64_t* count, wf_stream_t stream) {
    using namespace wf;
    if (B <= 0 || V <= 0) return WF_OK;
    dense::vertex_split_fwd_kernel<<<B, 128, 0, as_stream(stream)>>>(vf, B, V, coords, prob, reinterpret_cast<long long*>(count));
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_vertex_split_bwd(const float* d_coords, const float* d_prob, const float* prob, int B, int V, float* d_vf,
                                   wf_stream_t stream) {
    using namespace wf;
    const int n = B * V;
    if (n <= 0) return WF_OK;
    dense::vertex_split_bwd_kernel<<<cdiv(n, 256), 256, 0, as_stream(stream)>>>(d_coords, d_prob, prob, n, d_vf);
    WF_LAUNCH_CHECK();
    return WF_OK;
}
