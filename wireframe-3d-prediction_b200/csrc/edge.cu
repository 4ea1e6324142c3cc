// Edge head kernels on a ragged batch (all samples' vertices concatenated, CSR offsets):
// prefix gather/scatter, the 8-head self-attention core, the all-pairs first edge layer in its
// decomposed form, and the final 128->1 + sigmoid + zero-padding.
// Reference: models/EdgePredictor.py:91-140, models/PointCloudToWireframe.py:77-112.
#include "wf_common.cuh"

#include <math_constants.h>

namespace wf {
namespace edge {

__device__ __forceinline__ int find_segment(const int* __restrict__ off, int B, int t) {
    int lo = 0, hi = B;                       // largest b with off[b] <= t
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off[mid] <= t) lo = mid; else hi = mid; }
    return lo;
}
// index of pair (i, j), i < j, among the c(c-1)/2 pairs in row-major order
__device__ __forceinline__ long long pair_id(int i, int j, int c) {
    return (long long)i * (2 * c - i - 1) / 2 + (j - i - 1);
}

__global__ void gather_prefix_kernel(const float* __restrict__ verts, int B, int V, const int* __restrict__ v_off,
                                     int T, float* __restrict__ packed) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= T * 3) return;
    const int t = k / 3, d = k - t * 3;
    const int b = find_segment(v_off, B, t);
    packed[k] = verts[((size_t)b * V + (t - v_off[b])) * 3 + d];
}
__global__ void scatter_prefix_add_kernel(const float* __restrict__ d_packed, int B, int V, const int* __restrict__ v_off,
                                          int T, float* __restrict__ d_verts) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= T * 3) return;
    const int t = k / 3, d = k - t * 3;
    const int b = find_segment(v_off, B, t);
    d_verts[((size_t)b * V + (t - v_off[b])) * 3 + d] += d_packed[k];     // each target written by one thread
}

// ------------------------------------------------------------------------------------------
// attention: one CTA per (sample, head); everything for that head lives in shared memory
// ------------------------------------------------------------------------------------------
constexpr int HD_PAD = 1;

// 4 x 4 register tile of a product of two shared-memory matrices given by strides:
//   acc[u][v] = sum_{k < K} A[(i0 + u) * sai + k * sak] * B[k * sbk + (j0 + v) * sbj],   k ascending (one fmaf chain per output,
// the same order as a scalar loop).  8 shared loads per 16 FMAs instead of 2 per FMA.  Rows/columns beyond the matrices'
// logical extent are read (the arrays are padded to multiples of 4 rows) and their results discarded by the caller.
__device__ __forceinline__ void smem_mm_4x4(const float* __restrict__ A, int sai, int sak, const float* __restrict__ B, int sbk,
                                            int sbj, int K, int i0, int j0, float (&acc)[4][4]) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = 0.f;
    const float* a0 = A + i0 * sai;
    const float* b0 = B + j0 * sbj;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        float a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] = a0[u * sai + k * sak];
#pragma unroll
        for (int v = 0; v < 4; ++v) b[v] = b0[k * sbk + v * sbj];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
    }
}

template <int HD>
__global__ void __launch_bounds__(256)
attn_fwd_kernel(const float* __restrict__ qkv, const int* __restrict__ v_off, const long long* __restrict__ p_off,
                int heads, float* __restrict__ out, float* __restrict__ probs, const uint8_t* __restrict__ keep,
                float keep_scale) {
    extern __shared__ float sm[];
    const int b = blockIdx.x / heads, h = blockIdx.x % heads;
    const int t0 = v_off[b], c = v_off[b + 1] - t0;
    if (c <= 0) return;
    const int cp = (c + 3) & ~3;                       // rows padded for the 4 x 4 tiles
    const int E = heads * HD, LD = HD + HD_PAD, LS = cp + 1;
    float* Q = sm; float* K = Q + cp * LD; float* Vv = K + cp * LD; float* S = Vv + cp * LD;
    const float scale = rsqrtf((float)HD);
    for (int idx = threadIdx.x; idx < c * HD; idx += blockDim.x) {
        const int i = idx / HD, d = idx - i * HD;
        const float* row = qkv + (size_t)(t0 + i) * 3 * E + h * HD + d;
        Q[i * LD + d] = row[0] * scale; K[i * LD + d] = row[E]; Vv[i * LD + d] = row[2 * E];
    }
    __syncthreads();
    const int tc = cp >> 2;
    for (int tile = threadIdx.x; tile < tc * tc; tile += blockDim.x) {              // S = Qs K^T
        const int i0 = (tile / tc) * 4, j0 = (tile % tc) * 4;
        float acc[4][4];
        smem_mm_4x4(Q, LD, 1, K, 1, LD, HD, i0, j0, acc);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) S[(i0 + u) * LS + j0 + v] = acc[u][v];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long pbase = probs || keep ? p_off[b] + (long long)h * c * c : 0;
    for (int i = warp; i < c; i += (blockDim.x >> 5)) {
        float m = -CUDART_INF_F;
        for (int j = lane; j < c; j += 32) m = fmaxf(m, S[i * LS + j]);
        m = warp_max(m);
        float s = 0.f;
        for (int j = lane; j < c; j += 32) { const float e = expf(S[i * LS + j] - m); S[i * LS + j] = e; s += e; }
        const float inv = 1.0f / warp_sum(s);
        for (int j = lane; j < c; j += 32) {
            float pv = S[i * LS + j] * inv;
            if (probs) probs[pbase + (long long)i * c + j] = pv;
            if (keep) pv = keep[pbase + (long long)i * c + j] ? pv * keep_scale : 0.f;
            S[i * LS + j] = pv;
        }
    }
    __syncthreads();
    constexpr int TD = HD / 4;
    for (int tile = threadIdx.x; tile < tc * TD; tile += blockDim.x) {              // out = P V
        const int i0 = (tile / TD) * 4, d0 = (tile % TD) * 4;
        float acc[4][4];
        smem_mm_4x4(S, LS, 1, Vv, LD, 1, c, i0, d0, acc);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i0 + u >= c) break;
            *reinterpret_cast<float4*>(out + (size_t)(t0 + i0 + u) * E + h * HD + d0) = make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
        }
    }
}

template <int HD>
__global__ void __launch_bounds__(256)
attn_bwd_kernel(const float* __restrict__ d_out, const float* __restrict__ qkv, const float* __restrict__ probs,
                const int* __restrict__ v_off, const long long* __restrict__ p_off, int heads,
                float* __restrict__ d_qkv, const uint8_t* __restrict__ keep, float keep_scale) {
    extern __shared__ float sm[];
    const int b = blockIdx.x / heads, h = blockIdx.x % heads;
    const int t0 = v_off[b], c = v_off[b + 1] - t0;
    if (c <= 0) return;
    const int cp = (c + 3) & ~3;
    const int E = heads * HD, LD = HD + HD_PAD, LS = cp + 1;
    float* Q = sm; float* K = Q + cp * LD; float* Vv = K + cp * LD; float* dO = Vv + cp * LD;
    float* P = dO + cp * LD;        // softmax probabilities (pre-dropout)
    float* G = P + cp * LS;         // dropped probabilities first, then dS
    const float scale = rsqrtf((float)HD);
    const long long pbase = p_off[b] + (long long)h * c * c;
    for (int idx = threadIdx.x; idx < c * HD; idx += blockDim.x) {
        const int i = idx / HD, d = idx - i * HD;
        const float* row = qkv + (size_t)(t0 + i) * 3 * E + h * HD + d;
        Q[i * LD + d] = row[0] * scale; K[i * LD + d] = row[E]; Vv[i * LD + d] = row[2 * E];
        dO[i * LD + d] = d_out[(size_t)(t0 + i) * E + h * HD + d];
    }
    for (int idx = threadIdx.x; idx < c * c; idx += blockDim.x) {
        const int i = idx / c, j = idx - i * c;
        const float pv = probs[pbase + idx];
        P[i * LS + j] = pv;
        G[i * LS + j] = keep ? (keep[pbase + idx] ? pv * keep_scale : 0.f) : pv;
    }
    __syncthreads();
    const int tc = cp >> 2;
    constexpr int TD = HD / 4;
    // dV[j][d] = sum_i Pd[i][j] dO[i][d]
    for (int tile = threadIdx.x; tile < tc * TD; tile += blockDim.x) {
        const int j0 = (tile / TD) * 4, d0 = (tile % TD) * 4;
        float acc[4][4];
        smem_mm_4x4(G, 1, LS, dO, LD, 1, c, j0, d0, acc);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (j0 + u >= c) break;
            *reinterpret_cast<float4*>(d_qkv + (size_t)(t0 + j0 + u) * 3 * E + 2 * E + h * HD + d0) = make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
        }
    }
    __syncthreads();
    // dP[i][j] = keep * scale * sum_d dO[i][d] V[j][d]   (overwrites G)
    for (int tile = threadIdx.x; tile < tc * tc; tile += blockDim.x) {
        const int i0 = (tile / tc) * 4, j0 = (tile % tc) * 4;
        float acc[4][4];
        smem_mm_4x4(dO, LD, 1, Vv, 1, LD, HD, i0, j0, acc);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const int i = i0 + u, j = j0 + v;
                float a = acc[u][v];
                if (keep && i < c && j < c) a = keep[pbase + (long long)i * c + j] ? a * keep_scale : 0.f;
                G[i * LS + j] = a;
            }
    }
    __syncthreads();
    // dS = P * (dP - rowsum(dP * P))
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = warp; i < c; i += (blockDim.x >> 5)) {
        float s = 0.f;
        for (int j = lane; j < c; j += 32) s = fmaf(G[i * LS + j], P[i * LS + j], s);
        s = warp_sum(s);
        for (int j = lane; j < c; j += 32) G[i * LS + j] = P[i * LS + j] * (G[i * LS + j] - s);
    }
    __syncthreads();
    // dQ[i][d] = scale * sum_j dS[i][j] K[j][d];  dK[j][d] = sum_i dS[i][j] Qs[i][d]
    for (int tile = threadIdx.x; tile < tc * TD; tile += blockDim.x) {
        const int i0 = (tile / TD) * 4, d0 = (tile % TD) * 4;
        float acc[4][4];
        smem_mm_4x4(G, LS, 1, K, LD, 1, c, i0, d0, acc);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i0 + u >= c) break;
            *reinterpret_cast<float4*>(d_qkv + (size_t)(t0 + i0 + u) * 3 * E + h * HD + d0) =
                make_float4(acc[u][0] * scale, acc[u][1] * scale, acc[u][2] * scale, acc[u][3] * scale);
        }
        smem_mm_4x4(G, 1, LS, Q, LD, 1, c, i0, d0, acc);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i0 + u >= c) break;
            *reinterpret_cast<float4*>(d_qkv + (size_t)(t0 + i0 + u) * 3 * E + E + h * HD + d0) = make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// all-pairs first layer, one CTA per vertex (b, i): its row of pairs (i, j>i) is contiguous
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
edge_pair_fwd_kernel(const float* __restrict__ P, const float* __restrict__ Q, const float* __restrict__ verts,
                     const float* __restrict__ wd, const float* __restrict__ bias, const int* __restrict__ v_off,
                     const long long* __restrict__ e_off, int B, int C, int ld, float* __restrict__ z1, float* __restrict__ dist) {
    const int t = blockIdx.x;
    const int b = find_segment(v_off, B, t);
    const int t0 = v_off[b], c = v_off[b + 1] - t0, i = t - t0;
    const int nj = c - 1 - i;
    if (nj <= 0) return;
    const long long e0 = e_off[b] + pair_id(i, i + 1, c);
    const float vx = verts[(size_t)t * 3], vy = verts[(size_t)t * 3 + 1], vz = verts[(size_t)t * 3 + 2];
    const int C4 = C >> 2;
    if ((C & 3) == 0 && (ld & 3) == 0 && C4 <= (int)blockDim.x && (blockDim.x % C4) == 0 &&
        ((reinterpret_cast<uintptr_t>(P) | reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(wd) |
          reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(z1)) & 15) == 0) {
        // a thread keeps four channels of this vertex's P row, the distance weights and the bias in registers and walks the
        // partner vertices: one float4 of Q in, one float4 of z1 out per pair (the distance is formed once per 4 channels)
        const int c4 = threadIdx.x % C4, j0 = threadIdx.x / C4, jstep = blockDim.x / C4;
        const float4 p4 = reinterpret_cast<const float4*>(P + (size_t)t * ld)[c4];
        const float4 w4 = reinterpret_cast<const float4*>(wd)[c4];
        const float4 b4 = reinterpret_cast<const float4*>(bias)[c4];
        for (int jj = j0; jj < nj; jj += jstep) {
            const int tj = t + 1 + jj;
            const float dx = vx - verts[(size_t)tj * 3], dy = vy - verts[(size_t)tj * 3 + 1], dz = vz - verts[(size_t)tj * 3 + 2];
            const float dd = sqrtf(dx * dx + dy * dy + dz * dz);
            if (c4 == 0) dist[e0 + jj] = dd;
            const float4 q4 = reinterpret_cast<const float4*>(Q + (size_t)tj * ld)[c4];
            // same association as the scalar form: ((P + Q) + wd * d) + bias
            float4 o;
            o.x = p4.x + q4.x + w4.x * dd + b4.x; o.y = p4.y + q4.y + w4.y * dd + b4.y;
            o.z = p4.z + q4.z + w4.z * dd + b4.z; o.w = p4.w + q4.w + w4.w * dd + b4.w;
            reinterpret_cast<float4*>(z1 + (size_t)(e0 + jj) * C)[c4] = o;
        }
        return;
    }
    for (long long idx = threadIdx.x; idx < (long long)nj * C; idx += blockDim.x) {
        const int jj = (int)(idx / C), ch = (int)(idx - (long long)jj * C);
        const int tj = t + 1 + jj;
        const float dx = vx - verts[(size_t)tj * 3], dy = vy - verts[(size_t)tj * 3 + 1], dz = vz - verts[(size_t)tj * 3 + 2];
        const float dd = sqrtf(dx * dx + dy * dy + dz * dz);
        if (ch == 0) dist[e0 + jj] = dd;
        z1[(size_t)(e0 + jj) * C + ch] = P[(size_t)t * ld + ch] + Q[(size_t)tj * ld + ch] + wd[ch] * dd + bias[ch];
    }
}

template <int CPL>     // channels per lane = C / 32
__global__ void __launch_bounds__(256)
edge_pair_bwd_kernel(const float* __restrict__ dz1, const float* __restrict__ dist, const float* __restrict__ verts,
                     const float* __restrict__ wd, const int* __restrict__ v_off, const long long* __restrict__ e_off, int B,
                     int ld, float* __restrict__ dP, float* __restrict__ dQ, float* __restrict__ d_verts, float* __restrict__ d_wd) {
    constexpr int C = CPL * 32;
    __shared__ float red[8][C];
    const int t = blockIdx.x;
    const int b = find_segment(v_off, B, t);
    const int t0 = v_off[b], c = v_off[b + 1] - t0, i = t - t0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long eb = e_off[b];
    float accP[CPL], accW[CPL], accQ[CPL], wl[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) { accP[k] = accW[k] = accQ[k] = 0.f; wl[k] = wd[lane + 32 * k]; }
    const float vx = verts[(size_t)t * 3], vy = verts[(size_t)t * 3 + 1], vz = verts[(size_t)t * 3 + 2];
    float dvx = 0.f, dvy = 0.f, dvz = 0.f;              // lane 0 accumulates this vertex's coordinate gradient
    // row pairs (i, j), j > i
    for (int j = i + 1 + warp; j < c; j += 8) {
        const long long e = eb + pair_id(i, j, c);
        const float dd = dist[e];
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const float g = dz1[(size_t)e * C + lane + 32 * k];
            accP[k] += g; accW[k] = fmaf(g, dd, accW[k]); dot = fmaf(g, wl[k], dot);
        }
        dot = warp_sum(dot);
        if (lane == 0 && dd > 0.f) {
            const int tj = t0 + j;
            const float s = dot / dd;
            const float gx = s * (vx - verts[(size_t)tj * 3]), gy = s * (vy - verts[(size_t)tj * 3 + 1]), gz = s * (vz - verts[(size_t)tj * 3 + 2]);
            dvx += gx; dvy += gy; dvz += gz;
            atomicAdd(d_verts + (size_t)tj * 3, -gx); atomicAdd(d_verts + (size_t)tj * 3 + 1, -gy); atomicAdd(d_verts + (size_t)tj * 3 + 2, -gz);
        }
    }
    // column pairs (i', i), i' < i
    for (int ip = warp; ip < i; ip += 8) {
        const long long e = eb + pair_id(ip, i, c);
#pragma unroll
        for (int k = 0; k < CPL; ++k) accQ[k] += dz1[(size_t)e * C + lane + 32 * k];
    }
    if (lane == 0) { atomicAdd(d_verts + (size_t)t * 3, dvx); atomicAdd(d_verts + (size_t)t * 3 + 1, dvy); atomicAdd(d_verts + (size_t)t * 3 + 2, dvz); }
    // cross-warp reductions (three rounds through one buffer)
    auto reduce_store = [&](float (&acc)[CPL], float* dst, bool atomic) {
#pragma unroll
        for (int k = 0; k < CPL; ++k) red[warp][lane + 32 * k] = acc[k];
        __syncthreads();
        for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += red[w][ch];
            if (atomic) atomicAdd(dst + ch, s); else dst[ch] = s;
        }
        __syncthreads();
    };
    reduce_store(accP, dP + (size_t)t * ld, false);
    reduce_store(accQ, dQ + (size_t)t * ld, false);
    reduce_store(accW, d_wd, true);
}

// final layer: warp per (sample, padded edge slot)
template <int KPL>
__global__ void edge_out_fwd_kernel(const float* __restrict__ h, const float* __restrict__ w, const float* __restrict__ bias,
                                    const long long* __restrict__ e_off, int B, int max_e, float* __restrict__ probs) {
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= (long long)B * max_e) return;
    const int b = (int)(gw / max_e), k = (int)(gw - (long long)b * max_e);
    const long long ne = e_off[b + 1] - e_off[b];
    float p = 0.f;
    if (k < ne) {
        const float* row = h + (size_t)(e_off[b] + k) * (KPL * 32);
        float a = 0.f;
#pragma unroll
        for (int q = 0; q < KPL; ++q) a = fmaf(row[lane + 32 * q], w[lane + 32 * q], a);
        a = warp_sum(a) + bias[0];
        p = 1.0f / (1.0f + expf(-a));
    }
    if (lane == 0) probs[gw] = p;
}

template <int KPL>
__global__ void edge_out_bwd_kernel(const float* __restrict__ d_probs, const float* __restrict__ probs,
                                    const float* __restrict__ h, const float* __restrict__ w,
                                    const long long* __restrict__ e_off, int B, int max_e, float* __restrict__ dh,
                                    float* __restrict__ dw, float* __restrict__ db) {
    const int lane = threadIdx.x & 31;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    float aw[KPL], ab = 0.f, wl[KPL];
#pragma unroll
    for (int q = 0; q < KPL; ++q) { aw[q] = 0.f; wl[q] = w[lane + 32 * q]; }
    for (long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; gw < (long long)B * max_e; gw += nw) {
        const int b = (int)(gw / max_e), k = (int)(gw - (long long)b * max_e);
        if (k >= e_off[b + 1] - e_off[b]) continue;
        const float p = probs[gw];
        const float dl = d_probs[gw] * p * (1.0f - p);
        const size_t e = (size_t)(e_off[b] + k);
#pragma unroll
        for (int q = 0; q < KPL; ++q) {
            dh[e * (KPL * 32) + lane + 32 * q] = dl * wl[q];
            aw[q] = fmaf(dl, h[e * (KPL * 32) + lane + 32 * q], aw[q]);
        }
        ab += dl;
    }
#pragma unroll
    for (int q = 0; q < KPL; ++q) atomicAdd(dw + lane + 32 * q, aw[q]);
    if (lane == 0) atomicAdd(db, ab);
}

}  // namespace edge
}  // namespace wf

extern "C" int wf_gather_prefix(const float* verts, int B, int V, const int32_t* v_off, int T, float* packed, wf_stream_t stream) {
    using namespace wf;
    if (B <= 0 || T <= 0) return WF_OK;
    edge::gather_prefix_kernel<<<cdiv((long long)T * 3, 256), 256, 0, as_stream(stream)>>>(verts, B, V, v_off, T, packed);
    WF_LAUNCH_CHECK();
    return WF_OK;
}
extern "C" int wf_scatter_prefix_add(const float* d_packed, int B, int V, const int32_t* v_off, int T, float* d_verts, wf_stream_t stream) {
    using namespace wf;
    if (B <= 0 || T <= 0) return WF_OK;
    edge::scatter_prefix_add_kernel<<<cdiv((long long)T * 3, 256), 256, 0, as_stream(stream)>>>(d_packed, B, V, v_off, T, d_verts);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

static int attn_smem(int max_c, int hd, bool bwd, size_t* out) {
    const size_t cp = ((size_t)max_c + 3) & ~(size_t)3;          // rows padded for the kernels' 4 x 4 register tiles
    const size_t ld = hd + wf::edge::HD_PAD, ls = cp + 1;
    size_t fl = (bwd ? 4 : 3) * cp * ld + (bwd ? 2 : 1) * cp * ls;
    *out = fl * sizeof(float);
    if (*out > 227 * 1024) { wf::set_error("attention: %d vertices per sample exceed shared memory (%zu B)", max_c, *out); return WF_ETOOBIG; }
    return WF_OK;
}

extern "C" int wf_attn_fwd(const float* qkv, const int32_t* v_off, const int64_t* p_off, int B, int heads, int head_dim,
                           int max_c, float* out, float* probs, const uint8_t* keep, float keep_scale, wf_stream_t stream) {
    using namespace wf;
    if (B <= 0 || max_c <= 0) return WF_OK;
    WF_CHECK_ARG(head_dim == 16 || head_dim == 32 || head_dim == 64 || head_dim == 128, "wf_attn_fwd: head_dim %d not built (16, 32, 64, 128)", head_dim);
    WF_CHECK_ARG(!((probs || keep) && !p_off), "wf_attn_fwd: probs/keep need p_off");
    size_t smem; int rc = attn_smem(max_c, head_dim, false, &smem); if (rc) return rc;
#define WF_ATTN_FWD(HD_)                                                                                                      \
    case HD_:                                                                                                                 \
        WF_CUDA(cudaFuncSetAttribute(edge::attn_fwd_kernel<HD_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));   \
        edge::attn_fwd_kernel<HD_><<<B * heads, 256, smem, as_stream(stream)>>>(                                              \
            qkv, v_off, reinterpret_cast<const long long*>(p_off), heads, out, probs, keep, keep_scale);                      \
        break;
    switch (head_dim) { WF_ATTN_FWD(16) WF_ATTN_FWD(32) WF_ATTN_FWD(64) WF_ATTN_FWD(128) }
#undef WF_ATTN_FWD
    WF_LAUNCH_CHECK();
    return WF_OK;
}
extern "C" int wf_attn_bwd(const float* d_out, const float* qkv, const float* probs, const int32_t* v_off, const int64_t* p_off,
                           int B, int heads, int head_dim, int max_c, float* d_qkv, const uint8_t* keep, float keep_scale,
                           wf_stream_t stream) {
    using namespace wf;
    if (B <= 0 || max_c <= 0) return WF_OK;
    WF_CHECK_ARG(head_dim == 16 || head_dim == 32 || head_dim == 64 || head_dim == 128, "wf_attn_bwd: head_dim %d not built (16, 32, 64, 128)", head_dim);
    WF_CHECK_ARG(probs && p_off, "wf_attn_bwd: needs saved probabilities");
    size_t smem; int rc = attn_smem(max_c, head_dim, true, &smem); if (rc) return rc;
#define WF_ATTN_BWD(HD_)                                                                                                      \
    case HD_:                                                                                                                 \
        WF_CUDA(cudaFuncSetAttribute(edge::attn_bwd_kernel<HD_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));   \
        edge::attn_bwd_kernel<HD_><<<B * heads, 256, smem, as_stream(stream)>>>(                                              \
            d_out, qkv, probs, v_off, reinterpret_cast<const long long*>(p_off), heads, d_qkv, keep, keep_scale);             \
        break;
    switch (head_dim) { WF_ATTN_BWD(16) WF_ATTN_BWD(32) WF_ATTN_BWD(64) WF_ATTN_BWD(128) }
#undef WF_ATTN_BWD
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_edge_pair_fwd(const float* P, const float* Q, const float* verts, const float* wd, const float* bias,
                                const int32_t* v_off, const int64_t* e_off, int B, int T, int C, int ld, float* z1, float* dist,
                                wf_stream_t stream) {
    using namespace wf;
    if (B <= 0 || T <= 0) return WF_OK;
    WF_CHECK_ARG(ld >= C, "wf_edge_pair_fwd: row stride %d < C = %d", ld, C);
    edge::edge_pair_fwd_kernel<<<T, 256, 0, as_stream(stream)>>>(P, Q, verts, wd, bias, v_off, reinterpret_cast<const long long*>(e_off), B, C, ld, z1, dist);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_edge_pair_bwd(const float* dz1, const float* dist, const float* verts, const float* wd, const int32_t* v_off,
                                const int64_t* e_off, int B, int T, int C, int ld, float* dP, float* dQ, float* d_verts, float* d_wd,
                                wf_stream_t stream) {
    using namespace wf;
    if (B <= 0 || T <= 0) return WF_OK;
    WF_CHECK_ARG(ld >= C, "wf_edge_pair_bwd: row stride %d < C = %d", ld, C);
    WF_CHECK_ARG(C == 128 || C == 256 || C == 512 || C == 1024, "wf_edge_pair_bwd: C=%d not built (128, 256, 512, 1024)", C);
#define WF_PAIR_BWD(CPL_)                                                                                                     \
    case CPL_ * 32:                                                                                                           \
        edge::edge_pair_bwd_kernel<CPL_><<<T, 256, 0, as_stream(stream)>>>(                                                   \
            dz1, dist, verts, wd, v_off, reinterpret_cast<const long long*>(e_off), B, ld, dP, dQ, d_verts, d_wd);            \
        break;
    switch (C) { WF_PAIR_BWD(4) WF_PAIR_BWD(8) WF_PAIR_BWD(16) WF_PAIR_BWD(32) }
#undef WF_PAIR_BWD
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_edge_out_fwd(const float* h, const float* w, const float* bias, const int64_t* e_off, int B, int K, int max_e,
                               float* probs, wf_stream_t stream) {
    using namespace wf;
    if (B <= 0 || max_e <= 0) return WF_OK;
    WF_CHECK_ARG(K == 32 || K == 64 || K == 128 || K == 256, "wf_edge_out_fwd: K=%d not built (32, 64, 128, 256)", K);
    const long long warps = (long long)B * max_e;
#define WF_OUT_FWD(KPL_)                                                                                                      \
    case KPL_ * 32:                                                                                                           \
        edge::edge_out_fwd_kernel<KPL_><<<cdiv(warps * 32, 256), 256, 0, as_stream(stream)>>>(                                \
            h, w, bias, reinterpret_cast<const long long*>(e_off), B, max_e, probs);                                          \
        break;
    switch (K) { WF_OUT_FWD(1) WF_OUT_FWD(2) WF_OUT_FWD(4) WF_OUT_FWD(8) }
#undef WF_OUT_FWD
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_edge_out_bwd(const float* d_probs, const float* probs, const float* h, const float* w, const int64_t* e_off,
                               int B, int K, int max_e, float* dh, float* dw, float* db, wf_stream_t stream) {
    using namespace wf;
    if (B <= 0 || max_e <= 0) return WF_OK;
    WF_CHECK_ARG(K == 32 || K == 64 || K == 128 || K == 256, "wf_edge_out_bwd: K=%d not built (32, 64, 128, 256)", K);
    const long long warps = (long long)B * max_e;
    const int grid = (int)(cdiv(warps * 32, 256) < sm_count() * 8 ? cdiv(warps * 32, 256) : sm_count() * 8);
#define WF_OUT_BWD(KPL_)                                                                                                      \
    case KPL_ * 32:                                                                                                           \
        edge::edge_out_bwd_kernel<KPL_><<<grid, 256, 0, as_stream(stream)>>>(                                                 \
            d_probs, probs, h, w, reinterpret_cast<const long long*>(e_off), B, max_e, dh, dw, db);                           \
        break;
    switch (K) { WF_OUT_BWD(1) WF_OUT_BWD(2) WF_OUT_BWD(4) WF_OUT_BWD(8) }
#undef WF_OUT_BWD
    WF_LAUNCH_CHECK();
    return WF_OK;
}
