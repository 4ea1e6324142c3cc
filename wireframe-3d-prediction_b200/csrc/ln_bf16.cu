// bf16 LayerNorm+ReLU passes between the tensor-core layers of the encoder
// (models/PointNetEncoder.py:38-39 and their backward), HBM-bound, 16-byte vector accesses.
//
//   forward : h = relu((z - mean) * rstd * gamma + beta)                  2 B read + 2 B write per element
//   backward: g = dh * [y > 0];  gh = g * gamma;  c1 = mean(gh), c2 = mean(gh * xhat)
//             dz = rstd * (gh - c1 - xhat * c2);  dgamma += g * xhat;  dbeta += g;  dbias += dz
//             single pass: 4 B read + 2 B write per element.  One CTA = C/8 threads, each owning 8 consecutive channels
//             (column partial sums stay in 24 registers); rows are processed 4 at a time with one block reduction
//             per group; a persistent grid keeps the final atomics at 3*C per CTA.
#include "wf_common.cuh"

namespace wf {
namespace lnb {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return u;
}
__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

__global__ void __launch_bounds__(256)
ln_relu_fwd_kernel(const uint4* __restrict__ z, const float* __restrict__ mean, const float* __restrict__ rstd,
                   const float* __restrict__ gamma, const float* __restrict__ beta, uint4* __restrict__ h, long long M, int C8) {
    const long long total = M * C8;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long row = idx / C8;
        const int c8 = (int)(idx - row * C8);
        float v[8], g[8], b[8];
        unpack8(z[idx], v);
        load8f(gamma + c8 * 8, g); load8f(beta + c8 * 8, b);
        const float mu = mean[row], rs = rstd[row];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaxf((v[i] - mu) * rs * g[i] + b[i], 0.f);
        h[idx] = pack8(v);
    }
}

// Same pass for the LAST LayerNorm of the per-point MLP, which also produces per-cloud column sums of its output h
// (all rows, and rows with mask != 0).  The final Linear is affine, so the two mean pools of its output
// (models/PointNetEncoder.py:103-105, models/VertexPredictor.py:86) are that Linear applied to the mean of h: the
// (B,N,512) point-feature tensor never has to exist for them.  Deterministic: a CTA owns CS_R consecutive rows, a thread
// owns 8 channels, partial sums go to part[row block][segment][kind][C] (segment 1 = rows of the next cloud when the
// block straddles a cloud boundary) and are added in block order by seg_mean_kernel.
constexpr int CS_R = 128;

template <int C8>
__global__ void __launch_bounds__(256)
ln_relu_fwd_colsum_kernel(const uint4* __restrict__ z, const float* __restrict__ mean, const float* __restrict__ rstd,
                          const float* __restrict__ gamma, const float* __restrict__ beta, uint4* __restrict__ h,
                          const uint8_t* __restrict__ mask, int M, int pool_n, int row_off, float* __restrict__ part) {
    constexpr int C = C8 * 8, RS = 256 / C8;
    __shared__ float red[RS][2][C];
    const int tid = threadIdx.x, c8 = tid % C8, rsub = tid / C8;
    const int blk_row0 = blockIdx.x * CS_R;
    const int rows = min(CS_R, M - blk_row0);
    const int g0 = row_off + blk_row0;                       // global row of this block's first row
    const int rb = (g0 / pool_n + 1) * pool_n - g0;          // local rows >= rb belong to the next cloud
    const size_t gblk = (size_t)(g0 / CS_R);
    float gm[8], bt[8];
    load8f(gamma + c8 * 8, gm); load8f(beta + c8 * 8, bt);
#pragma unroll 1
    for (int seg = 0; seg < 2; ++seg) {
        const int r_lo = seg == 0 ? 0 : rb, r_hi = seg == 0 ? min(rb, rows) : rows;
        if (r_lo >= r_hi) break;                             // block-uniform
        float su[8], sm[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) su[i] = sm[i] = 0.f;
        for (int r = r_lo + rsub; r < r_hi; r += RS) {
            const size_t row = (size_t)blk_row0 + r;
            float v[8];
            unpack8(z[row * C8 + c8], v);
            const float mu = mean[row], rs = rstd[row];
            const bool mk = mask == nullptr || mask[row] != 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaxf((v[i] - mu) * rs * gm[i] + bt[i], 0.f);
            const uint4 pk = pack8(v);
            h[row * C8 + c8] = pk;
            unpack8(pk, v);                                  // sum what the next GEMM will read (bf16-rounded)
#pragma unroll
            for (int i = 0; i < 8; ++i) { su[i] += v[i]; if (mk) sm[i] += v[i]; }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { red[rsub][0][c8 * 8 + i] = su[i]; red[rsub][1][c8 * 8 + i] = sm[i]; }
        __syncthreads();
        float* dst = part + ((gblk * 2 + seg) * 2) * C;
        for (int i = tid; i < 2 * C; i += 256) {
            const int kind = i / C, c = i - kind * C;
            float t = 0.f;
#pragma unroll
            for (int q = 0; q < RS; ++q) t += red[q][kind][c];
            dst[(size_t)kind * C + c] = t;
        }
        __syncthreads();
    }
}

// hbar[kind][b][c] = (sum over the row blocks of cloud b, in order) * (kind 0: 1/N, kind 1: 1/valid[b])
__global__ void seg_mean_kernel(const float* __restrict__ part, const float* __restrict__ valid, int B, int pool_n, int C,
                                float* __restrict__ hbar) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y, kind = blockIdx.z;
    if (c >= C) return;
    const long long lo = (long long)b * pool_n, hi = lo + pool_n - 1;
    float t = 0.f;
    for (long long blk = lo / CS_R; blk <= hi / CS_R; ++blk) {
        const int seg = b - (int)((blk * CS_R) / pool_n);
        t += part[((blk * 2 + seg) * 2 + kind) * C + c];
    }
    hbar[((size_t)kind * B + b) * C + c] = t * (kind == 0 ? 1.0f / (float)pool_n : 1.0f / valid[b]);
}

constexpr int RG = 4;      // rows per group

template <int NW>          // warps per CTA = C / 256
__global__ void __launch_bounds__(NW * 32, NW <= 4 ? 3 : 2)
ln_relu_bwd_kernel(const uint4* __restrict__ dh, const uint4* __restrict__ z, const float* __restrict__ mean,
                   const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                   uint4* __restrict__ dz, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dcolsum,
                   long long M) {
    constexpr int C8 = NW * 32, C = C8 * 8;
    __shared__ float red[2][NW][2 * RG];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float gm[8], bt[8];
    load8f(gamma + tid * 8, gm); load8f(beta + tid * 8, bt);
    float acc_g[8], acc_gx[8], acc_dz[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc_g[i] = acc_gx[i] = acc_dz[i] = 0.f;
    const long long groups = (M + RG - 1) / RG;
    int buf = 0;
    for (long long grp = blockIdx.x; grp < groups; grp += gridDim.x, buf ^= 1) {
        const long long r0 = grp * RG;
        uint4 ud[RG], uz[RG];
#pragma unroll
        for (int r = 0; r < RG; ++r) {
            if (r0 + r < M) { ud[r] = dh[(r0 + r) * C8 + tid]; uz[r] = z[(r0 + r) * C8 + tid]; }
            else { ud[r] = make_uint4(0, 0, 0, 0); uz[r] = make_uint4(0, 0, 0, 0); }
        }
        float part[2 * RG];
        float mu[RG], rs[RG];
#pragma unroll
        for (int r = 0; r < RG; ++r) {
            const bool ok = r0 + r < M;
            mu[r] = ok ? mean[r0 + r] : 0.f; rs[r] = ok ? rstd[r0 + r] : 0.f;
            float d[8], x[8];
            unpack8(ud[r], d); unpack8(uz[r], x);
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float xh = (x[i] - mu[r]) * rs[r];
                const float y = xh * gm[i] + bt[i];
                const float gh = (y > 0.f ? d[i] : 0.f) * gm[i];
                a += gh; b = fmaf(gh, xh, b);
            }
            part[2 * r] = a; part[2 * r + 1] = b;
        }
#pragma unroll
        for (int k = 0; k < 2 * RG; ++k) part[k] = warp_sum(part[k]);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 2 * RG; ++k) red[buf][warp][k] = part[k];
        }
        __syncthreads();                       // double-buffered: the next group's writes go to the other buffer
#pragma unroll
        for (int r = 0; r < RG; ++r) {
            if (r0 + r >= M) break;
            float c1 = 0.f, c2 = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) { c1 += red[buf][w][2 * r]; c2 += red[buf][w][2 * r + 1]; }
            c1 *= (1.0f / C); c2 *= (1.0f / C);
            float d[8], x[8], o[8];
            unpack8(ud[r], d); unpack8(uz[r], x);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float xh = (x[i] - mu[r]) * rs[r];
                const float y = xh * gm[i] + bt[i];
                const float g = y > 0.f ? d[i] : 0.f;
                const float dzv = rs[r] * (g * gm[i] - c1 - xh * c2);
                acc_g[i] += g; acc_gx[i] = fmaf(g, xh, acc_gx[i]); acc_dz[i] += dzv;
                o[i] = dzv;
            }
            dz[(r0 + r) * C8 + tid] = pack8(o);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        atomicAdd(dgamma + tid * 8 + i, acc_gx[i]);
        atomicAdd(dbeta + tid * 8 + i, acc_g[i]);
        atomicAdd(dcolsum + tid * 8 + i, acc_dz[i]);
    }
}

}  // namespace lnb
}  // namespace wf

extern "C" int wf_ln_relu_bf16_fwd(const void* z, const float* mean, const float* rstd, const float* gamma, const float* beta,
                                   void* h, int M, int C, wf_stream_t stream) {
    using namespace wf;
    if (M <= 0 || C <= 0) return WF_OK;
    WF_CHECK_ARG(C % 8 == 0, "wf_ln_relu_bf16_fwd: C %% 8 != 0");
    WF_CHECK_ARG(((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(gamma) |
                   reinterpret_cast<uintptr_t>(beta)) & 15) == 0, "wf_ln_relu_bf16_fwd: 16-byte alignment required");
    const long long total = (long long)M * (C / 8);
    const int grid = (int)(cdiv(total, 256) < 16LL * sm_count() ? cdiv(total, 256) : 16LL * sm_count());
    lnb::ln_relu_fwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(static_cast<const uint4*>(z), mean, rstd, gamma, beta,
                                                                static_cast<uint4*>(h), M, C / 8);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_ln_relu_bf16_fwd_colsum(const void* z, const float* mean, const float* rstd, const float* gamma,
                                          const float* beta, void* h, const uint8_t* mask, int M, int C, int points_per_cloud,
                                          int row_offset, float* part, wf_stream_t stream) {
    using namespace wf;
    if (M <= 0) return WF_OK;
    WF_CHECK_ARG(C == 512 || C == 1024 || C == 2048, "wf_ln_relu_bf16_fwd_colsum: C=%d not built (512/1024/2048)", C);
    WF_CHECK_ARG(points_per_cloud >= lnb::CS_R && row_offset >= 0 && row_offset % lnb::CS_R == 0,
                 "wf_ln_relu_bf16_fwd_colsum: points_per_cloud >= %d and row_offset %% %d == 0 required", lnb::CS_R, lnb::CS_R);
    WF_CHECK_ARG(((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(gamma) |
                   reinterpret_cast<uintptr_t>(beta)) & 15) == 0, "wf_ln_relu_bf16_fwd_colsum: 16-byte alignment required");
    const int grid = cdiv(M, lnb::CS_R);
    cudaStream_t s = as_stream(stream);
#define WF_LNC(C8) lnb::ln_relu_fwd_colsum_kernel<C8><<<grid, 256, 0, s>>>(static_cast<const uint4*>(z), mean, rstd, gamma, beta, static_cast<uint4*>(h), mask, M, points_per_cloud, row_offset, part)
    if (C == 512) WF_LNC(64); else if (C == 1024) WF_LNC(128); else WF_LNC(256);
#undef WF_LNC
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_seg_mean(const float* part, const float* valid, int B, int points_per_cloud, int C, float* hbar,
                           wf_stream_t stream) {
    using namespace wf;
    if (B <= 0 || C <= 0) return WF_OK;
    WF_CHECK_ARG(points_per_cloud >= lnb::CS_R, "wf_seg_mean: points_per_cloud >= %d required", lnb::CS_R);
    dim3 grid(cdiv(C, 256), B, 2);
    lnb::seg_mean_kernel<<<grid, 256, 0, as_stream(stream)>>>(part, valid, B, points_per_cloud, C, hbar);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_seg_part_floats(int total_rows, int C) {      // floats in `part` for total_rows rows (all chunks)
    return (int)(((long long)(total_rows + wf::lnb::CS_R - 1) / wf::lnb::CS_R) * 4 * C);
}

extern "C" int wf_ln_relu_bf16_bwd(const void* dh, const void* z, const float* mean, const float* rstd, const float* gamma,
                                   const float* beta, void* dz, float* dgamma, float* dbeta, float* dcolsum, int M, int C,
                                   wf_stream_t stream) {
    using namespace wf;
    if (M <= 0 || C <= 0) return WF_OK;
    WF_CHECK_ARG(C == 512 || C == 1024 || C == 2048, "wf_ln_relu_bf16_bwd: C=%d not built (512/1024/2048)", C);
    WF_CHECK_ARG(dgamma && dbeta && dcolsum, "wf_ln_relu_bf16_bwd: gradient accumulators required");
    WF_CHECK_ARG(((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(dh) | reinterpret_cast<uintptr_t>(dz) |
                   reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0,
                 "wf_ln_relu_bf16_bwd: 16-byte alignment required");
    const long long groups = ((long long)M + lnb::RG - 1) / lnb::RG;
    const int nw = C / 256;
    const int per_sm = nw <= 4 ? 3 : 2;
    const int grid = (int)(groups < (long long)per_sm * sm_count() ? groups : (long long)per_sm * sm_count());
    cudaStream_t s = as_stream(stream);
#define WF_LNB(NW) lnb::ln_relu_bwd_kernel<NW><<<grid, NW * 32, 0, s>>>(static_cast<const uint4*>(dh), static_cast<const uint4*>(z), mean, rstd, gamma, beta, static_cast<uint4*>(dz), dgamma, dbeta, dcolsum, M)
    if (nw == 2) WF_LNB(2); else if (nw == 4) WF_LNB(4); else WF_LNB(8);
#undef WF_LNB
    WF_LAUNCH_CHECK();
    return WF_OK;
}
