// bf16 LayerNorm+ReLU passes between the tensor-core layers of the encoder
// (models/PointNetEncoder.py:38-39 and their backward), HBM-bound, 16-byte vector accesses.
//
//   forward : h = relu((z - mean) * rstd * gamma + beta)                  2 B read + 2 B write per element
//   backward: g = dh * [y > 0];  gh = g * gamma;  c1 = mean(gh), c2 = mean(gh * xhat)
//             dz = rstd * (gh - c1 - xhat * c2);  dgamma += g * xhat;  dbeta += g;  dbias += dz
//             single pass: 4 B read + 2 B write per element.
//
// Both kernels keep a thread on 8 consecutive channels (gamma/beta and the column partial sums stay in registers) and
// do their arithmetic on channel PAIRS with the packed fp32 instructions (fma/mul/add.f32x2 -> FFMA2/FMUL2/FADD2): the
// 3-register scalar forms issue at half rate on sm_100, and the backward pass at ~24 scalar fp32 instructions per
// element was bound by the fma pipe, not by HBM.
//
// The backward streams its operands through a shared-memory ring filled with cp.async (16 bytes per thread per row):
// every thread copies exactly the bytes it will consume, so the only wait is its own cp.async group -- no register
// double buffering and three row groups of loads in flight per CTA.
#include "wf_common.cuh"
#include "ln_side.cuh"

#include <mutex>

namespace wf {
namespace lnb {

// forward, stand-alone: a CTA of 256 threads per block of CS_R rows (the body lives in ln_side.cuh, shared with the side-job
// form inside the GEMM kernel)
template <int C8, bool COLSUM>
__global__ void __launch_bounds__(256)
ln_relu_fwd_kernel(const FwdArgs a) {
    __shared__ float red[COLSUM ? (256 / C8) * 2 * C8 * 8 : 1];
    ln_fwd_block<C8, COLSUM, 256>(a, blockIdx.x, threadIdx.x, red, 0);
}

// hbar[kind][b][c] = (sum over the row blocks of cloud b) * (kind 0: 1/N, kind 1: 1/valid[b]).
// A cloud of a million points has 7 813 row blocks: one thread walking them in order is a chain of 7 813 dependent L2 loads
// (4.7 ms for ONE cloud -- 28 % of the whole 1M-point encoder pass).  CTA = 32 channels x SM_RG block ranges; a thread sums
// the blocks blk = first + r, first + r + SM_RG, ... of its range r on four interleaved accumulators, the ranges are added
// through shared memory in range order: a fixed summation order (deterministic), ~60 dependent loads at a million points.
constexpr int SM_RG = 32;

__global__ void __launch_bounds__(32 * SM_RG)
seg_mean_kernel(const float* __restrict__ part, const float* __restrict__ valid, int B, int pool_n, int C,
                float* __restrict__ hbar) {
    __shared__ float red[SM_RG][33];
    const int cl = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl, b = blockIdx.y, kind = blockIdx.z;
    const long long lo = (long long)b * pool_n, hi = lo + pool_n - 1;
    const long long b0 = lo / CS_R, b1 = hi / CS_R;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < C) {
        int u = 0;
        for (long long blk = b0 + r; blk <= b1; blk += SM_RG, u = (u + 1) & 3) {
            const int seg = b - (int)((blk * CS_R) / pool_n);
            acc[u] += part[((blk * 2 + seg) * 2 + kind) * C + c];
        }
    }
    red[r][cl] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    __syncthreads();
    if (r == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < SM_RG; ++q) t += red[q][cl];
        hbar[((size_t)kind * B + b) * C + c] = t * (kind == 0 ? 1.0f / (float)pool_n : 1.0f / valid[b]);
    }
}

// ------------------------------------------------------------------------------------------
// backward.  CTA = C/8 threads (NW warps), rows in groups of RG, STAGES groups in flight through cp.async.
// ------------------------------------------------------------------------------------------
constexpr int RG = 4, STAGES = 3;

template <int NW>
__global__ void __launch_bounds__(NW * 32, NW <= 4 ? 4 : 2)
ln_relu_bwd_kernel(const uint4* __restrict__ dh, const uint4* __restrict__ z, const float* __restrict__ mean,
                   const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                   uint4* __restrict__ dz, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dcolsum,
                   long long M) {
    constexpr int C8 = NW * 32, C = C8 * 8, NV = 2 * RG;    // NV row sums per group: (sum gh, sum gh*xhat) x RG rows
    extern __shared__ uint4 ring[];                       // [STAGES][2 (dh, z)][RG][C8]
    // Row sums: every thread deposits its NV partials ([value][thread], conflict-free), warp w then reduces values
    // w, w+NW, ... (NW consecutive partials per lane, one warp reduction each) -- 1/6 of the instructions of reducing all NV
    // values in every warp and combining the warps' results in every thread.
    __shared__ __align__(16) float partial[NV][C8];
    __shared__ __align__(16) float rowsum[NV];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t ring_s = static_cast<uint32_t>(__cvta_generic_to_shared(ring)) + tid * 16;
    u64 gm[4], bt[4];
    load_pairs(gamma + tid * 8, gm); load_pairs(beta + tid * 8, bt);
    u64 acc_g[4], acc_gx[4], acc_dz[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc_g[i] = acc_gx[i] = acc_dz[i] = 0ull;
    const long long groups = (M + RG - 1) / RG;

    auto issue = [&](long long grp, int stage) {
        if (grp < groups) {
            const long long r0 = grp * RG;
#pragma unroll
            for (int r = 0; r < RG; ++r) {
                const bool ok = r0 + r < M;
                const long long row = ok ? r0 + r : 0;                      // size 0 -> zero fill, address stays valid
                const uint32_t d = ring_s + ((stage * 2 + 0) * RG + r) * C8 * 16;
                cp_async16(d, dh + row * C8 + tid, ok ? 16u : 0u);
                cp_async16(d + RG * C8 * 16, z + row * C8 + tid, ok ? 16u : 0u);
            }
        }
        cp_async_commit();
    };

    // row statistics of a group (RG = 4 rows -> one aligned float4 each); fetched one group AHEAD into registers: read on
    // demand they cost every group a full global-memory latency that nothing else in the loop covers
    auto load_stats = [&](long long g, float4& m4, float4& r4) {
        m4 = make_float4(0.f, 0.f, 0.f, 0.f); r4 = m4;
        if (g < groups) {
            const long long r0 = g * RG;
            if (r0 + RG <= M) {
                m4 = __ldg(reinterpret_cast<const float4*>(mean + r0)); r4 = __ldg(reinterpret_cast<const float4*>(rstd + r0));
            } else {
                float* pm = reinterpret_cast<float*>(&m4); float* pr = reinterpret_cast<float*>(&r4);
                for (int r = 0; r < RG; ++r) if (r0 + r < M) { pm[r] = mean[r0 + r]; pr[r] = rstd[r0 + r]; }
            }
        }
    };
    static_assert(RG == 4, "statistics are fetched as one float4 per group");

    long long grp = blockIdx.x;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issue(grp + (long long)s * gridDim.x, s);
    float4 mu_next, rs_next;
    load_stats(grp, mu_next, rs_next);
    int stage = 0;
    for (; grp < groups; grp += gridDim.x) {
        issue(grp + (long long)(STAGES - 1) * gridDim.x, stage == 0 ? STAGES - 1 : stage - 1);
        const float mu[RG] = {mu_next.x, mu_next.y, mu_next.z, mu_next.w};
        const float rs[RG] = {rs_next.x, rs_next.y, rs_next.z, rs_next.w};
        load_stats(grp + gridDim.x, mu_next, rs_next);
        cp_async_wait<STAGES - 1>();                                         // this thread's copies of `grp` have landed
        const long long r0 = grp * RG;
        uint4* sd = ring + ((stage * 2 + 0) * RG) * C8 + tid;                // this thread's own 16 bytes of each row
        const uint4* sz = ring + ((stage * 2 + 1) * RG) * C8 + tid;
#pragma unroll
        for (int r = 0; r < RG; ++r) {
            const uint4 ud = sd[r * C8], uz = sz[r * C8];
            const uint32_t wd[4] = {ud.x, ud.y, ud.z, ud.w}, wz[4] = {uz.x, uz.y, uz.z, uz.w};
            const u64 rs2 = pk2(rs[r], rs[r]), nm2 = pk2(-mu[r] * rs[r], -mu[r] * rs[r]);
            u64 a = 0ull, b = 0ull;
            uint32_t gw[4];                                                  // g = dh * [y > 0], still exact in bf16
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const u64 xh = fma2(bf2(wz[i]), rs2, nm2);
                float y0, y1;
                up2(fma2(xh, gm[i], bt[i]), y0, y1);
                gw[i] = wd[i] & ((y0 > 0.f ? 0x0000FFFFu : 0u) | (y1 > 0.f ? 0xFFFF0000u : 0u));
                const u64 gh = mul2(bf2(gw[i]), gm[i]);
                a = add2(a, gh); b = fma2(gh, xh, b);
            }
            sd[r * C8] = make_uint4(gw[0], gw[1], gw[2], gw[3]);             // the second pass reads g, not dh
            float lo, hi;
            up2(a, lo, hi); partial[2 * r][tid] = lo + hi;
            up2(b, lo, hi); partial[2 * r + 1][tid] = lo + hi;
        }
        __syncthreads();
        for (int k = warp; k < NV; k += NW) {
            float t = 0.f;
            if (NW % 4 == 0) {
#pragma unroll
                for (int j = 0; j < NW / 4; ++j) {
                    const float4 v = reinterpret_cast<const float4*>(&partial[k][lane * NW])[j];
                    t += (v.x + v.y) + (v.z + v.w);
                }
            } else {
                const float2 v = *reinterpret_cast<const float2*>(&partial[k][lane * NW]);
                t = v.x + v.y;
            }
            t = warp_sum(t);
            if (lane == 0) rowsum[k] = t;
        }
        __syncthreads();
        float cs[NV];
#pragma unroll
        for (int k = 0; k < NV / 4; ++k) {
            const float4 v = reinterpret_cast<const float4*>(rowsum)[k];
            cs[4 * k] = v.x; cs[4 * k + 1] = v.y; cs[4 * k + 2] = v.z; cs[4 * k + 3] = v.w;
        }
#pragma unroll
        for (int r = 0; r < RG; ++r) {
            if (r0 + r >= M) break;
            const float c1 = cs[2 * r] * (1.0f / C), c2 = cs[2 * r + 1] * (1.0f / C);
            const uint4 ug = sd[r * C8], uz = sz[r * C8];
            const uint32_t wg[4] = {ug.x, ug.y, ug.z, ug.w}, wz[4] = {uz.x, uz.y, uz.z, uz.w};
            const u64 rs2 = pk2(rs[r], rs[r]), nm2 = pk2(-mu[r] * rs[r], -mu[r] * rs[r]);
            const u64 k1 = pk2(-c1 * rs[r], -c1 * rs[r]), k2 = pk2(-c2 * rs[r], -c2 * rs[r]);
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const u64 xh = fma2(bf2(wz[i]), rs2, nm2);
                const u64 g = bf2(wg[i]);
                // dz = rstd * (g*gamma - c1 - xhat*c2)
                const u64 dzv = fma2(xh, k2, fma2(mul2(g, gm[i]), rs2, k1));
                acc_g[i] = add2(acc_g[i], g); acc_gx[i] = fma2(g, xh, acc_gx[i]); acc_dz[i] = add2(acc_dz[i], dzv);
                o[i] = to_bf2(dzv);
            }
            dz[(r0 + r) * C8 + tid] = make_uint4(o[0], o[1], o[2], o[3]);
        }
        stage = stage + 1 == STAGES ? 0 : stage + 1;
    }
    cp_async_wait<0>();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float lo, hi;
        up2(acc_gx[i], lo, hi); atomicAdd(dgamma + tid * 8 + 2 * i, lo); atomicAdd(dgamma + tid * 8 + 2 * i + 1, hi);
        up2(acc_g[i], lo, hi); atomicAdd(dbeta + tid * 8 + 2 * i, lo); atomicAdd(dbeta + tid * 8 + 2 * i + 1, hi);
        up2(acc_dz[i], lo, hi); atomicAdd(dcolsum + tid * 8 + 2 * i, lo); atomicAdd(dcolsum + tid * 8 + 2 * i + 1, hi);
    }
}

template <int NW>
static int launch_bwd(const void* dh, const void* z, const float* mean, const float* rstd, const float* gamma, const float* beta,
                      void* dz, float* dgamma, float* dbeta, float* dcolsum, long long M, cudaStream_t s) {
    constexpr int SMEM = STAGES * 2 * RG * NW * 32 * 16;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] { attr_err = cudaFuncSetAttribute(ln_relu_bwd_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM); });
    WF_CUDA(attr_err);
    const long long groups = (M + RG - 1) / RG;
    const int per_sm = NW <= 4 ? 4 : 2;
    const int grid = (int)(groups < (long long)per_sm * sm_count() ? groups : (long long)per_sm * sm_count());
    ln_relu_bwd_kernel<NW><<<grid, NW * 32, SMEM, s>>>(static_cast<const uint4*>(dh), static_cast<const uint4*>(z), mean, rstd, gamma,
                                                      beta, static_cast<uint4*>(dz), dgamma, dbeta, dcolsum, M);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

}  // namespace lnb
}  // namespace wf

extern "C" int wf_ln_relu_bf16_fwd(const void* z, const float* mean, const float* rstd, const float* gamma, const float* beta,
                                   void* h, int M, int C, wf_stream_t stream) {
    using namespace wf;
    if (M <= 0 || C <= 0) return WF_OK;
    WF_CHECK_ARG(C == 512 || C == 1024 || C == 2048, "wf_ln_relu_bf16_fwd: C=%d not built (512/1024/2048)", C);
    WF_CHECK_ARG(((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(gamma) |
                   reinterpret_cast<uintptr_t>(beta)) & 15) == 0, "wf_ln_relu_bf16_fwd: 16-byte alignment required");
    const int grid = cdiv(M, lnb::CS_R);
    cudaStream_t s = as_stream(stream);
    const lnb::FwdArgs a{static_cast<const uint4*>(z), mean, rstd, gamma, beta, static_cast<uint4*>(h), nullptr, M, 1, 0, nullptr};
#define WF_LNF(C8) lnb::ln_relu_fwd_kernel<C8, false><<<grid, 256, 0, s>>>(a)
    if (C == 512) WF_LNF(64); else if (C == 1024) WF_LNF(128); else WF_LNF(256);
#undef WF_LNF
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
    const lnb::FwdArgs a{static_cast<const uint4*>(z), mean, rstd, gamma, beta, static_cast<uint4*>(h), mask, M, points_per_cloud, row_offset, part};
#define WF_LNC(C8) lnb::ln_relu_fwd_kernel<C8, true><<<grid, 256, 0, s>>>(a)
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
    dim3 grid(cdiv(C, 32), B, 2);
    lnb::seg_mean_kernel<<<grid, 32 * lnb::SM_RG, 0, as_stream(stream)>>>(part, valid, B, points_per_cloud, C, hbar);
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
    cudaStream_t s = as_stream(stream);
    if (C == 512) return lnb::launch_bwd<2>(dh, z, mean, rstd, gamma, beta, dz, dgamma, dbeta, dcolsum, M, s);
    if (C == 1024) return lnb::launch_bwd<4>(dh, z, mean, rstd, gamma, beta, dz, dgamma, dbeta, dcolsum, M, s);
    return lnb::launch_bwd<8>(dh, z, mean, rstd, gamma, beta, dz, dgamma, dbeta, dcolsum, M, s);
}
