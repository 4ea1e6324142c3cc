// Pooled reductions of the per-point features WITHOUT the (B,N,512) tensor (production bf16 path).
//
// forward  (models/PointNetEncoder.py:94-111, models/VertexPredictor.py:86-87)
//   max pools : wf_gemm_bf16_pool (gemm_tc.cu) leaves packed (value, first row) maxima per (cloud, channel);
//   mean pools: the final Linear is affine, so mean_n(W h_n + b) = W mean_n(h_n) + b -- wf_ln_relu_bf16_fwd_colsum
//               produces the per-cloud means of h (all rows / valid rows), a (2B x 1024) x (1024 x 512) product
//               maps them through W;
//   pool_finalize_kernel decodes the packed maxima and adds the bias (a cloud without valid points gets 0 for
//   both masked pools, as the reference's clamp(count, 1) / non-finite -> 0 rules give).
//
// backward: the gradient of the pools w.r.t. the point features is (a per-cloud constant row) + (a per-cloud
//   constant row on valid points) + (<= 2*512 single entries at the argmax rows).  Pushed through the final
//   Linear analytically instead of as two dense GEMMs over all points:
//     d h[n,:]  = du[b,:]/N + mask[n] * dm[b,:]/valid[b] + sum_{c: argmax(b,c) = n} g(b,c) * W[c,:]
//     d W[c,:]  = (G^T hbar)[c,:] + sum_b g(b,c) * h[b, argmax(b,c), :]
//   with du, dm = [g_mean; g_avg] W (small GEMMs, done by the caller).  pool_bwd_prepare_kernel groups the
//   argmax entries of each cloud by row (deterministic order), pool_bwd_fill_kernel writes the constant part,
//   pool_bwd_rows_kernel rewrites the hit rows, pool_bwd_dw_kernel gathers the weight gradient.
#include "wf_common.cuh"

#include <math_constants.h>

namespace wf {
namespace pf {

__device__ __forceinline__ float unordered_bits(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}

__global__ void pool_finalize_kernel(const unsigned long long* __restrict__ pk_u, const unsigned long long* __restrict__ pk_m,
                                     const float* __restrict__ lin, const float* __restrict__ bias, int B, int C,
                                     float* __restrict__ max_m, int* __restrict__ arg_m, float* __restrict__ avg_m,
                                     float* __restrict__ max_u, int* __restrict__ arg_u, float* __restrict__ mean_u) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * C) return;
    const int c = i % C;
    const unsigned long long u = pk_u[i], m = pk_m[i];
    const float bc = bias ? bias[c] : 0.f;
    max_u[i] = u ? unordered_bits((uint32_t)(u >> 32)) : 0.f;
    arg_u[i] = u ? (int)(0xFFFFFFFFu - (uint32_t)(u & 0xFFFFFFFFu)) : -1;
    mean_u[i] = lin[i] + bc;
    const float vm = m ? unordered_bits((uint32_t)(m >> 32)) : 0.f;
    const bool fin = m != 0ull && isfinite(vm);                        // models/PointNetEncoder.py:111
    max_m[i] = fin ? vm : 0.f;
    arg_m[i] = fin ? (int)(0xFFFFFFFFu - (uint32_t)(m & 0xFFFFFFFFu)) : -1;
    avg_m[i] = lin[(size_t)B * C + i] + (m != 0ull ? bc : 0.f);
}

// One CTA per cloud, one thread per (channel, kind) entry: kind 0 = masked max, kind 1 = unmasked max.
// Output (per cloud, stride E = 2*C): n_uniq, u_row[j], u_start[j] (CSR over entries grouped by row, j < n_uniq,
// u_start[n_uniq] = #active entries), ent_c[pos], ent_g[pos].
__global__ void __launch_bounds__(1024)
pool_bwd_prepare_kernel(const float* __restrict__ g_max_m, const float* __restrict__ g_max_u, const int* __restrict__ arg_m,
                        const int* __restrict__ arg_u, int C, int* __restrict__ n_uniq, int* __restrict__ u_row,
                        int* __restrict__ u_start, int* __restrict__ ent_c, float* __restrict__ ent_g) {
    __shared__ int key[1024], aux[1024], wsum[32];
    const int b = blockIdx.x, e = threadIdx.x, E = 2 * C;
    const int lane = e & 31, warp = e >> 5;
    int k = -1; float g = 0.f; int c = 0;
    if (e < E) {
        c = e < C ? e : e - C;
        const size_t o = (size_t)b * C + c;
        if (e < C) { if (g_max_m) { k = arg_m[o]; g = g_max_m[o]; } }
        else       { if (g_max_u) { k = arg_u[o]; g = g_max_u[o]; } }
    }
    key[e] = k;
    __syncthreads();
    int first = e, rank = 0, cnt = 0;
    if (k >= 0) {
        first = -1;
        for (int j = 0; j < E; ++j) {
            if (key[j] == k) { if (first < 0) first = j; rank += (j < e); ++cnt; }
        }
    }
    const bool leader = k >= 0 && first == e;
    // exclusive scan of leader flags -> slot of each unique row
    auto block_scan = [&](int v, int& total) {
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        __syncthreads();
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        int base = 0; total = 0;
        for (int w = 0; w < 32; ++w) { const int t = wsum[w]; if (w < warp) base += t; total += t; }
        return base + inc - v;
    };
    int nu;
    const int slot = block_scan(leader ? 1 : 0, nu);
    aux[e] = leader ? slot : -1;                     // slot of the leader entry
    __syncthreads();
    const int myslot = k >= 0 ? aux[first] : -1;
    __syncthreads();
    // counts per slot, then their exclusive scan = CSR starts
    aux[e] = 0;
    __syncthreads();
    if (leader) aux[slot] = cnt;
    __syncthreads();
    int total_active;
    const int start = block_scan(aux[e], total_active);      // thread j handles slot j
    __syncthreads();
    key[e] = start;                                   // reuse: key[j] = start of slot j
    __syncthreads();
    const size_t cb = (size_t)b * E;
    if (e < nu) u_start[(size_t)b * (E + 1) + e] = start;
    if (e == 0) { u_start[(size_t)b * (E + 1) + nu] = total_active; n_uniq[b] = nu; }
    if (leader) u_row[cb + slot] = k;
    if (k >= 0) {
        const int pos = key[myslot] + rank;
        ent_c[cb + pos] = c; ent_g[cb + pos] = g;
    }
}

// constant part of d h: thread owns 8 consecutive channels, block = (K/8) x (256/(K/8)) threads
__global__ void __launch_bounds__(256)
pool_bwd_fill_kernel(const float* __restrict__ dbar, const uint8_t* __restrict__ mask, const float* __restrict__ valid, int B,
                     int N, int K, uint4* __restrict__ dh) {
    const int b = blockIdx.z;
    const int k0 = threadIdx.x * 8;
    const float invN = 1.0f / (float)N, invV = 1.0f / valid[b];
    float du[8], dm[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        du[i] = dbar[(size_t)b * K + k0 + i] * invN;
        dm[i] = dbar[((size_t)B + b) * K + k0 + i] * invV;
    }
    uint4 with_m, without_m;
    {
        __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&with_m);
        __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&without_m);
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            h1[i >> 1] = __floats2bfloat162_rn(du[i] + dm[i], du[i + 1] + dm[i + 1]);
            h0[i >> 1] = __floats2bfloat162_rn(du[i], du[i + 1]);
        }
    }
    const int K8 = K >> 3;
    for (int n = blockIdx.y * blockDim.y + threadIdx.y; n < N; n += gridDim.y * blockDim.y) {
        const bool mk = mask[(size_t)b * N + n] != 0;
        dh[((size_t)b * N + n) * K8 + threadIdx.x] = mk ? with_m : without_m;
    }
}

// hit rows: CTA (slot j, cloud b), thread owns 4 consecutive k
__global__ void __launch_bounds__(256)
pool_bwd_rows_kernel(const float* __restrict__ dbar, const uint8_t* __restrict__ mask, const float* __restrict__ valid,
                     const float* __restrict__ W, const int* __restrict__ n_uniq, const int* __restrict__ u_row,
                     const int* __restrict__ u_start, const int* __restrict__ ent_c, const float* __restrict__ ent_g, int B,
                     int N, int C, int K, __nv_bfloat16* __restrict__ dh) {
    const int b = blockIdx.y, j = blockIdx.x, E = 2 * C;
    if (j >= n_uniq[b]) return;
    const int row = u_row[(size_t)b * E + j];
    const int e0 = u_start[(size_t)b * (E + 1) + j], e1 = u_start[(size_t)b * (E + 1) + j + 1];
    const bool mk = mask[(size_t)b * N + row] != 0;
    const float invN = 1.0f / (float)N, invV = mk ? 1.0f / valid[b] : 0.f;
    for (int k = threadIdx.x * 4; k < K; k += blockDim.x * 4) {
        const float4 u = *reinterpret_cast<const float4*>(dbar + (size_t)b * K + k);
        const float4 m = *reinterpret_cast<const float4*>(dbar + ((size_t)B + b) * K + k);
        float a0 = u.x * invN + m.x * invV, a1 = u.y * invN + m.y * invV, a2 = u.z * invN + m.z * invV, a3 = u.w * invN + m.w * invV;
        for (int e = e0; e < e1; ++e) {
            const float g = ent_g[(size_t)b * E + e];
            const float4 w = __ldg(reinterpret_cast<const float4*>(W + (size_t)ent_c[(size_t)b * E + e] * K + k));
            a0 = fmaf(g, w.x, a0); a1 = fmaf(g, w.y, a1); a2 = fmaf(g, w.z, a2); a3 = fmaf(g, w.w, a3);
        }
        __nv_bfloat162 p0 = __floats2bfloat162_rn(a0, a1), p1 = __floats2bfloat162_rn(a2, a3);
        uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
        *reinterpret_cast<uint2*>(dh + ((size_t)b * N + row) * K + k) = pk;
    }
}

// dW[c,:] += sum_b gm(b,c) h[b, am(b,c), :] + gx(b,c) h[b, au(b,c), :];  db[c] = closed form over the four pools.
// The 2 B (argmax row, gradient) pairs of the channel are staged in shared memory first, so that the gathers of the h rows
// are independent loads issued eight at a time: walking b with the index, the gradient and the row fetched one after the
// other was a chain of 2 B dependent loads per thread (0.15 ms for a 2 MB gather).  Same summation order as before.
constexpr int DW_MAXB = 512;                       // staged pairs per pass: 2 kinds x 256 clouds

__global__ void __launch_bounds__(256)
pool_bwd_dw_kernel(const float* __restrict__ g_max_m, const float* __restrict__ g_avg_m, const float* __restrict__ g_max_u,
                   const float* __restrict__ g_mean_u, const int* __restrict__ arg_m, const int* __restrict__ arg_u,
                   const __nv_bfloat16* __restrict__ h, int B, int N, int C, int K, float* __restrict__ dW,
                   float* __restrict__ db) {
    __shared__ long long s_row[DW_MAXB];           // element offset of the gathered row (b * N + n) * K, or -1
    __shared__ float s_g[DW_MAXB];
    const int c = blockIdx.x;
    float4 acc[4];                                  // up to 4 column groups per thread (K <= 4096)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b0 = 0; b0 < B; b0 += DW_MAXB / 2) {
        const int nb = min(DW_MAXB / 2, B - b0);
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * nb; i += blockDim.x) {         // pair order: (b, kind 0), (b, kind 1), (b + 1, kind 0) ...
            const int b = b0 + (i >> 1), kind = i & 1;
            const size_t o = (size_t)b * C + c;
            const float* gp = kind == 0 ? g_max_m : g_max_u;
            const int n = gp == nullptr ? -1 : (kind == 0 ? arg_m[o] : arg_u[o]);
            s_row[i] = n < 0 ? -1ll : ((long long)b * N + n) * K;
            s_g[i] = n < 0 ? 0.f : gp[o];
        }
        __syncthreads();
        int gi = 0;
        for (int k = threadIdx.x * 4; k < K && gi < 4; k += blockDim.x * 4, ++gi) {
            float a0 = acc[gi].x, a1 = acc[gi].y, a2 = acc[gi].z, a3 = acc[gi].w;
            for (int i0 = 0; i0 < 2 * nb; i0 += 8) {
                uint2 raw[8]; float g[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = i0 + u;
                    const long long r = i < 2 * nb ? s_row[i] : -1ll;
                    g[u] = r < 0 ? 0.f : s_g[i];
                    raw[u] = r < 0 ? make_uint2(0u, 0u) : *reinterpret_cast<const uint2*>(h + r + k);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float2 x0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw[u].x));
                    const float2 x1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw[u].y));
                    if (g[u] != 0.f || raw[u].x != 0u || raw[u].y != 0u) {      // skipped pairs contribute nothing (and no -0 / NaN * 0)
                        a0 = fmaf(g[u], x0.x, a0); a1 = fmaf(g[u], x0.y, a1); a2 = fmaf(g[u], x1.x, a2); a3 = fmaf(g[u], x1.y, a3);
                    }
                }
            }
            acc[gi] = make_float4(a0, a1, a2, a3);
        }
    }
    {
        int gi = 0;
        for (int k = threadIdx.x * 4; k < K && gi < 4; k += blockDim.x * 4, ++gi) {
            float4* dst = reinterpret_cast<float4*>(dW + (size_t)c * K + k);
            float4 t = *dst;
            t.x += acc[gi].x; t.y += acc[gi].y; t.z += acc[gi].z; t.w += acc[gi].w;
            *dst = t;
        }
    }
    if (threadIdx.x == 0 && db != nullptr) {
        float t = 0.f;
        for (int b = 0; b < B; ++b) {
            const size_t o = (size_t)b * C + c;
            if (g_mean_u) t += g_mean_u[o];
            if (arg_m[o] >= 0) { if (g_avg_m) t += g_avg_m[o]; if (g_max_m) t += g_max_m[o]; }
            if (arg_u[o] >= 0 && g_max_u) t += g_max_u[o];
        }
        db[c] = t;
    }
}

}  // namespace pf
}  // namespace wf

extern "C" int wf_pool_finalize(const uint64_t* packed_u, const uint64_t* packed_m, const float* lin, const float* bias, int B,
                                int C, float* max_m, int32_t* arg_m, float* avg_m, float* max_u, int32_t* arg_u,
                                float* mean_u, wf_stream_t stream) {
    using namespace wf;
    if (B <= 0 || C <= 0) return WF_OK;
    pf::pool_finalize_kernel<<<cdiv((long long)B * C, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const unsigned long long*>(packed_u), reinterpret_cast<const unsigned long long*>(packed_m), lin, bias,
        B, C, max_m, arg_m, avg_m, max_u, arg_u, mean_u);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_pool_fused_bwd(const float* g_max_m, const float* g_avg_m, const float* g_max_u, const float* g_mean_u,
                                 const int32_t* arg_m, const int32_t* arg_u, const uint8_t* mask, const float* valid,
                                 const float* dbar, const float* W, const void* h, int B, int N, int C, int K,
                                 int32_t* work, void* dh, float* dW, float* db, wf_stream_t stream) {
    using namespace wf;
    if (B <= 0) return WF_OK;
    WF_CHECK_ARG(2 * C <= 1024 && K % 8 == 0 && K / 8 <= 256 && 256 % (K / 8) == 0,
                 "wf_pool_fused_bwd: built for C <= 512 channels and K in {64..2048} (got C=%d K=%d)", C, K);
    WF_CHECK_ARG(B <= 65535, "wf_pool_fused_bwd: B > 65535");
    WF_CHECK_ARG(((reinterpret_cast<uintptr_t>(dh) | reinterpret_cast<uintptr_t>(dbar) | reinterpret_cast<uintptr_t>(W) |
                   reinterpret_cast<uintptr_t>(dW) | reinterpret_cast<uintptr_t>(h)) & 15) == 0,
                 "wf_pool_fused_bwd: 16-byte alignment required");
    cudaStream_t s = as_stream(stream);
    const int E = 2 * C;
    // work: n_uniq[B] | u_row[B*E] | u_start[B*(E+1)] | ent_c[B*E] | ent_g[B*E]   (wf_pool_fused_bwd_work_ints)
    int* n_uniq = work;
    int* u_row = n_uniq + B;
    int* u_start = u_row + (size_t)B * E;
    int* ent_c = u_start + (size_t)B * (E + 1);
    float* ent_g = reinterpret_cast<float*>(ent_c + (size_t)B * E);
    pf::pool_bwd_prepare_kernel<<<B, 1024, 0, s>>>(g_max_m, g_max_u, arg_m, arg_u, C, n_uniq, u_row, u_start, ent_c, ent_g);
    WF_LAUNCH_CHECK();
    {
        dim3 block(K / 8, 256 / (K / 8));
        const int want = cdiv(N, (int)block.y);
        dim3 grid(1, want < 64 ? want : 64, B);
        pf::pool_bwd_fill_kernel<<<grid, block, 0, s>>>(dbar, mask, valid, B, N, K, static_cast<uint4*>(dh));
        WF_LAUNCH_CHECK();
    }
    {
        dim3 grid(E, B);
        pf::pool_bwd_rows_kernel<<<grid, 256, 0, s>>>(dbar, mask, valid, W, n_uniq, u_row, u_start, ent_c, ent_g, B, N, C, K,
                                                      static_cast<__nv_bfloat16*>(dh));
        WF_LAUNCH_CHECK();
    }
    pf::pool_bwd_dw_kernel<<<C, 256, 0, s>>>(g_max_m, g_avg_m, g_max_u, g_mean_u, arg_m, arg_u,
                                             static_cast<const __nv_bfloat16*>(h), B, N, C, K, dW, db);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_pool_fused_bwd_work_ints(int B, int C) { return B * (1 + 4 * 2 * C + 1); }
