// Fused WireframeLoss forward/backward: matched SmoothL1 on vertices, BCE on existence
// probabilities (slot order), BCE on the zero-padded edge block, weighted sum.
// Reference: losses/WireframeLoss.py:38-104 (forward), :248-283 (_compute_matched_vertex_loss).
// Deterministic: a fixed grid writes per-CTA partial sums, a second tiny kernel adds them in order.
#include "wf_common.cuh"

namespace wf {
namespace loss {

constexpr int NPART = 64;          // partial-sum CTAs
// out layout (floats): [0..3] total, vertex, existence, edge; [4] match count; [8 + 4*p + k] partials
constexpr int OUT_FLOATS = 8 + 4 * NPART;

__device__ __forceinline__ float bce_term(float p, float y) {
    // torch.nn.functional.binary_cross_entropy clamps both logs at -100
    const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(logf(1.0f - p), -100.f);
    return -(y * lp + (1.0f - y) * lq);
}
__device__ __forceinline__ float bce_grad(float p, float y) {
    return (p - y) / fmaxf((1.0f - p) * p, 1e-12f);
}

__device__ __forceinline__ float block_sum(float v, float* scratch) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float t = 0.f;
    if (warp == 0) {
        t = lane < (blockDim.x >> 5) ? scratch[lane] : 0.f;
        t = warp_sum(t);
    }
    __syncthreads();
    return t;          // valid in warp 0
}

__global__ void __launch_bounds__(256)
loss_partial_kernel(const float* __restrict__ pred_v, const float* __restrict__ pred_e, const float* __restrict__ edge_p,
                    const float* __restrict__ tgt_v, const float* __restrict__ tgt_e, const float* __restrict__ edge_l,
                    const int* __restrict__ col_of_row, const long long* __restrict__ counts, int B, int V, int Vt, int Ep,
                    int El, int min_e, float* __restrict__ out) {
    __shared__ float scratch[8];
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    float sv = 0.f, nm = 0.f, sx = 0.f, se = 0.f;
    for (long long k = tid; k < (long long)B * V; k += nth) {
        const int b = (int)(k / V);
        sx += bce_term(pred_e[k], tgt_e[k]);
        const int col = col_of_row[k];
        if (col >= 0 && col < counts[b]) {
            const float* p = pred_v + k * 3;
            const float* t = tgt_v + ((size_t)b * Vt + col) * 3;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float df = p[d] - t[d], a = fabsf(df);
                sv += a < 1.0f ? 0.5f * df * df : a - 0.5f;
            }
            nm += 1.0f;
        }
    }
    for (long long k = tid; k < (long long)B * min_e; k += nth) {
        const int b = (int)(k / min_e), j = (int)(k - (long long)b * min_e);
        se += bce_term(edge_p[(size_t)b * Ep + j], edge_l[(size_t)b * El + j]);
    }
    sv = block_sum(sv, scratch); nm = block_sum(nm, scratch); sx = block_sum(sx, scratch); se = block_sum(se, scratch);
    if (threadIdx.x == 0) {
        float* o = out + 8 + 4 * blockIdx.x;
        o[0] = sv; o[1] = nm; o[2] = sx; o[3] = se;
    }
}

__global__ void loss_final_kernel(float* __restrict__ out, int B, int V, int min_e, float wv, float we, float wx) {
    if (threadIdx.x != 0) return;
    float sv = 0.f, nm = 0.f, sx = 0.f, se = 0.f;
    for (int p = 0; p < NPART; ++p) { const float* o = out + 8 + 4 * p; sv += o[0]; nm += o[1]; sx += o[2]; se += o[3]; }
    const float lv = nm > 0.f ? sv / (3.0f * nm) : 0.f;
    const float lx = sx / ((float)B * (float)V);
    const float le = min_e > 0 ? se / ((float)B * (float)min_e) : 0.f;
    out[0] = wv * lv + wx * lx + we * le; out[1] = lv; out[2] = lx; out[3] = le; out[4] = nm;
}

__global__ void __launch_bounds__(256)
loss_bwd_kernel(const float* __restrict__ g_out, const float* __restrict__ pred_v, const float* __restrict__ pred_e,
                const float* __restrict__ edge_p, const float* __restrict__ tgt_v, const float* __restrict__ tgt_e,
                const float* __restrict__ edge_l, const int* __restrict__ col_of_row, const long long* __restrict__ counts,
                const float* __restrict__ fwd_out, int B, int V, int Vt, int Ep, int El, int min_e, float wv, float we,
                float wx, float* __restrict__ d_pred_v, float* __restrict__ d_pred_e, float* __restrict__ d_edge_p) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    const float gt = g_out[0];
    const float nm = fwd_out[4];
    const float cv = nm > 0.f ? (gt * wv + g_out[1]) / (3.0f * nm) : 0.f;
    const float cx = (gt * wx + g_out[2]) / ((float)B * (float)V);
    const float ce = min_e > 0 ? (gt * we + g_out[3]) / ((float)B * (float)min_e) : 0.f;
    for (long long k = tid; k < (long long)B * V; k += nth) {
        const int b = (int)(k / V);
        d_pred_e[k] = cx * bce_grad(pred_e[k], tgt_e[k]);
        const int col = col_of_row[k];
        float g0 = 0.f, g1 = 0.f, g2 = 0.f;
        if (col >= 0 && col < counts[b]) {
            const float* p = pred_v + k * 3;
            const float* t = tgt_v + ((size_t)b * Vt + col) * 3;
            float df = p[0] - t[0]; g0 = cv * (fabsf(df) < 1.0f ? df : (df > 0.f ? 1.0f : -1.0f));
            df = p[1] - t[1]; g1 = cv * (fabsf(df) < 1.0f ? df : (df > 0.f ? 1.0f : -1.0f));
            df = p[2] - t[2]; g2 = cv * (fabsf(df) < 1.0f ? df : (df > 0.f ? 1.0f : -1.0f));
        }
        d_pred_v[k * 3] = g0; d_pred_v[k * 3 + 1] = g1; d_pred_v[k * 3 + 2] = g2;
    }
    for (long long k = tid; k < (long long)B * Ep; k += nth) {
        const int b = (int)(k / Ep), j = (int)(k - (long long)b * Ep);
        d_edge_p[k] = j < min_e ? ce * bce_grad(edge_p[k], edge_l[(size_t)b * El + j]) : 0.f;
    }
}

}  // namespace loss
}  // namespace wf

extern "C" int wf_loss_out_floats(void) { return wf::loss::OUT_FLOATS; }

extern "C" int wf_loss_fwd(const float* pred_v, const float* pred_e, const float* edge_p, const float* tgt_v, const float* tgt_e,
                           const float* edge_l, const int32_t* col_of_row, const int64_t* counts, int B, int V, int Vt, int Ep,
                           int El, float w_vertex, float w_edge, float w_exist, float* out, wf_stream_t stream) {
    using namespace wf;
    WF_CHECK_ARG(B > 0 && V > 0, "wf_loss_fwd: empty batch");
    const int min_e = (Ep < El ? Ep : El);
    loss::loss_partial_kernel<<<loss::NPART, 256, 0, as_stream(stream)>>>(pred_v, pred_e, edge_p, tgt_v, tgt_e, edge_l, col_of_row,
        reinterpret_cast<const long long*>(counts), B, V, Vt, Ep, El, min_e, out);
    WF_LAUNCH_CHECK();
    loss::loss_final_kernel<<<1, 32, 0, as_stream(stream)>>>(out, B, V, min_e, w_vertex, w_edge, w_exist);
    WF_LAUNCH_CHECK();
    return WF_OK;
}

extern "C" int wf_loss_bwd(const float* g_out, const float* pred_v, const float* pred_e, const float* edge_p, const float* tgt_v,
                           const float* tgt_e, const float* edge_l, const int32_t* col_of_row, const int64_t* counts,
                           const float* fwd_out, int B, int V, int Vt, int Ep, int El, float w_vertex, float w_edge,
                           float w_exist, float* d_pred_v, float* d_pred_e, float* d_edge_p, wf_stream_t stream) {
    using namespace wf;
    WF_CHECK_ARG(B > 0 && V > 0, "wf_loss_bwd: empty batch");
    const int min_e = (Ep < El ? Ep : El);
    const long long work = (long long)B * (V > Ep ? V : Ep);
    const int grid = (int)(cdiv(work, 256) < 4 * sm_count() ? cdiv(work, 256) : 4 * sm_count());
    loss::loss_bwd_kernel<<<grid < 1 ? 1 : grid, 256, 0, as_stream(stream)>>>(g_out, pred_v, pred_e, edge_p, tgt_v, tgt_e, edge_l,
        col_of_row, reinterpret_cast<const long long*>(counts), fwd_out, B, V, Vt, Ep, El, min_e, w_vertex, w_edge, w_exist,
        d_pred_v, d_pred_e, d_edge_p);
    WF_LAUNCH_CHECK();
    return WF_OK;
}
