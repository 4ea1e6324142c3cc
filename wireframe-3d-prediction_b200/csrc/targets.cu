// Target preparation on the device (the step immediately before the hot path).
// Reference: train.py:48-88,112-115 builds, per batch and with O(c^2) Python loops per sample, the zero-padded ground-truth
// vertices, the existence labels, the vertex counts and one 0/1 label per candidate vertex pair (i < j < count, row-major --
// the order of models/EdgePredictor.py:84-89), padded to the batch maximum.  Here the host concatenates the ragged inputs
// once (one pinned H2D copy each) and ONE launch writes all four tensors.
#include "wf_common.cuh"

namespace wf {
namespace tgt {

__device__ __forceinline__ int find_segment(const int* __restrict__ off, int B, int t) {
    int lo = 0, hi = B;                       // largest b with off[b] <= t
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off[mid] <= t) lo = mid; else hi = mid; }
    return lo;
}

// grid-stride over three index ranges: [0, B*V) vertex slots, then the packed edges.  edge_labels is pre-zeroed.
__global__ void pack_targets_kernel(const float* __restrict__ verts, const int* __restrict__ v_off, const float* __restrict__ edges,
                                    const int* __restrict__ e_off, int B, int V, int max_e, float* __restrict__ tv,
                                    float* __restrict__ te, long long* __restrict__ counts, float* __restrict__ labels) {
    const int n_slots = B * V, n_edges = e_off[B];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_slots + n_edges; i += gridDim.x * blockDim.x) {
        if (i < n_slots) {
            const int b = i / V, k = i - b * V;
            const int cnt = v_off[b + 1] - v_off[b];                       // len(wf_vertices[b]) (train.py:56)
            const bool live = k < cnt;
            const float* src = verts + (size_t)(v_off[b] + k) * 3;
            tv[(size_t)i * 3 + 0] = live ? src[0] : 0.f;                   // train.py:112-115
            tv[(size_t)i * 3 + 1] = live ? src[1] : 0.f;
            tv[(size_t)i * 3 + 2] = live ? src[2] : 0.f;
            te[i] = live ? 1.f : 0.f;                                      // train.py:58
            if (k == 0) counts[b] = cnt;
        } else {
            const int e = i - n_slots;
            const int b = find_segment(e_off, B, e);
            const int cnt = min(v_off[b + 1] - v_off[b], V);
            // the dataset stores edge endpoints as float32 (datasets/building3d.py:180-183); train.py:70 reads them with .item()
            const int a = (int)edges[(size_t)e * 2], c = (int)edges[(size_t)e * 2 + 1];
            const int lo = min(a, c), hi = max(a, c);                      // train.py:71
            if (lo >= 0 && lo < hi && hi < cnt) {                          // only pairs (j, k), j < k < count get a label (train.py:74)
                const long long idx = (long long)lo * (2 * cnt - lo - 1) / 2 + (hi - lo - 1);
                if (idx < max_e) labels[(size_t)b * max_e + idx] = 1.f;    // duplicates write the same value
            }
        }
    }
}

}  // namespace tgt
}  // namespace wf

extern "C" int wf_pack_targets(const float* verts, const int32_t* v_off, const float* edges, const int32_t* e_off, int B, int V,
                               int max_e, int total_edges, float* tgt_vertices, float* tgt_existence, int64_t* counts,
                               float* edge_labels, wf_stream_t stream) {
    using namespace wf;
    if (B <= 0 || V <= 0) return WF_OK;
    WF_CHECK_ARG(max_e >= 0 && total_edges >= 0, "wf_pack_targets: negative sizes");
    cudaStream_t s = as_stream(stream);
    if (max_e > 0) WF_CUDA(cudaMemsetAsync(edge_labels, 0, (size_t)B * max_e * sizeof(float), s));
    const long long n = (long long)B * V + total_edges;
    const int grid = (int)(cdiv(n, 256) < 4LL * sm_count() ? cdiv(n, 256) : 4LL * sm_count());
    tgt::pack_targets_kernel<<<grid, 256, 0, s>>>(verts, v_off, edges, e_off, B, V, max_e, tgt_vertices, tgt_existence,
                                                 reinterpret_cast<long long*>(counts), edge_labels);
    WF_LAUNCH_CHECK();
    return WF_OK;
}
