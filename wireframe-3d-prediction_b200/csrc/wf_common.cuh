// Shared helpers for libwf_b200: error plumbing, dtype-generic loads/stores, warp reductions.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/wf_b200.h"

namespace wf {

void set_error(const char* fmt, ...);

#define WF_CHECK_ARG(cond, ...)                        \
    do {                                               \
        if (!(cond)) {                                 \
            wf::set_error(__VA_ARGS__);                \
            return WF_EINVAL;                          \
        }                                              \
    } while (0)

#define WF_CUDA(expr)                                                                     \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            wf::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return WF_ECUDA;                                                              \
        }                                                                                 \
    } while (0)

#define WF_LAUNCH_CHECK() WF_CUDA(cudaGetLastError())

static inline cudaStream_t as_stream(wf_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

int sm_count();

// ---- dtype-generic element access (fp32 / bf16 storage, fp32 math) -------------------------
template <int DT> struct elem;
template <> struct elem<WF_F32> {
    using type = float;
    __device__ static __forceinline__ float ld(const void* p, size_t i) { return static_cast<const float*>(p)[i]; }
    __device__ static __forceinline__ void st(void* p, size_t i, float v) { static_cast<float*>(p)[i] = v; }
};
template <> struct elem<WF_BF16> {
    using type = __nv_bfloat16;
    __device__ static __forceinline__ float ld(const void* p, size_t i) {
        return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
    }
    __device__ static __forceinline__ void st(void* p, size_t i, float v) {
        static_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
    }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// exact-erf GELU (nn.GELU() default) and its derivative
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
    const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}
__device__ __forceinline__ float act_f(int act, float y) {
    return act == WF_ACT_RELU ? fmaxf(y, 0.0f) : (act == WF_ACT_GELU ? gelu_f(y) : y);
}
__device__ __forceinline__ float act_grad_f(int act, float y) {
    return act == WF_ACT_RELU ? (y > 0.0f ? 1.0f : 0.0f) : (act == WF_ACT_GELU ? gelu_grad_f(y) : 1.0f);
}

}  // namespace wf
