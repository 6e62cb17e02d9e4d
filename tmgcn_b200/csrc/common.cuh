// Shared helpers for libtmgcn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/tmgcn.h"

namespace tmgcn {

void set_error(const char *fmt, ...);
extern std::atomic<int64_t> g_launches;

// call right after a <<<>>> launch
int after_launch(const char *what);
int sm_count();
size_t l2_bytes();

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
__device__ __forceinline__ int64_t ceil_div_dev(int64_t a, int64_t b) { return (a + b - 1) / b; }

#define TMGCN_REQUIRE(cond, ...)           \
    do {                                   \
        if (!(cond)) {                     \
            ::tmgcn::set_error(__VA_ARGS__); \
            return 1;                      \
        }                                  \
    } while (0)

#define TMGCN_CUDA(call)                                                              \
    do {                                                                              \
        cudaError_t e__ = (call);                                                     \
        if (e__ != cudaSuccess) {                                                     \
            ::tmgcn::set_error("%s failed: %s", #call, cudaGetErrorString(e__));      \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

// ---- activation (ref: ehf:284-289) ---------------------------------------
#define TMGCN_SELU_ALPHA 1.6732632423543772848170429916717f
#define TMGCN_SELU_SCALE 1.0507009873554804934193349852946f

template <int ACT>
__device__ __forceinline__ float act_apply(float x) {
    if (ACT == TMGCN_ACT_RELU) return x > 0.f ? x : 0.f;
    if (ACT == TMGCN_ACT_LEAKY) return x > 0.f ? x : 0.01f * x;
    if (ACT == TMGCN_ACT_SELU) return TMGCN_SELU_SCALE * (x > 0.f ? x : TMGCN_SELU_ALPHA * expm1f(x));
    return x;
}
// derivative expressed through the OUTPUT y = act(x)
template <int ACT>
__device__ __forceinline__ float act_grad_from_out(float y) {
    if (ACT == TMGCN_ACT_RELU) return y > 0.f ? 1.f : 0.f;
    if (ACT == TMGCN_ACT_LEAKY) return y > 0.f ? 1.f : 0.01f;
    if (ACT == TMGCN_ACT_SELU) return y > 0.f ? TMGCN_SELU_SCALE : y + TMGCN_SELU_SCALE * TMGCN_SELU_ALPHA;
    return 1.f;
}
__device__ __forceinline__ float act_apply_rt(float x, int act) {
    switch (act) {
        case TMGCN_ACT_RELU: return act_apply<TMGCN_ACT_RELU>(x);
        case TMGCN_ACT_LEAKY: return act_apply<TMGCN_ACT_LEAKY>(x);
        case TMGCN_ACT_SELU: return act_apply<TMGCN_ACT_SELU>(x);
        default: return x;
    }
}
__device__ __forceinline__ float act_grad_rt(float y, int act) {
    switch (act) {
        case TMGCN_ACT_RELU: return act_grad_from_out<TMGCN_ACT_RELU>(y);
        case TMGCN_ACT_LEAKY: return act_grad_from_out<TMGCN_ACT_LEAKY>(y);
        case TMGCN_ACT_SELU: return act_grad_from_out<TMGCN_ACT_SELU>(y);
        default: return 1.f;
    }
}

// streaming (read-once / write-once) 128-bit accesses: keep them out of L1
__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_f4(float4 *p, const float4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

}  // namespace tmgcn
