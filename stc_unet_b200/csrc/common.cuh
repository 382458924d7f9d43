// Common helpers for the stc_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/stc_b200.h"

namespace stc {

// ---- error plumbing (C ABI returns int; message kept per thread) ----------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define STC_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            stc::set_error(__VA_ARGS__);       \
            return STC_ERR_INVALID;            \
        }                                      \
    } while (0)

#define STC_CUDA(call)                                                          \
    do {                                                                        \
        cudaError_t e__ = (call);                                               \
        if (e__ != cudaSuccess) {                                               \
            stc::set_error("%s failed: %s", #call, cudaGetErrorString(e__));    \
            return STC_ERR_CUDA;                                                \
        }                                                                       \
    } while (0)

// ---- element access --------------------------------------------------------
typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// 8-element vector (the unit all NHWC kernels use along C; C % 8 == 0 fast path)
template <typename T> struct Vec8;
template <> struct Vec8<float> {
    float v[8];
    __device__ __forceinline__ void load(const float* p) {
        float4 a = *reinterpret_cast<const float4*>(p);
        float4 b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    __device__ __forceinline__ void store(float* p) const {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
};
template <> struct Vec8<bf16> {
    float v[8];
    __device__ __forceinline__ void load(const bf16* p) {
        uint4 r = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __bfloat1622float2(h[i]);
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
    __device__ __forceinline__ void store(bf16* p) const {
        uint4 r;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = r;
    }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Virtual channel concatenation (torch.cat([skip, up]) of Up.forward / UpConvBlock.forward / the UNet++ decoder blocks is never
// materialised): up to 5 NHWC tensors that share (N, H, W) act as ONE tensor whose channel axis is their concatenation.  As a conv INPUT
// the K loop walks the sources' 64-channel chunks through one TMA descriptor per source; as a dgrad OUTPUT every 64-channel chunk of the
// result is stored through the descriptor of the tensor that owns it.  n == 0 / nullptr = a plain single tensor.
constexpr int kMaxCat = 5;
struct ChanCat {
    int n;
    const void* ptr[kMaxCat];
    int c[kMaxCat];
    int total() const { int t = 0; for (int i = 0; i < n; ++i) t += c[i]; return t; }
};

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
int num_sms();
int max_cta_pairs(const void* func, int threads, size_t smem);

// dtype dispatch: calls f(T{}) with T = float or bf16
#define STC_DISPATCH_DTYPE(dtype, ...)                                  \
    do {                                                                \
        if ((dtype) == STC_F32) { typedef float T; __VA_ARGS__; }       \
        else if ((dtype) == STC_BF16) { typedef stc::bf16 T; __VA_ARGS__; } \
        else { stc::set_error("bad dtype %d", (int)(dtype)); return STC_ERR_INVALID; } \
    } while (0)

}  // namespace stc
