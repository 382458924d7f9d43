// Per-channel reductions over the rows of an NHWC tensor ([P][C], C % 8 == 0, C/8 a power of two <= 256).
// Each thread owns one 8-channel vector lane and strides over rows, so the inner loop is pure 16/32-byte
// coalesced loads; the block reduces through shared memory and writes one fp32 partial row per block.
#pragma once
#include "common.cuh"

namespace stc {

inline bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
inline bool vec_ok(int C) { return C % 8 == 0 && pow2(C / 8) && C / 8 <= 256; }

// NV = number of reduced quantities.  acc[q][8] are this thread's sums for vector lane tid % lanes.
// emit(q, c, sum) is called once per (quantity, channel) per block.  smem must hold 256*NV*8 floats.
template <int NV, typename Emit>
__device__ __forceinline__ void block_reduce_lanes_emit(float (&acc)[NV][8], int lanes, float* smem, int C, Emit emit) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int q = 0; q < NV; ++q)
#pragma unroll
        for (int k = 0; k < 8; ++k) smem[(q * 8 + k) * 256 + tid] = acc[q][k];
    __syncthreads();
    const int per = 256 / lanes;
    for (int o = tid; o < NV * C; o += 256) {
        int q = o / C, c = o - q * C;
        int lv = c >> 3, k = c & 7;
        float s = 0.f;
        for (int r = 0; r < per; ++r) s += smem[(q * 8 + k) * 256 + r * lanes + lv];
        emit(q, c, s);
    }
}

// Writes partial[blockIdx.x][q][C] (fp32).
template <int NV>
__device__ __forceinline__ void block_reduce_lanes(float (&acc)[NV][8], int lanes, int lane_v, float* smem, float* partial, int C) {
    (void)lane_v;
    block_reduce_lanes_emit<NV>(acc, lanes, smem, C, [&](int q, int c, float s) {
        partial[((size_t)blockIdx.x * NV + q) * C + c] = s;
    });
}

// sums partial[g][j] over g in fp64
static __global__ void reduce_partials_kernel(const float* __restrict__ partial, double* __restrict__ out, int G, int len) {
    // launched with 32 threads per output element (ceil_div(len * 32, blockDim))
    int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (j >= len) return;
    double s = 0.0;
    int g = lane;
    for (; g + 7 * 32 < G; g += 8 * 32) {   // 8 independent loads in flight per lane
        float a[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] = partial[(size_t)(g + u * 32) * len + j];
#pragma unroll
        for (int u = 0; u < 8; ++u) s += (double)a[u];
    }
    for (; g < G; g += 32) s += (double)partial[(size_t)g * len + j];
    s = warp_sum(s);
    if (lane == 0) out[j] = s;
}

inline int reduce_blocks(long long P, int lanes) {
    long long rows_per_block_iter = 256 / lanes;
    long long want = (P + rows_per_block_iter * 8 - 1) / (rows_per_block_iter * 8);  // >= 8 rows per thread
    long long cap = (long long)num_sms() * 4;
    long long g = want < cap ? want : cap;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace stc
