// Microbenchmark / semantics probe (not part of the product library): the A operand of tcgen05.mma from TENSOR MEMORY.
//
//   1. correctness: D1 = A B^T with both operands in shared memory (SS) against D2 = the same product with A copied shared -> tensor memory
//      by tcgen05.cp.128x256b (one K = 16 slice = 8 TMEM columns per copy, SW128 K-major descriptor, also with the halo kernels' SHIFTED
//      start address) and read by the MMA from TMEM (TS): bit-identical, and both against a host fp32 reference;
//   2. throughput: clocks per M = 128, K = 16 MMA for N = 64 / 128 / 256, SS vs TS, on every SM at once - does taking A out of shared
//      memory remove the operand bound of the N = 64 layers? - and with one copy per `reuse` MMAs in the loop.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_tmem_a_probe tools/tmem_a_probe.cu ; run on a B200.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#include "../stc_unet_b200/csrc/umma.cuh"

using namespace stc;

__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ uint32_t make_idesc_bf16_dev(int M, int N) {   // kind::f16, bf16 x bf16 -> fp32, both operands K-major
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// element (row, k) of a [rows][64] bf16 K-major tile in the 128B swizzle (1024-byte aligned base): 8-row groups of 1024 B
__host__ __device__ inline uint32_t sw128_off(int row, int k) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2);
}

constexpr int kARows = 136;   // 128 + shift room: the shifted descriptor starts `shift` pixel rows (128 B each) into the tile

struct Result {
    float d_ss[128 * 64];
    float d_ts[128 * 64];
    long long clk[16];
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, Result* out, int shift,
                                                     int iters) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                       // [kARows -> 17 groups of 8 rows][64] bf16, SW128
    uint8_t* sB = smem + 18 * 1024;           // [256][64] bf16, SW128 (N up to 256)
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 18 * 1024 + 32 * 1024);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < kARows * 64; i += 128) {
        int r = i >> 6, k = i & 63;
        *reinterpret_cast<__nv_bfloat16*>(sA + sw128_off(r, k)) = A[r * 64 + k];
    }
    for (int i = tid; i < 256 * 64; i += 128) {
        int r = i >> 6, k = i & 63;
        *reinterpret_cast<__nv_bfloat16*>(sB + sw128_off(r, k)) = B[r * 64 + k];
    }
    ptx::fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async proxy
    if (tid == 0) { ptx::mbar_init(ptx::smem_u32(bar), 1); ptx::fence_barrier_init(); }
    if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(tslot), 512); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tbase = *tslot;
    const uint32_t bar_a = ptx::smem_u32(bar);
    const uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint64_t a_desc0 = desc_hi | (uint64_t)(((ptx::smem_u32(sA) >> 4) & 0x3FFF) + shift * 8);   // + shift pixel rows of 128 B
    const uint64_t b_desc0 = desc_hi | (uint64_t)((ptx::smem_u32(sB) >> 4) & 0x3FFF);
    // TMEM columns: D_ss [0, 64), D_ts [64, 128), timing accumulator [128, 384), A slices [448, 480) (4 K steps x 8 columns)
    const uint32_t tD1 = tbase, tD2 = tbase + 64, tDt = tbase + 128, tA = tbase + 448;
    uint32_t phase = 0;
    auto commit_wait = [&]() {
        if (warp == 0) {
            if (ptx::elect_one_sync()) ptx::tc_commit(bar_a);
            __syncwarp();
        }
        ptx::mbar_wait(bar_a, phase);
        phase ^= 1;
        ptx::tc_fence_after();
    };
    // ---- 1. correctness ----
    const uint32_t idesc64 = make_idesc_bf16_dev(128, 64);
    if (warp == 0) {
        if (ptx::elect_one_sync()) {
            for (int ks = 0; ks < 4; ++ks) ptx::mma_bf16_ss(tD1, a_desc0 + ks * 2, b_desc0 + ks * 2, idesc64, ks ? 1u : 0u);
            for (int ks = 0; ks < 4; ++ks) tmem_cp_128x256b(tA + ks * 8, a_desc0 + ks * 2);
            for (int ks = 0; ks < 4; ++ks) mma_bf16_ts(tD2, tA + ks * 8, b_desc0 + ks * 2, idesc64, ks ? 1u : 0u);
        }
        __syncwarp();
    }
    commit_wait();
    if (blockIdx.x == 0) {
        for (int c = 0; c < 64; c += 32) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(tD1 + c + ((uint32_t)(warp * 32) << 16), v);
            ptx::tmem_ld_wait();
            for (int j = 0; j < 32; ++j) out->d_ss[(warp * 32 + lane) * 64 + c + j] = __uint_as_float(v[j]);
            ptx::tmem_ld_32x32(tD2 + c + ((uint32_t)(warp * 32) << 16), v);
            ptx::tmem_ld_wait();
            for (int j = 0; j < 32; ++j) out->d_ts[(warp * 32 + lane) * 64 + c + j] = __uint_as_float(v[j]);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    // ---- 2. throughput: `iters` x 4 K steps; variants: [0..2] SS N = 64 / 128 / 256, [3..5] TS, [6..8] TS with one copy per 3 MMAs (reuse 3) ----
    for (int var = 0; var < 9; ++var) {
        const int N = 64 << (var % 3);
        const uint32_t idesc = make_idesc_bf16_dev(128, N);
        const int kind = var / 3;
        __syncthreads();
        const long long t0 = clock64();
        if (warp == 0) {
            if (ptx::elect_one_sync()) {
                for (int it = 0; it < iters; ++it) {
                    if (kind == 0) {
                        for (int ks = 0; ks < 4; ++ks) ptx::mma_bf16_ss(tDt, a_desc0 + ks * 2, b_desc0 + ks * 2, idesc, 1u);
                    } else if (kind == 1) {
                        for (int ks = 0; ks < 4; ++ks) mma_bf16_ts(tDt, tA + ks * 8, b_desc0 + ks * 2, idesc, 1u);
                    } else {
                        if (it % 3 == 0) for (int ks = 0; ks < 4; ++ks) tmem_cp_128x256b(tA + ks * 8, a_desc0 + ks * 2);
                        for (int ks = 0; ks < 4; ++ks) mma_bf16_ts(tDt, tA + ks * 8, b_desc0 + ks * 2, idesc, 1u);
                    }
                }
            }
            __syncwarp();
        }
        commit_wait();
        if (tid == 0 && blockIdx.x == 0) out->clk[var] = clock64() - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tbase, 512); }
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 3000;
    std::vector<__nv_bfloat16> hA(kARows * 64), hB(256 * 64);
    srand(7);
    for (auto& v : hA) v = __float2bfloat16((rand() % 2001 - 1000) / 1000.f);
    for (auto& v : hB) v = __float2bfloat16((rand() % 2001 - 1000) / 1000.f);
    __nv_bfloat16 *dA, *dB;
    Result* dR;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dR, sizeof(Result));
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    const size_t smem = 18 * 1024 + 32 * 1024 + 64 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    std::vector<Result> hR(1);
    for (int shift : {0, 3}) {
        cudaMemset(dR, 0, sizeof(Result));
        probe_kernel<<<sms, 128, smem>>>(dA, dB, dR, shift, iters);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("shift %d: CUDA error %s\n", shift, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hR.data(), dR, sizeof(Result), cudaMemcpyDeviceToHost);
        double max_ref = 0, max_diff = 0; int mismatch = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 64; ++n) {
                double ref = 0;
                for (int k = 0; k < 64; ++k) ref += (double)__bfloat162float(hA[(m + shift) * 64 + k]) * (double)__bfloat162float(hB[n * 64 + k]);
                max_ref = fmax(max_ref, fabs(hR[0].d_ss[m * 64 + n] - ref));
                max_diff = fmax(max_diff, fabs((double)hR[0].d_ss[m * 64 + n] - (double)hR[0].d_ts[m * 64 + n]));
                mismatch += hR[0].d_ss[m * 64 + n] != hR[0].d_ts[m * 64 + n];
            }
        printf("shift %d: SS vs host reference max |diff| %.3e ; TS (A via tcgen05.cp.128x256b) vs SS: max |diff| %.3e, %d of 8192 elements differ\n", shift,
               max_ref, max_diff, mismatch);
        const char* names[3] = {"SS (A, B in shared memory)", "TS (A in tensor memory)", "TS + one 4-slice copy per 3 K-blocks"};
        for (int kind = 0; kind < 3; ++kind)
            printf("  %-38s clocks per MMA (M=128, K=16), all %d SMs busy: N=64 %.1f  N=128 %.1f  N=256 %.1f\n", names[kind], sms,
                   hR[0].clk[kind * 3 + 0] / (4.0 * iters), hR[0].clk[kind * 3 + 1] / (4.0 * iters), hR[0].clk[kind * 3 + 2] / (4.0 * iters));
    }
    return 0;
}
