// Microbenchmark (not part of the product library): does reading accumulators out of tensor memory (tcgen05.ld) overlap tcgen05.mma?
// One CTA per SM; warp 4 issues `iters` x 4 SS MMAs (M = 128, K = 16, N = 64 or 256) into TMEM columns [0, 256); warps 0..3 (one per lane
// quarter) read columns [256, 512) `passes` times with tcgen05.ld.32x32b.x32 / .x64 / .x128.  Times: MMAs alone, loads alone, both at once.
// If the two overlapped, "both" would be max(MMA, ld); DESIGN 3.1 claims it is close to the SUM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_tmem_ld_probe tools/tmem_ld_probe.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../stc_unet_b200/csrc/umma.cuh"
using namespace stc;

__device__ __forceinline__ uint32_t idesc_bf16(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

template <int X>
__device__ __forceinline__ uint32_t tmem_ld_x(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t tmem_ld_x<32>(uint32_t taddr) {
    uint32_t v[32];
    ptx::tmem_ld_32x32(taddr, v);
    ptx::tmem_ld_wait();
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) s ^= v[j];
    return s;
}
template <>
__device__ __forceinline__ uint32_t tmem_ld_x<64>(uint32_t taddr) {
    uint32_t v[64];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
          "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
          "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]),
          "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]),
          "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),
          "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr)
        : "memory");
    ptx::tmem_ld_wait();
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 64; ++j) s ^= v[j];
    return s;
}
// two x32 loads in flight before one wait
template <>
__device__ __forceinline__ uint32_t tmem_ld_x<2>(uint32_t taddr) {
    uint32_t a[32], b[32];
    ptx::tmem_ld_32x32(taddr, a);
    ptx::tmem_ld_32x32(taddr + 32, b);
    ptx::tmem_ld_wait();
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) s ^= a[j] ^ b[j];
    return s;
}

struct Out { long long clk[32]; unsigned sink; };

template <int X>
__global__ void __launch_bounds__(160, 1) ld_probe(Out* out, int iters, int passes, int N, int slot, int M = 128) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 48 * 1024);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 48 * 1024 / 4; i += 160) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;   // arbitrary bf16 operands
    ptx::fence_proxy_async();
    if (tid == 0) { ptx::mbar_init(ptx::smem_u32(bar), 1); ptx::fence_barrier_init(); }
    if (warp == 4) { ptx::tmem_alloc(ptx::smem_u32(tslot), 512); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tbase = *tslot, bar_a = ptx::smem_u32(bar);
    const uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint64_t a_desc0 = desc_hi | (uint64_t)((ptx::smem_u32(smem) >> 4) & 0x3FFF);
    const uint64_t b_desc0 = desc_hi | (uint64_t)((ptx::smem_u32(smem + 16 * 1024) >> 4) & 0x3FFF);
    const uint32_t idesc = idesc_bf16(M, N);
    uint32_t phase = 0, sink = 0;
    constexpr int COLS = X == 2 ? 64 : X;   // columns per call
    for (int mode = 1; mode <= 3; ++mode) {   // 1 = MMAs only, 2 = loads only, 3 = both
        __syncthreads();
        const long long t0 = clock64();
        if (warp == 4) {
            if (mode & 1) {
                if (ptx::elect_one_sync()) {
                    for (int it = 0; it < iters; ++it)
                        for (int ks = 0; ks < 4; ++ks) ptx::mma_bf16_ss(tbase, a_desc0 + ks * 2, b_desc0 + ks * 2, idesc, 1u);
                }
                __syncwarp();
            }
            if (ptx::elect_one_sync()) ptx::tc_commit(bar_a);
            __syncwarp();
        } else if (mode & 2) {
            const uint32_t t_addr = tbase + 256 + ((uint32_t)(warp * 32) << 16);
            for (int ps = 0; ps < passes; ++ps)
                for (int c = 0; c < 256; c += COLS) sink ^= tmem_ld_x<X>(t_addr + c);
        }
        const long long t_ld = clock64() - t0;   // readers: their own time
        ptx::mbar_wait(bar_a, phase);
        phase ^= 1;
        ptx::tc_fence_after();
        __syncthreads();
        if (blockIdx.x == 0) {
            if (tid == 0) out->clk[slot * 8 + mode] = clock64() - t0;
            if (tid == 0 && (mode & 2)) out->clk[slot * 8 + 4 + mode] = t_ld;
        }
    }
    if (sink == 0x12345678u) out->sink = sink;
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 4) { ptx::tc_fence_after(); ptx::tmem_dealloc(tbase, 512); }
}

int main(int argc, char** argv) {
    const int passes = argc > 1 ? atoi(argv[1]) : 400;
    Out* d; cudaMalloc(&d, sizeof(Out)); cudaMemset(d, 0, sizeof(Out));
    const size_t smem = 48 * 1024 + 64 + 1024;
    cudaFuncSetAttribute(ld_probe<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(ld_probe<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(ld_probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const double bytes = (double)passes * 128 * 256 * 4;   // read per CTA
    for (int N : {64, 256}) {
        // MMA count chosen so that MMAs alone take about as long as the loads alone (64 B/clk assumed)
        const int mma_clk = N == 64 ? 48 : 128;
        const int iters = (int)(bytes / 64.0 / mma_clk / 4.0);
        for (int shape = 0; shape < 3; ++shape) {
            if (shape == 0) ld_probe<32><<<sms, 160, smem>>>(d, iters, passes, N, shape);
            if (shape == 1) ld_probe<64><<<sms, 160, smem>>>(d, iters, passes, N, shape);
            if (shape == 2) ld_probe<2><<<sms, 160, smem>>>(d, iters, passes, N, shape);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        }
        Out h; cudaMemcpy(&h, d, sizeof(Out), cudaMemcpyDeviceToHost);
        const char* names[3] = {"32x32b.x32, wait each", "32x32b.x64, wait each", "2 x (32x32b.x32), one wait"};
        for (int shape = 0; shape < 3; ++shape) {
            const long long* c = h.clk + shape * 8;
            printf("N=%3d  ld %-28s MMAs alone %8lld clk (%.1f per MMA) | loads alone %8lld clk (%.1f B/clk/SM) | both %8lld clk = %.2f x max, %.2f x sum; loads took %lld\n",
                   N, names[shape], c[1], c[1] / (4.0 * iters), c[2], bytes / c[2], c[3], (double)c[3] / (c[1] > c[2] ? c[1] : c[2]),
                   (double)c[3] / (c[1] + c[2]), c[4 + 3]);
        }
    }
    // M = 64 (weights-as-A formulations): clocks per MMA, MMAs alone
    for (int N : {128, 256}) {
        const int iters = 2000;
        cudaMemset(d, 0, sizeof(Out));
        ld_probe<32><<<sms, 160, smem>>>(d, iters, 1, N, 0, 64);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        Out h; cudaMemcpy(&h, d, sizeof(Out), cudaMemcpyDeviceToHost);
        printf("M= 64 N=%3d  SS MMAs alone: %.1f clocks per MMA (%.0f flop/clk/SM of 8192)\n", N, h.clk[1] / (4.0 * iters), 2.0 * 64 * N * 16 / (h.clk[1] / (4.0 * iters)));
    }
    return 0;
}
