// Microbenchmark (not part of the product library): clocks per tcgen05.mma.cta_group::2 (M = 256 over a CTA pair, K = 16, SS operands) for
// N = 64 / 128 / 256, every SM busy.  Each CTA supplies its 128 rows of A and N / 2 rows of B.  Question: is the pair form's operand cost
// per CTA A + B / 2 (the peer's half arrives over the pair link, one shared-memory read serves both tensor cores) or A + B?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_pair_mma_probe tools/pair_mma_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "../stc_unet_b200/csrc/umma.cuh"
using namespace stc;

__device__ __forceinline__ uint32_t idesc_bf16(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

struct Out { long long clk[8]; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) pair_probe(Out* out, int iters) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 48 * 1024);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int rank = (int)ptx::cluster_ctarank();
    for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
    ptx::fence_proxy_async();
    if (tid == 0) { ptx::mbar_init(ptx::smem_u32(bar), 1); ptx::fence_barrier_init(); }
    if (warp == 0) { ptx::tmem_alloc2(ptx::smem_u32(tslot), 512); ptx::tmem_relinquish2(); }
    ptx::tc_fence_before();
    ptx::cluster_sync_all();
    ptx::tc_fence_after();
    const uint32_t tbase = *tslot, bar_a = ptx::smem_u32(bar);
    const uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint64_t a_desc0 = desc_hi | (uint64_t)((ptx::smem_u32(smem) >> 4) & 0x3FFF);
    const uint64_t b_desc0 = desc_hi | (uint64_t)((ptx::smem_u32(smem + 16 * 1024) >> 4) & 0x3FFF);
    uint32_t phase = 0;
    for (int var = 0; var < 3; ++var) {
        const int N = 64 << var;
        const uint32_t idesc = idesc_bf16(256, N);
        ptx::cluster_sync_all();
        const long long t0 = clock64();
        if (warp == 0 && rank == 0) {
            if (ptx::elect_one_sync()) {
                for (int it = 0; it < iters; ++it)
                    for (int ks = 0; ks < 4; ++ks) ptx::mma_bf16_ss2(tbase, a_desc0 + ks * 2, b_desc0 + ks * 2, idesc, 1u);
                ptx::tc_commit2(bar_a, 3);
            }
            __syncwarp();
        }
        ptx::mbar_wait(bar_a, phase);
        phase ^= 1;
        ptx::tc_fence_after();
        if (tid == 0 && blockIdx.x == 0) out->clk[var] = clock64() - t0;
    }
    ptx::tc_fence_before();
    ptx::cluster_sync_all();
    if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc2(tbase, 512); }
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 3000;
    Out* d; cudaMalloc(&d, sizeof(Out)); cudaMemset(d, 0, sizeof(Out));
    const size_t smem = 48 * 1024 + 64 + 1024;
    cudaFuncSetAttribute(pair_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    pair_probe<<<(sms / 2) * 2, 128, smem>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    Out h; cudaMemcpy(&h, d, sizeof(Out), cudaMemcpyDeviceToHost);
    printf("pair MMA (cta_group::2, M = 256, K = 16, SS), %d CTAs: clocks per MMA  N=64 %.1f  N=128 %.1f  N=256 %.1f\n", (sms / 2) * 2,
           h.clk[0] / (4.0 * iters), h.clk[1] / (4.0 * iters), h.clk[2] / (4.0 * iters));
    printf("  per CTA operand bytes if the B half is read once for both tensor cores: N=64 5 KB (40 clk at 128 B/clk), N=128 6 KB (48 < math 64), N=256 8 KB (64 < math 128)\n");
    return 0;
}
