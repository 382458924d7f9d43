// C-ABI entry points of the dense contractions: engine selection between the tcgen05 kernel
// (umma.cu) and the fp32-accumulate SIMT kernel (conv_simt.cu).
#include "common.cuh"

namespace stc {
int conv_fprop_simt(const void*, const void*, const float*, const void*, void*, int, int, int, int, int, int, int, int, int,
                    cudaStream_t);
int conv_wgrad_simt(const void*, const void*, float*, int, int, int, int, int, int, int, int, cudaStream_t);
int gemm_simt(const void*, const void*, void*, const stc_gemm_desc*, int, cudaStream_t);
bool conv_umma_eligible(int Cin, int Cout, int dtype);
int conv_fprop_umma(const void*, const void*, const float*, const void*, void*, int, int, int, int, int, int, int, int,
                    cudaStream_t);
int conv_wgrad_umma(const void*, const void*, float*, int, int, int, int, int, int, int, cudaStream_t);
bool gemm_umma_eligible(const stc_gemm_desc*, int dtype);
bool conv_convh_eligible(int W, int Cin, int Cout, int R, int S, int dtype);
bool conv_wgradh_eligible(int W, int Cin, int Cout, int R, int S, int dtype);
int conv_wgrad_wgradh(const void*, const void*, float*, int, int, int, int, int, int, int, cudaStream_t);
int conv_fprop_convh(const void*, const void*, const float*, const void*, void*, int, int, int, int, int, int, int, int, cudaStream_t);
int gemm_umma(const void*, const void*, void*, const stc_gemm_desc*, int, cudaStream_t);
}  // namespace stc

using namespace stc;

static thread_local int g_last_engine = 0;
/* engine actually used by the last stc_conv_fprop / stc_conv_wgrad / stc_gemm call on this thread */
extern "C" int stc_dense_last_engine(void) { return g_last_engine; }

extern "C" int stc_conv_fprop(const void* x, const void* wp, const float* bias, const void* residual, void* y, int N, int H,
                              int W, int Cin, int Cout, int R, int S, int act, int dtype, int engine, void* stream) {
    STC_REQUIRE(N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && (R & 1) && (S & 1), "conv_fprop: bad shape N=%d H=%d W=%d Cin=%d Cout=%d R=%d S=%d",
                N, H, W, Cin, Cout, R, S);
    cudaStream_t st = (cudaStream_t)stream;
    bool elig = conv_umma_eligible(Cin, Cout, dtype);
    if (engine != STC_ENGINE_SIMT && elig && conv_convh_eligible(W, Cin, Cout, R, S, dtype)) {
        g_last_engine = STC_KERNEL_CONVH;  // halo-reuse variant of the tcgen05 engine
        return conv_fprop_convh(x, wp, bias, residual, y, N, H, W, Cin, Cout, R, S, act, st);
    }
    if (engine == STC_ENGINE_TCGEN05) {
        STC_REQUIRE(elig, "conv_fprop: tcgen05 engine requested but shape/dtype not eligible (Cin=%d Cout=%d dtype=%d)", Cin, Cout, dtype);
        g_last_engine = STC_ENGINE_TCGEN05;
        return conv_fprop_umma(x, wp, bias, residual, y, N, H, W, Cin, Cout, R, S, act, st);
    }
    if (engine == STC_ENGINE_AUTO && elig) {
        g_last_engine = STC_ENGINE_TCGEN05;
        return conv_fprop_umma(x, wp, bias, residual, y, N, H, W, Cin, Cout, R, S, act, st);
    }
    g_last_engine = STC_ENGINE_SIMT;
    return conv_fprop_simt(x, wp, bias, residual, y, N, H, W, Cin, Cout, R, S, act, dtype, st);
}

extern "C" int stc_conv_wgrad(const void* x, const void* dy, float* dw_ws, int N, int H, int W, int Cin, int Cout, int R,
                              int S, int dtype, int engine, void* stream) {
    STC_REQUIRE(N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && (R & 1) && (S & 1), "conv_wgrad: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    bool elig = dtype == STC_BF16 && Cin % 64 == 0 && Cout % 64 == 0;
    if (engine != STC_ENGINE_SIMT && elig && conv_wgradh_eligible(W, Cin, Cout, R, S, dtype)) {
        g_last_engine = STC_KERNEL_WGRADH;  // halo-reuse variant
        return conv_wgrad_wgradh(x, dy, dw_ws, N, H, W, Cin, Cout, R, S, st);
    }
    if (engine == STC_ENGINE_TCGEN05) {
        STC_REQUIRE(elig, "conv_wgrad: tcgen05 engine requested but shape/dtype not eligible");
        g_last_engine = STC_ENGINE_TCGEN05;
        return conv_wgrad_umma(x, dy, dw_ws, N, H, W, Cin, Cout, R, S, st);
    }
    if (engine == STC_ENGINE_AUTO && elig) {
        g_last_engine = STC_ENGINE_TCGEN05;
        return conv_wgrad_umma(x, dy, dw_ws, N, H, W, Cin, Cout, R, S, st);
    }
    g_last_engine = STC_ENGINE_SIMT;
    return conv_wgrad_simt(x, dy, dw_ws, N, H, W, Cin, Cout, R, S, dtype, st);
}

extern "C" int stc_gemm(const void* A, const void* B, void* C, const stc_gemm_desc* d, int dtype, int engine, void* stream) {
    STC_REQUIRE(d && d->M > 0 && d->N > 0 && d->K > 0 && d->batch1 > 0 && d->batch2 > 0, "gemm: bad descriptor");
    cudaStream_t st = (cudaStream_t)stream;
    bool elig = gemm_umma_eligible(d, dtype);
    if (engine == STC_ENGINE_TCGEN05) {
        STC_REQUIRE(elig, "gemm: tcgen05 engine requested but descriptor/dtype not eligible");
        g_last_engine = STC_ENGINE_TCGEN05;
        return gemm_umma(A, B, C, d, dtype, st);
    }
    if (engine == STC_ENGINE_AUTO && elig) {
        g_last_engine = STC_ENGINE_TCGEN05;
        return gemm_umma(A, B, C, d, dtype, st);
    }
    g_last_engine = STC_ENGINE_SIMT;
    return gemm_simt(A, B, C, d, dtype, st);
}

/* bf16 operands, fp32 result: tcgen05 engine only (small weight-gradient products of a folded Linear pair). */
extern "C" int stc_gemm_f32out(const void* A, const void* B, float* C, const stc_gemm_desc* d, void* stream) {
    STC_REQUIRE(d && d->M > 0 && d->N > 0 && d->K > 0 && d->batch1 > 0 && d->batch2 > 0, "gemm_f32out: bad descriptor");
    STC_REQUIRE(gemm_umma_eligible(d, STC_BF16), "gemm_f32out: descriptor not eligible for the tcgen05 engine");
    STC_REQUIRE(((uintptr_t)C & 15) == 0 && d->sCm % 4 == 0, "gemm_f32out: C must be 16-byte aligned with a row stride multiple of 4");
    g_last_engine = STC_ENGINE_TCGEN05;
    return gemm_umma(A, B, C, d, STC_F32, (cudaStream_t)stream);
}
