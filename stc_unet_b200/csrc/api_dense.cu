// C-ABI entry points of the dense contractions: engine selection between the tcgen05 kernel
// (umma.cu) and the fp32-accumulate SIMT kernel (conv_simt.cu).
#include <stdlib.h>

#include "common.cuh"

namespace stc {
int conv_fprop_simt(const void*, const void*, const float*, const void*, void*, int, int, int, int, int, int, int, int, int,
                    cudaStream_t);
int conv_wgrad_simt(const void*, const void*, float*, int, int, int, int, int, int, int, int, cudaStream_t);
int gemm_simt(const void*, const void*, void*, const stc_gemm_desc*, int, cudaStream_t);
bool conv_umma_eligible(int Cin, int Cout, int dtype);
int conv_fprop_umma(const void*, const void*, const float*, const void*, void*, int, int, int, int, int, int, int, int,
                    cudaStream_t, const ChanCat* src = nullptr, const ChanCat* dst = nullptr);
int conv_wgrad_umma(const void*, const void*, float*, int, int, int, int, int, int, int, cudaStream_t, const ChanCat* src = nullptr);
bool gemm_umma_eligible(const stc_gemm_desc*, int dtype);
bool umma_staged_ok(int N);
bool conv_convh_eligible(int W, int Cin, int Cout, int R, int S, int dtype);
bool conv_wgradh_eligible(int W, int Cin, int Cout, int R, int S, int dtype);
int conv_wgrad_wgradh(const void*, const void*, float*, int, int, int, int, int, int, int, cudaStream_t, const ChanCat* src = nullptr);
int conv_fprop_convh(const void*, const void*, const float*, const void*, void*, int, int, int, int, int, int, int, int, cudaStream_t, float* stats = nullptr,
                     int* stats_rows = nullptr, const ChanCat* src = nullptr, const ChanCat* dst = nullptr);
bool conv_convh_stats_ok(int Cout, int R);
int gemm_umma(const void*, const void*, void*, const stc_gemm_desc*, int, cudaStream_t, const void* mul_residual = nullptr, const float* rowvec = nullptr,
              int row_mode = 0, float sm_scale = 0.f);
bool gemm_umma_rowsoftmax_ok(const stc_gemm_desc* d, const void* C, const void* P);
}  // namespace stc

using namespace stc;

static thread_local int g_last_engine = 0;
/* engine actually used by the last stc_conv_fprop / stc_conv_wgrad / stc_gemm call on this thread */
extern "C" int stc_dense_last_engine(void) { return g_last_engine; }

static long long* g_debug_profile = nullptr;
namespace stc { long long* debug_profile_buffer() { return g_debug_profile; } }
extern "C" int stc_debug_profile(void* counters) { g_debug_profile = static_cast<long long*>(counters); return STC_OK; }

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

// ---- virtual channel concat (K9: the torch.cat of Up.forward / UpConvBlock.forward / UNet++ decoder blocks is never written) ----
static int make_cat(ChanCat& c, const void* p0, const void* p1, const void* p2, const void* p3, const void* p4, int c0, int c1, int c2, int c3,
                    int c4, const char* what) {
    const void* ps[kMaxCat] = {p0, p1, p2, p3, p4};
    const int cs[kMaxCat] = {c0, c1, c2, c3, c4};
    c.n = 0;
    for (int i = 0; i < kMaxCat; ++i) {
        if (cs[i] == 0) {
            for (int k = i; k < kMaxCat; ++k) STC_REQUIRE(cs[k] == 0, "%s: channel counts must be packed to the front", what);
            break;
        }
        STC_REQUIRE(cs[i] > 0 && cs[i] % 64 == 0 && ps[i] && ((uintptr_t)ps[i] & 15) == 0,
                    "%s: part %d needs a 16-byte aligned pointer and a multiple of 64 channels (has %d)", what, i, cs[i]);
        c.ptr[c.n] = ps[i];
        c.c[c.n++] = cs[i];
    }
    STC_REQUIRE(c.n >= 1, "%s: no parts", what);
    return STC_OK;
}

extern "C" int stc_conv_cat_ok(int c0, int c1, int c2, int c3, int c4, int Cother, int dtype, int engine) {
    static int off = -1;
    if (off < 0) { const char* e = getenv("STC_VCAT"); off = (e && e[0] == '0') ? 1 : 0; }
    const int cs[kMaxCat] = {c0, c1, c2, c3, c4};
    int total = 0;
    for (int c : cs) {
        if (c < 0 || c % 64) return 0;
        total += c;
    }
    // the concatenated side is Cin of fprop / wgrad and Cout of dgrad; Cother is the conv's other channel count
    return (!off && dtype == STC_BF16 && engine != STC_ENGINE_SIMT && total > 0 && conv_umma_eligible(total, Cother, dtype) &&
            conv_umma_eligible(Cother, total, dtype) && Cother % 64 == 0) ? 1 : 0;
}

extern "C" int stc_conv_fprop_cat(const void* x0, const void* x1, const void* x2, const void* x3, const void* x4, int c0, int c1, int c2, int c3,
                                  int c4, const void* wp, const float* bias, void* y, int N, int H, int W, int Cout, int R, int S, int act,
                                  int dtype, int engine, void* stream) {
    ChanCat src;
    if (int rc = make_cat(src, x0, x1, x2, x3, x4, c0, c1, c2, c3, c4, "conv_fprop_cat")) return rc;
    const int Cin = src.total();
    STC_REQUIRE(N > 0 && H > 0 && W > 0 && Cout > 0 && (R & 1) && (S & 1), "conv_fprop_cat: bad shape");
    STC_REQUIRE(dtype == STC_BF16 && engine != STC_ENGINE_SIMT && conv_umma_eligible(Cin, Cout, dtype),
                "conv_fprop_cat: tcgen05 engine only (bf16, Cin=%d, Cout=%d); materialise the concat otherwise", Cin, Cout);
    cudaStream_t st = (cudaStream_t)stream;
    if (conv_convh_eligible(W, Cin, Cout, R, S, dtype)) {
        g_last_engine = STC_KERNEL_CONVH;
        return conv_fprop_convh(nullptr, wp, bias, nullptr, y, N, H, W, Cin, Cout, R, S, act, st, nullptr, nullptr, &src, nullptr);
    }
    g_last_engine = STC_ENGINE_TCGEN05;
    return conv_fprop_umma(nullptr, wp, bias, nullptr, y, N, H, W, Cin, Cout, R, S, act, st, &src, nullptr);
}

extern "C" int stc_conv_dgrad_split(const void* dy, const void* wpt, void* dx0, void* dx1, void* dx2, void* dx3, void* dx4, int c0, int c1, int c2,
                                    int c3, int c4, int N, int H, int W, int Cdy, int R, int S, int dtype, int engine, void* stream) {
    ChanCat dst;
    if (int rc = make_cat(dst, dx0, dx1, dx2, dx3, dx4, c0, c1, c2, c3, c4, "conv_dgrad_split")) return rc;
    const int Cdx = dst.total();
    STC_REQUIRE(N > 0 && H > 0 && W > 0 && Cdy > 0 && (R & 1) && (S & 1), "conv_dgrad_split: bad shape");
    STC_REQUIRE(dtype == STC_BF16 && engine != STC_ENGINE_SIMT && conv_umma_eligible(Cdy, Cdx, dtype),
                "conv_dgrad_split: tcgen05 engine only (bf16, Cdy=%d, Cdx=%d)", Cdy, Cdx);
    cudaStream_t st = (cudaStream_t)stream;
    if (conv_convh_eligible(W, Cdy, Cdx, R, S, dtype)) {
        g_last_engine = STC_KERNEL_CONVH;
        return conv_fprop_convh(dy, wpt, nullptr, nullptr, nullptr, N, H, W, Cdy, Cdx, R, S, STC_ACT_NONE, st, nullptr, nullptr, nullptr, &dst);
    }
    g_last_engine = STC_ENGINE_TCGEN05;
    return conv_fprop_umma(dy, wpt, nullptr, nullptr, nullptr, N, H, W, Cdy, Cdx, R, S, STC_ACT_NONE, st, nullptr, &dst);
}

extern "C" int stc_conv_wgrad_cat(const void* x0, const void* x1, const void* x2, const void* x3, const void* x4, int c0, int c1, int c2, int c3,
                                  int c4, const void* dy, float* dw_ws, int N, int H, int W, int Cout, int R, int S, int dtype, int engine,
                                  void* stream) {
    ChanCat src;
    if (int rc = make_cat(src, x0, x1, x2, x3, x4, c0, c1, c2, c3, c4, "conv_wgrad_cat")) return rc;
    const int Cin = src.total();
    STC_REQUIRE(N > 0 && H > 0 && W > 0 && Cout > 0 && (R & 1) && (S & 1), "conv_wgrad_cat: bad shape");
    STC_REQUIRE(dtype == STC_BF16 && engine != STC_ENGINE_SIMT && Cout % 64 == 0, "conv_wgrad_cat: tcgen05 engine only (bf16, Cout=%d)", Cout);
    cudaStream_t st = (cudaStream_t)stream;
    if (conv_wgradh_eligible(W, Cin, Cout, R, S, dtype)) {
        g_last_engine = STC_KERNEL_WGRADH;
        return conv_wgrad_wgradh(nullptr, dy, dw_ws, N, H, W, Cin, Cout, R, S, st, &src);
    }
    g_last_engine = STC_ENGINE_TCGEN05;
    return conv_wgrad_umma(nullptr, dy, dw_ws, N, H, W, Cin, Cout, R, S, st, &src);
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

/* Row softmax inside the score product (K4, nn.MultiheadAttention forward / backward), two sweeps over a 128-row block's N tiles with the row
 * statistics kept in the epilogue threads (umma.cu, UmmaParams::epi_mode 2 / 3):
 *   stc_gemm_softmax:       P  = softmax_rows(scale * bf16(A B^T))                          - the L x L scores are never written
 *   stc_gemm_softmax_bwd:   dS = scale * P * (bf16(A B^T) - sum_j P dP / sum_j P)           - no dP tensor, no separate pass
 * Same arithmetic as stc_gemm + stc_softmax_rows_fwd / _bwd (products rounded to bf16 first, __expf, division by the actual row sum), up to
 * the order of the fp32 row sums.  tcgen05 engine, bf16, M % 128 == 0, N a multiple of its tile; stc_gemm_softmax_ok tells. */
extern "C" int stc_gemm_softmax_ok(const stc_gemm_desc* d, const void* C, const void* P, int dtype, int engine) {
    return (d && dtype == STC_BF16 && engine != STC_ENGINE_SIMT && gemm_umma_rowsoftmax_ok(d, C, P)) ? 1 : 0;
}
extern "C" int stc_gemm_softmax(const void* A, const void* B, void* P, const stc_gemm_desc* d, float scale, int dtype, int engine, void* stream) {
    STC_REQUIRE(d && A && B && P && stc_gemm_softmax_ok(d, P, nullptr, dtype, engine), "gemm_softmax: not eligible (see stc_gemm_softmax_ok)");
    g_last_engine = STC_ENGINE_TCGEN05;
    return gemm_umma(A, B, P, d, dtype, (cudaStream_t)stream, nullptr, nullptr, 2, scale);
}
extern "C" int stc_gemm_softmax_bwd(const void* A, const void* B, const void* P, void* dS, const stc_gemm_desc* d, float scale, int dtype,
                                    int engine, void* stream) {
    STC_REQUIRE(d && A && B && P && dS && stc_gemm_softmax_ok(d, dS, P, dtype, engine), "gemm_softmax_bwd: not eligible (see stc_gemm_softmax_ok)");
    g_last_engine = STC_ENGINE_TCGEN05;
    return gemm_umma(A, B, dS, d, dtype, (cudaStream_t)stream, P, nullptr, 3, scale);
}

/* Softmax backward inside the dP product (K4): C = alpha * P .* (A * B - D[row]) with P laid out like C (bf16) and D fp32 indexed
 * (batch1, batch2, m): dS = scale * P * (dO V^T - rowsum(dO * O)).  tcgen05 engine only (the caller keeps the separate softmax pass else). */
extern "C" int stc_gemm_dsoftmax(const void* A, const void* B, const void* P, const float* D, void* C, const stc_gemm_desc* d, int dtype,
                                 int engine, void* stream) {
    STC_REQUIRE(d && d->M > 0 && d->N > 0 && d->K > 0 && d->batch1 > 0 && d->batch2 > 0 && P && D, "gemm_dsoftmax: bad arguments");
    STC_REQUIRE(dtype == STC_BF16 && engine != STC_ENGINE_SIMT && gemm_umma_eligible(d, dtype), "gemm_dsoftmax: tcgen05 engine only (bf16, eligible strides)");
    g_last_engine = STC_ENGINE_TCGEN05;
    return gemm_umma(A, B, C, d, dtype, (cudaStream_t)stream, P, D);
}
extern "C" int stc_gemm_dsoftmax_ok(const stc_gemm_desc* d, int dtype, int engine) {
    // OPT-IN (STC_DSOFTMAX_FUSED=1).  Measured on the STC-UNet step (N=16, 512x512, same box A/B): 75.1 -> 74.5-75.0 ms, but D taken from the
    // bf16-rounded O does not cancel the tokens' common component exactly (the separate pass divides by the ACTUAL row sum of the stored
    // probabilities, so sum_j dS_ij = 0): the worst q / k gradient error of the flip-free 512x512 parity run rises from 2.6 to 11 (the
    // reference's own autocast path: 4.8), so the default keeps the separate softmax-backward pass.
    static int on = -1;
    if (on < 0) { const char* e = getenv("STC_DSOFTMAX_FUSED"); on = (e && e[0] == '1') ? 1 : 0; }
    const int off = !on;
    return (!off && d && dtype == STC_BF16 && engine != STC_ENGINE_SIMT && gemm_umma_eligible(d, dtype) && umma_staged_ok(d->N)) ? 1 : 0;
}

namespace stc {
// D[n, h, i] = sum_d a[n, i, h*hd + d] * b[n, i, h*hd + d]: one warp per (n, i, h), 8-element vectors, fp32 accumulation
template <typename T>
__global__ void __launch_bounds__(256) rowdot_heads_kernel(const T* __restrict__ a, const T* __restrict__ b, float* __restrict__ out, long long rows,
                                                           int L, int heads, int hd) {
    const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= rows * heads) return;
    const long long r = wid / heads;          // n * L + i
    const int h = (int)(wid - r * heads);
    const T* pa = a + (r * heads + h) * (long long)hd;
    const T* pb = b + (r * heads + h) * (long long)hd;
    float acc = 0.f;
    for (int c = lane * 8; c < hd; c += 256) {
        Vec8<T> x, y;
        x.load(pa + c);
        y.load(pb + c);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = fmaf(x.v[k], y.v[k], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        const long long n = r / L, i = r - n * L;
        out[(n * heads + h) * (long long)L + i] = acc;
    }
}
}  // namespace stc

extern "C" int stc_rowdot_heads(const void* a, const void* b, float* out, int N, int L, int heads, int hd, int dtype, void* stream) {
    STC_REQUIRE(a && b && out && N > 0 && L > 0 && heads > 0 && hd > 0 && hd % 8 == 0, "rowdot_heads: bad arguments (hd=%d must be a multiple of 8)", hd);
    const long long warps = (long long)N * L * heads;
    STC_DISPATCH_DTYPE(dtype, (rowdot_heads_kernel<T><<<ceil_div(warps * 32, 256), 256, 0, (cudaStream_t)stream>>>((const T*)a, (const T*)b, out,
                                                                                                                  (long long)N * L, L, heads, hd)));
    return check_launch("rowdot_heads");
}

/* bf16 operands, fp32 result: tcgen05 engine only (small weight-gradient products of a folded Linear pair). */
extern "C" int stc_gemm_f32out(const void* A, const void* B, float* C, const stc_gemm_desc* d, void* stream) {
    STC_REQUIRE(d && d->M > 0 && d->N > 0 && d->K > 0 && d->batch1 > 0 && d->batch2 > 0, "gemm_f32out: bad descriptor");
    STC_REQUIRE(gemm_umma_eligible(d, STC_BF16), "gemm_f32out: descriptor not eligible for the tcgen05 engine");
    STC_REQUIRE(((uintptr_t)C & 15) == 0 && d->sCm % 4 == 0, "gemm_f32out: C must be 16-byte aligned with a row stride multiple of 4");
    g_last_engine = STC_ENGINE_TCGEN05;
    return gemm_umma(A, B, C, d, STC_F32, (cudaStream_t)stream);
}

/* Conv2d + the BatchNorm batch statistics of its output in one pass: y = conv(x) + bias and sums = [sum_p y | sum_p y^2] (fp64, 2*Cout)
 * over all N*H*W pixels of the STORED outputs.  When the halo-reuse tcgen05 kernel takes the shape (and Cout is one N tile) the sums
 * come out of its epilogue; otherwise the conv is followed by stc_bn_reduce.  ws: stc_bn_ws_bytes(P, Cout) bytes of scratch. */
extern "C" int stc_bn_reduce(const void* y, double* sums, long long P, int C, void* ws, long long ws_bytes, int dtype, void* stream);
namespace stc { int reduce_partials_f64(const float* partial, double* out, int G, int len, cudaStream_t st); }
static bool bnstats_fused(int W, int Cin, int Cout, int R, int S, int dtype, int engine) {
    static int fused_off = -1;
    if (fused_off < 0) { const char* e = getenv("STC_BNSTATS_FUSED"); fused_off = (e && e[0] == '0') ? 1 : 0; }
    static int min_k = -1;   // the column sums double the epilogue work: they only hide behind the MMAs of a long-K tile (measured on the
    if (min_k < 0) { const char* e = getenv("STC_BNSTATS_MIN_K"); min_k = e ? atoi(e) : 1152; }   // 64-channel 3x3 layers: break-even)
    return !fused_off && engine != STC_ENGINE_SIMT && conv_umma_eligible(Cin, Cout, dtype) && conv_convh_eligible(W, Cin, Cout, R, S, dtype) &&
           conv_convh_stats_ok(Cout, R) && R * S * Cin >= min_k;
}
/* 1 if stc_conv_fprop_bnstats takes the statistics out of the conv epilogue for this shape (else it runs conv + stc_bn_reduce) */
extern "C" int stc_conv_bnstats_fused_ok(int W, int Cin, int Cout, int R, int S, int dtype, int engine) {
    return bnstats_fused(W, Cin, Cout, R, S, dtype, engine) ? 1 : 0;
}

extern "C" int stc_conv_fprop_bnstats(const void* x, const void* wp, const float* bias, void* y, int N, int H, int W, int Cin, int Cout, int R,
                                      int S, int dtype, int engine, double* sums, void* ws, long long ws_bytes, void* stream) {
    STC_REQUIRE(N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && (R & 1) && (S & 1) && sums && ws, "conv_fprop_bnstats: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const long long P = (long long)N * H * W;
    if (bnstats_fused(W, Cin, Cout, R, S, dtype, engine) && ws_bytes >= (long long)num_sms() * 4 * 2 * Cout * (long long)sizeof(float)) {
        g_last_engine = STC_KERNEL_CONVH;
        int rows = 0;
        int rc = conv_fprop_convh(x, wp, bias, nullptr, y, N, H, W, Cin, Cout, R, S, STC_ACT_NONE, st, (float*)ws, &rows);
        if (rc) return rc;
        return reduce_partials_f64((const float*)ws, sums, rows, 2 * Cout, st);
    }
    int rc = stc_conv_fprop(x, wp, bias, nullptr, y, N, H, W, Cin, Cout, R, S, STC_ACT_NONE, dtype, engine, stream);
    if (rc) return rc;
    return stc_bn_reduce(y, sums, P, Cout, ws, ws_bytes, dtype, stream);
}
