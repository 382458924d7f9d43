// fp32-accumulate SIMT implicit-GEMM kernels: the exact-precision engine (fp32 parity path,
// odd shapes such as Cin=3 or the tiny CoordAtt/KSA layers) and the on-device cross-check for
// the tcgen05 engine.  64x64x16 tiles, 256 threads, 4x4 register micro-tiles.
#include "common.cuh"

namespace stc {

constexpr int TM = 64, TN = 64, TK = 16;

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == STC_ACT_RELU) return v > 0.f ? v : 0.f;
    if (act == STC_ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
    if (act == STC_ACT_HSWISH) return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);
    return v;
}

// Two-level accumulation: each 16-deep K tile is summed in fp32 FMAs from zero and the tile sums are added in fp64, so the
// rounding error does not grow with K (a 1024-channel or 8192-pixel contraction keeps ~1e-7 relative accuracy).  This engine
// is the exact-parity path; operands whose products share a large common component (BN-cancelled sums such as CoordAtt's
// conv1 weight gradient) amplify accumulation error ~100x, which is what the reference's cuDNN/cuBLAS tiling also limits.
#define STC_SIMT_COMPUTE()                                            \
    {                                                                 \
        float part[4][4] = {};                                        \
        _Pragma("unroll") for (int kk = 0; kk < TK; ++kk) {           \
            float a[4], b[4];                                         \
            _Pragma("unroll") for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i]; \
            _Pragma("unroll") for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j]; \
            _Pragma("unroll") for (int i = 0; i < 4; ++i)             \
                _Pragma("unroll") for (int j = 0; j < 4; ++j) part[i][j] = fmaf(a[i], b[j], part[i][j]); \
        }                                                             \
        _Pragma("unroll") for (int i = 0; i < 4; ++i)                 \
            _Pragma("unroll") for (int j = 0; j < 4; ++j) acc[i][j] += (double)part[i][j]; \
    }

// ---------------------------------------------------------------- conv fprop / dgrad
// y[p][co] = act(sum_{tap,ci} x[p + tap][ci] * wp[tap][co][ci] + bias[co] + residual[p][co])
template <typename T>
__global__ void __launch_bounds__(256) conv_fprop_simt_kernel(const T* __restrict__ x, const T* __restrict__ wp,
                                                              const float* __restrict__ bias, const T* __restrict__ residual,
                                                              T* __restrict__ y, int N, int H, int W, int Cin, int Cout,
                                                              int R, int S, int act) {
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long P = (long long)N * H * W;
    const long long m0 = (long long)blockIdx.x * TM;
    const int n0 = blockIdx.y * TN;
    const int K = R * S * Cin;
    const int pr = R / 2, ps = S / 2;
    // loader mapping: k = tid % 16, rows (tid / 16) + 16 j
    const int lk = tid & 15, lr = tid >> 4;
    int rn[4], rh[4], rw[4];
    bool rok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        long long m = m0 + lr + 16 * j;
        rok[j] = m < P;
        long long mm = rok[j] ? m : 0;
        rw[j] = (int)(mm % W);
        rh[j] = (int)((mm / W) % H);
        rn[j] = (int)(mm / ((long long)W * H));
    }
    double acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += TK) {
        int k = k0 + lk;
        bool kok = k < K;
        int tap = kok ? k / Cin : 0, ci = kok ? k % Cin : 0;
        int r = tap / S, s = tap % S;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v = 0.f;
            int hh = rh[j] + r - pr, ww = rw[j] + s - ps;
            if (kok && rok[j] && hh >= 0 && hh < H && ww >= 0 && ww < W)
                v = ldf(x + (((long long)rn[j] * H + hh) * W + ww) * Cin + ci);
            As[lk][lr + 16 * j] = v;
            int n = n0 + lr + 16 * j;
            float wv = 0.f;
            if (kok && n < Cout) wv = ldf(wp + ((long long)tap * Cout + n) * Cin + ci);
            Bs[lk][lr + 16 * j] = wv;
        }
        __syncthreads();
        STC_SIMT_COMPUTE();
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        long long m = m0 + ty * 4 + i;
        if (m >= P) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= Cout) continue;
            float v = (float)acc[i][j];
            if (bias) v += bias[n];
            if (residual) v += ldf(residual + m * Cout + n);
            stf(y + m * Cout + n, apply_act(v, act));
        }
    }
}

// ---------------------------------------------------------------- conv wgrad
// ws[(tap*Cin+ci)][co] += sum_p x[p + tap][ci] * dy[p][co]   (split over gridDim.z pixel ranges)
template <typename T>
__global__ void __launch_bounds__(256) conv_wgrad_simt_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                              float* __restrict__ ws, int N, int H, int W, int Cin,
                                                              int Cout, int R, int S, long long chunk) {
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long P = (long long)N * H * W;
    const int Mtot = R * S * Cin;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
    const long long p_begin = (long long)blockIdx.z * chunk, p_end = min(P, p_begin + chunk);
    const int pr = R / 2, ps = S / 2;
    // loader mapping: m/n = tid % 64 (contiguous channels), k = tid / 64 + 4 j
    const int lm = tid & 63, lkb = tid >> 6;
    const int m = m0 + lm;
    const bool mok = m < Mtot;
    const int tap = mok ? m / Cin : 0, ci = mok ? m % Cin : 0;
    const int r = tap / S - pr, s = tap % S - ps;
    const int n = n0 + lm;
    const bool nok = n < Cout;
    double acc[4][4] = {};
    for (long long k0 = p_begin; k0 < p_end; k0 += TK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int kk = lkb + 4 * j;
            long long p = k0 + kk;
            float av = 0.f, bv = 0.f;
            if (p < p_end) {
                int w_ = (int)(p % W), h_ = (int)((p / W) % H);
                long long n_ = p / ((long long)W * H);
                int hh = h_ + r, ww = w_ + s;
                if (mok && hh >= 0 && hh < H && ww >= 0 && ww < W) av = ldf(x + ((n_ * H + hh) * W + ww) * Cin + ci);
                if (nok) bv = ldf(dy + p * Cout + n);
            }
            As[kk][lm] = av;
            Bs[kk][lm] = bv;
        }
        __syncthreads();
        STC_SIMT_COMPUTE();
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int mm = m0 + ty * 4 + i;
        if (mm >= Mtot) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int nn = n0 + tx * 4 + j;
            if (nn < Cout) atomicAdd(ws + (long long)mm * Cout + nn, (float)acc[i][j]);
        }
    }
}

// ---------------------------------------------------------------- strided batched GEMM
template <typename T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const T* __restrict__ A, const T* __restrict__ B, T* __restrict__ C,
                                                        stc_gemm_desc d) {
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int b = blockIdx.z, b1 = b / d.batch2, b2 = b % d.batch2;
    A += b1 * d.sA1 + b2 * d.sA2;
    B += b1 * d.sB1 + b2 * d.sB2;
    C += b1 * d.sC1 + b2 * d.sC2;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
    const bool a_kfast = d.sAk == 1, b_kfast = d.sBk == 1;
    double acc[4][4] = {};
    for (int k0 = 0; k0 < d.K; k0 += TK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int am, ak, bn, bk;
            if (a_kfast) { ak = tid & 15; am = (tid >> 4) + 16 * j; } else { am = tid & 63; ak = (tid >> 6) + 4 * j; }
            if (b_kfast) { bk = tid & 15; bn = (tid >> 4) + 16 * j; } else { bn = tid & 63; bk = (tid >> 6) + 4 * j; }
            float av = 0.f, bv = 0.f;
            if (m0 + am < d.M && k0 + ak < d.K) av = ldf(A + (long long)(m0 + am) * d.sAm + (long long)(k0 + ak) * d.sAk);
            if (n0 + bn < d.N && k0 + bk < d.K) bv = ldf(B + (long long)(k0 + bk) * d.sBk + (long long)(n0 + bn) * d.sBn);
            As[ak][am] = av;
            Bs[bk][bn] = bv;
        }
        __syncthreads();
        STC_SIMT_COMPUTE();
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= d.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= d.N) continue;
            T* c = C + (long long)m * d.sCm + n;
            float v = d.alpha * (float)acc[i][j];
            if (d.beta != 0.f) v += d.beta * ldf(c);
            stf(c, v);
        }
    }
}

int conv_fprop_simt(const void* x, const void* wp, const float* bias, const void* residual, void* y, int N, int H, int W,
                    int Cin, int Cout, int R, int S, int act, int dtype, cudaStream_t st) {
    long long P = (long long)N * H * W;
    dim3 grid(ceil_div(P, TM), ceil_div(Cout, TN));
    STC_DISPATCH_DTYPE(dtype, (conv_fprop_simt_kernel<T><<<grid, 256, 0, st>>>((const T*)x, (const T*)wp, bias,
                                                                              (const T*)residual, (T*)y, N, H, W, Cin,
                                                                              Cout, R, S, act)));
    return check_launch("conv_fprop_simt");
}

int conv_wgrad_simt(const void* x, const void* dy, float* ws, int N, int H, int W, int Cin, int Cout, int R, int S,
                    int dtype, cudaStream_t st) {
    long long P = (long long)N * H * W;
    int gx = ceil_div(R * S * Cin, TM), gy = ceil_div(Cout, TN);
    // enough splits to fill the machine, at least 256 pixels per split
    long long want = max(1LL, (long long)num_sms() * 4 / ((long long)gx * gy));
    long long splits = max(1LL, min(want, (P + 255) / 256));
    long long chunk = ((P + splits - 1) / splits + TK - 1) / TK * TK;
    splits = (P + chunk - 1) / chunk;
    dim3 grid(gx, gy, (unsigned)splits);
    STC_DISPATCH_DTYPE(dtype, (conv_wgrad_simt_kernel<T><<<grid, 256, 0, st>>>((const T*)x, (const T*)dy, ws, N, H, W, Cin,
                                                                              Cout, R, S, chunk)));
    return check_launch("conv_wgrad_simt");
}

int gemm_simt(const void* A, const void* B, void* C, const stc_gemm_desc* d, int dtype, cudaStream_t st) {
    dim3 grid(ceil_div(d->M, TM), ceil_div(d->N, TN), d->batch1 * d->batch2);
    STC_DISPATCH_DTYPE(dtype, (gemm_simt_kernel<T><<<grid, 256, 0, st>>>((const T*)A, (const T*)B, (T*)C, *d)));
    return check_launch("gemm_simt");
}

}  // namespace stc
