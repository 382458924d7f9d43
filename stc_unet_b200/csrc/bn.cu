// BatchNorm / SyncBatchNorm building blocks (K5/K6; C1/C2 exchange happens between reduce and finalize/apply).
#include "reduce.cuh"

namespace stc {

__device__ __forceinline__ float act_fwd(float z, int act) {
    if (act == STC_ACT_RELU) return fmaxf(z, 0.f);
    if (act == STC_ACT_HSWISH) return z * fminf(fmaxf(z + 3.f, 0.f), 6.f) * (1.f / 6.f);
    return z;
}
__device__ __forceinline__ float act_grad(float z, int act) {
    if (act == STC_ACT_RELU) return z > 0.f ? 1.f : 0.f;
    if (act == STC_ACT_HSWISH) {
        // d/dz [ z * relu6(z+3)/6 ] with hardtanh's open-interval gradient
        float r6 = fminf(fmaxf(z + 3.f, 0.f), 6.f);
        float inside = (z + 3.f > 0.f && z + 3.f < 6.f) ? 1.f : 0.f;
        return r6 * (1.f / 6.f) + z * inside * (1.f / 6.f);
    }
    return 1.f;
}

// ---------------------------------------------------------------- forward statistics
template <typename T>
__global__ void __launch_bounds__(256) bn_reduce_kernel(const T* __restrict__ y, float* __restrict__ partial, long long P, int C) {
    __shared__ float smem[256 * 16];
    const int lanes = C >> 3, lv = threadIdx.x % lanes, r0 = threadIdx.x / lanes, rstep = 256 / lanes;
    // Shifted sums: accumulate (y - s) and (y - s)^2 with s = y[0][c] (the same shift in every block), so the fp32
    // partials do not suffer the E[y^2] - mean^2 cancellation when |mean| >> std; bn_unshift_kernel undoes the
    // shift in fp64.
    Vec8<T> sh;
    sh.load(y + lv * 8);
    float acc[2][8] = {};
    for (long long p = (long long)blockIdx.x * rstep + r0; p < P; p += (long long)gridDim.x * rstep) {
        Vec8<T> v;
        v.load(y + p * C + lv * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float d = v.v[k] - sh.v[k];
            acc[0][k] += d;
            acc[1][k] = fmaf(d, d, acc[1][k]);
        }
    }
    block_reduce_lanes<2>(acc, lanes, lv, smem, partial, C);
}

// sums[c] = sum y, sums[C+c] = sum y^2 in fp64 from the shifted fp32 block partials
template <typename T>
__global__ void bn_unshift_kernel(const float* __restrict__ partial, const T* __restrict__ y, double* __restrict__ sums, int G,
                                  long long P, int C) {
    // one warp per channel: lanes stride over the G block partials
    int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= C) return;
    double s1 = 0.0, s2 = 0.0;
    // the G partial rows are read with 8 independent loads in flight per lane (this tiny kernel is pure load latency otherwise)
    int g = lane;
    for (; g + 7 * 32 < G; g += 8 * 32) {
        float a[8], b[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a[u] = partial[(size_t)(g + u * 32) * 2 * C + c];
            b[u] = partial[(size_t)(g + u * 32) * 2 * C + C + c];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { s1 += (double)a[u]; s2 += (double)b[u]; }
    }
    for (; g < G; g += 32) {
        s1 += (double)partial[(size_t)g * 2 * C + c];
        s2 += (double)partial[(size_t)g * 2 * C + C + c];
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane) return;
    const double s = (double)ldf(y + c), n = (double)P;
    sums[c] = s1 + n * s;
    sums[C + c] = s2 + 2.0 * s * s1 + n * s * s;
}

// generic fallback (any C): one block column of 32 channels, fp64 atomics
template <typename T>
__global__ void bn_reduce_generic_kernel(const T* __restrict__ y, double* __restrict__ sums, long long P, int C) {
    int c = blockIdx.x * 32 + (threadIdx.x & 31);
    int ry = threadIdx.x >> 5;
    double s = 0, s2 = 0;
    if (c < C)
        for (long long p = (long long)blockIdx.y * 8 + ry; p < P; p += (long long)gridDim.y * 8) {
            float v = ldf(y + p * C + c);
            s += v;
            s2 += (double)v * v;
        }
    if (c < C) {
        atomicAdd(sums + c, s);
        atomicAdd(sums + C + c, s2);
    }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, float* __restrict__ mean,
                                   float* __restrict__ invstd, float* __restrict__ rm, float* __restrict__ rv,
                                   long long* __restrict__ nbt, float momentum, float eps, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double m = sums[c] / count;
    double var = sums[C + c] / count - m * m;
    if (var < 0) var = 0;
    mean[c] = (float)m;
    invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (rm) {
        double unb = count > 1 ? var * (count / (count - 1.0)) : var;
        rm[c] = (1.f - momentum) * rm[c] + momentum * (float)m;
        rv[c] = (1.f - momentum) * rv[c] + momentum * (float)unb;
    }
    if (nbt && c == 0) *nbt += 1;
}

__global__ void bn_eval_stats_kernel(const float* __restrict__ rm, const float* __restrict__ rv, float* __restrict__ mean,
                                     float* __restrict__ invstd, float eps, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    mean[c] = rm[c];
    invstd[c] = rsqrtf(rv[c] + eps);
}

// ---------------------------------------------------------------- apply
template <typename T>
__global__ void __launch_bounds__(256) bn_apply_kernel(const T* __restrict__ y, const float* __restrict__ mean,
                                                       const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, T* __restrict__ a, long long nvec, int lanes,
                                                       int act) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < nvec; i += stride) {
        int c0 = (int)(i % lanes) * 8;
        Vec8<T> v;
        v.load(y + i * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float sc = gamma[c0 + k] * invstd[c0 + k];
            float z = fmaf(v.v[k] - mean[c0 + k], sc, beta[c0 + k]);
            v.v[k] = act_fwd(z, act);
        }
        v.store(a + i * 8);
    }
}

template <typename T>
__global__ void bn_apply_generic_kernel(const T* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ invstd,
                                        const float* __restrict__ gamma, const float* __restrict__ beta, T* __restrict__ a,
                                        long long total, int C, int act) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int c = (int)(i % C);
        float z = fmaf(ldf(y + i) - mean[c], gamma[c] * invstd[c], beta[c]);
        stf(a + i, act_fwd(z, act));
    }
}

// ---------------------------------------------------------------- backward
// g = dout * act'(z) with the activation resolved at compile time for ReLU (every DoubleConv / KSA branch): a select, no multiply
template <int ACT>
__device__ __forceinline__ float act_mask(float d, float z, int act) {
    if (ACT == STC_ACT_RELU) return z > 0.f ? d : 0.f;
    return d * act_grad(z, act);
}

// ACT = STC_ACT_RELU: specialised; ACT = -1: run-time `act`.  Per element: z = fma(v, sc, sh), xh = fma(v, is, -mu*is), g, two sums.
// AFF: the upstream gradient is given implicitly as  d' = ua[n][c] * dout + ub[n][c] * ub_scale  per image n (blockIdx.y) of
// P = rows-per-image rows — KernelSelectAttention's branch gradients df_k = w_k[n,c] * dout + dS[n,c]/HW are never materialised.
template <typename T, int ACT, bool AFF>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const T* __restrict__ y, const T* __restrict__ dout,
                                                            const float* __restrict__ mean, const float* __restrict__ invstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ partial, long long P, int C, int act,
                                                            const float* __restrict__ ua, const float* __restrict__ ub, float ub_scale) {
    __shared__ float smem[256 * 16];
    const int lanes = C >> 3, lv = threadIdx.x % lanes, r0 = threadIdx.x / lanes, rstep = 256 / lanes;
    constexpr bool kExact = sizeof(T) == 4;   // fp32 storage: subtract the mean first (see bn_bwd_apply_rows_kernel)
    float is[8], nm[8], sc[8], sh[8];
    float ua_r[8], ub_r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float mu = mean[lv * 8 + k];
        is[k] = invstd[lv * 8 + k];
        nm[k] = kExact ? mu : -mu * is[k];
        sc[k] = gamma[lv * 8 + k] * is[k];
        sh[k] = kExact ? beta[lv * 8 + k] : beta[lv * 8 + k] - mu * sc[k];
        ua_r[k] = AFF ? ua[(long long)blockIdx.y * C + lv * 8 + k] : 1.f;
        ub_r[k] = AFF ? ub[(long long)blockIdx.y * C + lv * 8 + k] * ub_scale : 0.f;
    }
    if (AFF) {
        y += (long long)blockIdx.y * P * C;
        dout += (long long)blockIdx.y * P * C;
    }
    auto up = [&](float d, int k) -> float { return AFF ? fmaf(ua_r[k], d, ub_r[k]) : d; };
    auto zx = [&](float v, int k, float& z, float& xh) {
        if (kExact) {
            const float vc = v - nm[k];
            z = fmaf(vc, sc[k], sh[k]);
            xh = vc * is[k];
        } else {
            z = fmaf(v, sc[k], sh[k]);
            xh = fmaf(v, is[k], nm[k]);
        }
    };
    float acc[2][8] = {};
    // two rows per iteration: four independent 16/32-byte loads in flight per thread
    const long long step = (long long)gridDim.x * rstep;
    long long p = (long long)blockIdx.x * rstep + r0;
    const T* yp = y + p * C + lv * 8;
    const T* dp = dout + p * C + lv * 8;
    const long long bump = step * C;
    for (; p + step < P; p += 2 * step, yp += 2 * bump, dp += 2 * bump) {
        Vec8<T> v0, d0, v1, d1;
        v0.load(yp);
        d0.load(dp);
        v1.load(yp + bump);
        d1.load(dp + bump);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float z0, x0, z1, x1;
            zx(v0.v[k], k, z0, x0);
            zx(v1.v[k], k, z1, x1);
            const float g0 = act_mask<ACT>(up(d0.v[k], k), z0, act), g1 = act_mask<ACT>(up(d1.v[k], k), z1, act);
            acc[0][k] += g0 + g1;
            acc[1][k] = fmaf(g0, x0, fmaf(g1, x1, acc[1][k]));
        }
    }
    for (; p < P; p += step, yp += bump, dp += bump) {
        Vec8<T> v, d;
        v.load(yp);
        d.load(dp);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float z, xh;
            zx(v.v[k], k, z, xh);
            const float g = act_mask<ACT>(up(d.v[k], k), z, act);
            acc[0][k] += g;
            acc[1][k] = fmaf(g, xh, acc[1][k]);
        }
    }
    const size_t block_row = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    block_reduce_lanes_emit<2>(acc, lanes, smem, C, [&](int q, int c, float sv) { partial[(block_row * 2 + q) * C + c] = sv; });
}

template <typename T>
__global__ void bn_bwd_reduce_generic_kernel(const T* __restrict__ y, const T* __restrict__ dout, const float* __restrict__ mean,
                                             const float* __restrict__ invstd, const float* __restrict__ gamma,
                                             const float* __restrict__ beta, double* __restrict__ sums, long long P, int C,
                                             int act) {
    int c = blockIdx.x * 32 + (threadIdx.x & 31);
    int ry = threadIdx.x >> 5;
    double s = 0, s2 = 0;
    if (c < C)
        for (long long p = (long long)blockIdx.y * 8 + ry; p < P; p += (long long)gridDim.y * 8) {
            float xh = (ldf(y + p * C + c) - mean[c]) * invstd[c];
            float g = ldf(dout + p * C + c) * act_grad(fmaf(xh, gamma[c], beta[c]), act);
            s += g;
            s2 += (double)g * xh;
        }
    if (c < C) {
        atomicAdd(sums + c, s);
        atomicAdd(sums + C + c, s2);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const T* __restrict__ y, const T* __restrict__ dout,
                                                           const float* __restrict__ mean, const float* __restrict__ invstd,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const double* __restrict__ sums, float inv_count, T* __restrict__ dy,
                                                           long long total, int C, int vec, int act, int eval) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (vec) {
        const int lanes = C >> 3;
        const long long nvec = total >> 3;
        for (; i < nvec; i += stride) {
            int c0 = (int)(i % lanes) * 8;
            Vec8<T> v, d;
            v.load(y + i * 8);
            d.load(dout + i * 8);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                int c = c0 + k;
                float xh = (v.v[k] - mean[c]) * invstd[c];
                float g = d.v[k] * act_grad(fmaf(xh, gamma[c], beta[c]), act);
                float r = eval ? g : g - (float)sums[c] * inv_count - xh * (float)sums[C + c] * inv_count;
                v.v[k] = gamma[c] * invstd[c] * r;
            }
            v.store(dy + i * 8);
        }
    } else {
        for (; i < total; i += stride) {
            int c = (int)(i % C);
            float xh = (ldf(y + i) - mean[c]) * invstd[c];
            float g = ldf(dout + i) * act_grad(fmaf(xh, gamma[c], beta[c]), act);
            float r = eval ? g : g - (float)sums[c] * inv_count - xh * (float)sums[C + c] * inv_count;
            stf(dy + i, gamma[c] * invstd[c] * r);
        }
    }
}

__global__ void bn_param_grads_kernel(const double* __restrict__ sums, float* __restrict__ dgamma, float* __restrict__ dbeta, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (dbeta) dbeta[c] = (float)sums[c];
    if (dgamma) dgamma[c] = (float)sums[C + c];
}


// ---------------------------------------------------------------- row-strided fast paths (C/8 a power of two <= 256)
// Each thread owns one 8-channel lane for the whole launch, so the per-channel constants live in registers and the
// inner loop is load -> fma -> store with no index arithmetic beyond a pointer bump.
template <typename T>
__global__ void __launch_bounds__(256) bn_apply_rows_kernel(const T* __restrict__ y, const float* __restrict__ mean,
                                                            const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, T* __restrict__ a, long long P, int C,
                                                            int act) {
    const int lanes = C >> 3, lv = threadIdx.x % lanes, r0 = threadIdx.x / lanes, rstep = 256 / lanes;
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        int c = lv * 8 + k;
        sc[k] = gamma[c] * invstd[c];
        sh[k] = beta[c] - mean[c] * sc[k];
    }
    const long long step = (long long)gridDim.x * rstep;
    for (long long p = (long long)blockIdx.x * rstep + r0; p < P; p += step) {
        Vec8<T> v;
        v.load(y + p * C + lv * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) v.v[k] = act_fwd(fmaf(v.v[k], sc[k], sh[k]), act);
        v.store(a + p * C + lv * 8);
    }
}

// dy = sc * (g - c1 - xh * c2) = sc * g + A * v + B  with A = -sc*c2*is, B = sc*(c2*is*mu - c1): z, select, two FMAs per element
template <typename T, int ACT, bool AFF>
__global__ void __launch_bounds__(256) bn_bwd_apply_rows_kernel(const T* __restrict__ y, const T* __restrict__ dout,
                                                                const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                const double* __restrict__ sums, float inv_count, T* __restrict__ dy,
                                                                long long P, int C, int act, int eval,
                                                                const float* __restrict__ ua, const float* __restrict__ ub, float ub_scale) {
    const int lanes = C >> 3, lv = threadIdx.x % lanes, r0 = threadIdx.x / lanes, rstep = 256 / lanes;
    float ua_r[8], ub_r[8];   // AFF: see bn_bwd_reduce_kernel (P = rows per image, blockIdx.y = image)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        ua_r[k] = AFF ? ua[(long long)blockIdx.y * C + lv * 8 + k] : 1.f;
        ub_r[k] = AFF ? ub[(long long)blockIdx.y * C + lv * 8 + k] * ub_scale : 0.f;
    }
    if (AFF) {
        y += (long long)blockIdx.y * P * C;
        dout += (long long)blockIdx.y * P * C;
        dy += (long long)blockIdx.y * P * C;
    }
    // fp32 storage keeps the subtract-first form (v - mu) * is: the folded A*v + B loses |mu|/std digits, which only bf16 storage hides
    constexpr bool kExact = sizeof(T) == 4;
    float sc[8], sh[8], A[8], B[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = lv * 8 + k;
        const float mu = mean[c], is = invstd[c];
        sc[k] = gamma[c] * is;
        sh[k] = beta[c] - mu * sc[k];
        const float c1 = eval ? 0.f : (float)sums[c] * inv_count;
        const float c2 = eval ? 0.f : (float)sums[C + c] * inv_count;
        if (kExact) {          // out = sc*g + (v - mu) * A + B
            A[k] = -sc[k] * c2 * is;
            B[k] = -sc[k] * c1;
            sh[k] = mu;        // reused as the mean; z is recomputed from (v - mu) below
        } else {
            A[k] = -sc[k] * c2 * is;
            B[k] = sc[k] * (c2 * is * mu - c1);
        }
    }
    const float* be = beta + lv * 8;
    float bek[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) bek[k] = kExact ? be[k] : 0.f;
    auto elem = [&](float v, float d, int k) -> float {
        if (AFF) d = fmaf(ua_r[k], d, ub_r[k]);
        if (kExact) {
            const float vc = v - sh[k];
            const float g = act_mask<ACT>(d, fmaf(vc, sc[k], bek[k]), act);
            return fmaf(sc[k], g, fmaf(A[k], vc, B[k]));
        }
        const float g = act_mask<ACT>(d, fmaf(v, sc[k], sh[k]), act);
        return fmaf(sc[k], g, fmaf(A[k], v, B[k]));
    };
    const long long step = (long long)gridDim.x * rstep;
    long long p = (long long)blockIdx.x * rstep + r0;
    const long long bump = step * C;
    const T* yp = y + p * C + lv * 8;
    const T* dp = dout + p * C + lv * 8;
    T* op = dy + p * C + lv * 8;
    for (; p + step < P; p += 2 * step, yp += 2 * bump, dp += 2 * bump, op += 2 * bump) {
        Vec8<T> v0, d0, v1, d1;
        v0.load(yp);
        d0.load(dp);
        v1.load(yp + bump);
        d1.load(dp + bump);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            v0.v[k] = elem(v0.v[k], d0.v[k], k);
            v1.v[k] = elem(v1.v[k], d1.v[k], k);
        }
        v0.store(op);
        v1.store(op + bump);
    }
    for (; p < P; p += step, yp += bump, dp += bump, op += bump) {
        Vec8<T> v, d;
        v.load(yp);
        d.load(dp);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            v.v[k] = elem(v.v[k], d.v[k], k);
        }
        v.store(op);
    }
}

inline int rows_blocks(long long P, int lanes) {
    long long rstep = 256 / lanes;
    long long want = (P + rstep * 4 - 1) / (rstep * 4);
    long long cap = (long long)num_sms() * 8;
    long long g = want < cap ? want : cap;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace stc

using namespace stc;

namespace stc {
__global__ void bn_fold_conv_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                                    float* __restrict__ wf, float* __restrict__ bf, int Cout, long long per_out) {
    const long long total = (long long)Cout * per_out;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i / per_out);
        const float s = gamma[co] * rsqrtf(rv[co] + eps);
        wf[i] = w[i] * s;
        if (i - co * per_out == 0) bf[co] = fmaf((b ? b[co] : 0.f) - rm[co], s, beta[co]);
    }
}
}  // namespace stc

/* Inference: Conv2d followed by an eval-mode BatchNorm is ONE conv with w' = w * s (per output channel), b' = (b - running_mean) * s + beta,
 * s = gamma / sqrt(running_var + eps); the activation goes into the conv epilogue, so the BN-apply pass over the feature map disappears. */
extern "C" int stc_bn_fold_conv(const float* w, const float* b, const float* gamma, const float* beta, const float* running_mean,
                                const float* running_var, float eps, float* w_folded, float* b_folded, int Cout, long long per_out, void* stream) {
    STC_REQUIRE(w && gamma && beta && running_mean && running_var && w_folded && b_folded && Cout > 0 && per_out > 0, "bn_fold_conv: bad arguments");
    const long long total = (long long)Cout * per_out;
    int blocks = (int)min((long long)num_sms() * 8, (long long)ceil_div(total, 256));
    stc::bn_fold_conv_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, b, gamma, beta, running_mean, running_var, eps, w_folded, b_folded, Cout, per_out);
    return check_launch("bn_fold_conv");
}

namespace stc {
// sums partial[g][j] over g in fp64 (shared with the conv epilogue statistics, api_dense.cu)
int reduce_partials_f64(const float* partial, double* out, int G, int len, cudaStream_t st) {
    reduce_partials_kernel<<<ceil_div((long long)len * 32, 128), 128, 0, st>>>(partial, out, G, len);
    return check_launch("reduce_partials");
}
}  // namespace stc

extern "C" long long stc_bn_ws_bytes(long long P, int C) {
    (void)P;
    return (long long)num_sms() * 4 * 3 * (long long)C * sizeof(float) + 256;
}

extern "C" int stc_bn_reduce(const void* y, double* sums, long long P, int C, void* ws, long long ws_bytes, int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    STC_REQUIRE(P > 0 && C > 0, "bn_reduce: bad shape");
    if (vec_ok(C) && (((uintptr_t)y) & 15) == 0) {
        int lanes = C / 8, G = reduce_blocks(P, lanes);
        STC_REQUIRE(ws && ws_bytes >= (long long)G * 2 * C * (long long)sizeof(float), "bn_reduce: workspace too small");
        STC_DISPATCH_DTYPE(dtype, (bn_reduce_kernel<T><<<G, 256, 0, st>>>((const T*)y, (float*)ws, P, C)));
        STC_DISPATCH_DTYPE(dtype, (bn_unshift_kernel<T><<<ceil_div((long long)C * 32, 128), 128, 0, st>>>((const float*)ws, (const T*)y, sums, G, P, C)));
    } else {
        STC_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
        dim3 grid(ceil_div(C, 32), (unsigned)max(1LL, min((long long)num_sms() * 2, (P + 63) / 64)));
        STC_DISPATCH_DTYPE(dtype, (bn_reduce_generic_kernel<T><<<grid, 256, 0, st>>>((const T*)y, sums, P, C)));
    }
    return check_launch("bn_reduce");
}

extern "C" int stc_bn_finalize(const double* sums, double count, float* mean, float* invstd, float* running_mean,
                               float* running_var, int64_t* num_batches_tracked, float momentum, float eps, int C, void* stream) {
    STC_REQUIRE(count > 0 && C > 0, "bn_finalize: bad count");
    bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(sums, count, mean, invstd, running_mean, running_var,
                                                                           (long long*)num_batches_tracked, momentum, eps, C);
    return check_launch("bn_finalize");
}

extern "C" int stc_bn_eval_stats(const float* running_mean, const float* running_var, float* mean, float* invstd, float eps,
                                 int C, void* stream) {
    bn_eval_stats_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(running_mean, running_var, mean, invstd, eps, C);
    return check_launch("bn_eval_stats");
}

extern "C" int stc_bn_apply(const void* y, const float* mean, const float* invstd, const float* gamma, const float* beta, void* a,
                            long long P, int C, int act, int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    long long total = P * C;
    if (total <= 0) return STC_OK;
    bool vec = C % 8 == 0 && ((((uintptr_t)y) | ((uintptr_t)a)) & 15) == 0;
    if (vec && vec_ok(C)) {
        STC_DISPATCH_DTYPE(dtype, (bn_apply_rows_kernel<T><<<rows_blocks(P, C / 8), 256, 0, st>>>((const T*)y, mean, invstd, gamma, beta,
                                                                                                (T*)a, P, C, act)));
    } else if (vec) {
        long long nvec = total / 8;
        int blocks = (int)min((long long)num_sms() * 8, (long long)ceil_div(nvec, 256));
        STC_DISPATCH_DTYPE(dtype, (bn_apply_kernel<T><<<blocks, 256, 0, st>>>((const T*)y, mean, invstd, gamma, beta, (T*)a, nvec, C / 8, act)));
    } else {
        int blocks = (int)min((long long)num_sms() * 8, (long long)ceil_div(total, 256));
        STC_DISPATCH_DTYPE(dtype, (bn_apply_generic_kernel<T><<<blocks, 256, 0, st>>>((const T*)y, mean, invstd, gamma, beta, (T*)a, total, C, act)));
    }
    return check_launch("bn_apply");
}

extern "C" int stc_bn_bwd_reduce(const void* y, const void* dout, const float* mean, const float* invstd, const float* gamma,
                                 const float* beta, double* sums, long long P, int C, int act, void* ws, long long ws_bytes,
                                 int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    STC_REQUIRE(P > 0 && C > 0, "bn_bwd_reduce: bad shape");
    if (vec_ok(C) && ((((uintptr_t)y) | ((uintptr_t)dout)) & 15) == 0) {
        int lanes = C / 8, G = reduce_blocks(P, lanes);
        STC_REQUIRE(ws && ws_bytes >= (long long)G * 2 * C * (long long)sizeof(float), "bn_bwd_reduce: workspace too small");
        if (act == STC_ACT_RELU) {
            STC_DISPATCH_DTYPE(dtype, (bn_bwd_reduce_kernel<T, STC_ACT_RELU, false><<<G, 256, 0, st>>>((const T*)y, (const T*)dout, mean, invstd, gamma,
                                                                                                      beta, (float*)ws, P, C, act, nullptr, nullptr, 0.f)));
        } else {
            STC_DISPATCH_DTYPE(dtype, (bn_bwd_reduce_kernel<T, -1, false><<<G, 256, 0, st>>>((const T*)y, (const T*)dout, mean, invstd, gamma, beta,
                                                                                            (float*)ws, P, C, act, nullptr, nullptr, 0.f)));
        }
        reduce_partials_kernel<<<ceil_div((long long)2 * C * 32, 128), 128, 0, st>>>((const float*)ws, sums, G, 2 * C);
    } else {
        STC_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
        dim3 grid(ceil_div(C, 32), (unsigned)max(1LL, min((long long)num_sms() * 2, (P + 63) / 64)));
        STC_DISPATCH_DTYPE(dtype, (bn_bwd_reduce_generic_kernel<T><<<grid, 256, 0, st>>>((const T*)y, (const T*)dout, mean, invstd,
                                                                                        gamma, beta, sums, P, C, act)));
    }
    return check_launch("bn_bwd_reduce");
}

extern "C" int stc_bn_bwd_apply(const void* y, const void* dout, const float* mean, const float* invstd, const float* gamma,
                                const float* beta, const double* sums, double count, void* dy,
                                long long P, int C, int act, int eval, int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    long long total = P * C;
    if (total <= 0) return STC_OK;
    int vec = (C % 8 == 0 && ((((uintptr_t)y) | ((uintptr_t)dout) | ((uintptr_t)dy)) & 15) == 0) ? 1 : 0;
    if (vec && vec_ok(C)) {
        if (act == STC_ACT_RELU) {
            STC_DISPATCH_DTYPE(dtype, (bn_bwd_apply_rows_kernel<T, STC_ACT_RELU, false><<<rows_blocks(P, C / 8), 256, 0, st>>>(
                                          (const T*)y, (const T*)dout, mean, invstd, gamma, beta, sums, (float)(1.0 / count), (T*)dy, P, C, act, eval,
                                          nullptr, nullptr, 0.f)));
        } else {
            STC_DISPATCH_DTYPE(dtype, (bn_bwd_apply_rows_kernel<T, -1, false><<<rows_blocks(P, C / 8), 256, 0, st>>>(
                                          (const T*)y, (const T*)dout, mean, invstd, gamma, beta, sums, (float)(1.0 / count), (T*)dy, P, C, act, eval,
                                          nullptr, nullptr, 0.f)));
        }
        return check_launch("bn_bwd_apply");
    }
    long long work = vec ? total / 8 : total;
    int blocks = (int)min((long long)num_sms() * 8, (long long)ceil_div(work, 256));
    STC_DISPATCH_DTYPE(dtype, (bn_bwd_apply_kernel<T><<<blocks, 256, 0, st>>>((const T*)y, (const T*)dout, mean, invstd, gamma, beta,
                                                                            sums, (float)(1.0 / count), (T*)dy, total, C, vec, act,
                                                                            eval)));
    return check_launch("bn_bwd_apply");
}

extern "C" int stc_bn_param_grads(const double* sums, float* dgamma, float* dbeta, int C, void* stream) {
    bn_param_grads_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(sums, dgamma, dbeta, C);
    return check_launch("bn_param_grads");
}

/* BN backward with an IMPLICIT upstream gradient d' = up_scale[n][c] * dout + up_shift[n][c] * shift_scale (per image n of
 * rows_per_image rows): KernelSelectAttention's branch gradients without the three df tensors.  ReLU, train mode, C/8 a power of two. */
extern "C" int stc_bn_bwd_aff_ok(int C) { return vec_ok(C) ? 1 : 0; }

extern "C" int stc_bn_bwd_reduce_aff(const void* y, const void* dout, const float* up_scale, const float* up_shift, float shift_scale,
                                     long long rows_per_image, int N, const float* mean, const float* invstd, const float* gamma,
                                     const float* beta, double* sums, int C, void* ws, long long ws_bytes, int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    STC_REQUIRE(rows_per_image > 0 && N > 0 && vec_ok(C) && ((((uintptr_t)y) | ((uintptr_t)dout)) & 15) == 0, "bn_bwd_reduce_aff: shape/alignment not supported");
    const int lanes = C / 8;
    int gx = reduce_blocks(rows_per_image * N, lanes) / N;
    if (gx < 1) gx = 1;
    const int G = gx * N;
    STC_REQUIRE(ws && ws_bytes >= (long long)G * 2 * C * (long long)sizeof(float), "bn_bwd_reduce_aff: workspace too small");
    dim3 grid(gx, N);
    STC_DISPATCH_DTYPE(dtype, (bn_bwd_reduce_kernel<T, STC_ACT_RELU, true><<<grid, 256, 0, st>>>((const T*)y, (const T*)dout, mean, invstd, gamma, beta,
                                                                                                (float*)ws, rows_per_image, C, STC_ACT_RELU,
                                                                                                up_scale, up_shift, shift_scale)));
    reduce_partials_kernel<<<ceil_div((long long)2 * C * 32, 128), 128, 0, st>>>((const float*)ws, sums, G, 2 * C);
    return check_launch("bn_bwd_reduce_aff");
}

extern "C" int stc_bn_bwd_apply_aff(const void* y, const void* dout, const float* up_scale, const float* up_shift, float shift_scale,
                                    long long rows_per_image, int N, const float* mean, const float* invstd, const float* gamma,
                                    const float* beta, const double* sums, double count, void* dy, int C, int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    STC_REQUIRE(rows_per_image > 0 && N > 0 && vec_ok(C) && ((((uintptr_t)y) | ((uintptr_t)dout) | ((uintptr_t)dy)) & 15) == 0,
                "bn_bwd_apply_aff: shape/alignment not supported");
    // whole waves: the grid is (at most) two times the number of blocks the chip holds at once (this variant needs more registers
    // than the plain one, so num_sms * 8 blocks would leave a two-thirds-empty third wave)
    int occ = 0;
    if (dtype == STC_BF16) {
        STC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bn_bwd_apply_rows_kernel<bf16, STC_ACT_RELU, true>, 256, 0));
    } else {
        STC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bn_bwd_apply_rows_kernel<float, STC_ACT_RELU, true>, 256, 0));
    }
    const long long want = (rows_per_image + (256 / (C / 8)) * 4 - 1) / ((256 / (C / 8)) * 4);   // >= 4 rows per thread
    long long gxl = (long long)occ * num_sms() * 2 / N;
    if (gxl > want) gxl = want;
    const int gx = gxl < 1 ? 1 : (int)gxl;
    dim3 grid(gx, N);
    STC_DISPATCH_DTYPE(dtype, (bn_bwd_apply_rows_kernel<T, STC_ACT_RELU, true><<<grid, 256, 0, st>>>(
                                  (const T*)y, (const T*)dout, mean, invstd, gamma, beta, sums, (float)(1.0 / count), (T*)dy, rows_per_image, C,
                                  STC_ACT_RELU, 0, up_scale, up_shift, shift_scale)));
    return check_launch("bn_bwd_apply_aff");
}
