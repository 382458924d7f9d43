// Bandwidth-bound spatial kernels: MaxPool2d(2), bilinear x2 upsample fused with the skip concat,
// CoordAtt pooling / apply, KernelSelectAttention fuse.  All NHWC, 8-channel vectors, fp32 math.
#include "reduce.cuh"

namespace stc {

// ---------------------------------------------------------------- MaxPool2d(2)
template <typename T>
__global__ void maxpool2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int H, int W, int C, long long total) {
    const int lanes = C >> 3, OH = H >> 1, OW = W >> 1;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int lv = (int)(i % lanes);
        long long p = i / lanes;
        int ow = (int)(p % OW), oh = (int)((p / OW) % OH);
        long long n = p / ((long long)OW * OH);
        const T* b = x + (((n * H + 2 * oh) * W + 2 * ow) * (long long)C) + lv * 8;
        Vec8<T> a0, a1, a2, a3;
        a0.load(b); a1.load(b + C); a2.load(b + (long long)W * C); a3.load(b + (long long)W * C + C);
#pragma unroll
        for (int k = 0; k < 8; ++k) a0.v[k] = fmaxf(fmaxf(a0.v[k], a1.v[k]), fmaxf(a2.v[k], a3.v[k]));
        a0.store(y + i * 8);
    }
}

template <typename T>
__global__ void maxpool2_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, int H, int W, int C,
                                    long long total) {
    const int lanes = C >> 3, OH = H >> 1, OW = W >> 1;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int lv = (int)(i % lanes);
        long long p = i / lanes;
        int ow = (int)(p % OW), oh = (int)((p / OW) % OH);
        long long n = p / ((long long)OW * OH);
        long long base = (((n * H + 2 * oh) * W + 2 * ow) * (long long)C) + lv * 8;
        Vec8<T> a[4], g, o[4];
        a[0].load(x + base); a[1].load(x + base + C); a[2].load(x + base + (long long)W * C); a[3].load(x + base + (long long)W * C + C);
        g.load(dy + i * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            // first maximum in window scan order (PyTorch: `val > maxval`, NaN propagates)
            int best = 0;
            float m = a[0].v[k];
#pragma unroll
            for (int j = 1; j < 4; ++j)
                if (a[j].v[k] > m || a[j].v[k] != a[j].v[k]) { m = a[j].v[k]; best = j; }
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j].v[k] = (j == best) ? g.v[k] : 0.f;
        }
        o[0].store(dx + base); o[1].store(dx + base + C); o[2].store(dx + base + (long long)W * C); o[3].store(dx + base + (long long)W * C + C);
    }
}

// ---------------------------------------------------------------- bilinear x2 + pad + concat
struct Interp {
    int i0, i1;
    float l0, l1;
};
__device__ __forceinline__ Interp interp_src(int o, int in, int out, int align) {
    Interp r;
    float src;
    if (align) {
        float scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
        src = scale * (float)o;
    } else {
        src = 0.5f * ((float)o + 0.5f) - 0.5f;  // scale_factor 2 -> 1/2
        if (src < 0.f) src = 0.f;
    }
    r.i0 = (int)src;
    if (r.i0 > in - 1) r.i0 = in - 1;
    r.i1 = r.i0 + (r.i0 < in - 1 ? 1 : 0);
    r.l1 = src - (float)r.i0;
    r.l0 = 1.f - r.l1;
    return r;
}

template <typename T>
__global__ void upcat_fwd_kernel(const T* __restrict__ skip, const T* __restrict__ low, T* __restrict__ out, int H, int W, int Cs,
                                 int h, int w, int Cu, int align, long long total) {
    const int Ct = Cs + Cu, lanes = Ct >> 3, ls = Cs >> 3;
    const int py = (H - 2 * h) / 2, px = (W - 2 * w) / 2;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int lv = (int)(i % lanes);
        long long p = i / lanes;
        Vec8<T> v;
        if (lv < ls) {
            v.load(skip + p * Cs + lv * 8);
        } else {
            int ox = (int)(p % W) - px, oy = (int)((p / W) % H) - py;
            long long n = p / ((long long)W * H);
            if (ox < 0 || oy < 0 || ox >= 2 * w || oy >= 2 * h) {
#pragma unroll
                for (int k = 0; k < 8; ++k) v.v[k] = 0.f;
            } else {
                Interp iy = interp_src(oy, h, 2 * h, align), ix = interp_src(ox, w, 2 * w, align);
                const T* b = low + n * (long long)h * w * Cu + (lv - ls) * 8;
                Vec8<T> a00, a01, a10, a11;
                a00.load(b + ((long long)iy.i0 * w + ix.i0) * Cu);
                a01.load(b + ((long long)iy.i0 * w + ix.i1) * Cu);
                a10.load(b + ((long long)iy.i1 * w + ix.i0) * Cu);
                a11.load(b + ((long long)iy.i1 * w + ix.i1) * Cu);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    v.v[k] = iy.l0 * (ix.l0 * a00.v[k] + ix.l1 * a01.v[k]) + iy.l1 * (ix.l0 * a10.v[k] + ix.l1 * a11.v[k]);
            }
        }
        v.store(out + i * 8);
    }
}

template <typename T>
__global__ void upcat_bwd_skip_kernel(const T* __restrict__ dout, T* __restrict__ dskip, int Cs, int Ct, long long total) {
    const int ls = Cs >> 3;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int lv = (int)(i % ls);
        long long p = i / ls;
        Vec8<T> v;
        v.load(dout + p * Ct + lv * 8);
        v.store(dskip + i * 8);
    }
}

// gather form of the adjoint: each low-res element sums the (<= ~4x4) up-sampled pixels that read it
template <typename T>
__global__ void upcat_bwd_low_kernel(const T* __restrict__ dout, T* __restrict__ dlow, int H, int W, int Cs, int h, int w, int Cu,
                                     int align, long long total) {
    const int Ct = Cs + Cu, lu = Cu >> 3;
    const int py = (H - 2 * h) / 2, px = (W - 2 * w) / 2;
    const int OH = 2 * h, OW = 2 * w;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int lv = (int)(i % lu);
        long long p = i / lu;
        int jx = (int)(p % w), jy = (int)((p / w) % h);
        long long n = p / ((long long)w * h);
        // candidate output rows/cols: src index within (j-1, j+1)
        int y_lo, y_hi, x_lo, x_hi;
        if (align) {
            float sy = OH > 1 ? (float)(h - 1) / (float)(OH - 1) : 0.f, sx = OW > 1 ? (float)(w - 1) / (float)(OW - 1) : 0.f;
            y_lo = sy > 0.f ? max(0, (int)floorf((jy - 1) / sy) - 1) : 0;
            y_hi = sy > 0.f ? min(OH - 1, (int)ceilf((jy + 1) / sy) + 1) : OH - 1;
            x_lo = sx > 0.f ? max(0, (int)floorf((jx - 1) / sx) - 1) : 0;
            x_hi = sx > 0.f ? min(OW - 1, (int)ceilf((jx + 1) / sx) + 1) : OW - 1;
        } else {
            y_lo = max(0, 2 * jy - 2); y_hi = min(OH - 1, 2 * jy + 3);
            x_lo = max(0, 2 * jx - 2); x_hi = min(OW - 1, 2 * jx + 3);
        }
        float acc[8] = {};
        for (int oy = y_lo; oy <= y_hi; ++oy) {
            Interp iy = interp_src(oy, h, OH, align);
            float wy = (iy.i0 == jy ? iy.l0 : 0.f) + (iy.i1 == jy ? iy.l1 : 0.f);
            if (wy == 0.f) continue;
            for (int ox = x_lo; ox <= x_hi; ++ox) {
                Interp ix = interp_src(ox, w, OW, align);
                float wx = (ix.i0 == jx ? ix.l0 : 0.f) + (ix.i1 == jx ? ix.l1 : 0.f);
                if (wx == 0.f) continue;
                Vec8<T> g;
                g.load(dout + ((n * H + (oy + py)) * (long long)W + (ox + px)) * Ct + Cs + lv * 8);
                float ww = wy * wx;
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = fmaf(ww, g.v[k], acc[k]);
            }
        }
        Vec8<T> o;
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] = acc[k];
        o.store(dlow + i * 8);
    }
}

// ---------------------------------------------------------------- CoordAtt pieces
// acc += sum_t x[t*xs] (* w[t*ws]); four independent 16-byte loads in flight per thread (the loop is latency-bound otherwise)
template <typename T, bool WEIGHTED>
__device__ __forceinline__ void rowcol_accumulate(const T* __restrict__ x, long long xs, const T* __restrict__ w, long long ws, int cnt,
                                                  float (&acc)[8]) {
    float a1[8] = {}, a2[8] = {}, a3[8] = {};
    // unweighted (pooling) sums are taken relative to the first element and un-shifted at the end: see pool_skip_kernel
    Vec8<T> sh;
    if (!WEIGHTED) sh.load(x);
    int t = 0;
    for (; t + 3 < cnt; t += 4) {
        Vec8<T> v0, v1, v2, v3;
        v0.load(x + t * xs); v1.load(x + (t + 1) * xs); v2.load(x + (t + 2) * xs); v3.load(x + (t + 3) * xs);
        if (WEIGHTED) {
            Vec8<T> w0, w1, w2, w3;
            w0.load(w + t * ws); w1.load(w + (t + 1) * ws); w2.load(w + (t + 2) * ws); w3.load(w + (t + 3) * ws);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                acc[k] = fmaf(v0.v[k], w0.v[k], acc[k]); a1[k] = fmaf(v1.v[k], w1.v[k], a1[k]);
                a2[k] = fmaf(v2.v[k], w2.v[k], a2[k]); a3[k] = fmaf(v3.v[k], w3.v[k], a3[k]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                acc[k] += v0.v[k] - sh.v[k]; a1[k] += v1.v[k] - sh.v[k]; a2[k] += v2.v[k] - sh.v[k]; a3[k] += v3.v[k] - sh.v[k];
            }
        }
    }
    for (; t < cnt; ++t) {
        Vec8<T> v0;
        v0.load(x + t * xs);
        if (WEIGHTED) {
            Vec8<T> w0;
            w0.load(w + t * ws);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = fmaf(v0.v[k], w0.v[k], acc[k]);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += v0.v[k] - sh.v[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += a1[k] + a2[k] + a3[k];
    if (!WEIGHTED) {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(sh.v[k], (float)cnt, acc[k]);
    }
}


// row part: thread per (n,h,lane) loops over w;  col part: thread per (n,w,lane) loops over h
template <typename T, bool WEIGHTED>
__global__ void rowcol_reduce_kernel(const T* __restrict__ x, const T* __restrict__ a, T* __restrict__ y, int N, int H, int W, int C) {
    const int lanes = C >> 3;
    // image-major work order: the column pass of image n runs right after its row pass and finds the image still in L2
    long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long per_img = (long long)(H + W) * lanes;
    if (i0 >= per_img * N) return;
    const long long n_ = i0 / per_img;
    const long long rem = i0 - n_ * per_img;
    const long long n_row = (long long)N * H * lanes;
    // map back onto the original (row items first, then column items) numbering used below
    const long long i = rem < (long long)H * lanes ? n_ * H * lanes + rem : n_row + n_ * W * lanes + (rem - (long long)H * lanes);
    float acc[8] = {};
    if (i < n_row) {
        int lv = (int)(i % lanes);
        long long nh = i / lanes;
        long long n = nh / H;
        int hh = (int)(nh % H);
        const T* b = x + nh * (long long)W * C + lv * 8;
        const T* ab = WEIGHTED ? a + (n * (H + W) + H) * (long long)C + lv * 8 : nullptr;
        rowcol_accumulate<T, WEIGHTED>(b, C, ab, C, W, acc);
        Vec8<T> o;
        float sc = WEIGHTED ? 1.f : 1.f / (float)W;
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] = acc[k] * sc;
        o.store(y + (n * (H + W) + hh) * (long long)C + lv * 8);
    } else {
        long long j = i - n_row;
        int lv = (int)(j % lanes);
        long long nw = j / lanes;
        long long n = nw / W;
        int ww = (int)(nw % W);
        const T* b = x + (n * H * (long long)W + ww) * C + lv * 8;
        const T* ab = WEIGHTED ? a + (n * (H + W)) * (long long)C + lv * 8 : nullptr;
        rowcol_accumulate<T, WEIGHTED>(b, (long long)W * C, ab, C, H, acc);
        Vec8<T> o;
        float sc = WEIGHTED ? 1.f : 1.f / (float)H;
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] = acc[k] * sc;
        o.store(y + (n * (H + W) + H + ww) * (long long)C + lv * 8);
    }
}

// MODE 0: out = x + ah*aw ; MODE 1: out = x + dyh/W + dyw/H
template <typename T, int MODE>
__global__ void coordatt_ew_kernel(const T* __restrict__ x, const T* __restrict__ a, T* __restrict__ out, int H, int W, int C,
                                   long long total) {
    const int lanes = C >> 3;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float iw = 1.f / (float)W, ih = 1.f / (float)H;
    for (; i < total; i += stride) {
        int lv = (int)(i % lanes);
        long long p = i / lanes;
        int ww = (int)(p % W), hh = (int)((p / W) % H);
        long long n = p / ((long long)W * H);
        Vec8<T> v, ah, aw;
        v.load(x + i * 8);
        ah.load(a + (n * (H + W) + hh) * (long long)C + lv * 8);
        aw.load(a + (n * (H + W) + H + ww) * (long long)C + lv * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) v.v[k] += MODE == 0 ? ah.v[k] * aw.v[k] : ah.v[k] * iw + aw.v[k] * ih;
        v.store(out + i * 8);
    }
}

// ---------------------------------------------------------------- KernelSelectAttention fuse
// NQ = 1: S += scale * sum_hw (f0+f1+f2) ; NQ = 3: dw[k] += sum_hw dout * f_k      (grid.y = image)
template <typename T, int NQ>
__global__ void __launch_bounds__(256) ksa_reduce_kernel(const T* __restrict__ dout, const T* __restrict__ f0, const T* __restrict__ f1,
                                                         const T* __restrict__ f2, float* __restrict__ out, long long HW, int C, int N,
                                                         float scale) {
    __shared__ float smem[256 * 8 * NQ];
    const int lanes = C >> 3, lv = threadIdx.x % lanes, r0 = threadIdx.x / lanes, rstep = 256 / lanes;
    const long long n = blockIdx.y;
    const long long base = n * HW * C;
    float acc[NQ][8] = {};
    for (long long p = (long long)blockIdx.x * rstep + r0; p < HW; p += (long long)gridDim.x * rstep) {
        Vec8<T> a, b, c;
        long long o = base + p * C + lv * 8;
        a.load(f0 + o); b.load(f1 + o); c.load(f2 + o);
        if (NQ == 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[0][k] += a.v[k] + b.v[k] + c.v[k];
        } else {
            Vec8<T> g;
            g.load(dout + o);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                acc[0][k] = fmaf(g.v[k], a.v[k], acc[0][k]);
                acc[NQ > 1 ? 1 : 0][k] = fmaf(g.v[k], b.v[k], acc[NQ > 1 ? 1 : 0][k]);
                acc[NQ > 2 ? 2 : 0][k] = fmaf(g.v[k], c.v[k], acc[NQ > 2 ? 2 : 0][k]);
            }
        }
    }
    block_reduce_lanes_emit<NQ>(acc, lanes, smem, C, [&](int q, int c, float s) {
        atomicAdd(out + ((long long)q * N + n) * C + c, s * scale);
    });
}

template <typename T>
__global__ void ksa_combine_kernel(const T* __restrict__ x, const T* __restrict__ f0, const T* __restrict__ f1, const T* __restrict__ f2,
                                   const float* __restrict__ w, T* __restrict__ out, long long HW, int C, int N, long long total) {
    const int lanes = C >> 3;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int lv = (int)(i % lanes);
        long long n = (i / lanes) / HW;
        const float* wp = w + n * C + lv * 8;
        const long long ks = (long long)N * C;
        Vec8<T> v, a, b, c;
        v.load(x + i * 8); a.load(f0 + i * 8); b.load(f1 + i * 8); c.load(f2 + i * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) v.v[k] += wp[k] * a.v[k] + wp[ks + k] * b.v[k] + wp[2 * ks + k] * c.v[k];
        v.store(out + i * 8);
    }
}

template <typename T>
__global__ void ksa_df_kernel(const T* __restrict__ dout, const float* __restrict__ w, const float* __restrict__ dS, T* __restrict__ d0,
                              T* __restrict__ d1, T* __restrict__ d2, long long HW, int C, int N, long long total) {
    const int lanes = C >> 3;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float ihw = 1.f / (float)HW;
    for (; i < total; i += stride) {
        int lv = (int)(i % lanes);
        long long n = (i / lanes) / HW;
        const float* wp = w + n * C + lv * 8;
        const float* sp = dS + n * C + lv * 8;
        const long long ks = (long long)N * C;
        Vec8<T> g, a, b, c;
        g.load(dout + i * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float s = sp[k] * ihw;
            a.v[k] = fmaf(wp[k], g.v[k], s);
            b.v[k] = fmaf(wp[ks + k], g.v[k], s);
            c.v[k] = fmaf(wp[2 * ks + k], g.v[k], s);
        }
        a.store(d0 + i * 8); b.store(d1 + i * 8); c.store(d2 + i * 8);
    }
}

__global__ void softmax3_fwd_kernel(const float* __restrict__ a, float* __restrict__ w, long long NC) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= NC) return;
    float x0 = a[i], x1 = a[NC + i], x2 = a[2 * NC + i];
    float m = fmaxf(x0, fmaxf(x1, x2));
    float e0 = expf(x0 - m), e1 = expf(x1 - m), e2 = expf(x2 - m);
    float inv = 1.f / (e0 + e1 + e2);
    w[i] = e0 * inv; w[NC + i] = e1 * inv; w[2 * NC + i] = e2 * inv;
}
__global__ void softmax3_bwd_kernel(const float* __restrict__ w, const float* __restrict__ dw, float* __restrict__ da, long long NC) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= NC) return;
    float w0 = w[i], w1 = w[NC + i], w2 = w[2 * NC + i];
    float g0 = dw[i], g1 = dw[NC + i], g2 = dw[2 * NC + i];
    float dot = w0 * g0 + w1 * g1 + w2 * g2;
    da[i] = w0 * (g0 - dot); da[NC + i] = w1 * (g1 - dot); da[2 * NC + i] = w2 * (g2 - dot);
}

// ---------------------------------------------------------------- row softmax (MHA)
template <typename T> __device__ __forceinline__ float row_exp(float x);
template <> __device__ __forceinline__ float row_exp<float>(float x) { return expf(x); }     // fp32 parity path: full precision
template <> __device__ __forceinline__ float row_exp<bf16>(float x) { return __expf(x); }    // bf16 path: SFU ex2
template <typename T>
__global__ void __launch_bounds__(256) softmax_rows_fwd_kernel(const T* __restrict__ S, T* __restrict__ P, int L, float scale) {
    __shared__ float red[8];
    const T* s = S + (long long)blockIdx.x * L;
    T* p = P + (long long)blockIdx.x * L;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float m = -INFINITY;
    for (int j = tid; j < L; j += 256) m = fmaxf(m, ldf(s + j) * scale);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) m = fmaxf(m, red[k]);
    __syncthreads();
    float sum = 0.f;
    for (int j = tid; j < L; j += 256) sum += row_exp<T>(ldf(s + j) * scale - m);
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += red[k];
    const float inv = 1.f / sum;
    for (int j = tid; j < L; j += 256) stf(p + j, row_exp<T>(ldf(s + j) * scale - m) * inv);
}

template <typename T>
__global__ void __launch_bounds__(256) softmax_rows_bwd_kernel(const T* __restrict__ P, const T* __restrict__ dP, T* __restrict__ dS, int L,
                                                               float scale) {
    __shared__ float red[8];
    const T* p = P + (long long)blockIdx.x * L;
    const T* g = dP + (long long)blockIdx.x * L;
    T* o = dS + (long long)blockIdx.x * L;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // dot = sum(dP * P) / sum(P): dividing by the ACTUAL row sum of the (possibly bf16-rounded) probabilities keeps
    // sum_j dS_j == 0, so a large common component of K/Q cannot leak into dQ/dK.
    float dot = 0.f, psum = 0.f;
    for (int j = tid; j < L; j += 256) {
        float pj = ldf(p + j);
        dot = fmaf(pj, ldf(g + j), dot);
        psum += pj;
    }
    dot = warp_sum(dot);
    psum = warp_sum(psum);
    __shared__ float red2[8];
    if (lane == 0) { red[warp] = dot; red2[warp] = psum; }
    __syncthreads();
    dot = 0.f; psum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { dot += red[k]; psum += red2[k]; }
    dot /= psum;
    for (int j = tid; j < L; j += 256) stf(o + j, scale * ldf(p + j) * (ldf(g + j) - dot));
}


// single-pass variants: the row (L <= 256*8*NV elements, L % 8 == 0) is held in registers, one global read per element
__device__ __forceinline__ float block_max_256(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) r = fmaxf(r, red[k]);
    __syncthreads();
    return r;
}
__device__ __forceinline__ float block_sum_256(float v, float* red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += red[k];
    __syncthreads();
    return r;
}

template <typename T, int NV>
__global__ void __launch_bounds__(256) softmax_rows_fwd_fast_kernel(const T* __restrict__ S, T* __restrict__ P, int L, float scale) {
    __shared__ float red[8];
    const T* s = S + (long long)blockIdx.x * L;
    T* p = P + (long long)blockIdx.x * L;
    const int nvec = L >> 3;
    Vec8<T> v[NV];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        int j = threadIdx.x + 256 * i;
        if (j < nvec) {
            v[i].load(s + j * 8);
#pragma unroll
            for (int k = 0; k < 8; ++k) { v[i].v[k] *= scale; m = fmaxf(m, v[i].v[k]); }
        }
    }
    m = block_max_256(m, red);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        int j = threadIdx.x + 256 * i;
        if (j < nvec) {
#pragma unroll
            for (int k = 0; k < 8; ++k) { v[i].v[k] = row_exp<T>(v[i].v[k] - m); sum += v[i].v[k]; }
        }
    }
    sum = block_sum_256(sum, red);
    const float inv = 1.f / sum;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        int j = threadIdx.x + 256 * i;
        if (j < nvec) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[i].v[k] *= inv;
            v[i].store(p + j * 8);
        }
    }
}

template <typename T, int NV>
__global__ void __launch_bounds__(256) softmax_rows_bwd_fast_kernel(const T* __restrict__ P, const T* __restrict__ dP, T* __restrict__ dS,
                                                                    int L, float scale) {
    __shared__ float red[8];
    const T* p = P + (long long)blockIdx.x * L;
    const T* g = dP + (long long)blockIdx.x * L;
    T* o = dS + (long long)blockIdx.x * L;
    const int nvec = L >> 3;
    Vec8<T> pv[NV], gv[NV];
    float dot = 0.f, psum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        int j = threadIdx.x + 256 * i;
        if (j < nvec) {
            pv[i].load(p + j * 8);
            gv[i].load(g + j * 8);
#pragma unroll
            for (int k = 0; k < 8; ++k) { dot = fmaf(pv[i].v[k], gv[i].v[k], dot); psum += pv[i].v[k]; }
        }
    }
    dot = block_sum_256(dot, red);
    psum = block_sum_256(psum, red);
    dot /= psum;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        int j = threadIdx.x + 256 * i;
        if (j < nvec) {
#pragma unroll
            for (int k = 0; k < 8; ++k) gv[i].v[k] = scale * pv[i].v[k] * (gv[i].v[k] - dot);
            gv[i].store(o + j * 8);
        }
    }
}

// ---------------------------------------------------------------- tiny fp32 dense layers
__global__ void linear_f32_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b,
                                      float* __restrict__ y, int rows, int in, int out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * out) return;
    int r = i / out, o = i - r * out;
    float s = b ? b[o] : 0.f;
    for (int k = 0; k < in; ++k) s = fmaf(x[(long long)r * in + k], W[(long long)o * in + k], s);
    y[i] = s;
}
__global__ void linear_f32_bwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ dy,
                                      float* __restrict__ dx, float* __restrict__ dW, float* __restrict__ db, int rows, int in, int out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n_dx = dx ? rows * in : 0, n_dw = dW ? out * in : 0, n_db = db ? out : 0;
    if (i < n_dx) {
        int r = i / in, k = i - r * in;
        float s = 0.f;
        for (int o = 0; o < out; ++o) s = fmaf(dy[(long long)r * out + o], W[(long long)o * in + k], s);
        dx[i] = s;
    } else if (i < n_dx + n_dw) {
        int j = i - n_dx;
        int o = j / in, k = j - o * in;
        float s = 0.f;
        for (int r = 0; r < rows; ++r) s = fmaf(dy[(long long)r * out + o], x[(long long)r * in + k], s);
        dW[j] += s;
    } else if (i < n_dx + n_dw + n_db) {
        int o = i - n_dx - n_dw;
        float s = 0.f;
        for (int r = 0; r < rows; ++r) s += dy[(long long)r * out + o];
        db[o] += s;
    }
}

static inline int ew_blocks(long long work) {
    long long b = (work + 255) / 256, cap = (long long)num_sms() * 8;
    return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

}  // namespace stc

using namespace stc;

// ---------------------------------------------------------------- token centring (fp32 parity path of the attention blocks)
// softmax_j(q_i . k_j) is unchanged when one vector is subtracted from every k_j, and dS (rows sum to 0) is unchanged when one
// vector is subtracted from every v_j in dP = dO V^T / every k_j in dQ = dS K.  Removing the token mean takes the large common
// component out of those products, so the fp32 rounding of S and dP no longer dominates dS (random init: tokens nearly equal).
// mean[n][e] = mean_l x[n][l][e]: block = 8 row lanes x 32 channels, one block per (32-channel group, image), no atomics.
template <typename T>
__global__ void __launch_bounds__(256) token_mean_kernel(const T* __restrict__ x, float* __restrict__ mean, int L, int E) {
    __shared__ float red[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5, e = blockIdx.x * 32 + cx;
    const long long base = (long long)blockIdx.y * L * E;
    float acc = 0.f;
    if (e < E)
        for (int l = ry; l < L; l += 8) acc += ldf(x + base + (long long)l * E + e);
    red[ry][cx] = acc;
    __syncthreads();
    if (ry == 0 && e < E) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += red[k][cx];
        mean[(long long)blockIdx.y * E + e] = s / (float)L;
    }
}
template <typename T>
__global__ void token_center_kernel(const T* __restrict__ x, const float* __restrict__ mean, T* __restrict__ out, long long LE, int E,
                                    long long total) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const long long n = i / LE;
        const int e = (int)(i % E);
        stf(out + i, ldf(x + i) - mean[n * E + e]);
    }
}

extern "C" int stc_upcat_fused_ok(int N, int H, int W, int Cs, int h, int w, int Cu);
extern "C" int stc_upcat_apply_fwd(const void* skip, const void* low, const void* a, void* out, int N, int H, int W, int Cs, int h, int w, int Cu,
                                   int align_corners, int dtype, void* stream);
extern "C" int stc_upcat_apply_bwd(const void* dout, const void* dyhw, void* dskip, void* dlow, int N, int H, int W, int Cs, int h, int w, int Cu,
                                   int align_corners, int dtype, void* stream);

#define REQ_VEC(C, name) STC_REQUIRE((C) % 8 == 0, name ": channel count %d must be a multiple of 8", (int)(C))

extern "C" int stc_maxpool2_fwd(const void* x, void* y, int N, int H, int W, int C, int dtype, void* stream) {
    REQ_VEC(C, "maxpool2_fwd");
    long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
    if (total <= 0) return STC_OK;
    STC_DISPATCH_DTYPE(dtype, (maxpool2_fwd_kernel<T><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, H, W, C, total)));
    return check_launch("maxpool2_fwd");
}

extern "C" int stc_maxpool2_bwd(const void* x, const void* dy, void* dx, int N, int H, int W, int C, int dtype, void* stream) {
    REQ_VEC(C, "maxpool2_bwd");
    cudaStream_t st = (cudaStream_t)stream;
    if ((H & 1) || (W & 1))
        STC_CUDA(cudaMemsetAsync(dx, 0, (size_t)N * H * W * C * (dtype == STC_F32 ? 4 : 2), st));
    long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
    if (total <= 0) return STC_OK;
    STC_DISPATCH_DTYPE(dtype, (maxpool2_bwd_kernel<T><<<ew_blocks(total), 256, 0, st>>>((const T*)x, (const T*)dy, (T*)dx, H, W, C, total)));
    return check_launch("maxpool2_bwd");
}

extern "C" int stc_upcat_fwd(const void* skip, const void* low, void* out, int N, int H, int W, int Cs, int h, int w, int Cu,
                             int align_corners, int dtype, void* stream) {
    REQ_VEC(Cs, "upcat_fwd");
    REQ_VEC(Cu, "upcat_fwd");
    STC_REQUIRE(H >= 2 * h && W >= 2 * w, "upcat_fwd: skip (%d,%d) smaller than upsampled (%d,%d)", H, W, 2 * h, 2 * w);
    if (stc_upcat_fused_ok(N, H, W, Cs, h, w, Cu))   // row-structured kernels (upcat_fused.cu); the ones below are the wide-channel fallback
        return stc_upcat_apply_fwd(skip, low, nullptr, out, N, H, W, Cs, h, w, Cu, align_corners, dtype, stream);
    long long total = (long long)N * H * W * ((Cs + Cu) / 8);
    STC_DISPATCH_DTYPE(dtype, (upcat_fwd_kernel<T><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const T*)skip, (const T*)low, (T*)out,
                                                                                                     H, W, Cs, h, w, Cu, align_corners, total)));
    return check_launch("upcat_fwd");
}

extern "C" int stc_upcat_bwd(const void* dout, void* dskip, void* dlow, int N, int H, int W, int Cs, int h, int w, int Cu,
                             int align_corners, int dtype, void* stream) {
    REQ_VEC(Cs, "upcat_bwd");
    REQ_VEC(Cu, "upcat_bwd");
    if (stc_upcat_fused_ok(N, H, W, Cs, h, w, Cu))
        return stc_upcat_apply_bwd(dout, nullptr, dskip, dlow, N, H, W, Cs, h, w, Cu, align_corners, dtype, stream);
    cudaStream_t st = (cudaStream_t)stream;
    if (dskip && Cs > 0) {
        long long total = (long long)N * H * W * (Cs / 8);
        STC_DISPATCH_DTYPE(dtype, (upcat_bwd_skip_kernel<T><<<ew_blocks(total), 256, 0, st>>>((const T*)dout, (T*)dskip, Cs, Cs + Cu, total)));
    }
    if (dlow) {
        long long total = (long long)N * h * w * (Cu / 8);
        STC_DISPATCH_DTYPE(dtype, (upcat_bwd_low_kernel<T><<<ew_blocks(total), 256, 0, st>>>((const T*)dout, (T*)dlow, H, W, Cs, h, w, Cu,
                                                                                           align_corners, total)));
    }
    return check_launch("upcat_bwd");
}

extern "C" int stc_rowcol_mean(const void* x, void* y, int N, int H, int W, int C, int dtype, void* stream) {
    REQ_VEC(C, "rowcol_mean");
    long long total = (long long)N * (H + W) * (C / 8);
    STC_DISPATCH_DTYPE(dtype, (rowcol_reduce_kernel<T, false><<<ceil_div(total, 128), 128, 0, (cudaStream_t)stream>>>((const T*)x, nullptr, (T*)y, N, H, W, C)));
    return check_launch("rowcol_mean");
}

extern "C" int stc_coordatt_apply(const void* x, const void* a, void* out, int N, int H, int W, int C, int dtype, void* stream) {
    REQ_VEC(C, "coordatt_apply");
    long long total = (long long)N * H * W * (C / 8);
    STC_DISPATCH_DTYPE(dtype, (coordatt_ew_kernel<T, 0><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const T*)x, (const T*)a, (T*)out, H, W, C, total)));
    return check_launch("coordatt_apply");
}

extern "C" int stc_coordatt_apply_bwd(const void* dout, const void* a, void* da, int N, int H, int W, int C, int dtype, void* stream) {
    REQ_VEC(C, "coordatt_apply_bwd");
    long long total = (long long)N * (H + W) * (C / 8);
    STC_DISPATCH_DTYPE(dtype, (rowcol_reduce_kernel<T, true><<<ceil_div(total, 128), 128, 0, (cudaStream_t)stream>>>((const T*)dout, (const T*)a, (T*)da, N, H, W, C)));
    return check_launch("coordatt_apply_bwd");
}

extern "C" int stc_coordatt_dx(const void* dout, const void* dy, void* dx, int N, int H, int W, int C, int dtype, void* stream) {
    REQ_VEC(C, "coordatt_dx");
    long long total = (long long)N * H * W * (C / 8);
    STC_DISPATCH_DTYPE(dtype, (coordatt_ew_kernel<T, 1><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const T*)dout, (const T*)dy, (T*)dx, H, W, C, total)));
    return check_launch("coordatt_dx");
}

extern "C" int stc_ksa_pool(const void* f0, const void* f1, const void* f2, float* S, int N, long long HW, int C, int dtype, void* stream) {
    STC_REQUIRE(vec_ok(C), "ksa_pool: C=%d must be 8*2^k", C);
    cudaStream_t st = (cudaStream_t)stream;
    STC_CUDA(cudaMemsetAsync(S, 0, sizeof(float) * N * C, st));
    int lanes = C / 8;
    dim3 grid(max(1, min(reduce_blocks(HW, lanes), num_sms() * 4 / max(N, 1) + 1)), N);
    STC_DISPATCH_DTYPE(dtype, (ksa_reduce_kernel<T, 1><<<grid, 256, 0, st>>>(nullptr, (const T*)f0, (const T*)f1, (const T*)f2, S, HW, C, N, 1.f / (float)HW)));
    return check_launch("ksa_pool");
}

extern "C" int stc_ksa_dw(const void* dout, const void* f0, const void* f1, const void* f2, float* dw, int N, long long HW, int C,
                          int dtype, void* stream) {
    STC_REQUIRE(vec_ok(C), "ksa_dw: C=%d must be 8*2^k", C);
    cudaStream_t st = (cudaStream_t)stream;
    STC_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 3 * N * C, st));
    int lanes = C / 8;
    dim3 grid(max(1, min(reduce_blocks(HW, lanes), num_sms() * 4 / max(N, 1) + 1)), N);
    STC_DISPATCH_DTYPE(dtype, (ksa_reduce_kernel<T, 3><<<grid, 256, 0, st>>>((const T*)dout, (const T*)f0, (const T*)f1, (const T*)f2, dw, HW, C, N, 1.f)));
    return check_launch("ksa_dw");
}

extern "C" int stc_ksa_combine(const void* x, const void* f0, const void* f1, const void* f2, const float* w, void* out, int N,
                               long long HW, int C, int dtype, void* stream) {
    REQ_VEC(C, "ksa_combine");
    long long total = (long long)N * HW * (C / 8);
    STC_DISPATCH_DTYPE(dtype, (ksa_combine_kernel<T><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const T*)x, (const T*)f0, (const T*)f1,
                                                                                                       (const T*)f2, w, (T*)out, HW, C, N, total)));
    return check_launch("ksa_combine");
}

extern "C" int stc_ksa_df(const void* dout, const float* w, const float* dS, void* df0, void* df1, void* df2, int N, long long HW,
                          int C, int dtype, void* stream) {
    REQ_VEC(C, "ksa_df");
    long long total = (long long)N * HW * (C / 8);
    STC_DISPATCH_DTYPE(dtype, (ksa_df_kernel<T><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const T*)dout, w, dS, (T*)df0, (T*)df1, (T*)df2,
                                                                                                  HW, C, N, total)));
    return check_launch("ksa_df");
}

extern "C" int stc_softmax3_fwd(const float* a, float* w, long long NC, void* stream) {
    softmax3_fwd_kernel<<<ceil_div(NC, 256), 256, 0, (cudaStream_t)stream>>>(a, w, NC);
    return check_launch("softmax3_fwd");
}
extern "C" int stc_softmax3_bwd(const float* w, const float* dw, float* da, long long NC, void* stream) {
    softmax3_bwd_kernel<<<ceil_div(NC, 256), 256, 0, (cudaStream_t)stream>>>(w, dw, da, NC);
    return check_launch("softmax3_bwd");
}

extern "C" int stc_center_tokens(const void* x, void* out, float* mean_ws, int N, int L, int E, int dtype, void* stream) {
    if (N <= 0 || L <= 0 || E <= 0) return STC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)N * L * E;
    dim3 grid(ceil_div(E, 32), N);
    STC_DISPATCH_DTYPE(dtype, (token_mean_kernel<T><<<grid, 256, 0, st>>>((const T*)x, mean_ws, L, E)));
    STC_DISPATCH_DTYPE(dtype, (token_center_kernel<T><<<ew_blocks(total), 256, 0, st>>>((const T*)x, mean_ws, (T*)out, (long long)L * E, E, total)));
    return check_launch("center_tokens");
}

extern "C" int stc_softmax_rows_fwd(const void* S, void* P, long long rows, int L, float scale, int dtype, void* stream) {
    if (rows <= 0) return STC_OK;
    STC_REQUIRE(rows < (1LL << 31), "softmax_rows: too many rows");
    cudaStream_t st = (cudaStream_t)stream;
    bool al = ((((uintptr_t)S) | ((uintptr_t)P)) & 15) == 0 && L % 8 == 0;
    if (al && L <= 2048) {
        STC_DISPATCH_DTYPE(dtype, (softmax_rows_fwd_fast_kernel<T, 1><<<(unsigned)rows, 256, 0, st>>>((const T*)S, (T*)P, L, scale)));
    } else if (al && L <= 4096) {
        STC_DISPATCH_DTYPE(dtype, (softmax_rows_fwd_fast_kernel<T, 2><<<(unsigned)rows, 256, 0, st>>>((const T*)S, (T*)P, L, scale)));
    } else if (al && L <= 8192) {
        STC_DISPATCH_DTYPE(dtype, (softmax_rows_fwd_fast_kernel<T, 4><<<(unsigned)rows, 256, 0, st>>>((const T*)S, (T*)P, L, scale)));
    } else {
        STC_DISPATCH_DTYPE(dtype, (softmax_rows_fwd_kernel<T><<<(unsigned)rows, 256, 0, st>>>((const T*)S, (T*)P, L, scale)));
    }
    return check_launch("softmax_rows_fwd");
}
extern "C" int stc_softmax_rows_bwd(const void* P, const void* dP, void* dS, long long rows, int L, float scale, int dtype, void* stream) {
    if (rows <= 0) return STC_OK;
    STC_REQUIRE(rows < (1LL << 31), "softmax_rows: too many rows");
    cudaStream_t st = (cudaStream_t)stream;
    bool al = ((((uintptr_t)P) | ((uintptr_t)dP) | ((uintptr_t)dS)) & 15) == 0 && L % 8 == 0;
    if (al && L <= 2048) {
        STC_DISPATCH_DTYPE(dtype, (softmax_rows_bwd_fast_kernel<T, 1><<<(unsigned)rows, 256, 0, st>>>((const T*)P, (const T*)dP, (T*)dS, L, scale)));
    } else if (al && L <= 4096) {
        STC_DISPATCH_DTYPE(dtype, (softmax_rows_bwd_fast_kernel<T, 2><<<(unsigned)rows, 256, 0, st>>>((const T*)P, (const T*)dP, (T*)dS, L, scale)));
    } else {
        STC_DISPATCH_DTYPE(dtype, (softmax_rows_bwd_kernel<T><<<(unsigned)rows, 256, 0, st>>>((const T*)P, (const T*)dP, (T*)dS, L, scale)));
    }
    return check_launch("softmax_rows_bwd");
}

extern "C" int stc_linear_f32_fwd(const float* x, const float* W, const float* b, float* y, int rows, int in, int out, void* stream) {
    linear_f32_fwd_kernel<<<ceil_div((long long)rows * out, 128), 128, 0, (cudaStream_t)stream>>>(x, W, b, y, rows, in, out);
    return check_launch("linear_f32_fwd");
}
extern "C" int stc_linear_f32_bwd(const float* x, const float* W, const float* dy, float* dx, float* dW, float* db, int rows, int in,
                                  int out, void* stream) {
    long long total = (dx ? (long long)rows * in : 0) + (dW ? (long long)out * in : 0) + (db ? out : 0);
    if (total <= 0) return STC_OK;
    linear_f32_bwd_kernel<<<ceil_div(total, 128), 128, 0, (cudaStream_t)stream>>>(x, W, dy, dx, dW, db, rows, in, out);
    return check_launch("linear_f32_bwd");
}

// ---------------------------------------------------------------- n-ary channel concat with optional nearest x2 on input 0
// UNet++ DecoderBlock (segmentation_models_pytorch 0.2.0, decoder blocks x_i_j): x = interpolate(x, 2, 'nearest');
// x = cat([x, skip]) where skip is itself a cat of dense features.  out[n,h,w,:] = [in0 | in1 | ... ] with in0 read at
// (h/2, w/2) when up0 != 0.  The adjoint sums the 2x2 block for in0.
namespace stc {
struct CatN {
    const void* in[5];
    void* din[5];
    int ch[5];
    int n;
};

template <typename T>
__global__ void catn_fwd_kernel(CatN a, T* __restrict__ out, int H, int W, int Ct, int up0, long long total) {
    const int lanes = Ct >> 3;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int lv = (int)(i % lanes);
        long long p = i / lanes;
        int c = lv * 8, k = 0;
        while (k < a.n - 1 && c >= a.ch[k]) { c -= a.ch[k]; ++k; }
        long long src = p;
        if (k == 0 && up0) {
            int w_ = (int)(p % W), h_ = (int)((p / W) % H);
            long long n_ = p / ((long long)W * H);
            src = (n_ * (H >> 1) + (h_ >> 1)) * (long long)(W >> 1) + (w_ >> 1);
        }
        Vec8<T> v;
        v.load(reinterpret_cast<const T*>(a.in[k]) + src * a.ch[k] + c);
        v.store(out + i * 8);
    }
}

// one thread per (input-k pixel, 8 channels): gathers from dout
template <typename T>
__global__ void catn_bwd_kernel(CatN a, const T* __restrict__ dout, int H, int W, int Ct, int up0, int k, int coff, long long total) {
    const int lanes = a.ch[k] >> 3;
    T* dst = reinterpret_cast<T*>(a.din[k]);
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int lv = (int)(i % lanes);
        long long p = i / lanes;
        Vec8<T> v;
        if (k == 0 && up0) {
            const int w2 = W >> 1, h2 = H >> 1;
            int w_ = (int)(p % w2), h_ = (int)((p / w2) % h2);
            long long n_ = p / ((long long)w2 * h2);
            float acc[8] = {};
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    Vec8<T> g;
                    g.load(dout + ((n_ * H + 2 * h_ + dy) * (long long)W + 2 * w_ + dx) * Ct + coff + lv * 8);
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[e] += g.v[e];
                }
#pragma unroll
            for (int e = 0; e < 8; ++e) v.v[e] = acc[e];
        } else {
            v.load(dout + p * Ct + coff + lv * 8);
        }
        v.store(dst + i * 8);
    }
}
}  // namespace stc

extern "C" int stc_catn_fwd(const void* in0, const void* in1, const void* in2, const void* in3, const void* in4, int c0, int c1, int c2,
                            int c3, int c4, void* out, int N, int H, int W, int up0, int dtype, void* stream) {
    stc::CatN a;
    const void* ins[5] = {in0, in1, in2, in3, in4};
    int chs[5] = {c0, c1, c2, c3, c4};
    a.n = 0;
    int Ct = 0;
    for (int k = 0; k < 5; ++k) {
        if (!ins[k] || chs[k] <= 0) break;
        STC_REQUIRE(chs[k] % 8 == 0, "catn_fwd: channel counts must be multiples of 8");
        a.in[a.n] = ins[k]; a.din[a.n] = nullptr; a.ch[a.n] = chs[k]; Ct += chs[k]; ++a.n;
    }
    STC_REQUIRE(a.n >= 1 && (!up0 || (H % 2 == 0 && W % 2 == 0)), "catn_fwd: bad arguments");
    long long total = (long long)N * H * W * (Ct / 8);
    if (total <= 0) return STC_OK;
    STC_DISPATCH_DTYPE(dtype, (catn_fwd_kernel<T><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(a, (T*)out, H, W, Ct, up0, total)));
    return check_launch("catn_fwd");
}

extern "C" int stc_catn_bwd(const void* dout, void* d0, void* d1, void* d2, void* d3, void* d4, int c0, int c1, int c2, int c3, int c4,
                            int N, int H, int W, int up0, int dtype, void* stream) {
    stc::CatN a;
    void* ds[5] = {d0, d1, d2, d3, d4};
    int chs[5] = {c0, c1, c2, c3, c4};
    a.n = 0;
    int Ct = 0;
    for (int k = 0; k < 5; ++k) {
        if (chs[k] <= 0) break;
        a.in[a.n] = nullptr; a.din[a.n] = ds[k]; a.ch[a.n] = chs[k]; Ct += chs[k]; ++a.n;
    }
    int coff = 0;
    for (int k = 0; k < a.n; ++k) {
        if (a.din[k]) {
            long long pix = (k == 0 && up0) ? (long long)N * (H / 2) * (W / 2) : (long long)N * H * W;
            long long total = pix * (a.ch[k] / 8);
            STC_DISPATCH_DTYPE(dtype, (catn_bwd_kernel<T><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(a, (const T*)dout, H, W, Ct, up0, k,
                                                                                                            coff, total)));
        }
        coff += a.ch[k];
    }
    return check_launch("catn_bwd");
}

// ------------------------------------------------------------------------------------------------------------
// depth-to-space (pixel shuffle by 2) on NHWC: hi[n, 2h+py, 2w+px, c] <-> lo[n, h, w, (py*2+px)*C + c]
// (second half of the ConvTranspose2d(k=4,s=2,p=1) = [3x3 conv to 4C sub-pixel channels] decomposition; its adjoint is the inverse move)
// ------------------------------------------------------------------------------------------------------------
namespace stc {
template <typename T>
__global__ void d2s_kernel(const T* __restrict__ src, T* __restrict__ dst, int H, int W, int C, int inverse, long long total) {
    const int lanes = C >> 3;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {      // i indexes the hi-res tensor (N,2H,2W,C) in 8-channel lanes: coalesced on that side
        int lv = (int)(i % lanes);
        long long p = i / lanes;
        int x = (int)(p % (2 * W)), y = (int)((p / (2 * W)) % (2 * H));
        long long n = p / (4LL * W * H);
        long long lo = (((n * H + (y >> 1)) * W + (x >> 1)) * 4 + ((y & 1) * 2 + (x & 1))) * C + lv * 8;
        Vec8<T> v;
        if (inverse) { v.load(src + i * 8); v.store(dst + lo); }
        else         { v.load(src + lo);    v.store(dst + i * 8); }
    }
}
}  // namespace stc

extern "C" int stc_depth_to_space2(const void* src, void* dst, int N, int H, int W, int C, int inverse, int dtype, void* stream) {
    STC_REQUIRE(src && dst && C % 8 == 0 && N >= 0 && H >= 0 && W >= 0, "depth_to_space2: channels must be a multiple of 8");
    long long total = (long long)N * H * W * 4 * (C / 8);
    if (total <= 0) return STC_OK;
    STC_DISPATCH_DTYPE(dtype, (d2s_kernel<T><<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>((const T*)src, (T*)dst, H, W, C, inverse, total)));
    return check_launch("depth_to_space2");
}
