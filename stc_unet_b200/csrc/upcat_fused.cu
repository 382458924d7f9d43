// Up.forward (decode_heads/unet_head.py:50-60) as three row-structured passes that never materialise the concatenated tensor:
//
//   cat = [skip, pad(bilinear_x2(low))]                       (virtual)
//   y   = row / column means of cat                           stc_upcat_pool        (CoordAtt pool_h / pool_w, unet_head.py:134-136)
//   out = cat + a_h * a_w                                     stc_upcat_apply_fwd   (`self.ca(x) + x`, unet_head.py:57); a == NULL: out = cat
//   g   = dout + dy_h / W + dy_w / H  ->  dskip, dlow         stc_upcat_apply_bwd   (adjoint of cat and of the two means in one pass)
//
// Every block owns one image row (n, y): a thread keeps its 8-channel lane for the whole launch, the row-constant terms
// (vertical interpolation weights, a_h / dy_h) live in registers and the x loop is load -> fma -> store with pointer bumps only.
// HBM-bound; algorithmic bytes: fwd = |skip| + |low| + |out|, pool = |skip| + |low| per pass (two passes), bwd = |dout| + |dskip| + |dlow|.
#include "common.cuh"

namespace stc {

struct Lerp {
    int i0, i1;
    float l0, l1;
};
// nn.Upsample(scale_factor=2, bilinear): source position of output index o; scale = (in-1)/(out-1) for align_corners=True
__device__ __forceinline__ Lerp lerp_src(int o, int in, float scale, int align) {
    Lerp r;
    float src = align ? scale * (float)o : fmaxf(0.5f * ((float)o + 0.5f) - 0.5f, 0.f);
    r.i0 = min((int)src, in - 1);
    r.i1 = r.i0 + (r.i0 < in - 1 ? 1 : 0);
    r.l1 = src - (float)r.i0;
    r.l0 = 1.f - r.l1;
    return r;
}
static inline float lerp_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

struct UpCatGeom {
    int H, W, Cs, h, w, Cu, align, py, px;
    float sy, sx;
};
static UpCatGeom make_geom(int H, int W, int Cs, int h, int w, int Cu, int align) {
    UpCatGeom g{H, W, Cs, h, w, Cu, align, (H - 2 * h) / 2, (W - 2 * w) / 2, lerp_scale(h, 2 * h), lerp_scale(w, 2 * w)};
    return g;
}

// value of the virtual cat tensor at (n, y, x) for a LOW lane (channel offset cl inside low), given the row's vertical weights
template <typename T>
__device__ __forceinline__ void low_value(const T* __restrict__ lown, const UpCatGeom& g, bool row_valid, const Lerp& iy, int x, int cl,
                                          float (&v)[8]) {
    const int ox = x - g.px;
    if (!row_valid || ox < 0 || ox >= 2 * g.w) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0.f;
        return;
    }
    const Lerp ix = lerp_src(ox, g.w, g.sx, g.align);
    Vec8<T> a00, a01, a10, a11;
    const T* r0 = lown + (long long)iy.i0 * g.w * g.Cu + cl;
    const T* r1 = lown + (long long)iy.i1 * g.w * g.Cu + cl;
    a00.load(r0 + (long long)ix.i0 * g.Cu);
    a01.load(r0 + (long long)ix.i1 * g.Cu);
    a10.load(r1 + (long long)ix.i0 * g.Cu);
    a11.load(r1 + (long long)ix.i1 * g.Cu);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = iy.l0 * (ix.l0 * a00.v[k] + ix.l1 * a01.v[k]) + iy.l1 * (ix.l0 * a10.v[k] + ix.l1 * a11.v[k]);
}

// weights with which low index j receives from the up-sampled indices around it: candidates lo .. lo+STC_NC-1 (statically indexed so
// the table stays in registers; entries past the last real tap carry weight 0)
#define STC_NC 7
struct Taps {
    int lo;
    float wt[STC_NC];
};
__device__ __forceinline__ Taps adjoint_taps(int j, int in, float scale, int align) {
    Taps t;
    const int out = 2 * in;
    int hi;
    if (align) {
        t.lo = scale > 0.f ? max(0, (int)floorf((float)(j - 1) / scale)) : 0;
        hi = scale > 0.f ? min(out - 1, (int)ceilf((float)(j + 1) / scale)) : out - 1;
    } else {
        t.lo = max(0, 2 * j - 2);
        hi = min(out - 1, 2 * j + 3);
    }
#pragma unroll
    for (int c = 0; c < STC_NC; ++c) {
        const int o = t.lo + c;
        float wgt = 0.f;
        if (o <= hi) {
            const Lerp l = lerp_src(o, in, scale, align);
            wgt = (l.i0 == j ? l.l0 : 0.f) + (l.i1 == j ? l.l1 : 0.f);
        }
        t.wt[c] = wgt;
    }
    return t;
}

// ------------------------------------------------------------------------------------------------ forward
// Two launches, one per half of the channel range, so every warp runs ONE code path (copy or interpolate).
// grid (gx, H, N); block 256; lanes = (channels of this half)/8 <= 256; the half starts at channel `c0` of the output row.
template <typename T, bool HAS_A, bool IS_LOW>
__global__ void __launch_bounds__(256) upcat_apply_rows_kernel(const T* __restrict__ src, const T* __restrict__ a, T* __restrict__ out,
                                                               UpCatGeom g) {
    const int Ct = g.Cs + g.Cu, Ch = IS_LOW ? g.Cu : g.Cs, c0 = IS_LOW ? g.Cs : 0, lanes = Ch >> 3, rstep = 256 / lanes;
    if ((int)threadIdx.x >= rstep * lanes) return;
    const int lv = threadIdx.x % lanes, r0 = threadIdx.x / lanes;
    const int y = blockIdx.y;
    const long long n = blockIdx.z;
    const int oy = y - g.py;
    const bool row_valid = oy >= 0 && oy < 2 * g.h;
    const Lerp iy = lerp_src(row_valid ? oy : 0, g.h, g.sy, g.align);
    float ah[8];
    if (HAS_A) {
        Vec8<T> t;
        t.load(a + (n * (g.H + g.W) + y) * (long long)Ct + c0 + lv * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) ah[k] = t.v[k];
    }
    const T* srow = IS_LOW ? src + n * (long long)g.h * g.w * g.Cu : src + ((n * g.H + y) * (long long)g.W) * g.Cs + lv * 8;
    const T* aw = HAS_A ? a + (n * (g.H + g.W) + g.H) * (long long)Ct + c0 + lv * 8 : nullptr;
    T* orow = out + ((n * g.H + y) * (long long)g.W) * Ct + c0 + lv * 8;
    for (int x = blockIdx.x * rstep + r0; x < g.W; x += gridDim.x * rstep) {
        Vec8<T> v;
        if (!IS_LOW) v.load(srow + (long long)x * g.Cs);
        else low_value(srow, g, row_valid, iy, x, lv * 8, v.v);
        if (HAS_A) {
            Vec8<T> w_;
            w_.load(aw + (long long)x * Ct);
#pragma unroll
            for (int k = 0; k < 8; ++k) v.v[k] = fmaf(ah[k], w_.v[k], v.v[k]);
        }
        v.store(orow + (long long)x * Ct);
    }
}

// ------------------------------------------------------------------------------------------------ row / column means of the virtual cat
// skip half: plain sums.  thread per (n, y, lane) sums over x  [descriptor rows 0..H-1], thread per (n, x, lane) sums over y  [rows H..]
template <typename T>
__global__ void __launch_bounds__(128) pool_skip_kernel(const T* __restrict__ skip, T* __restrict__ yout, int N, UpCatGeom g) {
    const int Ct = g.Cs + g.Cu, lanes = g.Cs >> 3;
    // image-major work order: the column pass of image n follows its row pass closely enough to find the image still in L2
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long per_img = (long long)(g.H + g.W) * lanes;
    if (i >= per_img * N) return;
    float acc0[8] = {}, acc1[8] = {};
    const long long n = i / per_img;
    const long long rem = i - n * per_img;
    const bool rowpart = rem < (long long)g.H * lanes;
    const long long j = rowpart ? rem : rem - (long long)g.H * lanes;
    const int lv = (int)(j % lanes);
    const int r = (int)(j / lanes);                   // y (row pass) or x (column pass)
    const int cnt = rowpart ? g.W : g.H;
    const T* b = rowpart ? skip + (n * g.H + r) * (long long)g.W * g.Cs + lv * 8 : skip + (n * g.H * (long long)g.W + r) * g.Cs + lv * 8;
    const long long step = rowpart ? g.Cs : (long long)g.W * g.Cs;
    // shifted sums: mean = s + mean(v - s) with s = the first element.  Activations behind a BN sit on a common offset that the
    // CoordAtt BN removes again; summing the offset-free part keeps the descriptor's FLUCTUATION accurate (as in bn_reduce)
    Vec8<T> sh;
    sh.load(b);
    int t = 0;
    for (; t + 1 < cnt; t += 2) {                      // two independent accumulators: two loads in flight per thread
        Vec8<T> v0, v1;
        v0.load(b + t * step);
        v1.load(b + (t + 1) * step);
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc0[k] += v0.v[k] - sh.v[k]; acc1[k] += v1.v[k] - sh.v[k]; }
    }
    if (t < cnt) {
        Vec8<T> v0;
        v0.load(b + t * step);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc0[k] += v0.v[k] - sh.v[k];
    }
    Vec8<T> o;
    const float sc = 1.f / (float)cnt;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = fmaf(acc0[k] + acc1[k], sc, sh.v[k]);
    o.store(yout + (n * (g.H + g.W) + (rowpart ? r : g.H + r)) * (long long)Ct + lv * 8);
}

// low half, step 1: the means of the UP-SAMPLED tensor are linear in low, so reduce low itself with the adjoint interpolation weights:
//   R[n,i,c] = sum_j wx[j] low[n,i,j,c]   (wx[j] = total weight with which column j feeds the 2w up-sampled columns)
//   S[n,j,c] = sum_i wy[i] low[n,i,j,c]
// tmp is fp32 (N, h + w, Cu).  thread per (n, i, lane) / (n, j, lane) as above.
template <typename T>
__global__ void __launch_bounds__(128) pool_low_reduce_kernel(const T* __restrict__ low, float* __restrict__ tmp, int N, UpCatGeom g) {
    extern __shared__ float wtab[];   // [0, w): column weights wx[j];  [w, w + h): row weights wy[i]
    for (int t = threadIdx.x; t < g.w + g.h; t += blockDim.x) {
        const bool isx = t < g.w;
        const Taps tp = adjoint_taps(isx ? t : t - g.w, isx ? g.w : g.h, isx ? g.sx : g.sy, g.align);
        float wsum = 0.f;
#pragma unroll
        for (int c = 0; c < STC_NC; ++c) wsum += tp.wt[c];
        wtab[t] = wsum;
    }
    __syncthreads();
    const int lanes = g.Cu >> 3;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long n_row = (long long)N * g.h * lanes, n_col = (long long)N * g.w * lanes;
    if (i >= n_row + n_col) return;
    float acc0[8] = {}, acc1[8] = {};
    const bool rowpart = i < n_row;
    const long long j = rowpart ? i : i - n_row;
    const int lv = (int)(j % lanes);
    const long long q = j / lanes;
    const int cnt = rowpart ? g.w : g.h, dim = rowpart ? g.h : g.w;
    const long long n = q / dim;
    const int r = (int)(q % dim);
    const T* b = rowpart ? low + q * (long long)g.w * g.Cu + lv * 8 : low + (n * g.h * (long long)g.w + r) * g.Cu + lv * 8;
    const long long step = rowpart ? g.Cu : (long long)g.w * g.Cu;
    const float* wt = rowpart ? wtab : wtab + g.w;
    // shifted, as in pool_skip_kernel: sum_t w_t v_t = s * sum_t w_t + sum_t w_t (v_t - s)
    Vec8<T> sh;
    sh.load(b);
    float wsum = 0.f;
    int t = 0;
    for (; t + 1 < cnt; t += 2) {
        Vec8<T> v0, v1;
        v0.load(b + t * step);
        v1.load(b + (t + 1) * step);
        const float w0 = wt[t], w1 = wt[t + 1];
        wsum += w0 + w1;
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc0[k] = fmaf(w0, v0.v[k] - sh.v[k], acc0[k]); acc1[k] = fmaf(w1, v1.v[k] - sh.v[k], acc1[k]); }
    }
    if (t < cnt) {
        Vec8<T> v0;
        v0.load(b + t * step);
        const float w0 = wt[t];
        wsum += w0;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc0[k] = fmaf(w0, v0.v[k] - sh.v[k], acc0[k]);
    }
    float* o = tmp + (n * (g.h + g.w) + (rowpart ? r : g.h + r)) * (long long)g.Cu + lv * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = fmaf(sh.v[k], wsum, acc0[k] + acc1[k]);
}

// low half, step 2: y[n, y, Cs + c] = interp_y(R)[y] / W  and  y[n, H + x, Cs + c] = interp_x(S)[x] / H  (zero outside the padded window)
template <typename T>
__global__ void pool_low_finish_kernel(const float* __restrict__ tmp, T* __restrict__ yout, int N, UpCatGeom g) {
    const int Ct = g.Cs + g.Cu;
    const long long total = (long long)N * (g.H + g.W) * g.Cu;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % g.Cu);
    const long long q = i / g.Cu;
    const long long n = q / (g.H + g.W);
    const int r = (int)(q % (g.H + g.W));
    const bool rowpart = r < g.H;
    const int o = rowpart ? r - g.py : r - g.H - g.px;
    const int in = rowpart ? g.h : g.w;
    float v = 0.f;
    if (o >= 0 && o < 2 * in) {
        const Lerp l = lerp_src(o, in, rowpart ? g.sy : g.sx, g.align);
        const float* t = tmp + (n * (g.h + g.w) + (rowpart ? 0 : g.h)) * (long long)g.Cu + c;
        v = (l.l0 * t[(long long)l.i0 * g.Cu] + l.l1 * t[(long long)l.i1 * g.Cu]) / (float)(rowpart ? g.W : g.H);
    }
    stf(yout + q * Ct + g.Cs + c, v);
}

// ------------------------------------------------------------------------------------------------ backward
// skip half: dskip[n,y,x,:] = dout[n,y,x,:Cs] + dyh[n,y,:Cs]/W + dyw[n,x,:Cs]/H.   grid (gx, H, N), lanes = Cs/8
template <typename T, bool HAS_DY>
__global__ void __launch_bounds__(256) upcat_bwd_skip_rows_kernel(const T* __restrict__ dout, const T* __restrict__ dyhw, T* __restrict__ dskip,
                                                                  UpCatGeom g) {
    const int Ct = g.Cs + g.Cu, lanes = g.Cs >> 3, rstep = 256 / lanes;
    if ((int)threadIdx.x >= rstep * lanes) return;
    const int lv = threadIdx.x % lanes, r0 = threadIdx.x / lanes;
    const int y = blockIdx.y;
    const long long n = blockIdx.z;
    float dh[8];
    const float iw = 1.f / (float)g.W, ih = 1.f / (float)g.H;
    if (HAS_DY) {
        Vec8<T> t;
        t.load(dyhw + (n * (g.H + g.W) + y) * (long long)Ct + lv * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) dh[k] = t.v[k] * iw;
    }
    const T* drow = dout + ((n * g.H + y) * (long long)g.W) * Ct + lv * 8;
    const T* dw = HAS_DY ? dyhw + (n * (g.H + g.W) + g.H) * (long long)Ct + lv * 8 : nullptr;
    T* srow = dskip + ((n * g.H + y) * (long long)g.W) * g.Cs + lv * 8;
    for (int x = blockIdx.x * rstep + r0; x < g.W; x += gridDim.x * rstep) {
        Vec8<T> v;
        v.load(drow + (long long)x * Ct);
        if (HAS_DY) {
            Vec8<T> w_;
            w_.load(dw + (long long)x * Ct);
#pragma unroll
            for (int k = 0; k < 8; ++k) v.v[k] += dh[k] + w_.v[k] * ih;
        }
        v.store(srow + (long long)x * g.Cs);
    }
}

// low half (gather form of the adjoint).  grid (gx, h, N), lanes = Cu/8; one block per low-res row jy.  The per-column tap tables
// (first up-sampled column + STC_NC weights) are built once per block in shared memory.
template <typename T, bool HAS_DY>
__global__ void __launch_bounds__(256) upcat_bwd_low_rows_kernel(const T* __restrict__ dout, const T* __restrict__ dyhw, T* __restrict__ dlow,
                                                                 UpCatGeom g) {
    extern __shared__ float xtab[];   // per jx: [lo (as int bits), wt[0..STC_NC)]
    const int Ct = g.Cs + g.Cu, lanes = g.Cu >> 3, rstep = 256 / lanes;
    for (int jx = threadIdx.x; jx < g.w; jx += 256) {
        const Taps tx = adjoint_taps(jx, g.w, g.sx, g.align);
        xtab[jx * (STC_NC + 1)] = __int_as_float(tx.lo);
#pragma unroll
        for (int c = 0; c < STC_NC; ++c) xtab[jx * (STC_NC + 1) + 1 + c] = tx.wt[c];
    }
    __syncthreads();
    if ((int)threadIdx.x >= rstep * lanes) return;
    const int lv = threadIdx.x % lanes, r0 = threadIdx.x / lanes;
    const int jy = blockIdx.y;
    const long long n = blockIdx.z;
    const float iw = 1.f / (float)g.W, ih = 1.f / (float)g.H;
    const Taps ty = adjoint_taps(jy, g.h, g.sy, g.align);
    const int cofs = g.Cs + lv * 8;
    // row-constant part: sum_t wy_t * dyh[n, oy_t + py] / W   (multiplied by the sum of wx below)
    float dh[8] = {};
    float wysum = 0.f;
#pragma unroll
    for (int t = 0; t < STC_NC; ++t) {
        wysum += ty.wt[t];
        if (HAS_DY && ty.wt[t] != 0.f) {
            Vec8<T> v;
            v.load(dyhw + (n * (g.H + g.W) + ty.lo + t + g.py) * (long long)Ct + cofs);
#pragma unroll
            for (int k = 0; k < 8; ++k) dh[k] = fmaf(ty.wt[t] * iw, v.v[k], dh[k]);
        }
    }
    const T* dn = dout + (n * g.H + ty.lo + g.py) * (long long)g.W * Ct + cofs;
    const T* dwp = HAS_DY ? dyhw + (n * (g.H + g.W) + g.H) * (long long)Ct + cofs : nullptr;
    T* lrow = dlow + ((n * g.h + jy) * (long long)g.w) * g.Cu + lv * 8;
    const long long rowpitch = (long long)g.W * Ct;
    for (int jx = blockIdx.x * rstep + r0; jx < g.w; jx += gridDim.x * rstep) {
        const float* tx = xtab + jx * (STC_NC + 1);
        const int xlo = __float_as_int(tx[0]) + g.px;
        float acc[8] = {};
        float wxsum = 0.f;
#pragma unroll
        for (int s = 0; s < STC_NC; ++s) {
            const float wx = tx[1 + s];
            if (wx == 0.f) continue;
            wxsum += wx;
            const T* col = dn + (long long)(xlo + s) * Ct;
#pragma unroll
            for (int t = 0; t < STC_NC; ++t) {
                if (ty.wt[t] == 0.f) continue;
                Vec8<T> v;
                v.load(col + t * rowpitch);
                const float ww = ty.wt[t] * wx;
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = fmaf(ww, v.v[k], acc[k]);
            }
            if (HAS_DY) {
                Vec8<T> v;
                v.load(dwp + (long long)(xlo + s) * Ct);
                const float ww = wx * wysum * ih;
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = fmaf(ww, v.v[k], acc[k]);
            }
        }
        Vec8<T> o;
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] = HAS_DY ? fmaf(dh[k], wxsum, acc[k]) : acc[k];
        o.store(lrow + (long long)jx * g.Cu);
    }
}

static inline bool rows_ok(int C) { return C > 0 && C % 8 == 0 && C / 8 <= 256; }
static inline dim3 rows_grid(int Wd, int Hd, int N, int lanes) {
    int rstep = 256 / lanes;
    int gx = (Wd + rstep * 4 - 1) / (rstep * 4);   // ~4 pixels per thread
    return dim3((unsigned)(gx < 1 ? 1 : gx), (unsigned)Hd, (unsigned)N);
}

}  // namespace stc

using namespace stc;

extern "C" int stc_upcat_fused_ok(int N, int H, int W, int Cs, int h, int w, int Cu) {
    return (Cs % 8 == 0 && Cu % 8 == 0 && Cu > 0 && rows_ok(Cs + Cu) && H >= 2 * h && W >= 2 * w && H <= 65535 && N <= 65535 && w <= 1536 &&
            h + w <= 8192) ? 1 : 0;
}

extern "C" int stc_upcat_apply_fwd(const void* skip, const void* low, const void* a, void* out, int N, int H, int W, int Cs, int h, int w, int Cu,
                                   int align_corners, int dtype, void* stream) {
    STC_REQUIRE(stc_upcat_fused_ok(N, H, W, Cs, h, w, Cu), "upcat_apply_fwd: unsupported shape (N=%d H=%d W=%d Cs=%d h=%d w=%d Cu=%d)", N, H, W, Cs, h,
                w, Cu);
    STC_REQUIRE(low && out && (skip || Cs == 0), "upcat_apply_fwd: null pointer");
    if ((long long)N * H * W == 0) return STC_OK;
    UpCatGeom g = make_geom(H, W, Cs, h, w, Cu, align_corners);
    cudaStream_t st = (cudaStream_t)stream;
    if (Cs > 0) {
        dim3 grid = rows_grid(W, H, N, Cs / 8);
        if (a) {
            STC_DISPATCH_DTYPE(dtype, (upcat_apply_rows_kernel<T, true, false><<<grid, 256, 0, st>>>((const T*)skip, (const T*)a, (T*)out, g)));
        } else {
            STC_DISPATCH_DTYPE(dtype, (upcat_apply_rows_kernel<T, false, false><<<grid, 256, 0, st>>>((const T*)skip, nullptr, (T*)out, g)));
        }
    }
    dim3 grid = rows_grid(W, H, N, Cu / 8);
    if (a) {
        STC_DISPATCH_DTYPE(dtype, (upcat_apply_rows_kernel<T, true, true><<<grid, 256, 0, st>>>((const T*)low, (const T*)a, (T*)out, g)));
    } else {
        STC_DISPATCH_DTYPE(dtype, (upcat_apply_rows_kernel<T, false, true><<<grid, 256, 0, st>>>((const T*)low, nullptr, (T*)out, g)));
    }
    return check_launch("upcat_apply_fwd");
}

extern "C" long long stc_upcat_pool_ws_bytes(int N, int h, int w, int Cu) { return (long long)N * (h + w) * Cu * (long long)sizeof(float); }

extern "C" int stc_upcat_pool(const void* skip, const void* low, void* y, int N, int H, int W, int Cs, int h, int w, int Cu, int align_corners,
                              void* ws, long long ws_bytes, int dtype, void* stream) {
    STC_REQUIRE(stc_upcat_fused_ok(N, H, W, Cs, h, w, Cu), "upcat_pool: unsupported shape");
    STC_REQUIRE(low && y && (skip || Cs == 0), "upcat_pool: null pointer");
    STC_REQUIRE(ws && ws_bytes >= stc_upcat_pool_ws_bytes(N, h, w, Cu), "upcat_pool: workspace too small");
    UpCatGeom g = make_geom(H, W, Cs, h, w, Cu, align_corners);
    if ((long long)N * (H + W) == 0) return STC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (Cs > 0) {
        long long total = (long long)N * (H + W) * (Cs / 8);
        STC_DISPATCH_DTYPE(dtype, (pool_skip_kernel<T><<<ceil_div(total, 128), 128, 0, st>>>((const T*)skip, (T*)y, N, g)));
    }
    long long t1 = (long long)N * (h + w) * (Cu / 8), t2 = (long long)N * (H + W) * Cu;
    STC_DISPATCH_DTYPE(dtype, (pool_low_reduce_kernel<T><<<ceil_div(t1, 128), 128, sizeof(float) * (size_t)(h + w), st>>>((const T*)low, (float*)ws, N, g)));
    STC_DISPATCH_DTYPE(dtype, (pool_low_finish_kernel<T><<<ceil_div(t2, 256), 256, 0, st>>>((const float*)ws, (T*)y, N, g)));
    return check_launch("upcat_pool");
}

extern "C" int stc_upcat_apply_bwd(const void* dout, const void* dyhw, void* dskip, void* dlow, int N, int H, int W, int Cs, int h, int w, int Cu,
                                   int align_corners, int dtype, void* stream) {
    STC_REQUIRE(stc_upcat_fused_ok(N, H, W, Cs, h, w, Cu), "upcat_apply_bwd: unsupported shape");
    STC_REQUIRE(dout, "upcat_apply_bwd: null pointer");
    if ((long long)N * H * W == 0) return STC_OK;
    UpCatGeom g = make_geom(H, W, Cs, h, w, Cu, align_corners);
    cudaStream_t st = (cudaStream_t)stream;
    if (dskip && Cs > 0) {
        dim3 grid = rows_grid(W, H, N, Cs / 8);
        if (dyhw) {
            STC_DISPATCH_DTYPE(dtype, (upcat_bwd_skip_rows_kernel<T, true><<<grid, 256, 0, st>>>((const T*)dout, (const T*)dyhw, (T*)dskip, g)));
        } else {
            STC_DISPATCH_DTYPE(dtype, (upcat_bwd_skip_rows_kernel<T, false><<<grid, 256, 0, st>>>((const T*)dout, nullptr, (T*)dskip, g)));
        }
    }
    if (dlow) {
        dim3 grid = rows_grid(w, h, N, Cu / 8);
        const size_t sm = sizeof(float) * (size_t)w * (STC_NC + 1);
        STC_REQUIRE(sm <= 48 * 1024, "upcat_apply_bwd: w=%d too wide for the tap table", w);
        if (dyhw) {
            STC_DISPATCH_DTYPE(dtype, (upcat_bwd_low_rows_kernel<T, true><<<grid, 256, sm, st>>>((const T*)dout, (const T*)dyhw, (T*)dlow, g)));
        } else {
            STC_DISPATCH_DTYPE(dtype, (upcat_bwd_low_rows_kernel<T, false><<<grid, 256, sm, st>>>((const T*)dout, nullptr, (T*)dlow, g)));
        }
    }
    return check_launch("upcat_apply_bwd");
}
