// Error plumbing, device checks, layout/packing and elementwise kernels.
#include <stdarg.h>

#include "common.cuh"
#include "reduce.cuh"

namespace stc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return STC_ERR_CUDA;
    }
    return STC_OK;
}

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// How many 2-CTA clusters of `func` (1 CTA per SM at this shared-memory size) the device can hold at once: the persistent pair kernels
// launch exactly that many (a TPC with one SM fused off cannot host a pair; launching num_sms / 2 pairs would then run a second wave).
int max_cta_pairs(const void* func, int threads, size_t smem) {
    struct Entry { const void* f; size_t smem; int dev, n; };
    static Entry cache[32];
    static int used = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    for (int i = 0; i < used; ++i)
        if (cache[i].f == func && cache[i].smem == smem && cache[i].dev == dev) return cache[i].n;
    int n = num_sms() / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * n)); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int q = 0;
    if (cudaOccupancyMaxActiveClusters(&q, func, &cfg) == cudaSuccess && q > 0 && q < n) n = q;
    else cudaGetLastError();
    if (used < 32) cache[used++] = Entry{func, smem, dev, n};
    return n;
}

}  // namespace stc

using namespace stc;

extern "C" const char* stc_last_error(void) { return g_err; }
extern "C" int stc_version(void) { return 100; }
extern "C" int stc_num_sms(void) { return num_sms(); }

extern "C" int stc_check_device(void) {
    int dev = 0, major = 0;
    STC_CUDA(cudaGetDevice(&dev));
    STC_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) {
        set_error("libstc_b200 requires a compute-capability 10.x device (B200, sm_100a); got %d.x — no fallback path", major);
        return STC_ERR_ARCH;
    }
    return STC_OK;
}

// ------------------------------------------------------------------------------------
// layout / packing
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int C, long long HW, int Cpad,
                                    long long total) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;  // over N*HW*Cpad
    if (i >= total) return;
    int c = (int)(i % Cpad);
    long long p = i / Cpad;
    long long n = p / HW, hw = p % HW;
    float v = c < C ? src[(n * C + c) * HW + hw] : 0.f;
    stf(dst + i, v);
}

extern "C" int stc_nchw_to_nhwc(const float* src, void* dst, int N, int C, int H, int W, int Cpad, int dtype, void* stream) {
    STC_REQUIRE(Cpad >= C && N > 0 && C > 0, "nchw_to_nhwc: bad shape");
    long long HW = (long long)H * W, total = (long long)N * HW * Cpad;
    STC_DISPATCH_DTYPE(dtype, (nchw_to_nhwc_kernel<T><<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
                                   src, (T*)dst, C, HW, Cpad, total)));
    return check_launch("nchw_to_nhwc");
}

// Device side of the input pipeline: decoded 8-bit HWC pixels -> normalised NHWC activations.  Replaces the host-side
// Normalize + ImageToTensor/DefaultFormatBundle (mmseg/datasets/pipelines/transforms.py Normalize, formatting.py:179-217): the loader ships
// 3 B/pixel instead of 12.  dst[p][c] = (src[p][swap ? C-1-c : c] - mean[c]) * inv_std[c] for c < C, 0 for the padding channels.
template <typename T>
__global__ void image_u8_kernel(const uint8_t* __restrict__ src, T* __restrict__ dst, const float* __restrict__ mean,
                                const float* __restrict__ inv_std, long long P, int C, int Cpad, int swap_rb) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < P; p += stride) {
        for (int c = 0; c < Cpad; ++c) {
            float v = 0.f;
            if (c < C) v = ((float)src[p * C + (swap_rb ? C - 1 - c : c)] - mean[c]) * inv_std[c];
            stf(dst + p * Cpad + c, v);
        }
    }
}

extern "C" int stc_image_u8_to_nhwc(const uint8_t* src, void* dst, const float* mean, const float* inv_std, long long P, int C, int Cpad,
                                    int swap_rb, int dtype, void* stream) {
    STC_REQUIRE(src && dst && mean && inv_std && C >= 1 && C <= 4 && Cpad >= C, "image_u8_to_nhwc: bad arguments (C=%d, Cpad=%d)", C, Cpad);
    if (P <= 0) return STC_OK;
    int blocks = (int)min((long long)num_sms() * 16, (long long)ceil_div(P, 256));
    STC_DISPATCH_DTYPE(dtype, (image_u8_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, (T*)dst, mean, inv_std, P, C, Cpad, swap_rb)));
    return check_launch("image_u8_to_nhwc");
}

// RandomCrop -> RandomFlip(horizontal) -> Normalize(to_rgb) -> Pad of the training pipeline (my_config/STC-UNet.py:31-37;
// mmseg/datasets/pipelines/transforms.py RandomCrop.crop :610-614, RandomFlip :347-380, Normalize, Pad) on the device, for a batch of
// decoded 8-bit HWC images.  geom[n] = {y0, x0, flip}: the crop window starts at (y0, x0) of source image n (the host draws the offsets
// exactly as RandomCrop.get_crop_bbox does), is then mirrored horizontally when flip != 0; pixels of the H x W output that fall outside
// the source (the Pad step) get pad_val.  One thread per output pixel.
template <typename T>
__global__ void image_u8_crop_flip_kernel(const uint8_t* __restrict__ src, T* __restrict__ dst, const int* __restrict__ geom,
                                          const float* __restrict__ mean, const float* __restrict__ inv_std, int Hs, int Ws, int H, int W, int C,
                                          int Cpad, int swap_rb, float pad_val, long long total) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
        const int x = (int)(i % W);
        const int y = (int)((i / W) % H);
        const long long n = i / ((long long)W * H);
        const int y0 = geom[3 * n], x0 = geom[3 * n + 1], flip = geom[3 * n + 2];
        const int ch = min(H, Hs - y0), cw = min(W, Ws - x0);          // size of the cropped image (before the Pad step)
        const int xs = flip ? cw - 1 - x : x;                          // the flip mirrors the CROPPED image
        const bool inside = y < ch && x < cw;
        const uint8_t* sp = src + ((n * Hs + (y0 + y)) * (long long)Ws + (x0 + xs)) * C;
        for (int c = 0; c < Cpad; ++c) {
            float v = 0.f;
            if (c < C) v = inside ? ((float)sp[swap_rb ? C - 1 - c : c] - mean[c]) * inv_std[c] : pad_val;   // Pad comes AFTER Normalize
            stf(dst + i * Cpad + c, v);
        }
    }
}

__global__ void label_u8_crop_flip_kernel(const uint8_t* __restrict__ src, int64_t* __restrict__ dst, const int* __restrict__ geom, int Hs, int Ws,
                                          int H, int W, int seg_pad_val, long long total) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
        const int x = (int)(i % W);
        const int y = (int)((i / W) % H);
        const long long n = i / ((long long)W * H);
        const int y0 = geom[3 * n], x0 = geom[3 * n + 1], flip = geom[3 * n + 2];
        const int ch = min(H, Hs - y0), cw = min(W, Ws - x0);
        const int xs = flip ? cw - 1 - x : x;
        dst[i] = (y < ch && x < cw) ? (int64_t)src[(n * Hs + (y0 + y)) * (long long)Ws + (x0 + xs)] : (int64_t)seg_pad_val;
    }
}

extern "C" int stc_image_u8_crop_flip_to_nhwc(const uint8_t* src, void* dst, const int* geom, const float* mean, const float* inv_std, int N, int Hs,
                                              int Ws, int H, int W, int C, int Cpad, int swap_rb, float pad_val, int dtype, void* stream) {
    STC_REQUIRE(src && dst && geom && mean && inv_std && C >= 1 && C <= 4 && Cpad >= C && N > 0 && Hs > 0 && Ws > 0 && H > 0 && W > 0,
                "image_u8_crop_flip_to_nhwc: bad arguments");
    const long long total = (long long)N * H * W;
    int blocks = (int)min((long long)num_sms() * 16, (long long)ceil_div(total, 256));
    STC_DISPATCH_DTYPE(dtype, (image_u8_crop_flip_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, (T*)dst, geom, mean, inv_std, Hs, Ws, H, W, C,
                                                                                                    Cpad, swap_rb, pad_val, total)));
    return check_launch("image_u8_crop_flip_to_nhwc");
}

extern "C" int stc_label_u8_crop_flip_i64(const uint8_t* src, int64_t* dst, const int* geom, int N, int Hs, int Ws, int H, int W, int seg_pad_val,
                                          void* stream) {
    STC_REQUIRE(src && dst && geom && N > 0 && Hs > 0 && Ws > 0 && H > 0 && W > 0, "label_u8_crop_flip_i64: bad arguments");
    const long long total = (long long)N * H * W;
    int blocks = (int)min((long long)num_sms() * 16, (long long)ceil_div(total, 256));
    label_u8_crop_flip_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, geom, Hs, Ws, H, W, seg_pad_val, total);
    return check_launch("label_u8_crop_flip_i64");
}

// 8-bit label maps (as stored in the annotation PNGs) -> the int64 maps the loss / histogram kernels index with
__global__ void widen_u8_i64_kernel(const uint8_t* __restrict__ src, int64_t* __restrict__ dst, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

extern "C" int stc_widen_u8_i64(const uint8_t* src, int64_t* dst, long long n, void* stream) {
    if (n <= 0) return STC_OK;
    int blocks = (int)min((long long)num_sms() * 16, (long long)ceil_div(n, 256));
    widen_u8_i64_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, n);
    return check_launch("widen_u8_i64");
}

template <typename T>
__global__ void pack_w_kernel(const float* __restrict__ w, T* __restrict__ dst, int Cout, int Cin, int R, int S,
                              int inner_pad, int tf, long long total) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    int inner = (int)(i % inner_pad);
    long long t = i / inner_pad;
    int outer_n = tf ? Cin : Cout;
    int outer = (int)(t % outer_n);
    int tap = (int)(t / outer_n);
    int r = tap / S, s = tap % S;
    float v = 0.f;
    if (tf == 2) {
        // im2col pack: dst[co][k], k = tap*Cin + ci  (one "tap" of K = inner_pad >= R*S*Cin channels)
        int RS = R * S;
        long long co = i / inner_pad;
        int k = inner;
        if (k < RS * Cin) {
            int tp = k / Cin, ci = k - tp * Cin;
            v = w[((co * Cin + ci) * RS) + tp];
        }
        stf(dst + i, v);
        return;
    }
    if (!tf) {
        // dst[tap][co][ci] = W[co][ci][r][s]
        if (inner < Cin) v = w[(((long long)outer * Cin + inner) * R + r) * S + s];
    } else {
        // dst[(R-1-r)*S+(S-1-s)][ci][co] = W[co][ci][r][s]  <=> reading with flipped tap
        int rr = R - 1 - r, ss = S - 1 - s;
        if (inner < Cout) v = w[(((long long)inner * Cin + outer) * R + rr) * S + ss];
    }
    stf(dst + i, v);
}

extern "C" int stc_pack_conv_weight(const float* w, void* dst, int Cout, int Cin, int R, int S, int inner_pad,
                                    int transpose_flip, int dtype, void* stream) {
    int inner = transpose_flip == 1 ? Cout : (transpose_flip == 2 ? R * S * Cin : Cin);
    int outer = transpose_flip == 1 ? Cin : Cout;
    STC_REQUIRE(inner_pad >= inner, "pack_conv_weight: inner_pad %d < inner %d", inner_pad, inner);
    long long total = transpose_flip == 2 ? (long long)Cout * inner_pad : (long long)R * S * outer * inner_pad;
    STC_DISPATCH_DTYPE(dtype, (pack_w_kernel<T><<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
                                   w, (T*)dst, Cout, Cin, R, S, inner_pad, transpose_flip, total)));
    return check_launch("pack_conv_weight");
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ ws, float* __restrict__ dw, int Cout, int Cin, int RS,
                                    int accumulate, long long total) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;  // over OIHW
    if (i >= total) return;
    int tap = (int)(i % RS);
    long long t = i / RS;
    int ci = (int)(t % Cin);
    int co = (int)(t / Cin);
    float v = ws[((long long)tap * Cin + ci) * Cout + co];
    dw[i] = accumulate ? dw[i] + v : v;
}

extern "C" int stc_unpack_conv_wgrad(const float* ws, float* dw, int Cout, int Cin, int R, int S, int accumulate, void* stream) {
    long long total = (long long)Cout * Cin * R * S;
    unpack_wgrad_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(ws, dw, Cout, Cin, R * S, accumulate, total);
    return check_launch("unpack_conv_wgrad");
}

// ------------------------------------------------------------------------------------
// elementwise
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long long n8, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = i; j < n8; j += stride) {
        Vec8<T> x, y;
        x.load(a + j * 8);
        y.load(b + j * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) x.v[k] += y.v[k];
        x.store(out + j * 8);
    }
    for (long long j = n8 * 8 + i; j < n; j += stride) stf(out + j, ldf(a + j) + ldf(b + j));
}

extern "C" int stc_add(const void* a, const void* b, void* out, long long n, int dtype, void* stream) {
    if (n <= 0) return STC_OK;
    bool aligned = (((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) == 0;
    long long n8 = aligned ? n / 8 : 0;
    int blocks = (int)min((long long)num_sms() * 16, (long long)ceil_div(max(n8, 1LL), 256));
    if (blocks < 1) blocks = 1;
    STC_DISPATCH_DTYPE(dtype, (add_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>((const T*)a, (const T*)b, (T*)out, n8, n)));
    return check_launch("add");
}

__global__ void axpy_kernel(const float* __restrict__ x, float* __restrict__ y, float alpha, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) y[i] += alpha * x[i];
}

extern "C" int stc_axpy_f32(const float* x, float* y, float alpha, long long n, void* stream) {
    if (n <= 0) return STC_OK;
    int blocks = (int)min((long long)num_sms() * 8, (long long)ceil_div(n, 256));
    axpy_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, y, alpha, n);
    return check_launch("axpy");
}

template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ s, TD* __restrict__ d, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) stf(d + i, ldf(s + i));
}

extern "C" int stc_cast(const void* src, void* dst, long long n, int sd, int dd, void* stream) {
    if (n <= 0) return STC_OK;
    int blocks = (int)min((long long)num_sms() * 16, (long long)ceil_div(n, 256));
    cudaStream_t st = (cudaStream_t)stream;
    if (sd == STC_F32 && dd == STC_BF16) cast_kernel<float, bf16><<<blocks, 256, 0, st>>>((const float*)src, (bf16*)dst, n);
    else if (sd == STC_BF16 && dd == STC_F32) cast_kernel<bf16, float><<<blocks, 256, 0, st>>>((const bf16*)src, (float*)dst, n);
    else if (sd == STC_F32 && dd == STC_F32) cast_kernel<float, float><<<blocks, 256, 0, st>>>((const float*)src, (float*)dst, n);
    else if (sd == STC_BF16 && dd == STC_BF16) cast_kernel<bf16, bf16><<<blocks, 256, 0, st>>>((const bf16*)src, (bf16*)dst, n);
    else { set_error("cast: bad dtypes"); return STC_ERR_INVALID; }
    return check_launch("cast");
}

template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ yo, const T* __restrict__ dy, T* __restrict__ dx, long long n, int act) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float o = ldf(yo + i), g = ldf(dy + i), r;
        if (act == STC_ACT_SIGMOID) r = g * o * (1.f - o);
        else if (act == STC_ACT_RELU) r = o > 0.f ? g : 0.f;
        else r = g;
        stf(dx + i, r);
    }
}

extern "C" int stc_act_bwd(const void* y_out, const void* dy, void* dx, long long n, int act, int dtype, void* stream) {
    if (n <= 0) return STC_OK;
    STC_REQUIRE(act == STC_ACT_SIGMOID || act == STC_ACT_RELU || act == STC_ACT_NONE, "act_bwd: unsupported act %d", act);
    int blocks = (int)min((long long)num_sms() * 16, (long long)ceil_div(n, 256));
    STC_DISPATCH_DTYPE(dtype, (act_bwd_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>((const T*)y_out, (const T*)dy, (T*)dx, n, act)));
    return check_launch("act_bwd");
}

template <typename T>
__global__ void scale_channels_kernel(const T* __restrict__ x, const float* __restrict__ m, T* __restrict__ y, long long HW,
                                      int C, long long total) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int c = (int)(i % C);
        long long n = i / ((long long)C * HW);
        stf(y + i, ldf(x + i) * m[n * C + c]);
    }
}

extern "C" int stc_scale_channels(const void* x, const float* m, void* y, int N, long long HW, int C, int dtype, void* stream) {
    long long total = (long long)N * HW * C;
    if (total <= 0) return STC_OK;
    int blocks = (int)min((long long)num_sms() * 16, (long long)ceil_div(total, 256));
    STC_DISPATCH_DTYPE(dtype, (scale_channels_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>((const T*)x, m, (T*)y, HW, C, total)));
    return check_launch("scale_channels");
}

// column sums out[c] = sum_p x[p][c]; block = 256 threads = 8 row-lanes x 32 channel lanes
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, float* __restrict__ out, long long P, int C) {
    __shared__ float red[8][33];
    int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    int c = blockIdx.x * 32 + cx;
    float acc = 0.f;
    if (c < C) {
        long long rows_per = (P + gridDim.y - 1) / gridDim.y;
        long long p0 = blockIdx.y * rows_per, p1 = min(P, p0 + rows_per);
        for (long long p = p0 + ry; p < p1; p += 8) acc += ldf(x + p * C + c);
    }
    red[ry][cx] = acc;
    __syncthreads();
    if (ry == 0 && c < C) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += red[k][cx];
        atomicAdd(out + c, s);
    }
}

// vector path: a thread keeps one 8-channel lane and walks rows with two 16-byte loads in flight; one atomicAdd per (block, channel)
// Cc = channels of this column chunk (Cc/8 a power of two), ld = row stride; blockIdx.y selects the chunk, so widths such as
// 1536 = 3 x 512 (the folded q|k|v projection) keep the 16-byte row-structured path.  Four loads in flight per thread.
template <typename T>
__global__ void __launch_bounds__(256) colsum_rows_kernel(const T* __restrict__ x, float* __restrict__ out, long long P, int Cc, int ld) {
    __shared__ float smem[256 * 8];
    const int lanes = Cc >> 3, lv = threadIdx.x % lanes, r0 = threadIdx.x / lanes, rstep = 256 / lanes;
    x += (long long)blockIdx.y * Cc;
    out += (long long)blockIdx.y * Cc;
    float acc[1][8] = {};
    float acc1[8] = {}, acc2[8] = {}, acc3[8] = {};
    const long long step = (long long)gridDim.x * rstep;
    long long p = (long long)blockIdx.x * rstep + r0;
    for (; p + 3 * step < P; p += 4 * step) {
        Vec8<T> v0, v1, v2, v3;
        v0.load(x + p * ld + lv * 8);
        v1.load(x + (p + step) * ld + lv * 8);
        v2.load(x + (p + 2 * step) * ld + lv * 8);
        v3.load(x + (p + 3 * step) * ld + lv * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc[0][k] += v0.v[k]; acc1[k] += v1.v[k]; acc2[k] += v2.v[k]; acc3[k] += v3.v[k]; }
    }
    for (; p < P; p += step) {
        Vec8<T> v0;
        v0.load(x + p * ld + lv * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[0][k] += v0.v[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[0][k] += acc1[k] + acc2[k] + acc3[k];
    block_reduce_lanes_emit<1>(acc, lanes, smem, Cc, [&](int, int c, float s) { atomicAdd(out + c, s); });
}

// largest chunk width (8 * 2^k channels, at most 2048) that divides C; 0 if C is not a multiple of 8
static inline int colsum_chunk(int C) {
    if (C % 8) return 0;
    int c = 8;
    while (c * 2 <= 2048 && C % (c * 2) == 0) c *= 2;
    return c;
}

extern "C" int stc_colsum(const void* x, float* out, long long P, int C, int accumulate, int dtype, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!accumulate) STC_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * C, st));
    if (P <= 0) return STC_OK;
    const int Cc = colsum_chunk(C);
    if (Cc >= 64 && (((uintptr_t)x) & 15) == 0 && P >= 1024) {
        const int chunks = C / Cc;
        int gx = reduce_blocks(P, Cc / 8) / chunks;
        dim3 grid(gx < 1 ? 1 : gx, chunks);
        STC_DISPATCH_DTYPE(dtype, (colsum_rows_kernel<T><<<grid, 256, 0, st>>>((const T*)x, out, P, Cc, C)));
        return check_launch("colsum");
    }
    int gx = ceil_div(C, 32);
    int gy = (int)max(1LL, min((long long)ceil_div(P, 64), (long long)(num_sms() * 8 / gx + 1)));
    dim3 grid(gx, gy);
    STC_DISPATCH_DTYPE(dtype, (colsum_kernel<T><<<grid, 256, 0, st>>>((const T*)x, out, P, C)));
    return check_launch("colsum");
}

// ------------------------------------------------------------------------------------
// Adam (torch.optim.Adam, no amsgrad)
// ------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float gi = g[i], pi = p[i];
        if (wd != 0.f) gi += wd * pi;
        float mi = b1 * m[i] + (1.f - b1) * gi;
        float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - (lr / bc1) * (mi / denom);
    }
}

extern "C" int stc_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                             float eps, float weight_decay, int step, void* stream) {
    if (n <= 0) return STC_OK;
    STC_REQUIRE(step >= 1, "adam: step must be >= 1");
    float bc1 = 1.f - powf(beta1, (float)step);
    float bc2s = sqrtf(1.f - powf(beta2, (float)step));
    int blocks = (int)min((long long)num_sms() * 16, (long long)ceil_div(n, 256));
    adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2s);
    return check_launch("adam");
}

// Graph-capturable variant: the step counter and the learning rate live in device memory, so a captured training step can be
// replayed while the bias corrections and an LR schedule keep advancing.  dyn = {lr, beta1^t, beta2^t} (fp32), updated by the
// tick kernel that precedes the update.
__global__ void adam_tick_kernel(float* __restrict__ dyn, int* __restrict__ step, float b1, float b2) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const int t = *step + 1;
        *step = t;
        dyn[1] = powf(b1, (float)t);
        dyn[2] = powf(b2, (float)t);
    }
}

__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                long long n, const float* __restrict__ dyn, float b1, float b2, float eps, float wd) {
    const float lr = dyn[0], bc1 = 1.f - dyn[1], bc2_sqrt = sqrtf(1.f - dyn[2]);
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float gi = g[i], pi = p[i];
        if (wd != 0.f) gi += wd * pi;
        float mi = b1 * m[i] + (1.f - b1) * gi;
        float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - (lr / bc1) * (mi / denom);
    }
}

extern "C" int stc_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, float* dyn, int* step, float beta1, float beta2,
                                 float eps, float weight_decay, void* stream) {
    if (n <= 0) return STC_OK;
    STC_REQUIRE(dyn && step, "adam_step_dev: null state pointer");
    cudaStream_t st = (cudaStream_t)stream;
    adam_tick_kernel<<<1, 32, 0, st>>>(dyn, step, beta1, beta2);
    int blocks = (int)min((long long)num_sms() * 16, (long long)ceil_div(n, 256));
    adam_dev_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, n, dyn, beta1, beta2, eps, weight_decay);
    return check_launch("adam_dev");
}

// ------------------------------------------------------------------------------------
// row-block copy between (N, rows, C) tensors: dst[n, dst_off + r, :] = src[n, src_off + r, :], r < count
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void copy_rows_kernel(const T* __restrict__ src, T* __restrict__ dst, int src_rows, int dst_rows, int C, int src_off,
                                 int dst_off, int count, long long total) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int c = (int)(i % C);
        long long t = i / C;
        int r = (int)(t % count);
        long long n = t / count;
        dst[(n * dst_rows + dst_off + r) * C + c] = src[(n * src_rows + src_off + r) * C + c];
    }
}

extern "C" int stc_copy_rows(const void* src, void* dst, int N, int src_rows, int dst_rows, int C, int src_off, int dst_off,
                             int count, int dtype, void* stream) {
    STC_REQUIRE(src_off >= 0 && dst_off >= 0 && src_off + count <= src_rows && dst_off + count <= dst_rows, "copy_rows: range");
    long long total = (long long)N * count * C;
    if (total <= 0) return STC_OK;
    int blocks = (int)min((long long)num_sms() * 8, (long long)ceil_div(total, 256));
    STC_DISPATCH_DTYPE(dtype, (copy_rows_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>((const T*)src, (T*)dst, src_rows, dst_rows, C,
                                                                                           src_off, dst_off, count, total)));
    return check_launch("copy_rows");
}


// ------------------------------------------------------------------------------------
// im2col for small-Cin convolutions (the 3-channel image conv): out[p][k], k = tap*Cin + ci, zero padded to Kpad,
// so the layer runs as a K = Kpad 1x1 convolution on the tensor cores.
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void im2col_kernel(const T* __restrict__ x, T* __restrict__ out, int H, int W, int Cin, int R, int S, int Kpad, long long total) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;  // over P * Kpad/8
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int kv = Kpad >> 3, K = R * S * Cin, pr = R / 2, ps = S / 2;
    for (; i < total; i += stride) {
        int v8 = (int)(i % kv);
        long long p = i / kv;
        int w_ = (int)(p % W), h_ = (int)((p / W) % H);
        long long n = p / ((long long)W * H);
        Vec8<T> o;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            int k = v8 * 8 + e;
            float v = 0.f;
            if (k < K) {
                int tp = k / Cin, ci = k - tp * Cin;
                int hh = h_ + tp / S - pr, ww = w_ + tp % S - ps;
                if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = ldf(x + ((n * H + hh) * W + ww) * Cin + ci);
            }
            o.v[e] = v;
        }
        o.store(out + i * 8);
    }
}

// one block per image row: the R input rows it needs are staged in shared memory (zero halo included), a per-k offset table replaces
// the div/mod chain, and the (W x Kpad) output row is written with fully coalesced 16-byte stores
template <typename T>
__global__ void __launch_bounds__(256) im2col_rows_kernel(const T* __restrict__ x, T* __restrict__ out, int H, int W, int Cin, int R, int S,
                                                          int Kpad) {
    extern __shared__ __align__(16) unsigned char im2col_smem[];
    int* koff = reinterpret_cast<int*>(im2col_smem);
    T* tile = reinterpret_cast<T*>(koff + Kpad);
    const int rowlen = (W + S - 1) * Cin, K = R * S * Cin, pr = R / 2, ps = S / 2;
    const long long nh = blockIdx.x, n = nh / H;
    const int h_ = (int)(nh % H);
    for (int k = threadIdx.x; k < Kpad; k += 256) {
        int tp = k / Cin, ci = k - tp * Cin;
        koff[k] = k < K ? (tp / S) * rowlen + (tp % S) * Cin + ci : -1;
    }
    for (int idx = threadIdx.x; idx < R * rowlen; idx += 256) {
        const int r = idx / rowlen, c = idx - r * rowlen;
        const int hh = h_ + r - pr, cc = c - ps * Cin;
        float v = 0.f;
        if (hh >= 0 && hh < H && cc >= 0 && cc < W * Cin) v = ldf(x + ((n * H + hh) * (long long)W) * Cin + cc);
        stf(tile + idx, v);
    }
    __syncthreads();
    const int kv = Kpad >> 3;
    T* orow = out + nh * (long long)W * Kpad;
    for (int item = threadIdx.x; item < W * kv; item += 256) {
        const int w_ = item / kv, v8 = item - w_ * kv;
        Vec8<T> o;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int off = koff[v8 * 8 + e];
            o.v[e] = off >= 0 ? ldf(tile + off + w_ * Cin) : 0.f;
        }
        o.store(orow + (long long)item * 8);
    }
}

extern "C" int stc_im2col(const void* x, void* out, int N, int H, int W, int Cin, int R, int S, int Kpad, int dtype, void* stream) {
    STC_REQUIRE(Kpad % 8 == 0 && Kpad >= R * S * Cin, "im2col: Kpad=%d must be a multiple of 8 and >= R*S*Cin=%d", Kpad, R * S * Cin);
    long long total = (long long)N * H * W * (Kpad / 8);
    if (total <= 0) return STC_OK;
    const size_t smem = sizeof(int) * Kpad + (size_t)R * (W + S - 1) * Cin * (dtype == STC_BF16 ? 2 : 4);
    if (smem <= 48 * 1024) {
        STC_DISPATCH_DTYPE(dtype, (im2col_rows_kernel<T><<<(unsigned)((long long)N * H), 256, smem, (cudaStream_t)stream>>>((const T*)x, (T*)out, H, W,
                                                                                                                       Cin, R, S, Kpad)));
        return check_launch("im2col");
    }
    int blocks = (int)min((long long)num_sms() * 16, (long long)ceil_div(total, 256));
    STC_DISPATCH_DTYPE(dtype, (im2col_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)out, H, W, Cin, R, S, Kpad, total)));
    return check_launch("im2col");
}

__global__ void unpack_im2col_wgrad_kernel(const float* __restrict__ ws, float* __restrict__ dw, int Cout, int Cin, int RS, long long total) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;  // over OIHW
    if (i >= total) return;
    int tp = (int)(i % RS);
    long long t = i / RS;
    int ci = (int)(t % Cin);
    int co = (int)(t / Cin);
    dw[i] = ws[((long long)tp * Cin + ci) * Cout + co];   // ws is [Kpad][Cout], k = tap*Cin + ci
}

extern "C" int stc_unpack_im2col_wgrad(const float* ws, float* dw, int Cout, int Cin, int R, int S, void* stream) {
    long long total = (long long)Cout * Cin * R * S;
    unpack_im2col_wgrad_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(ws, dw, Cout, Cin, R * S, total);
    return check_launch("unpack_im2col_wgrad");
}


// out = a + b (+ c (+ d)) in one pass: fan-out gradient sums (KSA input: 4 consumers, transformer tokens: 4)
template <typename T>
__global__ void add_n_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ c, const T* __restrict__ d,
                             T* __restrict__ out, long long n8, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = i; j < n8; j += stride) {
        Vec8<T> x, y;
        x.load(a + j * 8);
        y.load(b + j * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) x.v[k] += y.v[k];
        if (c) {
            y.load(c + j * 8);
#pragma unroll
            for (int k = 0; k < 8; ++k) x.v[k] += y.v[k];
        }
        if (d) {
            y.load(d + j * 8);
#pragma unroll
            for (int k = 0; k < 8; ++k) x.v[k] += y.v[k];
        }
        x.store(out + j * 8);
    }
    for (long long j = n8 * 8 + i; j < n; j += stride) {
        float v = ldf(a + j) + ldf(b + j);
        if (c) v += ldf(c + j);
        if (d) v += ldf(d + j);
        stf(out + j, v);
    }
}

extern "C" int stc_add_n(const void* a, const void* b, const void* c, const void* d, void* out, long long n, int dtype, void* stream) {
    if (n <= 0) return STC_OK;
    STC_REQUIRE(a && b, "add_n: the first two inputs are mandatory");
    bool aligned = (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d | (uintptr_t)out) & 15) == 0;
    long long n8 = aligned ? n / 8 : 0;
    int blocks = (int)min((long long)num_sms() * 16, (long long)ceil_div(max(n8, 1LL), 256));
    if (blocks < 1) blocks = 1;
    STC_DISPATCH_DTYPE(dtype, (add_n_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>((const T*)a, (const T*)b, (const T*)c, (const T*)d,
                                                                                     (T*)out, n8, n)));
    return check_launch("add_n");
}


// ------------------------------------------------------------------------------------
// all weight packs of a training step in ONE launch.  table: n rows of 8 int64
// {src ptr, dst element offset, Cout, Cin, R, S, inner_pad, mode}; prefix: n+1 cumulative element counts.
// ------------------------------------------------------------------------------------
// Work item = one (co, ci) pair: the thread reads that pair's R*S contiguous source floats once and writes one element per tap.
// mode 0 (fprop, [tap][Cout][Cin]): ci fastest -> coalesced reads and writes; mode 1 (dgrad, [tap'][Cin][Cout], flipped taps): co fastest ->
// coalesced writes, each thread's R*S-float source run is consumed through L1.  mode 2 (im2col, [Cout][Kpad]): item = one output element.
// prefix[] counts ITEMS (Cout*Cin for modes 0/1, Cout*Kpad for mode 2).
template <typename T>
__global__ void pack_batched_kernel(const long long* __restrict__ table, const long long* __restrict__ prefix, int n,
                                    T* __restrict__ dst_base, long long total) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int lo = 0, hi = n;  // prefix[lo] <= i < prefix[hi]
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (prefix[mid] <= i) lo = mid; else hi = mid;
        }
        const long long* e = table + (long long)lo * 8;
        const float* w = reinterpret_cast<const float*>(e[0]);
        const int Cout = (int)e[2], Cin = (int)e[3], R = (int)e[4], S = (int)e[5], inner_pad = (int)e[6], tf = (int)e[7];
        const long long j = i - prefix[lo];
        T* dst = dst_base + e[1];
        const int RS = R * S;
        if (tf == 2) {
            const int inner = (int)(j % inner_pad);
            const long long co = j / inner_pad;
            float v = 0.f;
            if (inner < RS * Cin) {
                int tp = inner / Cin, ci = inner - tp * Cin;
                v = w[((co * Cin + ci) * RS) + tp];
            }
            stf(dst + j, v);
        } else if (tf == 0) {
            const int ci = (int)(j % Cin), co = (int)(j / Cin);
            const float* src = w + (long long)j * RS;                  // (co*Cin + ci) * RS
            const long long plane = (long long)Cout * Cin;
            for (int tap = 0; tap < RS; ++tap) stf(dst + tap * plane + (long long)co * Cin + ci, src[tap]);
        } else {
            const int co = (int)(j % Cout), ci = (int)(j / Cout);
            const float* src = w + ((long long)co * Cin + ci) * RS;
            const long long plane = (long long)Cout * Cin;
            for (int tap = 0; tap < RS; ++tap) stf(dst + (RS - 1 - tap) * plane + (long long)ci * Cout + co, src[tap]);
        }
    }
}

extern "C" int stc_pack_conv_weights_batched(const int64_t* table, const int64_t* prefix, int n, void* dst, long long total, int dtype,
                                             void* stream) {
    if (n <= 0 || total <= 0) return STC_OK;
    int blocks = (int)min((long long)num_sms() * 16, (long long)ceil_div(total, 256));
    STC_DISPATCH_DTYPE(dtype, (pack_batched_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>((const long long*)table, (const long long*)prefix,
                                                                                            n, (T*)dst, total)));
    return check_launch("pack_conv_weights_batched");
}

// ------------------------------------------------------------------------------------
// channel padding: y[p][0:Cdst) = x[p][0:min(Csrc, Cdst)), zero beyond Csrc.  Narrow layers (16 / 32 channels: UNet++'s decoder tail)
// are widened to the 64-channel granularity of the tcgen05 kernels' K chunks instead of falling back to the SIMT engine.
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void resize_channels_kernel(const T* __restrict__ x, T* __restrict__ y, int Csrc, int Cdst, long long total) {
    const int ld = Cdst >> 3, ls = Csrc >> 3;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const int lv = (int)(i % ld);
        const long long p = i / ld;
        Vec8<T> v;
        if (lv < ls) v.load(x + p * Csrc + lv * 8);
        else {
#pragma unroll
            for (int k = 0; k < 8; ++k) v.v[k] = 0.f;
        }
        v.store(y + i * 8);
    }
}

extern "C" int stc_resize_channels(const void* x, void* y, long long P, int Csrc, int Cdst, int dtype, void* stream) {
    STC_REQUIRE(Csrc > 0 && Cdst > 0 && Csrc % 8 == 0 && Cdst % 8 == 0, "resize_channels: channel counts must be multiples of 8 (%d -> %d)", Csrc, Cdst);
    const long long total = P * (Cdst / 8);
    if (total <= 0) return STC_OK;
    int blocks = (int)min((long long)num_sms() * 8, (long long)ceil_div(total, 256));
    STC_DISPATCH_DTYPE(dtype, (resize_channels_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, Csrc, Cdst, total)));
    return check_launch("resize_channels");
}

// ------------------------------------------------------------------------------------
// channel concat / split on NHWC rows: out[p] = [a[p] (Ca) | b[p] (Cb)]   (UpConvBlock.forward's torch.cat, up_conv_block.py:99)
// ------------------------------------------------------------------------------------
template <typename T, bool SPLIT>
__global__ void concat_channels_kernel(T* __restrict__ a, T* __restrict__ b, T* __restrict__ cat, int Ca, int Cb, long long total) {
    const int la = Ca >> 3, lt = (Ca + Cb) >> 3;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int lv = (int)(i % lt);
        long long p = i / lt;
        T* side = lv < la ? a + p * Ca + lv * 8 : b + p * Cb + (lv - la) * 8;
        Vec8<T> v;
        if (SPLIT) {
            if ((lv < la && a) || (lv >= la && b)) { v.load(cat + i * 8); v.store(side); }
        } else {
            v.load(side);
            v.store(cat + i * 8);
        }
    }
}

extern "C" int stc_concat_channels(const void* a, const void* b, void* out, long long P, int Ca, int Cb, int dtype, void* stream) {
    STC_REQUIRE(Ca % 8 == 0 && Cb % 8 == 0, "concat_channels: channel counts must be multiples of 8");
    long long total = P * ((Ca + Cb) / 8);
    if (total <= 0) return STC_OK;
    int blocks = (int)min((long long)num_sms() * 8, (long long)ceil_div(total, 256));
    STC_DISPATCH_DTYPE(dtype, (concat_channels_kernel<T, false><<<blocks, 256, 0, (cudaStream_t)stream>>>((T*)a, (T*)b, (T*)out, Ca, Cb, total)));
    return check_launch("concat_channels");
}

extern "C" int stc_split_channels(const void* cat, void* a, void* b, long long P, int Ca, int Cb, int dtype, void* stream) {
    STC_REQUIRE(Ca % 8 == 0 && Cb % 8 == 0, "split_channels: channel counts must be multiples of 8");
    long long total = P * ((Ca + Cb) / 8);
    if (total <= 0) return STC_OK;
    int blocks = (int)min((long long)num_sms() * 8, (long long)ceil_div(total, 256));
    STC_DISPATCH_DTYPE(dtype, (concat_channels_kernel<T, true><<<blocks, 256, 0, (cudaStream_t)stream>>>((T*)a, (T*)b, (T*)cat, Ca, Cb, total)));
    return check_launch("split_channels");
}
