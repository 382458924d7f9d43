// Classifier (1x1 conv to NCHW fp32 logits), fused CE + Dice + accuracy, inference post-processing
// and the integer confusion-matrix histogram.
#include "reduce.cuh"

namespace stc {

constexpr int kMaxCls = 32;

// ---------------------------------------------------------------- cls_seg
template <typename T>
__global__ void __launch_bounds__(256) cls_fwd_kernel(const T* __restrict__ x, const float* __restrict__ Wt, const float* __restrict__ b,
                                                      const float* __restrict__ mask, float* __restrict__ logits, long long HW, int Cin,
                                                      int Ccls, long long P) {
    extern __shared__ float sw[];  // [Ccls][Cin] + [Ccls]
    for (int i = threadIdx.x; i < Ccls * Cin; i += blockDim.x) sw[i] = Wt[i];
    for (int i = threadIdx.x; i < Ccls; i += blockDim.x) sw[Ccls * Cin + i] = b ? b[i] : 0.f;
    __syncthreads();
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= P) return;
    long long n = p / HW, hw = p - n * HW;
    const T* xr = x + p * Cin;
    for (int k0 = 0; k0 < Ccls; k0 += 4) {
        float acc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = (k0 + j < Ccls) ? sw[Ccls * Cin + k0 + j] : 0.f;
        for (int c = 0; c < Cin; c += 8) {
            Vec8<T> v;
            v.load(xr + c);
            if (mask) {  // Dropout2d: one factor per (image, channel)
#pragma unroll
                for (int e = 0; e < 8; ++e) v.v[e] *= mask[n * Cin + c + e];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (k0 + j < Ccls) {
                    const float* wr = sw + (k0 + j) * Cin + c;
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[j] = fmaf(v.v[e], wr[e], acc[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (k0 + j < Ccls) logits[(n * Ccls + k0 + j) * HW + hw] = acc[j];
    }
}

// thread = (pixel, 8-channel lane): a warp stores 32/lanes pixels x lanes*16 B as one contiguous run (a thread-per-pixel mapping
// writes 16 B into 32 different lines per store instruction).  grid (G, N): one image per blockIdx.y, its Dropout2d mask folded
// into the shared-memory copy of the weights.
template <typename T>
__global__ void __launch_bounds__(256) cls_bwd_dx_kernel(const float* __restrict__ dl, const float* __restrict__ Wt,
                                                         const float* __restrict__ mask, T* __restrict__ dx, long long HW, int Cin,
                                                         int Ccls) {
    extern __shared__ float sw[];  // [Ccls][Cin]
    const long long n = blockIdx.y;
    for (int i = threadIdx.x; i < Ccls * Cin; i += blockDim.x) sw[i] = Wt[i] * (mask ? mask[n * Cin + i % Cin] : 1.f);
    __syncthreads();
    const int lanes = Cin >> 3, lv = threadIdx.x % lanes, r0 = threadIdx.x / lanes, rstep = 256 / lanes;
    const float* dn = dl + n * Ccls * HW;
    T* on = dx + n * HW * Cin + lv * 8;
    const float* wl = sw + lv * 8;
    const long long step = (long long)gridDim.x * rstep;
    long long hw = (long long)blockIdx.x * rstep + r0;
    for (; hw + step < HW; hw += 2 * step) {      // two pixels per iteration
        Vec8<T> v0, v1;
#pragma unroll
        for (int e = 0; e < 8; ++e) v0.v[e] = v1.v[e] = 0.f;
        for (int k = 0; k < Ccls; ++k) {
            const float g0 = dn[k * HW + hw], g1 = dn[k * HW + hw + step];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float w = wl[k * Cin + e];
                v0.v[e] = fmaf(g0, w, v0.v[e]);
                v1.v[e] = fmaf(g1, w, v1.v[e]);
            }
        }
        v0.store(on + hw * Cin);
        v1.store(on + (hw + step) * Cin);
    }
    for (; hw < HW; hw += step) {
        Vec8<T> v;
#pragma unroll
        for (int e = 0; e < 8; ++e) v.v[e] = 0.f;
        for (int k = 0; k < Ccls; ++k) {
            const float g = dn[k * HW + hw];
#pragma unroll
            for (int e = 0; e < 8; ++e) v.v[e] = fmaf(g, wl[k * Cin + e], v.v[e]);
        }
        v.store(on + hw * Cin);
    }
}

// dW[k][c] += sum_p dl[k][p] x[p][c] for 4 classes starting at k0; db likewise.  grid (G, N): one image per blockIdx.y, so the pixel
// loop is a pointer walk and the Dropout2d mask (per image and channel) is applied once, after the sum.
template <typename T>
__global__ void __launch_bounds__(256) cls_bwd_dw_kernel(const float* __restrict__ dl, const T* __restrict__ x, const float* __restrict__ mask,
                                                         float* __restrict__ dW, float* __restrict__ db, long long HW, int Cin, int Ccls,
                                                         int k0) {
    __shared__ float smem[256 * 8 * 4];
    const int lanes = Cin >> 3, lv = threadIdx.x % lanes, r0 = threadIdx.x / lanes, rstep = 256 / lanes;
    const long long n = blockIdx.y;
    float acc[4][8] = {};
    float bs[4] = {};
    const T* xn = x + n * HW * Cin + lv * 8;
    const float* dn = dl + (n * Ccls + k0) * HW;
    const int nk = min(4, Ccls - k0);
    const long long step = (long long)gridDim.x * rstep;
    long long hw = (long long)blockIdx.x * rstep + r0;
    for (; hw + step < HW; hw += 2 * step) {      // two rows per iteration: two 16-byte loads in flight per thread
        Vec8<T> v0, v1;
        v0.load(xn + hw * Cin);
        v1.load(xn + (hw + step) * Cin);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < nk) {
                const float g0 = dn[j * HW + hw], g1 = dn[j * HW + hw + step];
                bs[j] += g0 + g1;
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[j][e] = fmaf(g0, v0.v[e], fmaf(g1, v1.v[e], acc[j][e]));
            }
        }
    }
    for (; hw < HW; hw += step) {
        Vec8<T> v;
        v.load(xn + hw * Cin);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < nk) {
                const float g = dn[j * HW + hw];
                bs[j] += g;
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[j][e] = fmaf(g, v.v[e], acc[j][e]);
            }
        }
    }
    if (mask) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float m = mask[n * Cin + lv * 8 + e];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j][e] *= m;
        }
    }
    block_reduce_lanes_emit<4>(acc, lanes, smem, Cin, [&](int q, int c, float s) {
        if (k0 + q < Ccls) atomicAdd(dW + (long long)(k0 + q) * Cin + c, s);
    });
    if (db) {   // block-level sum first: one global atomic per (block, class) instead of one per pixel-row thread
        __shared__ float sb[4];
        __syncthreads();
        if (threadIdx.x < 4) sb[threadIdx.x] = 0.f;
        __syncthreads();
        if (lv == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(&sb[j], bs[j]);
        }
        __syncthreads();
        if (threadIdx.x < 4 && k0 + (int)threadIdx.x < Ccls) atomicAdd(db + k0 + threadIdx.x, sb[threadIdx.x]);
    }
}

// ---------------------------------------------------------------- CE + Dice + accuracy
// stats layout (fp64): [n][c][3] = {sum p*t*m, sum p^2, sum t}; tail [4] = {ce_sum, n_correct, n_valid, unused}
template <int CM>
__global__ void __launch_bounds__(256) seg_loss_fwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ label,
                                                           double* __restrict__ stats, long long HW, int C, int N, int ignore) {
    const long long n = blockIdx.y;
    const float* lg = logits + n * C * HW;
    const int64_t* lb = label + n * HW;
    float spt[CM], sp2[CM], st[CM];
#pragma unroll
    for (int k = 0; k < CM; ++k) spt[k] = sp2[k] = st[k] = 0.f;
    float ce = 0.f, correct = 0.f, nvalid = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
        const long long y = lb[i];
        const bool valid = y != ignore;
        const int yc = (int)(y < 0 ? 0 : (y > C - 1 ? C - 1 : y));
        float z[CM];
        float m = -INFINITY, zy = 0.f;
        int arg = 0;
#pragma unroll
        for (int k = 0; k < CM; ++k) {
            if (k < C) {
                z[k] = lg[k * HW + i];
                if (z[k] > m) { m = z[k]; arg = k; }
                if (k == yc) zy = z[k];
            }
        }
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < CM; ++k)
            if (k < C) { z[k] = expf(z[k] - m); sum += z[k]; }
        const float inv = 1.f / sum;
#pragma unroll
        for (int k = 0; k < CM; ++k) {
            if (k < C) {
                float p = z[k] * inv;
                sp2[k] = fmaf(p, p, sp2[k]);
                if (k == yc) {
                    st[k] += 1.f;
                    if (valid) spt[k] += p;
                }
            }
        }
        if (valid) {
            nvalid += 1.f;
            if (arg == (int)y) correct += 1.f;
            ce += logf(sum) + (m - zy);  // -log softmax[y]
        }
    }
    // block reduction of 3*C + 3 values
    __shared__ float red[8][3 * CM + 3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < CM; ++k) {
        float a = warp_sum(spt[k]), b = warp_sum(sp2[k]), c = warp_sum(st[k]);
        if (lane == 0) { red[warp][3 * k] = a; red[warp][3 * k + 1] = b; red[warp][3 * k + 2] = c; }
    }
    {
        float a = warp_sum(ce), b = warp_sum(correct), c = warp_sum(nvalid);
        if (lane == 0) { red[warp][3 * CM] = a; red[warp][3 * CM + 1] = b; red[warp][3 * CM + 2] = c; }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 3 * CM + 3; j += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][j];
        if (j < 3 * CM) {
            int k = j / 3;
            if (k < C) atomicAdd(stats + (n * C + k) * 3 + (j - 3 * k), (double)s);
        } else {
            atomicAdd(stats + (long long)N * C * 3 + (j - 3 * CM), (double)s);
        }
    }
}

__global__ void seg_loss_finalize_kernel(const double* __restrict__ stats, float* __restrict__ out3, int N, int C, long long HW, float smooth) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double* tail = stats + (long long)N * C * 3;
    out3[0] = (float)(tail[0] / ((double)N * (double)HW));
    double total = 0.0;
    for (int k = 0; k < C; ++k) {
        double acc = 0.0;
        for (int n = 0; n < N; ++n) {
            const double* s = stats + ((long long)n * C + k) * 3;
            // reference arithmetic is fp32 per image: 1 - (2*spt + smooth) / (sp2 + st + smooth)
            float num = (float)s[0] * 2.f + smooth;
            float den = (float)(s[1] + s[2]) + smooth;
            acc += (double)(1.f - num / den);
        }
        total += acc / N;
    }
    out3[1] = (float)(total / C);
    const float eps = 1.1920928955078125e-07f;
    out3[2] = ((float)tail[1] + eps) * (100.0f / ((float)tail[2] + eps));
}

template <int CM>
__global__ void __launch_bounds__(256) seg_loss_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ label,
                                                           const double* __restrict__ stats, const float* __restrict__ g_ce,
                                                           const float* __restrict__ g_dice, float* __restrict__ dlogits, long long HW,
                                                           int C, int N, int ignore, float smooth) {
    const long long n = blockIdx.y;
    const float* lg = logits + n * C * HW;
    float* dl = dlogits + n * C * HW;
    const int64_t* lb = label + n * HW;
    __shared__ float sA[CM], sB[CM];
    if (threadIdx.x < C) {
        const double* s = stats + (n * C + threadIdx.x) * 3;
        sA[threadIdx.x] = (float)s[0] * 2.f + smooth;
        sB[threadIdx.x] = (float)(s[1] + s[2]) + smooth;
    }
    __syncthreads();
    const float gce = (g_ce ? *g_ce : 0.f) / ((float)N * (float)HW);
    const float gd = (g_dice ? *g_dice : 0.f) / ((float)C * (float)N);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
        float p[CM];
        float m = -INFINITY;
#pragma unroll
        for (int k = 0; k < CM; ++k)
            if (k < C) { p[k] = lg[k * HW + i]; m = fmaxf(m, p[k]); }
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < CM; ++k)
            if (k < C) { p[k] = expf(p[k] - m); sum += p[k]; }
        const float inv = 1.f / sum;
        const long long y = lb[i];
        const bool valid = y != ignore;
        const int yc = (int)(y < 0 ? 0 : (y > C - 1 ? C - 1 : y));
        // dDice/dp_k = -gd * ( 2 t_k m / B_k - A_k * 2 p_k / B_k^2 )
        float gp[CM];
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < CM; ++k) {
            if (k < C) {
                p[k] *= inv;
                float t = (k == yc) ? 1.f : 0.f;
                float tm = (valid && k == yc) ? 1.f : 0.f;
                float B = sB[k];
                gp[k] = -gd * (2.f * tm / B - sA[k] * 2.f * p[k] / (B * B));
                (void)t;
                dot = fmaf(gp[k], p[k], dot);
            }
        }
#pragma unroll
        for (int k = 0; k < CM; ++k) {
            if (k < C) {
                float d = p[k] * (gp[k] - dot);
                if (valid) d += gce * (p[k] - (k == yc ? 1.f : 0.f));
                dl[k * HW + i] = d;
            }
        }
    }
}

// ---------------------------------------------------------------- inference post-processing
__global__ void slide_accum_kernel(const float* __restrict__ crop, float* __restrict__ preds, float* __restrict__ count, int C, int H, int W,
                                   int hc, int wc, int y1, int x1, long long total) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    int x = (int)(i % wc), y = (int)((i / wc) % hc);
    long long n = i / ((long long)wc * hc);
    long long o = ((long long)(y1 + y)) * W + (x1 + x);
    for (int k = 0; k < C; ++k) preds[(n * C + k) * (long long)H * W + o] += crop[(n * C + k) * (long long)hc * wc + (long long)y * wc + x];
    count[n * (long long)H * W + o] += 1.f;
}

__global__ void argmax_kernel(const float* __restrict__ preds, const float* __restrict__ count, int64_t* __restrict__ pred, int C, long long HW,
                              long long total) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    long long n = i / HW, hw = i - n * HW;
    // softmax is monotone: argmax of softmax(preds / count) == argmax of preds / count (first max wins)
    float inv = count ? 1.f / count[i] : 1.f;
    float best = -INFINITY;
    int arg = 0;
    for (int k = 0; k < C; ++k) {
        float v = preds[(n * C + k) * HW + hw] * inv;
        if (v > best) { best = v; arg = k; }
    }
    pred[i] = arg;
}

// ---------------------------------------------------------------- confusion matrix / area histograms
// ONE shared-memory counter update per pixel: the (C+1) x (C+1) matrix M[y'][p'] with y' / p' = the class, or C when the label / the
// prediction is outside [0, C) (ignored pixels are skipped).  Everything metrics.py:75-87 asks for is a marginal of M:
//   cm[y][p] = M[y][p] (y, p < C) | area_label[c] = sum_p' M[c][p'] | area_pred[c] = sum_y' M[y'][c] | area_intersect[c] = M[c][c].
// Counters are PRIVATE PER WARP and the lanes of a warp that hit the same bin are combined first (__match_any_sync: one
// leader adds the population count), so the C = 3 case - at most 16 live bins - never serialises 32 lanes on one address.
// Loads: a warp iteration covers 256 consecutive pixels; a lane reads 4 pixel pairs as 16-byte (two int64 predictions) and 2-byte
// (two uint8 labels) vectors, consecutive lanes consecutive vectors.  Algorithmic traffic 9 B/pixel (int64 pred + uint8 label).
constexpr int kHistWarps = 8;
constexpr int kHistPrivBins = 1536;     // (C+1)^2 up to here: one matrix per warp (48 KB); above: one per CTA; above kHistBlockBins: global
constexpr int kHistBlockBins = 16384;

__device__ __forceinline__ void hist_add(unsigned int* h, int bin, int lane) {
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    if (bin >= 0 && lane == __ffs(peers) - 1) atomicAdd(h + bin, (unsigned)__popc(peers));
}

// mode 0: per-warp matrices in shared memory; 1: one matrix per CTA; 2: cm by global atomics, only the 3*C area counters in shared memory
template <bool VEC, bool L8>
__global__ void __launch_bounds__(kHistWarps * 32) confusion_hist_kernel(const int64_t* __restrict__ pred, const void* __restrict__ label, long long n,
                                                                        int C, int ignore, int mode, unsigned long long* __restrict__ cm,
                                                                        unsigned long long* __restrict__ areas) {
    extern __shared__ unsigned int s_hist[];
    const int B = C + 1, BB = B * B;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int copies = mode == 0 ? kHistWarps : 1;
    const int words = mode == 2 ? 3 * C : copies * BB;
    for (int i = threadIdx.x; i < words; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    unsigned int* mine = s_hist + (mode == 0 ? warp * BB : 0);
    const uint8_t* l8 = reinterpret_cast<const uint8_t*>(label);
    const int64_t* l64 = reinterpret_cast<const int64_t*>(label);
    auto count = [&](long long y, long long p) {     // executed by all 32 lanes (bin = -1: nothing to count)
        const bool live = y != (long long)ignore && y != -(1LL << 62);
        const int yy = (y >= 0 && y < C) ? (int)y : C, pp = (p >= 0 && p < C) ? (int)p : C;
        if (mode != 2) {
            hist_add(mine, live ? yy * B + pp : -1, lane);
        } else if (live) {
            if (yy < C) atomicAdd(s_hist + 2 * C + yy, 1u);
            if (pp < C) atomicAdd(s_hist + C + pp, 1u);
            if (pp < C && pp == yy) atomicAdd(s_hist + pp, 1u);
            if (yy < C && pp < C && cm) atomicAdd(cm + (size_t)yy * C + pp, 1ull);
        }
    };
    const long long kSkip = -(1LL << 62);             // marks a lane position past the end
    const long long warps_total = (long long)gridDim.x * kHistWarps, wid = (long long)blockIdx.x * kHistWarps + warp;
    for (long long base = wid * 256; base < n; base += warps_total * 256) {     // trip count is uniform across the warp
        if (VEC && base + 256 <= n) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long i = base + j * 64 + 2 * lane;
                const longlong2 pv = *reinterpret_cast<const longlong2*>(pred + i);
                long long y0, y1;
                if (L8) { const uchar2 lv = *reinterpret_cast<const uchar2*>(l8 + i); y0 = lv.x; y1 = lv.y; }
                else { const longlong2 lv = *reinterpret_cast<const longlong2*>(l64 + i); y0 = lv.x; y1 = lv.y; }
                count(y0, pv.x);
                count(y1, pv.y);
            }
        } else {
#pragma unroll 1
            for (int j = 0; j < 8; ++j) {
                const long long i = base + j * 32 + lane;
                const bool in = i < n;
                const long long y = in ? (L8 ? (long long)l8[i] : l64[i]) : kSkip;
                const long long p = in ? pred[i] : 0;
                count(y, p);
            }
        }
    }
    __syncthreads();
    if (mode == 2) {
        if (areas)
            for (int c = threadIdx.x; c < C; c += blockDim.x) {
                const unsigned long long I = s_hist[c], Pp = s_hist[C + c], Ll = s_hist[2 * C + c];
                if (I) atomicAdd(areas + c, I);
                if (Pp + Ll - I) atomicAdd(areas + C + c, Pp + Ll - I);
                if (Pp) atomicAdd(areas + 2 * C + c, Pp);
                if (Ll) atomicAdd(areas + 3 * C + c, Ll);
            }
        return;
    }
    // fold the per-warp copies into copy 0, then flush the marginals (global layout of areas: int64[4][C] = intersect, union, pred, label)
    for (int i = threadIdx.x; i < BB; i += blockDim.x) {
        unsigned int t = 0;
        for (int w = 0; w < copies; ++w) t += s_hist[w * BB + i];
        s_hist[i] = t;
    }
    __syncthreads();
    if (cm)
        for (int i = threadIdx.x; i < C * C; i += blockDim.x) {
            const unsigned int v = s_hist[(i / C) * B + (i % C)];
            if (v) atomicAdd(cm + i, (unsigned long long)v);
        }
    if (areas)
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            unsigned long long Ll = 0, Pp = 0;
            for (int k = 0; k < B; ++k) { Ll += s_hist[c * B + k]; Pp += s_hist[k * B + c]; }
            const unsigned long long I = s_hist[c * B + c];
            if (I) atomicAdd(areas + c, I);
            if (Pp + Ll - I) atomicAdd(areas + C + c, Pp + Ll - I);
            if (Pp) atomicAdd(areas + 2 * C + c, Pp);
            if (Ll) atomicAdd(areas + 3 * C + c, Ll);
        }
}

}  // namespace stc

using namespace stc;

extern "C" int stc_cls_fwd(const void* x, const float* W, const float* b, const float* mask, float* logits, int N, long long HW,
                           int Cin, int Ccls, int dtype, void* stream) {
    STC_REQUIRE(Cin % 8 == 0 && Ccls >= 1 && Ccls <= kMaxCls, "cls_fwd: Cin=%d (mult of 8) Ccls=%d (<=%d)", Cin, Ccls, kMaxCls);
    long long P = (long long)N * HW;
    size_t smem = sizeof(float) * ((size_t)Ccls * Cin + Ccls);
    STC_DISPATCH_DTYPE(dtype, (cls_fwd_kernel<T><<<ceil_div(P, 256), 256, smem, (cudaStream_t)stream>>>((const T*)x, W, b, mask, logits, HW, Cin, Ccls, P)));
    return check_launch("cls_fwd");
}

extern "C" long long stc_cls_bwd_ws_bytes(int, long long, int, int) { return 0; }

extern "C" int stc_cls_bwd(const float* dlogits, const void* x, const float* W, const float* mask, void* dx, float* dW, float* db, int N,
                           long long HW, int Cin, int Ccls, void* ws, long long ws_bytes, int dtype, void* stream) {
    (void)ws; (void)ws_bytes;
    STC_REQUIRE(vec_ok(Cin) && Ccls >= 1 && Ccls <= kMaxCls, "cls_bwd: Cin=%d must be 8*2^k, Ccls=%d (<=%d)", Cin, Ccls, kMaxCls);
    cudaStream_t st = (cudaStream_t)stream;
    long long P = (long long)N * HW;
    if (dx) {
        size_t smem = sizeof(float) * (size_t)Ccls * Cin;
        const int rstep = 256 / (Cin / 8);
        long long gx = (HW + (long long)rstep * 8 - 1) / ((long long)rstep * 8);          // ~8 pixels per thread
        const long long cap = max(1LL, (long long)num_sms() * 16 / max(N, 1));
        dim3 grid((unsigned)max(1LL, min(gx, cap)), (unsigned)N);
        STC_DISPATCH_DTYPE(dtype, (cls_bwd_dx_kernel<T><<<grid, 256, smem, st>>>(dlogits, W, mask, (T*)dx, HW, Cin, Ccls)));
    }
    if (dW) {
        // one full wave: as many blocks as the chip holds at once (register / shared-memory limited), split over the N images
        int occ = 0;
        if (dtype == STC_BF16) {
            STC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cls_bwd_dw_kernel<bf16>, 256, 0));
        } else {
            STC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cls_bwd_dw_kernel<float>, 256, 0));
        }
        const long long want = (HW + (256 / (Cin / 8)) * 8 - 1) / ((256 / (Cin / 8)) * 8);
        long long gxl = (long long)occ * num_sms() / max(N, 1);
        if (gxl > want) gxl = want;
        dim3 grid((unsigned)(gxl < 1 ? 1 : gxl), (unsigned)N);
        for (int k0 = 0; k0 < Ccls; k0 += 4)
            STC_DISPATCH_DTYPE(dtype, (cls_bwd_dw_kernel<T><<<grid, 256, 0, st>>>(dlogits, (const T*)x, mask, dW, db, HW, Cin, Ccls, k0)));
    }
    return check_launch("cls_bwd");
}

extern "C" long long stc_seg_loss_stats_len(int N, int C) { return (long long)N * C * 3 + 4; }

extern "C" int stc_seg_loss_fwd(const float* logits, const int64_t* label, double* stats, float* out3, int N, long long HW, int C,
                                int ignore_index, float smooth, void* stream) {
    STC_REQUIRE(C >= 1 && C <= kMaxCls && N >= 1, "seg_loss_fwd: C=%d must be in [1,%d]", C, kMaxCls);
    cudaStream_t st = (cudaStream_t)stream;
    STC_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * stc_seg_loss_stats_len(N, C), st));
    dim3 grid((unsigned)max(1LL, min((long long)ceil_div(HW, 256 * 4), (long long)num_sms() * 4 / N + 1)), N);
    if (C <= 4) seg_loss_fwd_kernel<4><<<grid, 256, 0, st>>>(logits, label, stats, HW, C, N, ignore_index);
    else if (C <= 8) seg_loss_fwd_kernel<8><<<grid, 256, 0, st>>>(logits, label, stats, HW, C, N, ignore_index);
    else seg_loss_fwd_kernel<kMaxCls><<<grid, 256, 0, st>>>(logits, label, stats, HW, C, N, ignore_index);
    seg_loss_finalize_kernel<<<1, 32, 0, st>>>(stats, out3, N, C, HW, smooth);
    return check_launch("seg_loss_fwd");
}

extern "C" int stc_seg_loss_bwd(const float* logits, const int64_t* label, const double* stats, const float* g_ce, const float* g_dice,
                                float* dlogits, int N, long long HW, int C, int ignore_index, float smooth, void* stream) {
    STC_REQUIRE(C >= 1 && C <= kMaxCls && N >= 1, "seg_loss_bwd: C=%d must be in [1,%d]", C, kMaxCls);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)max(1LL, min((long long)ceil_div(HW, 256 * 2), (long long)num_sms() * 8 / N + 1)), N);
    if (C <= 4) seg_loss_bwd_kernel<4><<<grid, 256, 0, st>>>(logits, label, stats, g_ce, g_dice, dlogits, HW, C, N, ignore_index, smooth);
    else if (C <= 8) seg_loss_bwd_kernel<8><<<grid, 256, 0, st>>>(logits, label, stats, g_ce, g_dice, dlogits, HW, C, N, ignore_index, smooth);
    else seg_loss_bwd_kernel<kMaxCls><<<grid, 256, 0, st>>>(logits, label, stats, g_ce, g_dice, dlogits, HW, C, N, ignore_index, smooth);
    return check_launch("seg_loss_bwd");
}

extern "C" int stc_slide_accum(const float* crop, float* preds, float* count, int N, int C, int H, int W, int hc, int wc, int y1, int x1,
                               void* stream) {
    STC_REQUIRE(y1 >= 0 && x1 >= 0 && y1 + hc <= H && x1 + wc <= W, "slide_accum: window out of range");
    long long total = (long long)N * hc * wc;
    slide_accum_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(crop, preds, count, C, H, W, hc, wc, y1, x1, total);
    return check_launch("slide_accum");
}

extern "C" int stc_argmax(const float* preds, const float* count, int64_t* pred, int N, int C, long long HW, void* stream) {
    long long total = (long long)N * HW;
    argmax_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(preds, count, pred, C, HW, total);
    return check_launch("argmax");
}

extern "C" int stc_confusion_hist(const int64_t* pred, const void* label, int label_is_u8, long long n, int C, int ignore_index,
                                  int64_t* cm, int64_t* areas, void* stream) {
    STC_REQUIRE(C >= 1 && C <= 4096, "confusion_hist: C=%d must be in [1,4096]", C);
    if (n <= 0) return STC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int BB = (C + 1) * (C + 1);
    const int mode = BB <= kHistPrivBins ? 0 : (BB <= kHistBlockBins ? 1 : 2);
    const size_t smem = sizeof(unsigned int) * (mode == 0 ? (size_t)kHistWarps * BB : (mode == 1 ? (size_t)BB : (size_t)3 * C));
    // vector path: 16-byte prediction pairs and 2-byte (or 16-byte) label pairs must be aligned; a warp iteration is 256 pixels
    const bool vec = ((uintptr_t)pred & 15) == 0 && ((uintptr_t)label & (label_is_u8 ? 1 : 15)) == 0;
    int blocks = (int)max(1LL, min((long long)num_sms() * 4, (long long)ceil_div(n, 256 * kHistWarps * 2)));
    auto go = [&](auto kern) -> int {
        if (smem > 48 * 1024) STC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, kHistWarps * 32, smem, st>>>(pred, label, n, C, ignore_index, mode, (unsigned long long*)cm, (unsigned long long*)areas);
        return STC_OK;
    };
    int rc;
    if (vec) rc = label_is_u8 ? go(confusion_hist_kernel<true, true>) : go(confusion_hist_kernel<true, false>);
    else rc = label_is_u8 ? go(confusion_hist_kernel<false, true>) : go(confusion_hist_kernel<false, false>);
    if (rc) return rc;
    return check_launch("confusion_hist");
}
