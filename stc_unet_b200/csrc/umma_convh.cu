// tcgen05 implicit-GEMM convolution with shared-memory HALO reuse (fprop, and dgrad with the flipped weight pack).
//
// The per-tap kernel in umma.cu re-fetches a 128-pixel x 64-channel A tile from L2 for every filter tap; on B200
// that fetch stream (~5.5 TB/s of distinct lines chip-wide) is what bounds it: every k-iteration costs the same
// ~0.4 us whatever the N tile is (profiles/r1_conv_sweep_v1.txt: 375 / 745 / 1300 TFLOP/s at Cout = 64 / 128 / 256).
// Here an input ROW SEGMENT of (128 + S - 1) pixels x 64 channels is loaded ONCE into a 128B-swizzled smem slot and
// every tap (r, s) that needs it reads it in place: the UMMA descriptor's start address is simply advanced by
// s * 128 bytes (one pixel row of the K-major tile).  Measured on B200: the 128B swizzle XOR is taken from the
// absolute shared-memory address bits [7,10), exactly as TMA wrote it, so the shifted descriptor needs NO base-offset
// correction (base_offset = (start >> 7) & 7 gives garbage; base_offset = 0 is bit-exact; tests/test_umma_gpu.py).  A work item is a strip of T vertically adjacent output rows x 128 pixels x one N tile with T accumulators
// in TMEM, so a segment is also shared by the R output rows that touch it:
//     A traffic per output tile:  R*S*16 KB  ->  (T+R-1)/T * 17 KB   (3x3, T=4: 144 KB -> 26 KB)
// Loop order inside a strip: for cin-chunk { for r { for s { B(r,s,chunk) once; for t: acc[t] += seg[t+r](+s) * B } } }.
// Warps: 0 = A (segment) producer, 1 = MMA issuer + TMEM owner, 2..5 = epilogue, 6 = B (weight) producer.
//
// CTA2 variant (cta_group::2): two CTAs of a cluster (one TPC) work on two strips of the SAME N tile with ONE tcgen05.mma of M = 256 per
// (segment, tap, k step): each CTA stages its own input segments and only HALF of the weight tile's output channels; the leader's MMA reads A
// from both shared memories and the two B halves, accumulator rows 0-127 in the leader's TMEM (its strip), 128-255 in the peer's.  The SS-mode
// MMA is paced by shared-memory operand reads (~90 B/clk/SM measured: 128x64 47 %, 128x128 68 % of the tensor peak, profiles/r2_convh_roles.txt);
// the pair form cuts them from A + B to A + B/2 per CTA.  Protocol as in umma.cu: `full` barriers in the leader receive the TMA bytes of both
// CTAs, `empty` / `tmem_full` are signalled in both CTAs by multicast commits, both epilogues release the leader's `tmem_empty`.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace stc {

struct alignas(64) ConvHParams {
    CUtensorMap tmA;  // NHWC activations {C, W, H, N}, box {64, 128+S-1, 1, 1}
    CUtensorMap tmB;  // packed weights {Cin, Cout, taps}, box {64, BN, 1}
    CUtensorMap tmC;  // output {Cout, W, H, N}, box {64, 32, 1, 1}: staged epilogue (see umma.cu), used when `staged`
    // virtual channel concat (common.cuh ChanCat): extra input sources / extra outputs (dgrad of a concat conv)
    CUtensorMap tmA2[kMaxCat - 1];
    CUtensorMap tmC2[kMaxCat - 1];
    int n_src, src_chunk_end[kMaxCat];
    int n_out, out_ch_end[kMaxCat];
    void* out2[kMaxCat - 1];
    int staged;
    int H, W, R, S, cin_chunks, T, BN, num_n_tiles;
    int strips_h, strips_w, num_strips;
    int a_slots, b_stages;
    uint32_t a_slot_bytes, a_box_bytes, b_stage_bytes;
    uint32_t idesc;
    int bo_mode;  // debugging only (STC_CONVH_BO=0 sets base_offset = (start >> 7) & 7, which is WRONG on B200)
    void* out;
    const float* bias;
    const void* residual;
    int act, Cout;
    float* stats;  // optional [gridDim.x * 4][2][Cout] fp32: per epilogue-warp column sums / sums of squares of the STORED (bf16) outputs
    long long* prof;  // diagnostics (stc_debug_profile): 16 clock counters per CTA, see tools/convh_prof.py; nullptr = off
};

constexpr int kConvHThreads = 224;

// Column sums of a 32-row x 32-column register tile (one row per lane, v[j] = column j) by recursive halving: after the five
// exchange steps lane l holds the total of column l.  31 shuffles instead of 32 x 5.
__device__ __forceinline__ float convh_colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = up ? v[i] : v[i + off];
            const float keep = up ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

// BatchNorm statistics fused into the conv epilogue ("BN stats = 0 extra if fused"): f[] are this lane's 32 output
// values of one pixel; they are rounded to bf16 exactly as stored, masked if the pixel is outside the image, and their per-column
// sum / sum of squares over the warp's 32 pixels is added to the lane's running totals for column chunk `ci` (static indexing).
__device__ __forceinline__ void convh_stats_add(const float (&f)[32], bool valid, int lane, int ci, float (&sacc)[8], float (&qacc)[8]) {
    float a[32], b[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const float r = valid ? __bfloat162float(__float2bfloat16_rn(f[j])) : 0.f;
        a[j] = r;
        b[j] = r * r;
    }
    const float s1 = convh_colsum32(a, lane), s2 = convh_colsum32(b, lane);
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (k == ci) { sacc[k] += s1; qacc[k] += s2; }
}
constexpr uint32_t kConvHStageBuf = 32 * 128;                   // 32 pixels x 64 bf16 channels, 128B-swizzled
constexpr uint32_t kConvHStagingBytes = 4 * 2 * kConvHStageBuf;  // two buffers per epilogue warp

__device__ __forceinline__ float convh_act(float v, int act) {
    if (act == STC_ACT_RELU) return fmaxf(v, 0.f);
    if (act == STC_ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
    if (act == STC_ACT_HSWISH) return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);
    return v;
}

struct Strip {
    int nt, n_img, h0, w0;
};
__device__ __forceinline__ const CUtensorMap* convh_map_a(const ConvHParams& p, int cc, int& local) {
    int j = 0;
    while (j + 1 < p.n_src && cc >= p.src_chunk_end[j]) ++j;
    local = cc - (j ? p.src_chunk_end[j - 1] : 0);
    return j ? &p.tmA2[j - 1] : &p.tmA;
}
__device__ __forceinline__ int convh_output(const ConvHParams& p, int gch, int& local, int& width) {
    int j = 0;
    while (j + 1 < p.n_out && gch >= p.out_ch_end[j]) ++j;
    const int start = j ? p.out_ch_end[j - 1] : 0;
    local = gch - start;
    width = p.out_ch_end[j] - start;
    return j;
}
// CTA2: idx counts strip PAIRS; the pair's two CTAs take spatially consecutive strips of the same N tile
template <bool CTA2>
__device__ __forceinline__ Strip decode_strip(const ConvHParams& p, int idx, int rank) {
    Strip s;
    s.nt = idx % p.num_n_tiles;
    int r = idx / p.num_n_tiles;
    if (CTA2) r = 2 * r + rank;
    int sw = r % p.strips_w;
    r /= p.strips_w;
    int sh = r % p.strips_h;
    s.n_img = r / p.strips_h;
    s.h0 = sh * p.T;
    s.w0 = sw * 128;
    return s;
}

template <int T, bool CTA2>
__global__ void __launch_bounds__(kConvHThreads, 1) umma_convh_kernel(const __grid_constant__ ConvHParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t a_bytes = (uint32_t)p.a_slots * p.a_slot_bytes;
    const uint32_t ring_bytes = (uint32_t)(a_bytes + (size_t)p.b_stages * p.b_stage_bytes);   // a multiple of 1024
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ring_bytes + (p.staged ? kConvHStagingBytes : 0u));
    // bars: a_full[a_slots], a_empty[a_slots], b_full[b_stages], b_empty[b_stages], tmem_full[2], tmem_empty[2]
    const int nb = 2 * p.a_slots + 2 * p.b_stages + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + nb);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = CTA2 ? (int)ptx::cluster_ctarank() : 0;              // 0 = leader of the pair
    const int item0 = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;     // persistent loop over strips (CTA2: strip pairs)
    const int istep = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const uint32_t smem_base = ptx::smem_u32(smem);
    const uint32_t b_base = smem_base + a_bytes;
    const uint32_t bar_base = ptx::smem_u32(bars);
    auto a_full = [&](int s) { return bar_base + 8u * s; };
    auto a_empty = [&](int s) { return bar_base + 8u * (p.a_slots + s); };
    auto b_full = [&](int s) { return bar_base + 8u * (2 * p.a_slots + s); };
    auto b_empty = [&](int s) { return bar_base + 8u * (2 * p.a_slots + p.b_stages + s); };
    auto tfull = [&](int s) { return bar_base + 8u * (2 * p.a_slots + 2 * p.b_stages + s); };
    auto tempty = [&](int s) { return bar_base + 8u * (2 * p.a_slots + 2 * p.b_stages + 2 + s); };

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&p.tmA);
        ptx::prefetch_tensormap(&p.tmB);
        for (int j = 1; j < p.n_src; ++j) ptx::prefetch_tensormap(&p.tmA2[j - 1]);
        for (int s = 0; s < p.a_slots; ++s) { ptx::mbar_init(a_full(s), 1); ptx::mbar_init(a_empty(s), 1); }
        for (int s = 0; s < p.b_stages; ++s) { ptx::mbar_init(b_full(s), 1); ptx::mbar_init(b_empty(s), 1); }
        for (int s = 0; s < 2; ++s) { ptx::mbar_init(tfull(s), 1); ptx::mbar_init(tempty(s), CTA2 ? 8 : 4); }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        if (CTA2) { ptx::tmem_alloc2(ptx::smem_u32(tmem_slot), 512); ptx::tmem_relinquish2(); }
        else { ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512); ptx::tmem_relinquish(); }
    }
    ptx::tc_fence_before();
    if (CTA2) ptx::cluster_sync_all(); else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tmem_base != 0) {
        if (threadIdx.x == 0) printf("stc_b200: unexpected TMEM base 0x%x\n", tmem_base);
        __trap();
    }
    const int segs_per_chunk = T + p.R - 1;
    const int pr = p.R / 2, ps = p.S / 2;

    if (warp == 0) {
        // ===================== A producer: input row segments (warp-uniform; elected lane issues) =====================
        {
            int slot = 0;          // ring position of the next segment
            uint32_t phase = 0;    // flips every time the ring wraps
            long long w_ae = 0;
            for (int st = item0; st < p.num_strips; st += istep) {
                Strip s = decode_strip<CTA2>(p, st, rank);
                for (int cc = 0; cc < p.cin_chunks; ++cc) {
                    int lc = cc;
                    const CUtensorMap* ma = p.n_src > 1 ? convh_map_a(p, cc, lc) : &p.tmA;
                    for (int i = 0; i < segs_per_chunk; ++i) {
                        const long long t0 = p.prof ? clock64() : 0;
                        ptx::mbar_wait(a_empty(slot), phase ^ 1);
                        if (p.prof) w_ae += clock64() - t0;
                        if (ptx::elect_one_sync()) {
                            if (CTA2) {   // both CTAs' bytes complete on the leader's barrier, which the leader arms for 2 boxes
                                if (rank == 0) ptx::mbar_arrive_expect_tx(a_full(slot), 2 * p.a_box_bytes);
                                ptx::tma_load_4d_2sm(smem_base + slot * p.a_slot_bytes, ma, ptx::mapa_shared(a_full(slot), 0), lc * 64, s.w0 - ps, s.h0 + i - pr, s.n_img);
                            } else {
                                ptx::mbar_arrive_expect_tx(a_full(slot), p.a_box_bytes);
                                ptx::tma_load_4d(smem_base + slot * p.a_slot_bytes, ma, a_full(slot), lc * 64, s.w0 - ps, s.h0 + i - pr, s.n_img);
                            }
                        }
                        __syncwarp();
                        if (++slot == p.a_slots) { slot = 0; phase ^= 1; }
                    }
                }
            }
            if (p.prof && lane == 0) p.prof[blockIdx.x * 16 + 8] = w_ae;
        }
    } else if (warp == 6) {
        // ===================== B producer: one weight tile per (chunk, tap) =====================
        {
            int stage = 0;
            uint32_t phase = 0;
            long long w_be = 0;
            for (int st = item0; st < p.num_strips; st += istep) {
                Strip s = decode_strip<CTA2>(p, st, rank);
                for (int cc = 0; cc < p.cin_chunks; ++cc) {
                    for (int tap = 0; tap < p.R * p.S; ++tap) {
                        const long long t0 = p.prof ? clock64() : 0;
                        ptx::mbar_wait(b_empty(stage), phase ^ 1);
                        if (p.prof) w_be += clock64() - t0;
                        if (ptx::elect_one_sync()) {
                            if (CTA2) {   // this CTA stages its half of the tile's output channels (b_stage_bytes = the half)
                                if (rank == 0) ptx::mbar_arrive_expect_tx(b_full(stage), 2 * p.b_stage_bytes);
                                ptx::tma_load_3d_2sm(b_base + stage * p.b_stage_bytes, &p.tmB, ptx::mapa_shared(b_full(stage), 0), cc * 64, s.nt * p.BN + rank * (p.BN / 2), tap);
                            } else {
                                ptx::mbar_arrive_expect_tx(b_full(stage), p.b_stage_bytes);
                                ptx::tma_load_3d(b_base + stage * p.b_stage_bytes, &p.tmB, b_full(stage), cc * 64, s.nt * p.BN, tap);
                            }
                        }
                        __syncwarp();
                        if (++stage == p.b_stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
            if (p.prof && lane == 0) p.prof[blockIdx.x * 16 + 9] = w_be;
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform; elected lane issues; CTA2: the leader only) =====================
        if (!CTA2 || rank == 0) {
            int a_head = 0;            // ring slot of segment 0 of the current chunk
            uint32_t a_phase = 0;      // parity of the ring pass that a_head belongs to
            int bstage = 0;
            uint32_t bphase = 0;
            int acc = 0;
            uint32_t acc_phase0 = 0, acc_phase1 = 0;
            // descriptor constants (16-byte units): SBO = 1024 B, version 1, SWIZZLE_128B; base_offset stays 0 (see header)
            const uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
            const uint32_t a_slot16 = p.a_slot_bytes >> 4, b_stage16 = p.b_stage_bytes >> 4;
            const uint32_t a_base16 = (smem_base >> 4) & 0x3FFF, b_base16 = (b_base >> 4) & 0x3FFF;
            // arrives once the MMAs issued so far have retired (CTA2: on the barrier at this offset in BOTH CTAs)
            auto commit = [&](uint32_t bar) { if (CTA2) ptx::tc_commit2(bar, 3); else ptx::tc_commit(bar); };
            long long w_te = 0, w_af = 0, w_bf = 0, n_st = 0;
            const long long t_begin = p.prof ? clock64() : 0;
            for (int st = item0; st < p.num_strips; st += istep) {
                long long t0 = p.prof ? clock64() : 0;
                ptx::mbar_wait(tempty(acc), (acc ? acc_phase1 : acc_phase0) ^ 1);
                if (p.prof) { w_te += clock64() - t0; ++n_st; }
                ptx::tc_fence_after();
                const uint32_t d_base = acc * 256;  // TMEM base is 0 (all 512 columns are ours; checked after the allocation)
                for (int cc = 0; cc < p.cin_chunks; ++cc) {
                    int ready = 0;  // segments of this chunk whose full barrier has been observed
                    for (int r = 0; r < p.R; ++r) {
                        // filter row r reads segments r .. r+T-1: at most one new segment per row (all T at r == 0)
                        while (ready < r + T) {
                            int sl = a_head + ready;
                            uint32_t ph = a_phase;
                            if (sl >= p.a_slots) { sl -= p.a_slots; ph ^= 1; }
                            t0 = p.prof ? clock64() : 0;
                            ptx::mbar_wait(a_full(sl), ph);
                            if (p.prof) w_af += clock64() - t0;
                            ++ready;
                        }
                        uint32_t a16[T];  // start-address fields (16 B units) of the T segments this row reads
#pragma unroll
                        for (int t = 0; t < T; ++t) {
                            int sl = a_head + r + t;
                            if (sl >= p.a_slots) sl -= p.a_slots;
                            a16[t] = a_base16 + sl * a_slot16;
                        }
                        for (int s = 0; s < p.S; ++s) {
                            t0 = p.prof ? clock64() : 0;
                            ptx::mbar_wait(b_full(bstage), bphase);
                            if (p.prof) w_bf += clock64() - t0;
                            ptx::tc_fence_after();
                            const uint64_t b_desc0 = desc_hi | (uint64_t)(b_base16 + bstage * b_stage16);
                            const uint32_t acc_flag = (cc | r | s) ? 1u : 0u;
                            if (ptx::elect_one_sync()) {
#pragma unroll
                                for (int t = 0; t < T; ++t) {
                                    const uint64_t a_desc0 = desc_hi | (uint64_t)(a16[t] + s * 8);
                                    const uint32_t d_addr = d_base + t * p.BN;
                                    if (CTA2) {
                                        ptx::mma_bf16_ss2(d_addr, a_desc0, b_desc0, p.idesc, acc_flag);
                                        ptx::mma_bf16_ss2(d_addr, a_desc0 + 2, b_desc0 + 2, p.idesc, 1u);
                                        ptx::mma_bf16_ss2(d_addr, a_desc0 + 4, b_desc0 + 4, p.idesc, 1u);
                                        ptx::mma_bf16_ss2(d_addr, a_desc0 + 6, b_desc0 + 6, p.idesc, 1u);
                                    } else {
                                        ptx::mma_bf16_ss(d_addr, a_desc0, b_desc0, p.idesc, acc_flag);
                                        ptx::mma_bf16_ss(d_addr, a_desc0 + 2, b_desc0 + 2, p.idesc, 1u);
                                        ptx::mma_bf16_ss(d_addr, a_desc0 + 4, b_desc0 + 4, p.idesc, 1u);
                                        ptx::mma_bf16_ss(d_addr, a_desc0 + 6, b_desc0 + 6, p.idesc, 1u);
                                    }
                                }
                                commit(b_empty(bstage));
                            }
                            __syncwarp();
                            if (++bstage == p.b_stages) { bstage = 0; bphase ^= 1; }
                        }
                        // segments whose last reader was this filter row can be recycled
                        if (ptx::elect_one_sync()) {
                            if (r < p.R - 1) {
                                int sl = a_head + r;
                                if (sl >= p.a_slots) sl -= p.a_slots;
                                commit(a_empty(sl));
                            } else {
                                for (int i = p.R - 1; i < segs_per_chunk; ++i) {
                                    int sl = a_head + i;
                                    if (sl >= p.a_slots) sl -= p.a_slots;
                                    commit(a_empty(sl));
                                }
                            }
                        }
                        __syncwarp();
                    }
                    a_head += segs_per_chunk;
                    if (a_head >= p.a_slots) { a_head -= p.a_slots; a_phase ^= 1; }
                }
                if (ptx::elect_one_sync()) commit(tfull(acc));
                __syncwarp();
                if (acc) acc_phase1 ^= 1; else acc_phase0 ^= 1;
                acc ^= 1;
            }
            if (p.prof && lane == 0) {
                long long* o = p.prof + blockIdx.x * 16;
                o[0] = clock64() - t_begin; o[1] = w_te; o[2] = w_af; o[3] = w_bf; o[4] = n_st;
            }
        }
    } else if (warp >= 2 && warp <= 5) {
        // ===================== epilogue =====================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        int acc = 0, stage_buf = 0;
        uint32_t acc_phase[2] = {0, 0};
        const uint32_t stg0 = smem_base + ring_bytes + (uint32_t)q * 2u * kConvHStageBuf;
        float sacc[8] = {}, qacc[8] = {};   // p.stats: running column sums of this warp (lane l <-> channel chunk*32 + l); num_n_tiles == 1
        long long w_tf = 0, t_work = 0;
        for (int st = item0; st < p.num_strips; st += istep) {
            Strip s = decode_strip<CTA2>(p, st, rank);
            const int w = s.w0 + row;
            const long long t0 = p.prof ? clock64() : 0;
            ptx::mbar_wait(tfull(acc), acc_phase[acc]);
            const long long t1 = p.prof ? clock64() : 0;
            w_tf += t1 - t0;
            ptx::tc_fence_after();
            for (int t = 0; t < T; ++t) {
                const int h = s.h0 + t;
                if (h >= p.H) break;  // uniform across the CTA
                const bool valid = w < p.W;
                const long long off = (((long long)s.n_img * p.H + h) * p.W + w) * p.Cout + (long long)s.nt * p.BN;
                const uint32_t t_addr = tmem_base + acc * 256 + t * p.BN + ((uint32_t)(q * 32) << 16);
                if (p.staged) {
                    // TMEM -> registers -> swizzled smem box -> TMA store (full 128-byte lines; out-of-range pixels clipped by TMA)
                    for (int c = 0; c < p.BN; c += 64) {
                        const uint32_t stg = stg0 + (uint32_t)stage_buf * kConvHStageBuf;
                        if (lane == 0) ptx::bulk_wait_read<1>();
                        __syncwarp();
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            uint32_t v[32];
                            ptx::tmem_ld_32x32(t_addr + c + half * 32, v);
                            ptx::tmem_ld_wait();
                            float f[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                            if (p.bias) {
                                const float* b = p.bias + s.nt * p.BN + c + half * 32;
#pragma unroll
                                for (int j = 0; j < 32; ++j) f[j] += __ldg(b + j);
                            }
                            if (p.residual && valid) {
                                const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.residual) + off + c + half * 32);
#pragma unroll
                                for (int g = 0; g < 4; ++g) {
                                    uint4 rv = __ldg(rp + g);
                                    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        float2 x = __bfloat1622float2(h2[e]);
                                        f[g * 8 + 2 * e] += x.x;
                                        f[g * 8 + 2 * e + 1] += x.y;
                                    }
                                }
                            }
                            if (p.act == STC_ACT_RELU) {   // uniform branch per chunk: a per-element switch compiles to a jump table
#pragma unroll
                                for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
                            } else if (p.act != STC_ACT_NONE) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) f[j] = convh_act(f[j], p.act);
                            }
                            if (p.stats) convh_stats_add(f, valid, lane, (c >> 5) + half, sacc, qacc);
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                uint4 ov;
                                __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
                                for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(f[g * 8 + 2 * e], f[g * 8 + 2 * e + 1]);
                                ptx::st_shared_v4(stg + (uint32_t)lane * 128u + (uint32_t)(((half * 4 + g) ^ (lane & 7)) << 4), ov);
                            }
                        }
                        ptx::fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            int oc = s.nt * p.BN + c, ow;
                            const CUtensorMap* mc = &p.tmC;
                            if (p.n_out > 1) { const int j = convh_output(p, oc, oc, ow); if (j) mc = &p.tmC2[j - 1]; }
                            ptx::tma_store_4d(mc, stg, oc, s.w0 + q * 32, h, s.n_img);
                            ptx::bulk_commit();
                        }
                        stage_buf ^= 1;
                    }
                    continue;
                }
                for (int c = 0; c < p.BN; c += 32) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(t_addr + c, v);
                    ptx::tmem_ld_wait();
                    if (!valid && !p.stats) continue;   // with statistics every lane takes part in the warp-wide column sums
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                    if (p.bias) {
                        const float* b = p.bias + s.nt * p.BN + c;
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] += __ldg(b + j);
                    }
                    if (p.residual && valid) {
                        const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.residual) + off + c);
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            uint4 rv = __ldg(rp + g);
                            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float2 x = __bfloat1622float2(h2[e]);
                                f[g * 8 + 2 * e] += x.x;
                                f[g * 8 + 2 * e + 1] += x.y;
                            }
                        }
                    }
                    if (p.act == STC_ACT_RELU) {   // uniform branch per chunk: a per-element switch compiles to a jump table
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
                    } else if (p.act != STC_ACT_NONE) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = convh_act(f[j], p.act);
                    }
                    if (p.stats) {
                        convh_stats_add(f, valid, lane, c >> 5, sacc, qacc);
                        if (!valid) continue;
                    }
                    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + off + c);
                    if (p.n_out > 1) {   // dgrad of a virtual concat: this 32-channel chunk belongs to one of the outputs
                        int lc, ow;
                        const int j = convh_output(p, s.nt * p.BN + c, lc, ow);
                        const long long pix = ((long long)s.n_img * p.H + h) * p.W + w;
                        o = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(j ? p.out2[j - 1] : p.out) + pix * ow + lc);
                    }
                    uint4 ov[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&ov[g]);
#pragma unroll
                        for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(f[g * 8 + 2 * e], f[g * 8 + 2 * e + 1]);
                    }
                    if ((reinterpret_cast<uintptr_t>(o) & 31) == 0) {   // two full 32-byte sectors per lane
                        ptx::st_global_v8(o, ov[0], ov[1]);
                        ptx::st_global_v8(o + 2, ov[2], ov[3]);
                    } else {
#pragma unroll
                        for (int g = 0; g < 4; ++g) o[g] = ov[g];
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CTA2) ptx::mbar_arrive_cluster(ptx::mapa_shared(tempty(acc), 0));
                else ptx::mbar_arrive(tempty(acc));
            }
            acc_phase[acc] ^= 1;
            acc ^= 1;
            if (p.prof) t_work += clock64() - t1;
        }
        if (p.prof && q == 0 && lane == 0) { p.prof[blockIdx.x * 16 + 5] = w_tf; p.prof[blockIdx.x * 16 + 6] = t_work; }
        if (p.staged && lane == 0) ptx::bulk_wait<0>();
        if (p.stats) {   // one partial row per epilogue warp: [sum | sum of squares] x Cout
            float* o = p.stats + ((size_t)blockIdx.x * 4 + q) * 2 * p.Cout;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k * 32 < p.BN) { o[k * 32 + lane] = sacc[k]; o[p.Cout + k * 32 + lane] = qacc[k]; }
        }
    }

    ptx::tc_fence_before();
    if (CTA2) ptx::cluster_sync_all(); else __syncthreads();   // CTA2: the peer's shared memory / barriers stay alive until both are done
    if (warp == 1) {
        ptx::tc_fence_after();
        if (CTA2) ptx::tmem_dealloc2(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------
int encode_map_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box);
int umma_pick_bn(int n);

// The staged (TMA store) epilogue costs 32 KB of shared memory; it pays where the tile's K is short (3x3: the write-back is a large
// share of the tile time) and is skipped where it would shrink the halo ring of the long-K 5x5 / 7x7 layers.  STC_CONVH_STAGED=<max R>.
static int convh_staged_max_r() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("STC_CONVH_STAGED"); v = e ? atoi(e) : 3; }
    return v;
}

// STC_CONVH_CTA2=0 keeps every halo conv on single CTAs
static bool convh_cta2_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("STC_CONVH_CTA2"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}

static int convh_plan(int Cout, int R, int& BN, int& T, int& a_slots, int& b_stages, int& staged, bool cta2 = false) {
    BN = umma_pick_bn(Cout);
    if (BN == 0 || BN > 256) return 0;
    staged = (BN % 64 == 0 && BN <= 128 && R <= convh_staged_max_r()) ? 1 : 0;   // BN = 256 needs the smem for its weight ring
    const size_t staging = staged ? kConvHStagingBytes : 0;
    const int tmax = 256 / BN;  // two accumulator stages of 256 TMEM columns
    const int cands[] = {4, 2, 1};
    static int bs_max = -1, extra_max = -1;   // experiment knobs: STC_CONVH_BS (weight-ring depth), STC_CONVH_EXTRA (spare segment slots)
    if (bs_max < 0) { const char* e = getenv("STC_CONVH_BS"); bs_max = e ? atoi(e) : 4; }
    if (extra_max < 0) { const char* e = getenv("STC_CONVH_EXTRA"); extra_max = e ? atoi(e) : 3; }
    for (int t : cands) {
        if (t > tmax) continue;
        for (int extra = extra_max; extra >= 1; --extra) {
            for (int bs = bs_max; bs >= 2; --bs) {
                int slots = t + R - 1 + extra;
                size_t smem = (size_t)slots * 17408 + (size_t)bs * (cta2 ? BN / 2 : BN) * 128 + staging + (2 * slots + 2 * bs + 4) * 8 + 16 + 1024;
                if (smem <= 225 * 1024) {
                    T = t; a_slots = slots; b_stages = bs;
                    return 1;
                }
            }
        }
    }
    return 0;
}

bool conv_convh_eligible(int W, int Cin, int Cout, int R, int S, int dtype) {
    static int disabled = -1;
    if (disabled < 0) { const char* e = getenv("STC_CONVH"); disabled = (e && e[0] == '0') ? 1 : 0; }
    if (disabled) return false;
    int BN, T, a, b, sg;
    return dtype == STC_BF16 && W >= 128 && Cin % 64 == 0 && R == S && (R == 3 || R == 5 || R == 7) && convh_plan(Cout, R, BN, T, a, b, sg);
}

// stats != nullptr: also writes *stats_rows partial rows of [sum | sumsq] x Cout (fp32) of the stored outputs; needs BN == Cout
bool conv_convh_stats_ok(int Cout, int R) {
    int BN, T, a, b, sg;
    return convh_plan(Cout, R, BN, T, a, b, sg) && BN == Cout;
}

int check_cat(const ChanCat* c, int total, const char* what);
long long* debug_profile_buffer();

static int convh_encode_nhwc(CUtensorMap* m, const void* base, int N, int H, int W, int C, uint32_t bw) {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[4] = {2, (uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {64, bw, 1, 1};
    return encode_map_bf16(m, base, 4, dims, str, box);
}

int conv_fprop_convh(const void* x, const void* wp, const float* bias, const void* residual, void* y, int N, int H, int W, int Cin,
                     int Cout, int R, int S, int act, cudaStream_t st, float* stats, int* stats_rows, const ChanCat* src, const ChanCat* dst) {
    ConvHParams p;
    memset(&p, 0, sizeof(p));
    STC_REQUIRE(convh_plan(Cout, R, p.BN, p.T, p.a_slots, p.b_stages, p.staged), "conv_fprop_convh: no plan for Cout=%d R=%d", Cout, R);
    // CTA pairs need an even number of spatial strips (the two CTAs of a pair always work on two real strips of one N tile)
    bool cta2 = convh_cta2_enabled() && ((long long)N * ((H + p.T - 1) / p.T) * ((W + 127) / 128)) % 2 == 0 && p.BN % 32 == 0;
    if (cta2) {
        int bn2, t2;
        cta2 = convh_plan(Cout, R, bn2, t2, p.a_slots, p.b_stages, p.staged, true) && bn2 == p.BN && t2 == p.T;
        if (!cta2) convh_plan(Cout, R, p.BN, p.T, p.a_slots, p.b_stages, p.staged);
    }
    if (src && src->n == 1) { x = src->ptr[0]; src = nullptr; }
    if (dst && dst->n == 1) { y = const_cast<void*>(dst->ptr[0]); dst = nullptr; }
    if (int rc = check_cat(src, Cin, "conv_fprop_convh input")) return rc;
    if (int rc = check_cat(dst, Cout, "conv_fprop_convh output")) return rc;
    STC_REQUIRE(!dst || (!bias && !residual && act == STC_ACT_NONE && !stats), "conv_fprop_convh: a split output takes no bias / residual / activation / statistics");
    if (src) x = src->ptr[0];
    if (dst) y = const_cast<void*>(dst->ptr[0]);
    STC_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)wp & 15) == 0 && ((uintptr_t)y & 15) == 0, "conv_fprop_convh: unaligned pointer");
    const int bwh = 128 + S - 1;
    p.n_src = 1; p.n_out = 1;
    if (src) {
        p.n_src = src->n;
        int acc = 0;
        for (int j = 0; j < src->n; ++j) {
            int rc = convh_encode_nhwc(j ? &p.tmA2[j - 1] : &p.tmA, src->ptr[j], N, H, W, src->c[j], (uint32_t)bwh);
            if (rc) return rc;
            acc += src->c[j] / 64;
            p.src_chunk_end[j] = acc;
        }
    } else {
        int rc = convh_encode_nhwc(&p.tmA, x, N, H, W, Cin, (uint32_t)bwh);
        if (rc) return rc;
    }
    if (dst) {
        p.n_out = dst->n;
        int acc = 0;
        for (int j = 0; j < dst->n; ++j) {
            acc += dst->c[j];
            p.out_ch_end[j] = acc;
            if (j) p.out2[j - 1] = const_cast<void*>(dst->ptr[j]);
        }
    }
    {
        uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)(R * S)};
        uint64_t str[3] = {2, (uint64_t)Cin * 2, (uint64_t)Cout * Cin * 2};
        uint32_t box[3] = {64, (uint32_t)(cta2 ? p.BN / 2 : p.BN), 1};
        int rc = encode_map_bf16(&p.tmB, wp, 3, dims, str, box);
        if (rc) return rc;
    }
    if (p.staged) {
        for (int j = 0; j < p.n_out; ++j) {
            int rc = convh_encode_nhwc(j ? &p.tmC2[j - 1] : &p.tmC, dst ? dst->ptr[j] : y, N, H, W, dst ? dst->c[j] : Cout, 32);
            if (rc) return rc;
        }
    }
    p.H = H; p.W = W; p.R = R; p.S = S; p.cin_chunks = Cin / 64;
    p.num_n_tiles = Cout / p.BN;
    p.strips_h = (H + p.T - 1) / p.T;
    p.strips_w = (W + 127) / 128;
    p.num_strips = N * p.strips_h * p.strips_w * p.num_n_tiles;   // work items of the persistent loop: strips, or (cta2) strip pairs
    if (cta2) p.num_strips /= 2;
    p.a_slot_bytes = 17408;
    p.a_box_bytes = (uint32_t)bwh * 128;
    p.b_stage_bytes = (uint32_t)(cta2 ? p.BN / 2 : p.BN) * 128;
    p.idesc = make_idesc_bf16(cta2 ? 256 : 128, p.BN, 0, 0);
    {
        static int bo = -1;
        if (bo < 0) { const char* e = getenv("STC_CONVH_BO"); bo = e ? atoi(e) : 1; }
        p.bo_mode = bo;
    }
    p.out = y; p.bias = bias; p.residual = residual; p.act = act; p.Cout = Cout;
    p.prof = debug_profile_buffer();
    size_t smem = (size_t)p.a_slots * p.a_slot_bytes + (size_t)p.b_stages * p.b_stage_bytes + (p.staged ? kConvHStagingBytes : 0) +
                  (2 * p.a_slots + 2 * p.b_stages + 4) * 8 + 16 + 1024;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        STC_CUDA(cudaFuncSetAttribute(umma_convh_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        STC_CUDA(cudaFuncSetAttribute(umma_convh_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        STC_CUDA(cudaFuncSetAttribute(umma_convh_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        STC_CUDA(cudaFuncSetAttribute(umma_convh_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        STC_CUDA(cudaFuncSetAttribute(umma_convh_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        STC_CUDA(cudaFuncSetAttribute(umma_convh_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[dev] = true;
    }
    int grid = p.num_strips < num_sms() ? p.num_strips : num_sms();
    if (cta2) {
        const void* fn = p.T == 4 ? (const void*)umma_convh_kernel<4, true> : p.T == 2 ? (const void*)umma_convh_kernel<2, true> : (const void*)umma_convh_kernel<1, true>;
        int pairs = max_cta_pairs(fn, kConvHThreads, smem);
        if (p.num_strips < pairs) pairs = p.num_strips;
        grid = 2 * pairs;
    }
    if (stats) {
        STC_REQUIRE(p.num_n_tiles == 1 && p.BN <= 256, "conv_fprop_convh: fused statistics need one N tile (Cout=%d BN=%d)", Cout, p.BN);
        p.stats = stats;
        *stats_rows = grid * 4;
    }
    if (cta2) {   // clusters of two CTAs (the two SMs of a TPC)
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kConvHThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (p.T == 4) STC_CUDA(cudaLaunchKernelEx(&cfg, umma_convh_kernel<4, true>, p));
        else if (p.T == 2) STC_CUDA(cudaLaunchKernelEx(&cfg, umma_convh_kernel<2, true>, p));
        else STC_CUDA(cudaLaunchKernelEx(&cfg, umma_convh_kernel<1, true>, p));
        return check_launch("umma_convh_kernel (cta_group::2)");
    }
    if (p.T == 4) umma_convh_kernel<4, false><<<grid, kConvHThreads, smem, st>>>(p);
    else if (p.T == 2) umma_convh_kernel<2, false><<<grid, kConvHThreads, smem, st>>>(p);
    else umma_convh_kernel<1, false><<<grid, kConvHThreads, smem, st>>>(p);
    return check_launch("umma_convh_kernel");
}

}  // namespace stc
