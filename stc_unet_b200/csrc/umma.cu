// tcgen05 / TMEM / TMA implicit-GEMM engine (sm_100a).
//
// One persistent, warp-specialised kernel serves the three dense contractions of the path:
//   MODE_CONV  : y[pixel][co] = sum_{tap,ci} x[pixel+tap][ci] * w[tap][co][ci]   (fprop, and dgrad with the
//                flipped/transposed weight pack).  A = NHWC activations fetched per tap by a 4-D TMA box whose
//                out-of-bounds rows/cols are zero-filled (this IS the conv zero padding), B = packed weights.
//   MODE_GEMM  : batched C = alpha * A * B with K-major or MN-major operands (attention QK^T, PV and their grads).
//   MODE_WGRAD : dW[(tap,ci)][co] += sum_pixels x[pixel+tap][ci] * dy[pixel][co]; both operands MN-major,
//                split-K over pixel tiles, fp32 red.add into the workspace.
// Roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread tcgen05.mma issuer,
// warps 2..5 = epilogue (TMEM -> registers -> global).  smem ring of `stages` A/B slots (128B-swizzled),
// two TMEM accumulator stages of 256 columns so the epilogue of tile i overlaps the main loop of tile i+1.
#include <cuda.h>
#include <cstdlib>

#include "common.cuh"
#include "umma.cuh"

namespace stc {

long long* debug_profile_buffer();

enum { MODE_CONV = 0, MODE_GEMM = 1, MODE_WGRAD = 2 };

struct alignas(64) UmmaParams {
    CUtensorMap tmA;
    CUtensorMap tmB;
    CUtensorMap tmC;   // output map for the staged (shared memory + TMA store) epilogue, see store_tma
    int n_src, n_out;   // virtual channel concat: number of A sources / outputs (1 = plain); the descriptors live at the END of the struct
    int mode;
    int num_tiles, num_n_tiles, num_m_tiles;
    int num_k_iters;  // conv/gemm: k iterations per tile; wgrad: total pixel tiles
    int BN, stages, ksteps;
    uint32_t a_stage_bytes, b_stage_bytes;
    int a_boxes, b_boxes;
    uint32_t a_box_bytes, b_box_bytes;
    uint32_t a_lbo, a_sbo, a_kstep_bytes;
    uint32_t b_lbo, b_sbo, b_kstep_bytes;
    int a_mn_major, b_mn_major;
    uint32_t idesc;
    // conv / wgrad geometry
    int H, W, BW, BH, bw_shift, tiles_w, tiles_h, cin_chunks, R, S;
    int num_atoms, k_per_split;
    // gemm
    int M, batch2;
    // epilogue
    void* out;
    const float* bias;
    const void* residual;
    int act, out_dtype, Cout;
    long long ldc, sC1, sC2;
    float alpha;
    // store_tma: each epilogue warp stages 32 rows x 64 bf16 columns in a 128B-swizzled 4 KB buffer (two per warp) and one lane
    // issues a TMA store of that box -> full 128-byte lines instead of 32 scattered 16-byte row segments per instruction
    int store_tma;
    // epi_mode 1 (GEMM, bf16 out): out = alpha * residual * (acc - rowvec[row]) - the softmax backward dS = scale * P * (dP - D) applied in
    // the epilogue of the dP = dO V^T product (residual = P, rowvec = D = rowsum(dO * O)); rowvec index = b1 * rv_s1 + b2 * rv_s2 + m
    // epi_mode 2 / 3 (GEMM, bf16 out, `sweeps` = 2): row softmax WITHOUT a score tensor.  A CTA owns a 128-row block over ALL its N tiles and
    // runs them twice: sweep 0 only reads the accumulators and keeps per-row statistics in the epilogue thread that owns the row (TMEM lane
    // = row, so no cross-thread reduction), sweep 1 recomputes the tiles and stores the final values.
    //   2: P = softmax(sm_scale * bf16(A B^T)) - statistics = running row max / sum of exponentials (the scores S are never written)
    //   3: dS = sm_scale * P * (bf16(A B^T) - D), D = sum(P * dP) / sum(P) over the row, P = `residual` (no dP tensor, no separate pass)
    int epi_mode;
    int sweeps;
    float sm_scale;
    const float* rowvec;
    long long rv_s1, rv_s2;
    // cta_group::2 variant (umma2_kernel): a CTA pair computes a 256 x BN tile; num_tiles counts PAIR tiles, b_* describe ONE CTA's half of B
    int cta2;
    long long* prof;   // diagnostics (stc_debug_profile): 16 clock counters per CTA; nullptr = off
    // ---- cold tail (virtual channel concat, common.cuh ChanCat): extra A sources (CONV / WGRAD) and extra outputs (CONV = dgrad of a concat
    // conv).  Kept BEHIND the hot fields: the roles read this struct through the small constant cache, and 1 KB of descriptors in front of
    // the per-iteration scalars cost the per-tap wgrad 25 % (measured: 0.297 -> 0.390 ms on 512->512 3x3 @64x64)
    int src_chunk_end[kMaxCat];            // cumulative 64-channel chunk counts of the A sources
    int out_ch_end[kMaxCat];               // cumulative channel counts of the outputs
    void* out2[kMaxCat - 1];
    CUtensorMap tmA2[kMaxCat - 1];
    CUtensorMap tmC2[kMaxCat - 1];
};

constexpr uint32_t kStageBufBytes = 32 * 128;   // 32 rows x 64 bf16
constexpr uint32_t kStagingBytes = 4 * 2 * kStageBufBytes;

struct TileInfo {
    int nt, mt;
    int n_img, h0, w0;  // conv
    int b1, b2;         // gemm
    int k0, k1;         // k-iteration range
};

__device__ __forceinline__ void pixel_tile_origin(const UmmaParams& p, int pt, int& n_img, int& h0, int& w0) {
    int per_img = p.tiles_h * p.tiles_w;
    n_img = pt / per_img;
    int rem = pt - n_img * per_img;
    int th = rem / p.tiles_w;
    h0 = th * p.BH;
    w0 = (rem - th * p.tiles_w) * p.BW;
}

// rank: this CTA's rank in its pair (CTA2), which owns M tile 2 * (pair tile) + rank
template <bool CTA2>
__device__ __forceinline__ TileInfo decode_tile(const UmmaParams& p, int tile, int rank) {
    TileInfo t;
    t.nt = tile % p.num_n_tiles;
    int t2 = tile / p.num_n_tiles;
    t.n_img = t.h0 = t.w0 = t.b1 = t.b2 = 0;
    t.k0 = 0;
    t.k1 = p.num_k_iters;
    if (p.mode == MODE_CONV) {
        t.mt = CTA2 ? 2 * t2 + rank : t2;
        pixel_tile_origin(p, t.mt, t.n_img, t.h0, t.w0);
    } else if (p.mode == MODE_GEMM) {
        const int mts = CTA2 ? p.num_m_tiles / 2 : p.num_m_tiles;
        t.mt = t2 % mts;
        int b = t2 / mts;
        if (CTA2) t.mt = 2 * t.mt + rank;
        t.b1 = b / p.batch2;
        t.b2 = b - t.b1 * p.batch2;
    } else {
        t.mt = t2 % p.num_m_tiles;
        int split = t2 / p.num_m_tiles;
        t.k0 = split * p.k_per_split;
        t.k1 = min(p.num_k_iters, t.k0 + p.k_per_split);
    }
    return t;
}

// chunk index (64 channels) within the virtual concat -> source index and chunk inside that source
__device__ __forceinline__ int cat_source(const UmmaParams& p, int cc, int& local) {
    int j = 0;
    while (j + 1 < p.n_src && cc >= p.src_chunk_end[j]) ++j;
    local = cc - (j ? p.src_chunk_end[j - 1] : 0);
    return j;
}
__device__ __forceinline__ const CUtensorMap* cat_map_a(const UmmaParams& p, int j) { return j ? &p.tmA2[j - 1] : &p.tmA; }
// global output channel -> output tensor index, channel inside it and its channel count
__device__ __forceinline__ int cat_output(const UmmaParams& p, int gch, int& local, int& width) {
    int j = 0;
    while (j + 1 < p.n_out && gch >= p.out_ch_end[j]) ++j;
    const int start = j ? p.out_ch_end[j - 1] : 0;
    local = gch - start;
    width = p.out_ch_end[j] - start;
    return j;
}

__device__ __forceinline__ float epi_act(float v, int act) {
    if (act == STC_ACT_RELU) return fmaxf(v, 0.f);
    if (act == STC_ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
    if (act == STC_ACT_HSWISH) return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);
    return v;
}

constexpr int kUmmaThreads = 192;
constexpr int kAccCols = 256;

template <bool CTA2, bool SWEEP2>
__device__ __forceinline__ void umma_body(const UmmaParams& p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024B alignment for the 128B swizzle atoms
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t stage_bytes = p.a_stage_bytes + p.b_stage_bytes;
    const uint32_t staging_bytes = p.store_tma ? kStagingBytes : 0u;   // right after the ring, 1024-byte aligned like it
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes + staging_bytes);
    // bars: [0,stages) full, [stages,2*stages) empty, then tmem_full[2], tmem_empty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = ptx::smem_u32(smem);
    const uint32_t bar_base = ptx::smem_u32(bars);
    const int rank = CTA2 ? (int)ptx::cluster_ctarank() : 0;                 // 0 = leader of the pair
    const int tile0 = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;        // persistent loop: pair (or CTA) index and stride
    const int tstep = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    // two-sweep row modes (epi_mode 2 / 3): the unit of the persistent loop is a row block over all its N tiles, visited twice
    constexpr bool two_sweeps = SWEEP2;   // its own instantiation (umma_sweep_kernel): the plain kernels compile without the extra loop level
    const int nsteps = two_sweeps ? 2 * p.num_n_tiles : 1;
    const int nsup = two_sweeps ? p.num_tiles / p.num_n_tiles : p.num_tiles;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * p.stages + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * p.stages + 2 + s); };

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&p.tmA);
        ptx::prefetch_tensormap(&p.tmB);
        if (p.store_tma) ptx::prefetch_tensormap(&p.tmC);
        for (int j = 1; j < p.n_src; ++j) ptx::prefetch_tensormap(&p.tmA2[j - 1]);
        if (p.store_tma) for (int j = 1; j < p.n_out; ++j) ptx::prefetch_tensormap(&p.tmC2[j - 1]);
        for (int s = 0; s < p.stages; ++s) {
            ptx::mbar_init(full_bar(s), 1);
            ptx::mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(tfull_bar(s), 1);
            ptx::mbar_init(tempty_bar(s), CTA2 ? 8 : 4);   // cta2: the epilogue warps of BOTH CTAs arrive on the leader's barrier
        }
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
    if (tmem_base != 0) {  // we allocate all 512 columns, so the base must be column 0 / lane 0
        if (threadIdx.x == 0) printf("stc_b200: unexpected TMEM base 0x%x\n", tmem_base);
        __trap();
    }

    if (warp == 0) {
        // ===================== TMA producer (warp-uniform; one elected lane issues) =====================
        // The k loop carries its coordinates incrementally: no integer division per iteration.  (With the tap / chunk / pixel-tile
        // decompositions recomputed every iteration the two dependent division chains of the per-tap wgrad took longer than the 8 MMAs of
        // a k iteration - the producer, not the tensor pipe, paced that kernel.)
        {
            int stage = 0;
            uint32_t phase = 0;
            const int bn_off = CTA2 ? rank * (p.BN / 2) : 0;   // cta2: this CTA stages its half of the N columns of B
            const int pw_end = p.tiles_w * p.BW, ph_end = p.tiles_h * p.BH;
            for (int sup = tile0; sup < nsup; sup += tstep) for (int step = 0; step < nsteps; ++step) {
                const int tile = two_sweeps ? sup * p.num_n_tiles + (step >= p.num_n_tiles ? step - p.num_n_tiles : step) : sup;
                TileInfo t = decode_tile<CTA2>(p, tile, rank);
                int cc = 0, r = 0, s = 0, tap = 0;            // CONV: k iteration = (tap = r * S + s, 64-channel chunk cc)
                int pn = 0, ph0 = 0, pw0 = 0;                 // WGRAD: k iteration = pixel tile (image pn, origin ph0 / pw0)
                int a_cc[2] = {0, 0}, a_r[2] = {0, 0}, a_s[2] = {0, 0};   // WGRAD: (chunk, tap row, tap column) of the tile's two 64-row atoms
                if (p.mode == MODE_WGRAD) {
                    pixel_tile_origin(p, t.k0, pn, ph0, pw0);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        int atom = 2 * t.mt + j;
                        if (atom >= p.num_atoms) atom = 0;  // dummy rows, dropped by the epilogue
                        const int atap = atom / p.cin_chunks;
                        a_cc[j] = atom - atap * p.cin_chunks;
                        a_r[j] = atap / p.S;
                        a_s[j] = atap - a_r[j] * p.S;
                    }
                }
                for (int kt = t.k0; kt < t.k1; ++kt) {
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t a_dst = smem_base + stage * stage_bytes;
                    const uint32_t b_dst = a_dst + p.a_stage_bytes;
                    // cta2: the TMA bytes of both CTAs complete on the LEADER's full barrier, which the leader arms for 2 x stage_bytes
                    const uint32_t fb = CTA2 ? ptx::mapa_shared(full_bar(stage), 0) : full_bar(stage);
                    auto ld4 = [&](uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, int c3) {
                        if (CTA2) ptx::tma_load_4d_2sm(dst, m, fb, c0, c1, c2, c3);
                        else ptx::tma_load_4d(dst, m, fb, c0, c1, c2, c3);
                    };
                    if (ptx::elect_one_sync()) {
                        if (!CTA2) ptx::mbar_arrive_expect_tx(fb, stage_bytes);
                        else if (rank == 0) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * stage_bytes);
                        if (p.mode == MODE_CONV) {
                            if (p.n_src > 1) {   // virtual concat: the chunk's source tensor has its own descriptor
                                int lc;
                                const CUtensorMap* ma = cat_map_a(p, cat_source(p, cc, lc));
                                ld4(a_dst, ma, lc * 64, t.w0 + s - p.S / 2, t.h0 + r - p.R / 2, t.n_img);
                            } else {
                                ld4(a_dst, &p.tmA, cc * 64, t.w0 + s - p.S / 2, t.h0 + r - p.R / 2, t.n_img);
                            }
                            if (CTA2) ptx::tma_load_3d_2sm(b_dst, &p.tmB, fb, cc * 64, t.nt * p.BN + bn_off, tap);
                            else ptx::tma_load_3d(b_dst, &p.tmB, fb, cc * 64, t.nt * p.BN, tap);
                        } else if (p.mode == MODE_GEMM) {
                            if (!p.a_mn_major) {
                                ld4(a_dst, &p.tmA, kt * 64, t.mt * 128, t.b2, t.b1);
                            } else {
                                for (int j = 0; j < p.a_boxes; ++j) ld4(a_dst + j * p.a_box_bytes, &p.tmA, t.mt * 128 + j * 64, kt * 64, t.b2, t.b1);
                            }
                            if (!p.b_mn_major) {
                                ld4(b_dst, &p.tmB, kt * 64, t.nt * p.BN + bn_off, t.b2, t.b1);
                            } else {
                                for (int j = 0; j < p.b_boxes; ++j)
                                    ld4(b_dst + j * p.b_box_bytes, &p.tmB, t.nt * p.BN + bn_off + j * 64, kt * 64, t.b2, t.b1);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                if (p.n_src > 1) {
                                    int lc;
                                    const CUtensorMap* ma = cat_map_a(p, cat_source(p, a_cc[j], lc));
                                    ld4(a_dst + j * p.a_box_bytes, ma, lc * 64, pw0 + a_s[j] - p.S / 2, ph0 + a_r[j] - p.R / 2, pn);
                                } else {
                                    ld4(a_dst + j * p.a_box_bytes, &p.tmA, a_cc[j] * 64, pw0 + a_s[j] - p.S / 2, ph0 + a_r[j] - p.R / 2, pn);
                                }
                            }
                            for (int j = 0; j < p.b_boxes; ++j) ld4(b_dst + j * p.b_box_bytes, &p.tmB, t.nt * p.BN + j * 64, pw0, ph0, pn);
                        }
                    }
                    __syncwarp();
                    if (p.mode == MODE_CONV) {
                        if (++cc == p.cin_chunks) {
                            cc = 0;
                            ++tap;
                            if (++s == p.S) { s = 0; ++r; }
                        }
                    } else if (p.mode == MODE_WGRAD) {
                        pw0 += p.BW;
                        if (pw0 >= pw_end) {
                            pw0 = 0;
                            ph0 += p.BH;
                            if (ph0 >= ph_end) { ph0 = 0; ++pn; }
                        }
                    }
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform; one elected lane issues; cta2: the leader CTA only) =====================
        if (!CTA2 || rank == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase[2] = {0, 0};
            long long w_te = 0, w_f = 0, n_t = 0;
            const long long t_begin = p.prof ? clock64() : 0;
            for (int sup = tile0; sup < nsup; sup += tstep) for (int step = 0; step < nsteps; ++step) {
                const int tile = two_sweeps ? sup * p.num_n_tiles + (step >= p.num_n_tiles ? step - p.num_n_tiles : step) : sup;
                TileInfo t = decode_tile<CTA2>(p, tile, rank);
                if (t.k1 <= t.k0) continue;
                long long t0 = p.prof ? clock64() : 0;
                ptx::mbar_wait(tempty_bar(acc), acc_phase[acc] ^ 1);
                if (p.prof) { w_te += clock64() - t0; ++n_t; }
                ptx::tc_fence_after();
                const uint32_t d_tmem = acc * kAccCols;  // TMEM base is 0: the CTA owns all 512 columns (checked above)
                for (int kt = t.k0; kt < t.k1; ++kt) {
                    t0 = p.prof ? clock64() : 0;
                    ptx::mbar_wait(full_bar(stage), phase);
                    if (p.prof) w_f += clock64() - t0;
                    ptx::tc_fence_after();
                    const uint32_t a_addr = smem_base + stage * stage_bytes;
                    const uint32_t b_addr = a_addr + p.a_stage_bytes;
                    const uint64_t a_desc0 = ptx::make_smem_desc_sw128(a_addr, p.a_lbo, p.a_sbo);
                    const uint64_t b_desc0 = ptx::make_smem_desc_sw128(b_addr, p.b_lbo, p.b_sbo);
                    if (ptx::elect_one_sync()) {
                        for (int ks = 0; ks < p.ksteps; ++ks) {
                            // advancing the 14-bit start-address field (16B units); never carries out of it (smem < 256 KB)
                            uint64_t a_desc = a_desc0 + (uint64_t)((ks * p.a_kstep_bytes) >> 4);
                            uint64_t b_desc = b_desc0 + (uint64_t)((ks * p.b_kstep_bytes) >> 4);
                            if (CTA2) ptx::mma_bf16_ss2(d_tmem, a_desc, b_desc, p.idesc, (kt > t.k0 || ks > 0) ? 1u : 0u);
                            else ptx::mma_bf16_ss(d_tmem, a_desc, b_desc, p.idesc, (kt > t.k0 || ks > 0) ? 1u : 0u);
                        }
                        // frees the smem slot (cta2: in both CTAs) once these MMAs retire
                        if (CTA2) ptx::tc_commit2(empty_bar(stage), 3); else ptx::tc_commit(empty_bar(stage));
                    }
                    __syncwarp();
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                if (ptx::elect_one_sync()) {   // accumulator ready for the epilogue (cta2: of both CTAs)
                    if (CTA2) ptx::tc_commit2(tfull_bar(acc), 3); else ptx::tc_commit(tfull_bar(acc));
                }
                __syncwarp();
                acc_phase[acc] ^= 1;
                acc ^= 1;
            }
            if (p.prof && lane == 0) {
                long long* o = p.prof + (blockIdx.x % 148) * 16;
                o[0] = clock64() - t_begin; o[1] = w_te; o[2] = w_f; o[4] = n_t;
            }
        }
    } else {
        // ===================== epilogue (4 warps = 128 TMEM lanes) =====================
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        int acc = 0, stage_buf = 0;
        uint32_t acc_phase[2] = {0, 0};
        const uint32_t tempty_leader0 = CTA2 ? ptx::mapa_shared(tempty_bar(0), 0) : 0u, tempty_leader1 = CTA2 ? ptx::mapa_shared(tempty_bar(1), 0) : 0u;
        long long w_tf = 0, t_work = 0;
        float rs0 = 0.f, rs1 = 0.f;   // two-sweep row modes: this thread's row statistics (max / sum, or sum(P dP) / sum(P))
        auto release_acc = [&](int a) {
            if (CTA2) ptx::mbar_arrive_cluster(a ? tempty_leader1 : tempty_leader0);
            else ptx::mbar_arrive(tempty_bar(a));
        };
        for (int sup = tile0; sup < nsup; sup += tstep) for (int step = 0; step < nsteps; ++step) {
                const int tile = two_sweeps ? sup * p.num_n_tiles + (step >= p.num_n_tiles ? step - p.num_n_tiles : step) : sup;
            TileInfo t = decode_tile<CTA2>(p, tile, rank);
            if (t.k1 <= t.k0) continue;
            bool valid;
            long long off;
            if (p.mode == MODE_CONV) {
                int hl = row >> p.bw_shift, wl = row & (p.BW - 1);
                int h = t.h0 + hl, w = t.w0 + wl;
                valid = (h < p.H) && (w < p.W);
                off = (((long long)t.n_img * p.H + h) * p.W + w) * p.Cout + (long long)t.nt * p.BN;
            } else if (p.mode == MODE_GEMM) {
                int m = t.mt * 128 + row;
                valid = m < p.M;
                off = t.b1 * p.sC1 + t.b2 * p.sC2 + (long long)m * p.ldc + (long long)t.nt * p.BN;
            } else {
                int atom = 2 * t.mt + (row >> 6);
                valid = atom < p.num_atoms;
                off = ((long long)atom * 64 + (row & 63)) * p.Cout + (long long)t.nt * p.BN;
            }
            // epi_mode 1: alpha * D[row], subtracted from alpha * acc before the product with the residual (P)
            const float dsub = (p.epi_mode == 1 && valid) ? p.alpha * __ldg(p.rowvec + t.b1 * p.rv_s1 + t.b2 * p.rv_s2 + (t.mt * 128 + row)) : 0.f;
            const long long pt0 = p.prof ? clock64() : 0;
            ptx::mbar_wait(tfull_bar(acc), acc_phase[acc]);
            const long long pt1 = p.prof ? clock64() : 0;
            w_tf += pt1 - pt0;
            ptx::tc_fence_after();
            const uint32_t t_addr = tmem_base + acc * kAccCols + ((uint32_t)(q * 32) << 16);
            if (SWEEP2) {
                // ---- two-sweep row softmax (see UmmaParams::epi_mode): this thread owns row `row` of the block for all N tiles and both sweeps ----
                const bool fwd = p.epi_mode == 2;
                const int sweep = step >= p.num_n_tiles ? 1 : 0;
                if (step == 0) { rs0 = fwd ? -INFINITY : 0.f; rs1 = 0.f; }
                if (step == p.num_n_tiles) {   // statistics complete: fwd 1 / sum of exponentials; bwd D = sum(P dP) / sum(P)
                    if (fwd) rs1 = 1.f / rs1; else rs0 = rs0 / rs1;
                }
                const int c1 = t.mt * 128 + q * 32, c2 = t.b2, c3 = t.b1;
                const uint32_t stg0 = smem_base + p.stages * stage_bytes + (uint32_t)q * 2u * kStageBufBytes;
                for (int c = 0; c < p.BN; c += 64) {
                    const uint32_t stg = stg0 + (uint32_t)stage_buf * kStageBufBytes;
                    if (sweep) {
                        if (lane == 0) ptx::bulk_wait_read<1>();
                        __syncwarp();
                    }
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint32_t v[32];
                        ptx::tmem_ld_32x32(t_addr + c + half * 32, v);
                        ptx::tmem_ld_wait();
                        float f[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j)   // the product rounded to bf16, exactly as the unfused path stores it
                            f[j] = __bfloat162float(__float2bfloat16_rn(__uint_as_float(v[j]) * p.alpha));
                        if (fwd) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] *= p.sm_scale;
                            if (!sweep) {
                                float cm = f[0];
#pragma unroll
                                for (int j = 1; j < 32; ++j) cm = fmaxf(cm, f[j]);
                                const float m_new = fmaxf(rs0, cm);
                                float add = 0.f;
#pragma unroll
                                for (int j = 0; j < 32; ++j) add += __expf(f[j] - m_new);
                                rs1 = rs1 * __expf(rs0 - m_new) + add;
                                rs0 = m_new;
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j) f[j] = __expf(f[j] - rs0) * rs1;
                            }
                        } else {
                            const uint4* r = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.residual) + off + c + half * 32);
                            float x[32];
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                uint4 rv = __ldg(r + g);
                                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 xx = __bfloat1622float2(h[e]);
                                    x[g * 8 + 2 * e] = xx.x;
                                    x[g * 8 + 2 * e + 1] = xx.y;
                                }
                            }
                            if (!sweep) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) { rs0 = fmaf(x[j], f[j], rs0); rs1 += x[j]; }
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j) f[j] = p.sm_scale * x[j] * (f[j] - rs0);
                            }
                        }
                        if (sweep) {
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                uint4 ov;
                                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
                                for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(f[g * 8 + 2 * e], f[g * 8 + 2 * e + 1]);
                                ptx::st_shared_v4(stg + (uint32_t)lane * 128u + (uint32_t)(((half * 4 + g) ^ (lane & 7)) << 4), ov);
                            }
                        }
                    }
                    if (sweep) {
                        ptx::fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            ptx::tma_store_4d(&p.tmC, stg, t.nt * p.BN + c, c1, c2, c3);
                            ptx::bulk_commit();
                        }
                        stage_buf ^= 1;
                    }
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) release_acc(acc);
                acc_phase[acc] ^= 1;
                acc ^= 1;
                if (p.prof) t_work += clock64() - pt1;
                continue;
            }
            if (p.store_tma) {
                // ---- staged epilogue: TMEM -> registers -> swizzled smem -> TMA store (bf16 out, CONV / GEMM modes) ----
                int c1, c2, c3;   // box origin below the channel coordinate
                if (p.mode == MODE_CONV) {
                    const int r0 = q * 32;
                    c1 = t.w0 + (r0 & (p.BW - 1));
                    c2 = t.h0 + (r0 >> p.bw_shift);
                    c3 = t.n_img;
                } else {
                    c1 = t.mt * 128 + q * 32;
                    c2 = t.b2;
                    c3 = t.b1;
                }
                const uint32_t stg0 = smem_base + p.stages * stage_bytes + (uint32_t)q * 2u * kStageBufBytes;
                for (int c = 0; c < p.BN; c += 64) {
                    const uint32_t stg = stg0 + (uint32_t)stage_buf * kStageBufBytes;
                    if (lane == 0) ptx::bulk_wait_read<1>();   // the store issued from this buffer two chunks ago has read it
                    __syncwarp();
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint32_t v[32];
                        ptx::tmem_ld_32x32(t_addr + c + half * 32, v);
                        ptx::tmem_ld_wait();
                        float f[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * p.alpha;
                        if (p.bias) {
                            const float* b = p.bias + t.nt * p.BN + c + half * 32;
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] += __ldg(b + j);
                        }
                        if (p.residual && valid) {
                            const uint4* r = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.residual) + off + c + half * 32);
                            float x[32];
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                uint4 rv = __ldg(r + g);
                                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 xx = __bfloat1622float2(h[e]);
                                    x[g * 8 + 2 * e] = xx.x;
                                    x[g * 8 + 2 * e + 1] = xx.y;
                                }
                            }
                            if (p.epi_mode == 1) {   // softmax backward: (alpha * acc - alpha * D[row]) * P
#pragma unroll
                                for (int j = 0; j < 32; ++j) f[j] = (f[j] - dsub) * x[j];
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j) f[j] += x[j];
                            }
                        }
                        if (p.act == STC_ACT_RELU) {   // uniform branch per chunk: a per-element switch compiles to a jump table
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
                        } else if (p.act != STC_ACT_NONE) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = epi_act(f[j], p.act);
                        }
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            uint4 ov;
                            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
                            for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(f[g * 8 + 2 * e], f[g * 8 + 2 * e + 1]);
                            // 128B swizzle: 16-byte chunk index XOR (row mod 8); the buffer is 1024-byte aligned
                            ptx::st_shared_v4(stg + (uint32_t)lane * 128u + (uint32_t)(((half * 4 + g) ^ (lane & 7)) << 4), ov);
                        }
                    }
                    ptx::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        int oc = t.nt * p.BN + c, ow;
                        const CUtensorMap* mc = &p.tmC;
                        if (p.n_out > 1) { const int j = cat_output(p, oc, oc, ow); if (j) mc = &p.tmC2[j - 1]; }
                        ptx::tma_store_4d(mc, stg, oc, c1, c2, c3);
                        ptx::bulk_commit();
                    }
                    stage_buf ^= 1;
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) release_acc(acc);
                acc_phase[acc] ^= 1;
                acc ^= 1;
                if (p.prof) t_work += clock64() - pt1;
                continue;
            }
            for (int c = 0; c < p.BN; c += 32) {
                uint32_t v[32];
                ptx::tmem_ld_32x32(t_addr + c, v);
                ptx::tmem_ld_wait();
                if (!valid) continue;
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * p.alpha;
                if (p.mode == MODE_WGRAD) {
                    float* o = reinterpret_cast<float*>(p.out) + off + c;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)   // 16-byte vector reductions: 4x fewer L2 atomic instructions
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + j), "f"(f[j]), "f"(f[j + 1]), "f"(f[j + 2]),
                                     "f"(f[j + 3])
                                     : "memory");
                    continue;
                }
                if (p.bias) {
                    const float* b = p.bias + t.nt * p.BN + c;
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] += __ldg(b + j);
                }
                if (p.residual) {
                    if (p.out_dtype == STC_BF16) {
                        const uint4* r = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.residual) + off + c);
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            uint4 rv = __ldg(r + g);
                            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float2 x = __bfloat1622float2(h[e]);
                                if (p.epi_mode) {
                                    f[g * 8 + 2 * e] = (f[g * 8 + 2 * e] - dsub) * x.x;
                                    f[g * 8 + 2 * e + 1] = (f[g * 8 + 2 * e + 1] - dsub) * x.y;
                                } else {
                                    f[g * 8 + 2 * e] += x.x;
                                    f[g * 8 + 2 * e + 1] += x.y;
                                }
                            }
                        }
                    } else {
                        const float* r = reinterpret_cast<const float*>(p.residual) + off + c;
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] += __ldg(r + j);
                    }
                }
                if (p.act == STC_ACT_RELU) {   // uniform branch per chunk: a per-element switch compiles to a jump table
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
                } else if (p.act != STC_ACT_NONE) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = epi_act(f[j], p.act);
                }
                if (p.out_dtype == STC_BF16) {
                    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + off + c);
                    if (p.n_out > 1) {   // dgrad of a virtual concat: this 32-channel chunk belongs to one of the outputs
                        int lc, ow;
                        const int j = cat_output(p, t.nt * p.BN + c, lc, ow);
                        const long long pix = (off - (long long)t.nt * p.BN) / p.Cout;
                        o = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(j ? p.out2[j - 1] : p.out) + pix * ow + lc);
                    }
                    uint4 ov[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&ov[g]);
#pragma unroll
                        for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(f[g * 8 + 2 * e], f[g * 8 + 2 * e + 1]);
                    }
                    if ((reinterpret_cast<uintptr_t>(o) & 31) == 0) {   // two full 32-byte sectors per lane
                        ptx::st_global_v8(o, ov[0], ov[1]);
                        ptx::st_global_v8(o + 2, ov[2], ov[3]);
                    } else {
#pragma unroll
                        for (int g = 0; g < 4; ++g) o[g] = ov[g];
                    }
                } else {
                    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off + c);
#pragma unroll
                    for (int g = 0; g < 8; ++g) o[g] = make_float4(f[4 * g], f[4 * g + 1], f[4 * g + 2], f[4 * g + 3]);
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) release_acc(acc);
            acc_phase[acc] ^= 1;
            acc ^= 1;
            if (p.prof) t_work += clock64() - pt1;
        }
        if (p.prof && q == 0 && lane == 0) { p.prof[(blockIdx.x % 148) * 16 + 5] = w_tf; p.prof[(blockIdx.x % 148) * 16 + 6] = t_work; }
        if (p.store_tma && lane == 0) ptx::bulk_wait<0>();   // all output boxes written before the CTA retires
    }

    ptx::tc_fence_before();
    if (CTA2) ptx::cluster_sync_all(); else __syncthreads();   // cta2: the peer's shared memory / barriers stay alive until both are done
    if (warp == 1) {
        ptx::tc_fence_after();
        if (CTA2) ptx::tmem_dealloc2(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512);
    }
}

__global__ void __launch_bounds__(kUmmaThreads, 1) umma_kernel(const __grid_constant__ UmmaParams p) { umma_body<false, false>(p); }
// the two-sweep row-softmax GEMM (epi_mode 2 / 3)
__global__ void __launch_bounds__(kUmmaThreads, 1) umma_sweep_kernel(const __grid_constant__ UmmaParams p) { umma_body<false, true>(p); }
// the cta_group::2 variant: CTA pairs (cluster of 2 = the two SMs of a TPC) computing 256 x BN tiles
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kUmmaThreads, 1) umma2_kernel(const __grid_constant__ UmmaParams p) { umma_body<true, false>(p); }

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// rank-`rank` bf16 tensor map with 128B swizzle; dims/strides innermost first; strides[0] is implicit (2 bytes)
static int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, bool swizzle64 = false) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return STC_ERR_CUDA;
    }
    cuuint64_t gdims[5], gstr[4];
    cuuint32_t gbox[5], estr[5];
    for (int i = 0; i < rank; ++i) {
        gdims[i] = dims[i];
        gbox[i] = box[i];
        estr[i] = 1;
        if (i > 0) gstr[i - 1] = strides_bytes[i];
    }
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr, gbox, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u] base %p", (int)r, rank,
                  (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
                  (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, base);
        return STC_ERR_CUDA;
    }
    return STC_OK;
}

static bool getenv_off(const char* name) {
    const char* v = getenv(name);
    return v && v[0] == '0';
}

static int pick_bn(int n) {
    const int cands[] = {256, 192, 128, 96, 64, 32};
    for (int c : cands)
        if (n % c == 0) return c;
    return 0;
}

int encode_map_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
    return encode_map(m, base, rank, dims, strides_bytes, box);
}
// 64-byte swizzle (inner box = 32 bf16): the half-width dy tiles of the paired halo wgrad
int encode_map_bf16_sw64(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
    return encode_map(m, base, rank, dims, strides_bytes, box, true);
}
int umma_pick_bn(int n) { return pick_bn(n); }

static int launch(UmmaParams& p, cudaStream_t st) {
    size_t smem = (size_t)p.stages * (p.a_stage_bytes + p.b_stage_bytes) + (p.store_tma ? kStagingBytes : 0) + (2 * p.stages + 4) * 8 + 16 + 1024;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        STC_CUDA(cudaFuncSetAttribute(umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[dev] = true;
    }
    if (smem > 227 * 1024) {
        set_error("umma: smem %zu too large", smem);
        return STC_ERR_INVALID;
    }
    p.prof = debug_profile_buffer();
    if (p.cta2) {
        static bool attr2[64] = {false};
        if (dev >= 0 && dev < 64 && !attr2[dev]) {
            STC_CUDA(cudaFuncSetAttribute(umma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            attr2[dev] = true;
        }
        int pairs = max_cta_pairs((const void*)umma2_kernel, kUmmaThreads, smem);
        if (p.num_tiles < pairs) pairs = p.num_tiles;
        if (pairs <= 0) return STC_OK;
        umma2_kernel<<<2 * pairs, kUmmaThreads, smem, st>>>(p);
        return check_launch("umma2_kernel");
    }
    int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
    if (grid <= 0) return STC_OK;
    if (p.sweeps == 2) {
        static bool attr3[64] = {false};
        if (dev >= 0 && dev < 64 && !attr3[dev]) {
            STC_CUDA(cudaFuncSetAttribute(umma_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            attr3[dev] = true;
        }
        const int blocks = p.num_tiles / p.num_n_tiles;   // row blocks: the unit of the persistent loop
        umma_sweep_kernel<<<blocks < grid ? blocks : grid, kUmmaThreads, smem, st>>>(p);
        return check_launch("umma_sweep_kernel");
    }
    umma_kernel<<<grid, kUmmaThreads, smem, st>>>(p);
    return check_launch("umma_kernel");
}

// cta_group::2 is used when the M tiles pair up and each CTA's half of B keeps the layout rules (K-major: BN/2 rows, a multiple of 16;
// MN-major: whole 64-column atoms).  STC_CTA2=0 switches it off (the 1-CTA kernel is the reference for the tests).
// Measured (profiles/r2_gemm_probe.txt): the pair wins where the main loop dominates (PV / dV / dQ / dK with K = 4096: +6 %, the W < 128
// 3x3 convolutions: +2-3 %) and loses on short-K tiles, which are bound by the accumulator read-out of the epilogue (QK^T, K = 256: -13 %),
// so it is taken from 16 k-iterations (K >= 1024) on.  STC_CTA2=0 never, STC_CTA2=2 whenever the shapes allow (tests).
static bool use_cta2(int num_m_tiles_per_batch, int BN, bool b_mn_major, int k_iters) {
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("STC_CTA2"); mode = e ? atoi(e) : 1; }
    if (mode == 0 || num_m_tiles_per_batch % 2 != 0 || BN % 32 != 0) return false;
    if (b_mn_major && BN % 128 != 0) return false;
    return mode == 2 || k_iters >= 16;
}

static void conv_geometry(UmmaParams& p, int H, int W, int R, int S, int Cin) {
    int bw = 1;
    while (bw * 2 <= W && bw * 2 <= 128) bw *= 2;
    p.BW = bw;
    p.BH = 128 / bw;
    p.bw_shift = 0;
    while ((1 << p.bw_shift) < bw) ++p.bw_shift;
    p.tiles_w = (W + p.BW - 1) / p.BW;
    p.tiles_h = (H + p.BH - 1) / p.BH;
    p.H = H;
    p.W = W;
    p.R = R;
    p.S = S;
    p.cin_chunks = Cin / 64;
}

static int pick_stages(uint32_t stage_bytes, bool staged_epilogue = false) {
    int s = (int)(((staged_epilogue ? 192 : 200) * 1024) / stage_bytes);
    if (s > 8) s = 8;
    return s;
}

bool conv_umma_eligible(int Cin, int Cout, int dtype) { return dtype == STC_BF16 && Cin % 64 == 0 && pick_bn(Cout) != 0; }

// NHWC activation map {C, W, H, N} with the given box
static int encode_nhwc(CUtensorMap* m, const void* base, int N, int H, int W, int C, uint32_t bc, uint32_t bw, uint32_t bh) {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[4] = {2, (uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {bc, bw, bh, 1};
    return encode_map(m, base, 4, dims, str, box);
}

// checks a ChanCat (every part a non-null 16-byte aligned pointer with a multiple of 64 channels, sum = total)
int check_cat(const ChanCat* c, int total, const char* what) {
    if (!c) return STC_OK;
    STC_REQUIRE(c->n >= 1 && c->n <= kMaxCat && c->total() == total, "%s: %d parts do not add up to %d channels", what, c->n, total);
    for (int i = 0; i < c->n; ++i)
        STC_REQUIRE(c->ptr[i] && ((uintptr_t)c->ptr[i] & 15) == 0 && c->c[i] > 0 && c->c[i] % 64 == 0,
                    "%s: part %d must be a 16-byte aligned tensor with a multiple of 64 channels (has %d)", what, i, c->c[i]);
    return STC_OK;
}

// src: the input is a virtual concat (x ignored); dst: the output is split over several tensors (y ignored; no bias / residual / act)
int conv_fprop_umma(const void* x, const void* wp, const float* bias, const void* residual, void* y, int N, int H, int W,
                    int Cin, int Cout, int R, int S, int act, cudaStream_t st, const ChanCat* src, const ChanCat* dst) {
    STC_REQUIRE(conv_umma_eligible(Cin, Cout, STC_BF16), "conv_fprop_umma: shape Cin=%d Cout=%d not eligible", Cin, Cout);
    if (src && src->n == 1) { x = src->ptr[0]; src = nullptr; }
    if (dst && dst->n == 1) { y = const_cast<void*>(dst->ptr[0]); dst = nullptr; }
    if (int rc = check_cat(src, Cin, "conv_fprop_umma input")) return rc;
    if (int rc = check_cat(dst, Cout, "conv_fprop_umma output")) return rc;
    STC_REQUIRE(!dst || (!bias && !residual && act == STC_ACT_NONE), "conv_fprop_umma: a split output takes no bias / residual / activation");
    if (src) x = src->ptr[0];
    if (dst) y = const_cast<void*>(dst->ptr[0]);
    STC_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)wp & 15) == 0 && ((uintptr_t)y & 15) == 0, "conv_fprop_umma: unaligned pointer");
    UmmaParams p;
    memset(&p, 0, sizeof(p));
    p.mode = MODE_CONV;
    conv_geometry(p, H, W, R, S, Cin);
    p.BN = pick_bn(Cout);
    p.n_src = 1; p.n_out = 1;
    p.num_n_tiles = Cout / p.BN;
    p.num_m_tiles = N * p.tiles_h * p.tiles_w;
    p.cta2 = use_cta2(p.num_m_tiles, p.BN, false, R * S * p.cin_chunks) ? 1 : 0;
    const int bn_cta = p.cta2 ? p.BN / 2 : p.BN;          // rows of B (output channels) staged by one CTA
    if (src) {
        p.n_src = src->n;
        int acc = 0;
        for (int j = 0; j < src->n; ++j) {
            int rc = encode_nhwc(j ? &p.tmA2[j - 1] : &p.tmA, src->ptr[j], N, H, W, src->c[j], 64, (uint32_t)p.BW, (uint32_t)p.BH);
            if (rc) return rc;
            acc += src->c[j] / 64;
            p.src_chunk_end[j] = acc;
        }
    } else {
        int rc = encode_nhwc(&p.tmA, x, N, H, W, Cin, 64, (uint32_t)p.BW, (uint32_t)p.BH);
        if (rc) return rc;
    }
    {
        uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)(R * S)};
        uint64_t str[3] = {2, (uint64_t)Cin * 2, (uint64_t)Cout * Cin * 2};
        uint32_t box[3] = {64, (uint32_t)bn_cta, 1};
        int rc = encode_map(&p.tmB, wp, 3, dims, str, box);
        if (rc) return rc;
    }
    p.num_tiles = (p.cta2 ? p.num_m_tiles / 2 : p.num_m_tiles) * p.num_n_tiles;
    p.num_k_iters = R * S * p.cin_chunks;
    p.ksteps = 4;
    p.a_boxes = p.b_boxes = 1;
    p.a_stage_bytes = p.a_box_bytes = 128 * 128;
    p.b_stage_bytes = p.b_box_bytes = (uint32_t)bn_cta * 128;
    p.a_lbo = 0; p.a_sbo = 1024; p.a_kstep_bytes = 32;
    p.b_lbo = 0; p.b_sbo = 1024; p.b_kstep_bytes = 32;
    p.idesc = make_idesc_bf16(p.cta2 ? 256 : 128, p.BN, 0, 0);
    p.store_tma = (p.BN % 64 == 0 && !getenv_off("STC_TMA_STORE")) ? 1 : 0;
    if (dst) {
        p.n_out = dst->n;
        int acc = 0;
        for (int j = 0; j < dst->n; ++j) {
            acc += dst->c[j];
            p.out_ch_end[j] = acc;
            if (j) p.out2[j - 1] = const_cast<void*>(dst->ptr[j]);
        }
    }
    if (p.store_tma) {   // per epilogue warp: 32 tile rows = a (bw32 x 32/bw32) pixel patch, 64 channels
        const uint32_t bw32 = (uint32_t)(p.BW < 32 ? p.BW : 32);
        for (int j = 0; j < p.n_out; ++j) {
            int rc = encode_nhwc(j ? &p.tmC2[j - 1] : &p.tmC, dst ? dst->ptr[j] : y, N, H, W, dst ? dst->c[j] : Cout, 64, bw32, 32 / bw32);
            if (rc) return rc;
        }
    }
    p.stages = pick_stages(p.a_stage_bytes + p.b_stage_bytes, p.store_tma);
    p.out = y; p.bias = bias; p.residual = residual; p.act = act; p.out_dtype = STC_BF16; p.Cout = Cout;
    p.alpha = 1.f;
    return launch(p, st);
}

int conv_wgrad_umma(const void* x, const void* dy, float* ws, int N, int H, int W, int Cin, int Cout, int R, int S,
                    cudaStream_t st, const ChanCat* src) {
    STC_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "conv_wgrad_umma: Cin=%d Cout=%d must be multiples of 64", Cin, Cout);
    if (src && src->n == 1) { x = src->ptr[0]; src = nullptr; }
    if (int rc = check_cat(src, Cin, "conv_wgrad_umma input")) return rc;
    UmmaParams p;
    memset(&p, 0, sizeof(p));
    p.mode = MODE_WGRAD;
    conv_geometry(p, H, W, R, S, Cin);
    p.BN = Cout % 128 == 0 ? 128 : 64;
    p.n_src = 1; p.n_out = 1;
    if (src) {
        p.n_src = src->n;
        int acc = 0;
        for (int j = 0; j < src->n; ++j) {
            int rc = encode_nhwc(j ? &p.tmA2[j - 1] : &p.tmA, src->ptr[j], N, H, W, src->c[j], 64, (uint32_t)p.BW, (uint32_t)p.BH);
            if (rc) return rc;
            acc += src->c[j] / 64;
            p.src_chunk_end[j] = acc;
        }
    } else {
        int rc = encode_nhwc(&p.tmA, x, N, H, W, Cin, 64, (uint32_t)p.BW, (uint32_t)p.BH);
        if (rc) return rc;
    }
    {
        uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[4] = {2, (uint64_t)Cout * 2, (uint64_t)W * Cout * 2, (uint64_t)H * W * Cout * 2};
        uint32_t box[4] = {64, (uint32_t)p.BW, (uint32_t)p.BH, 1};
        int rc = encode_map(&p.tmB, dy, 4, dims, str, box);
        if (rc) return rc;
    }
    p.num_atoms = R * S * p.cin_chunks;
    p.num_m_tiles = (p.num_atoms + 1) / 2;
    p.num_n_tiles = Cout / p.BN;
    p.num_k_iters = N * p.tiles_h * p.tiles_w;  // pixel tiles of 128
    int base_tiles = p.num_m_tiles * p.num_n_tiles;
    // split-K so that one wave of CTAs covers the tiles: every extra split costs a full 128 x BN fp32 reduction into the workspace
    int splits = num_sms() / base_tiles;
    if (const char* e = getenv("STC_WGRAD_WAVES")) splits = (num_sms() * atoi(e) + base_tiles - 1) / base_tiles;
    if (splits > p.num_k_iters) splits = p.num_k_iters;
    if (splits < 1) splits = 1;
    p.k_per_split = (p.num_k_iters + splits - 1) / splits;
    splits = (p.num_k_iters + p.k_per_split - 1) / p.k_per_split;
    p.num_tiles = base_tiles * splits;
    p.ksteps = 8;  // 128 pixels per stage / UMMA_K 16
    p.a_boxes = 2;
    p.a_box_bytes = 128 * 128;
    p.a_stage_bytes = 2 * p.a_box_bytes;
    p.b_boxes = p.BN / 64;
    p.b_box_bytes = 128 * 128;
    p.b_stage_bytes = p.b_boxes * p.b_box_bytes;
    // MN-major, 128B swizzle: SBO = 8 k-rows * 128 B, LBO = bytes between 64-wide MN atoms, k-step = 16 rows
    p.a_lbo = p.a_box_bytes; p.a_sbo = 1024; p.a_kstep_bytes = 16 * 128;
    p.b_lbo = p.b_box_bytes; p.b_sbo = 1024; p.b_kstep_bytes = 16 * 128;
    p.a_mn_major = p.b_mn_major = 1;
    p.idesc = make_idesc_bf16(128, p.BN, 1, 1);
    p.stages = pick_stages(p.a_stage_bytes + p.b_stage_bytes);
    p.out = ws; p.out_dtype = STC_F32; p.Cout = Cout; p.alpha = 1.f;
    return launch(p, st);
}

// Batched GEMM on the tensor cores.  Supported operand layouts (element strides):
//   A: K-major (sAk == 1) or M-major (sAm == 1);  B: K-major (sBk == 1) or N-major (sBn == 1).
// the epilogue modes that live in the staged (TMA store) path need BN % 64 == 0
bool umma_staged_ok(int N) { const int bn = pick_bn(N); return bn != 0 && bn % 64 == 0 && !getenv_off("STC_TMA_STORE"); }

bool gemm_umma_eligible(const stc_gemm_desc* d, int dtype) {
    if (dtype != STC_BF16 || d->beta != 0.f) return false;
    if (!(d->sAk == 1 || d->sAm == 1) || !(d->sBk == 1 || d->sBn == 1)) return false;
    int bn = pick_bn(d->N);
    if (bn == 0) return false;
    if (d->sBn == 1 && d->sBk != 1 && bn % 64 != 0) return false;
    auto ok8 = [](long long s) { return s % 8 == 0; };
    long long lda = d->sAk == 1 ? d->sAm : d->sAk, ldb = d->sBk == 1 ? d->sBn : d->sBk;
    if (!ok8(lda) || !ok8(ldb) || !ok8(d->sA1) || !ok8(d->sA2) || !ok8(d->sB1) || !ok8(d->sB2)) return false;
    if (!ok8(d->sCm) || !ok8(d->sC1) || !ok8(d->sC2)) return false;
    if (d->sAm == 1 && d->sAk != 1 && d->M % 64 != 0) return false;
    return true;
}

// row_mode 2 / 3: the two-sweep row softmax epilogues (UmmaParams::epi_mode); 3 takes P as `mul_residual` (no rowvec)
bool gemm_umma_rowsoftmax_ok(const stc_gemm_desc* d, const void* C, const void* P) {
    if (!gemm_umma_eligible(d, STC_BF16)) return false;
    const int BN = pick_bn(d->N);
    return BN % 64 == 0 && d->M % 128 == 0 && d->N % BN == 0 && ((uintptr_t)C & 15) == 0 && ((uintptr_t)P & 15) == 0 && d->sCm % 8 == 0 &&
           !getenv_off("STC_TMA_STORE");
}

int gemm_umma(const void* A, const void* B, void* C, const stc_gemm_desc* d, int out_dtype, cudaStream_t st, const void* mul_residual,
              const float* rowvec, int row_mode, float sm_scale) {
    STC_REQUIRE(gemm_umma_eligible(d, STC_BF16), "gemm_umma: descriptor not eligible");
    if (row_mode) {
        STC_REQUIRE((row_mode == 2 || row_mode == 3) && out_dtype == STC_BF16 && !rowvec && (row_mode == 3) == (mul_residual != nullptr) &&
                        gemm_umma_rowsoftmax_ok(d, C, mul_residual),
                    "gemm_umma: the two-sweep row softmax needs bf16 output, whole 128-row blocks and N tiles, and P for the backward form");
    } else {
        STC_REQUIRE((mul_residual == nullptr) == (rowvec == nullptr) && (!mul_residual || (out_dtype == STC_BF16 && ((uintptr_t)mul_residual & 15) == 0)),
                    "gemm_umma: the softmax-backward epilogue needs P (bf16, 16-byte aligned, laid out like C) and D together");
    }
    UmmaParams p;
    memset(&p, 0, sizeof(p));
    p.mode = MODE_GEMM;
    p.n_src = 1; p.n_out = 1;
    p.BN = pick_bn(d->N);
    p.a_mn_major = (d->sAk != 1);
    p.b_mn_major = (d->sBk != 1);
    p.cta2 = (!row_mode && d->M % 128 == 0 && use_cta2(d->M / 128, p.BN, p.b_mn_major, (d->K + 63) / 64)) ? 1 : 0;
    const int bn_cta = p.cta2 ? p.BN / 2 : p.BN;          // columns of B staged by one CTA
    uint64_t b1 = (uint64_t)d->batch1, b2 = (uint64_t)d->batch2;
    {
        // inner dim is the unit-stride one
        uint64_t inner = p.a_mn_major ? d->M : d->K, outer = p.a_mn_major ? d->K : d->M;
        uint64_t ld = p.a_mn_major ? d->sAk : d->sAm;
        uint64_t dims[4] = {inner, outer, b2, b1};
        uint64_t str[4] = {2, ld * 2, (uint64_t)(d->sA2 ? d->sA2 : 8) * 2, (uint64_t)(d->sA1 ? d->sA1 : 8) * 2};
        uint32_t box[4] = {64, (uint32_t)(p.a_mn_major ? 64 : 128), 1, 1};
        int rc = encode_map(&p.tmA, A, 4, dims, str, box);
        if (rc) return rc;
    }
    {
        uint64_t inner = p.b_mn_major ? d->N : d->K, outer = p.b_mn_major ? d->K : d->N;
        uint64_t ld = p.b_mn_major ? d->sBk : d->sBn;
        uint64_t dims[4] = {inner, outer, b2, b1};
        uint64_t str[4] = {2, ld * 2, (uint64_t)(d->sB2 ? d->sB2 : 8) * 2, (uint64_t)(d->sB1 ? d->sB1 : 8) * 2};
        uint32_t box[4] = {64, (uint32_t)(p.b_mn_major ? 64 : bn_cta), 1, 1};
        int rc = encode_map(&p.tmB, B, 4, dims, str, box);
        if (rc) return rc;
    }
    p.M = d->M;
    p.batch2 = d->batch2;
    p.num_m_tiles = (d->M + 127) / 128;
    p.num_n_tiles = d->N / p.BN;
    p.num_tiles = (p.cta2 ? p.num_m_tiles / 2 : p.num_m_tiles) * p.num_n_tiles * d->batch1 * d->batch2;
    p.num_k_iters = (d->K + 63) / 64;
    p.ksteps = 4;
    if (!p.a_mn_major) {
        p.a_boxes = 1; p.a_box_bytes = 128 * 128; p.a_lbo = 0; p.a_sbo = 1024; p.a_kstep_bytes = 32;
    } else {
        p.a_boxes = 2; p.a_box_bytes = 64 * 128; p.a_lbo = p.a_box_bytes; p.a_sbo = 1024; p.a_kstep_bytes = 16 * 128;
    }
    p.a_stage_bytes = 128 * 128;
    if (!p.b_mn_major) {
        p.b_boxes = 1; p.b_box_bytes = (uint32_t)bn_cta * 128; p.b_lbo = 0; p.b_sbo = 1024; p.b_kstep_bytes = 32;
    } else {
        p.b_boxes = bn_cta / 64; p.b_box_bytes = 64 * 128; p.b_lbo = p.b_box_bytes; p.b_sbo = 1024; p.b_kstep_bytes = 16 * 128;
    }
    p.b_stage_bytes = (uint32_t)bn_cta * 128;
    p.idesc = make_idesc_bf16(p.cta2 ? 256 : 128, p.BN, p.a_mn_major, p.b_mn_major);
    p.store_tma = (out_dtype == STC_BF16 && p.BN % 64 == 0 && ((uintptr_t)C & 15) == 0 && !getenv_off("STC_TMA_STORE")) ? 1 : 0;
    if (p.store_tma) {
        uint64_t dims[4] = {(uint64_t)d->N, (uint64_t)d->M, b2, b1};
        uint64_t str[4] = {2, (uint64_t)d->sCm * 2, (uint64_t)(d->sC2 ? d->sC2 : 8) * 2, (uint64_t)(d->sC1 ? d->sC1 : 8) * 2};
        uint32_t box[4] = {64, 32, 1, 1};
        int rc = encode_map(&p.tmC, C, 4, dims, str, box);
        if (rc) return rc;
    }
    p.stages = pick_stages(p.a_stage_bytes + p.b_stage_bytes, p.store_tma);
    p.out = C; p.out_dtype = out_dtype; p.Cout = d->N;
    p.ldc = d->sCm; p.sC1 = d->sC1; p.sC2 = d->sC2; p.alpha = d->alpha;
    p.sweeps = 1;
    if (row_mode) {
        STC_REQUIRE(p.store_tma, "gemm_umma: the two-sweep row softmax needs the staged (TMA store) epilogue");
        p.epi_mode = row_mode;
        p.sweeps = 2;
        p.sm_scale = sm_scale;
        p.residual = mul_residual;
    } else if (mul_residual) {
        p.epi_mode = 1;
        p.residual = mul_residual;
        p.rowvec = rowvec;
        p.rv_s1 = (long long)d->batch2 * d->M;
        p.rv_s2 = d->M;
    }
    return launch(p, st);
}

}  // namespace stc
