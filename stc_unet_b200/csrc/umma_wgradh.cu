// tcgen05 weight-gradient kernel with shared-memory HALO reuse.
//
//   dW[(r,s)][ci][co] += sum_{n,h,w} x[n, h+r-pr, w+s-ps, ci] * dy[n, h, w, co]
//
// GEMM view per output row h and 128-pixel block: D[(tap, ci)][co] += A[(tap, ci)][px] * B[px][co], K = 128 pixels.
// Both operands are MN-major in smem exactly as TMA delivers NHWC rows ([pixel][64 channels], 128B swizzle):
//   B = dy tile  : BN/64 atoms of [128 px][64 co]               (LBO = 16 KB between co atoms)
//   A = x segment: one input row, [(128 + S - 1) px][64 ci].  The M = 128 rows of one MMA are TWO TAPS (r, s) and
//       (r, s+1) of the same 64-channel chunk: the second 64-row atom is the same segment shifted by one pixel, i.e.
//       LBO = 128 bytes (overlapping atoms), and tap s starts s * 128 bytes into the segment.  A segment therefore feeds
//       all S taps of its filter row, and — rolling down the image — the R-1 following output rows as well.
//   S is odd: the taps (r, 0..S-2) of a filter row pair up as above (S/2 MMAs per row); the LAST COLUMN s = S-1 pairs ACROSS filter rows:
//   taps (r, S-1) and (r+1, S-1) read the segments of two consecutive input rows at the same pixel shift, i.e. LBO = one ring slot
//   (17408 B).  The ring keeps a mirror copy of slot 0 behind its last slot so that "the next slot" is always 17408 B further on.
//   Only an odd row count leaves one tap paired with a dummy (its 64 rows are dropped by the epilogue): 5 MMAs instead of 6 per pixel
//   block for 3x3, 13 instead of 15 for 5x5, 25 instead of 28 for 7x7 (Cout = 64).
// Accumulators (all taps of the CTA's filter-row group x 64 ci x BN co, fp32) stay in TMEM over the CTA's whole pixel
// range and are red.add-ed into the [tap][ci][co] workspace once at the end.
// A work item = (filter-row group, ci chunk, co tile, image n, column block, row range).
// Traffic per output row and CTA: one new x segment (17 KB) + one dy tile (BN*256 B) for rg*ceil(S/2)*8 MMAs.
// Warps: 0 = x-segment producer, 1 = MMA issuer + TMEM owner, 2..5 = epilogue, 6 = dy producer.
//
// CTA2 variant (cta_group::2): the two CTAs of a cluster work on the SAME (co tile, pixels) with one M = 256 MMA per tap pair, rows 0-127
// from the leader's x segments, 128-255 from the peer's.  What differs between the two is either the 64-channel ci chunk (pair_mode 0:
// two consecutive chunks, Cin % 128 == 0) or the filter-row group (pair_mode 1, Cin = 64 layers: the peer's ring holds the input rows RG
// further down, so the same descriptors address the next group's taps; a group that sticks out of the filter is computed and dropped).
// Each CTA stages only HALF of the dy tile's channels (dy traffic per pair halves): one 64-channel SW128 atom for BN = 128, a 32-channel
// tile in the 64-byte swizzle for BN = 64 (MN-major SW64: 64 B per pixel row, SBO = 512 B).  Barrier protocol as in umma_convh.cu.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace stc {

struct alignas(64) WgradHParams {
    CUtensorMap tmX;   // x  {Cin,  W, H, N}, box {64, 128+S-1, 1, 1}
    CUtensorMap tmDY;  // dy {Cout, W, H, N}, box {64, 128, 1, 1}
    int n_src;                       // virtual channel concat of the input: number of sources (descriptors at the end of the struct)
    int H, W, R, S, SPf, cin_chunks, BN, num_n_tiles;   // SPf = S / 2 full tap pairs per filter row
    int RG, num_groups;          // filter rows per group, number of groups
    int blocks_w;
    // Work = (filter-row group g, base item b = (image, column block, ci chunk, co tile), output row h).  The CTAs are divided among the
    // groups in proportion to a group's cost per row (its MMA count + a fixed per-row share), and the CTAs of one group split the
    // (base item, row) axis into equal contiguous slices.  Every CTA therefore finishes at the same time (round-robin over ~2 unequal
    // items per SM had left the SMs idle 22-27 % of the launch), AND the CTAs of different groups sweep the (base, row) axis in
    // step - the same x / dy rows at about the same time - so the re-reads of the later groups hit L2 instead of DRAM (laying the groups
    // out one after the other on a single cost axis read x and dy once per group from DRAM: 4.2 GB for a 1.07 GB request).
    long long rows_total;        // base items * H
    int grp_cta_begin[9];        // CTAs [grp_cta_begin[g], grp_cta_begin[g+1]) work on group g
    int a_slots, b_stages;
    uint32_t a_slot_bytes, a_box_bytes, b_stage_bytes;
    uint32_t idesc;
    float* ws;
    int Cin, Cout;
    int cta2;                    // CTA pairs
    int pair_mode;               // 0: the pair splits two ci chunks (`cin_chunks` counts PAIRS of chunks); 1: it splits two filter-row groups (`num_groups` counts PAIRS)
    int b_kstep16;               // descriptor advance of the dy tile per 16-pixel K step, 16 B units (128: SW128 atom, 64: SW64 half tile)
    // cold tail (virtual channel concat of the input, common.cuh ChanCat): sources 1..n_src-1
    int src_chunk_end[kMaxCat];
    CUtensorMap tmX2[kMaxCat - 1];
};

constexpr int kWgradHThreads = 224;

struct WItem {
    int g, cc, nt, n_img, w0, h_a, h_b, r0, rg;
};
struct WSched {
    const WgradHParams& p;
    long long pos, end;
    int g;
    int rank;
    __device__ __forceinline__ WSched(const WgradHParams& p_, int cta, int rank_ = 0) : p(p_), rank(rank_) {
        g = 0;
        while (g + 1 < p.num_groups && cta >= p.grp_cta_begin[g + 1]) ++g;
        const int j = cta - p.grp_cta_begin[g], n = p.grp_cta_begin[g + 1] - p.grp_cta_begin[g];
        pos = p.rows_total * j / n;
        end = p.rows_total * (j + 1) / n;
    }
    // next piece (a row range of one base item) of this CTA's slice; false when the slice is exhausted
    __device__ __forceinline__ bool next(WItem& it) {
        if (pos >= end) return false;
        const long long b = pos / p.H;
        const int h_a = (int)(pos - b * p.H);
        const long long left = end - pos;
        const int h_b = left < (long long)(p.H - h_a) ? h_a + (int)left : p.H;
        pos += h_b - h_a;
        long long idx = b;
        it.nt = (int)(idx % p.num_n_tiles); idx /= p.num_n_tiles;
        it.cc = (int)(idx % p.cin_chunks); idx /= p.cin_chunks;
        if (p.cta2 && p.pair_mode == 0) it.cc = 2 * it.cc + rank;
        it.w0 = (int)(idx % p.blocks_w) * 128;
        it.n_img = (int)(idx / p.blocks_w);
        it.g = g;
        it.h_a = h_a;
        it.h_b = h_b;
        if (p.cta2 && p.pair_mode == 1) {   // both CTAs run the leader group's row count; rows >= R are dropped by the epilogue
            it.rg = min(p.RG, p.R - 2 * g * p.RG);
            it.r0 = (2 * g + rank) * p.RG;
        } else {
            it.r0 = g * p.RG;
            it.rg = min(p.RG, p.R - it.r0);
        }
        return true;
    }
};

template <bool CTA2>
__global__ void __launch_bounds__(kWgradHThreads, 1) umma_wgradh_kernel(const __grid_constant__ WgradHParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t a_bytes = (uint32_t)(p.a_slots + 1) * p.a_slot_bytes;   // ring + mirror of slot 0
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a_bytes + (size_t)p.b_stages * p.b_stage_bytes);
    // bars: a_full[a_slots], a_empty[a_slots], b_full[b_stages], b_empty[b_stages], acc_full, acc_empty
    const int nb = 2 * p.a_slots + 2 * p.b_stages + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + nb);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = CTA2 ? (int)ptx::cluster_ctarank() : 0;          // 0 = leader of the pair
    const int unit = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;  // scheduling unit: CTA, or CTA pair
    const uint32_t smem_base = ptx::smem_u32(smem);
    const uint32_t b_base = smem_base + a_bytes;
    const uint32_t bar_base = ptx::smem_u32(bars);
    auto a_full = [&](int s) { return bar_base + 8u * s; };
    auto a_empty = [&](int s) { return bar_base + 8u * (p.a_slots + s); };
    auto b_full = [&](int s) { return bar_base + 8u * (2 * p.a_slots + s); };
    auto b_empty = [&](int s) { return bar_base + 8u * (2 * p.a_slots + p.b_stages + s); };
    const uint32_t acc_full = bar_base + 8u * (2 * p.a_slots + 2 * p.b_stages);
    const uint32_t acc_empty = acc_full + 8u;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&p.tmX);
        ptx::prefetch_tensormap(&p.tmDY);
        for (int j = 1; j < p.n_src; ++j) ptx::prefetch_tensormap(&p.tmX2[j - 1]);
        for (int s = 0; s < p.a_slots; ++s) { ptx::mbar_init(a_full(s), 1); ptx::mbar_init(a_empty(s), 1); }
        for (int s = 0; s < p.b_stages; ++s) { ptx::mbar_init(b_full(s), 1); ptx::mbar_init(b_empty(s), 1); }
        ptx::mbar_init(acc_full, 1);
        ptx::mbar_init(acc_empty, CTA2 ? 8 : 4);
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
    const int pr = p.R / 2, ps = p.S / 2;

    if (warp == 0) {
        // ===================== x-segment producer: input rows h_a + r0 - pr ... h_b - 1 + r0 + rg - 1 - pr =====================
        int slot = 0;
        uint32_t phase = 0;
        WSched sched(p, unit, rank);
        WItem it;
        while (sched.next(it)) {
            const int first = it.h_a + it.r0 - pr, count = (it.h_b - it.h_a) + it.rg - 1;
            int lc = it.cc;
            const CUtensorMap* mx = nullptr;      // non-null: a source of a virtual concat other than the first
            if (p.n_src > 1) {
                int j = 0;
                while (j + 1 < p.n_src && it.cc >= p.src_chunk_end[j]) ++j;
                lc = it.cc - (j ? p.src_chunk_end[j - 1] : 0);
                if (j) mx = &p.tmX2[j - 1];
            }
            for (int e = 0; e < count; ++e) {
                ptx::mbar_wait(a_empty(slot), phase ^ 1);
                if (ptx::elect_one_sync()) {
                    // slot 0 is loaded twice: in place and into the mirror behind the last ring slot (cross-row pairs starting in the last slot)
                    const uint32_t bytes = slot == 0 ? 2 * p.a_box_bytes : p.a_box_bytes;
                    const CUtensorMap* m = mx ? mx : &p.tmX;
                    if (CTA2) {   // both CTAs' bytes complete on the leader's barrier
                        const uint32_t fb = ptx::mapa_shared(a_full(slot), 0);
                        if (rank == 0) ptx::mbar_arrive_expect_tx(a_full(slot), 2 * bytes);
                        ptx::tma_load_4d_2sm(smem_base + slot * p.a_slot_bytes, m, fb, lc * 64, it.w0 - ps, first + e, it.n_img);
                        if (slot == 0) ptx::tma_load_4d_2sm(smem_base + p.a_slots * p.a_slot_bytes, m, fb, lc * 64, it.w0 - ps, first + e, it.n_img);
                    } else {
                        ptx::mbar_arrive_expect_tx(a_full(slot), bytes);
                        ptx::tma_load_4d(smem_base + slot * p.a_slot_bytes, m, a_full(slot), lc * 64, it.w0 - ps, first + e, it.n_img);
                        if (slot == 0) ptx::tma_load_4d(smem_base + p.a_slots * p.a_slot_bytes, m, a_full(slot), lc * 64, it.w0 - ps, first + e, it.n_img);
                    }
                }
                __syncwarp();
                if (++slot == p.a_slots) { slot = 0; phase ^= 1; }
            }
        }
    } else if (warp == 6) {
        // ===================== dy producer: one [128 px][BN] tile per output row =====================
        int stage = 0;
        uint32_t phase = 0;
        const int nbox = p.BN / 64;
        WSched sched(p, unit, rank);
        WItem it;
        while (sched.next(it)) {
            for (int h = it.h_a; h < it.h_b; ++h) {
                ptx::mbar_wait(b_empty(stage), phase ^ 1);
                if (ptx::elect_one_sync()) {
                    if (CTA2) {   // this CTA stages half `rank` of the tile's channels (b_stage_bytes = that half)
                        if (rank == 0) ptx::mbar_arrive_expect_tx(b_full(stage), 2 * p.b_stage_bytes);
                        ptx::tma_load_4d_2sm(b_base + stage * p.b_stage_bytes, &p.tmDY, ptx::mapa_shared(b_full(stage), 0), it.nt * p.BN + rank * (p.BN / 2), it.w0, h, it.n_img);
                    } else {
                        ptx::mbar_arrive_expect_tx(b_full(stage), p.b_stage_bytes);
                        for (int j = 0; j < nbox; ++j)
                            ptx::tma_load_4d(b_base + stage * p.b_stage_bytes + j * 16384, &p.tmDY, b_full(stage), it.nt * p.BN + j * 64, it.w0, h, it.n_img);
                    }
                }
                __syncwarp();
                if (++stage == p.b_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform; elected lane issues; CTA2: the leader only) =====================
        if (!CTA2 || rank == 0) {
        int a_head = 0;          // ring slot of the oldest live segment (x row h + r0 - pr)
        uint32_t a_phase = 0;
        int bstage = 0;
        uint32_t bphase = 0;
        uint32_t acc_phase = 0;
        // MN-major SW128 descriptors (16 B units): SBO = 1024 B between 8-pixel groups, version 1, layout 2
        const uint64_t desc_common = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        const uint64_t a_hi = desc_common | ((uint64_t)(128 >> 4) << 16);     // LBO = 128 B: second atom = next tap (one pixel on)
        const uint64_t a_hi_row = desc_common | ((uint64_t)(p.a_slot_bytes >> 4) << 16);   // LBO = one slot: second atom = same tap, next input row
        // dy tile: LBO = 16 KB between 64-wide co atoms; the 32-channel half tile (b_kstep16 == 64) is MN-major SW64: SBO = 512 B, layout 4
        const uint64_t b_hi = p.b_kstep16 == 64 ? (((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61))
                                                : (desc_common | ((uint64_t)(16384 >> 4) << 16));
        const uint64_t b_ks = (uint64_t)p.b_kstep16;
        const uint32_t a_slot16 = p.a_slot_bytes >> 4, b_stage16 = p.b_stage_bytes >> 4;
        const uint32_t a_base16 = (smem_base >> 4) & 0x3FFF, b_base16 = (b_base >> 4) & 0x3FFF;
        auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t flag) {
            if (CTA2) ptx::mma_bf16_ss2(d, a, b, p.idesc, flag); else ptx::mma_bf16_ss(d, a, b, p.idesc, flag);
        };
        // arrives once the MMAs issued so far have retired (CTA2: on the barrier at this offset in BOTH CTAs)
        auto commit = [&](uint32_t bar) { if (CTA2) ptx::tc_commit2(bar, 3); else ptx::tc_commit(bar); };
        WSched sched(p, unit, rank);
        WItem it;
        while (sched.next(it)) {
            ptx::mbar_wait(acc_empty, acc_phase ^ 1);
            ptx::tc_fence_after();
            int ready = 0;  // segments of this item observed full (relative to a_head at item start -> tracked via `seen`)
            int seen_slot = a_head;
            uint32_t seen_phase = a_phase;
            for (int h = it.h_a; h < it.h_b; ++h) {
                // output row h reads segments (h - h_a) .. (h - h_a) + rg - 1
                const int need = (h - it.h_a) + it.rg;
                while (ready < need) {
                    ptx::mbar_wait(a_full(seen_slot), seen_phase);
                    if (++seen_slot == p.a_slots) { seen_slot = 0; seen_phase ^= 1; }
                    ++ready;
                }
                ptx::mbar_wait(b_full(bstage), bphase);
                ptx::tc_fence_after();
                const uint64_t b_desc0 = b_hi | (uint64_t)(b_base16 + bstage * b_stage16);
                const uint32_t acc_flag = (h > it.h_a) ? 1u : 0u;
                if (ptx::elect_one_sync()) {
                    for (int rr = 0; rr < it.rg; ++rr) {
                        int sl = a_head + rr;
                        if (sl >= p.a_slots) sl -= p.a_slots;
                        const uint32_t seg16 = a_base16 + sl * a_slot16;
                        for (int sp = 0; sp < p.SPf; ++sp) {
                            const uint64_t a_desc0 = a_hi | (uint64_t)(seg16 + sp * 16);   // tap s = 2*sp starts 2*sp*128 B in
                            const uint32_t d_addr = (uint32_t)((rr * p.SPf + sp) * p.BN);
                            mma(d_addr, a_desc0, b_desc0, acc_flag);
#pragma unroll
                            for (int ks = 1; ks < 8; ++ks)   // K step = 16 pixels = 2048 B = 128 units
                                mma(d_addr, a_desc0 + (uint64_t)(ks * 128), b_desc0 + ks * b_ks, 1u);
                        }
                    }
                    // last filter column: rows (2j, 2j+1) share one MMA (second atom = the next ring slot, or the mirror of slot 0 behind
                    // the last one); a leftover row pairs with the dummy tap S of its own segment
                    for (int j = 0; 2 * j < it.rg; ++j) {
                        int sl = a_head + 2 * j;
                        if (sl >= p.a_slots) sl -= p.a_slots;
                        const uint64_t hi = (2 * j + 1 < it.rg) ? a_hi_row : a_hi;
                        const uint64_t a_desc0 = hi | (uint64_t)(a_base16 + sl * a_slot16 + (p.S - 1) * 8);   // tap S-1 starts (S-1)*128 B in
                        const uint32_t d_addr = (uint32_t)((it.rg * p.SPf + j) * p.BN);
                        mma(d_addr, a_desc0, b_desc0, acc_flag);
#pragma unroll
                        for (int ks = 1; ks < 8; ++ks)
                            mma(d_addr, a_desc0 + (uint64_t)(ks * 128), b_desc0 + ks * b_ks, 1u);
                    }
                    commit(b_empty(bstage));
                    commit(a_empty(a_head));   // x row h + r0 - pr is not read by later output rows
                }
                __syncwarp();
                if (++bstage == p.b_stages) { bstage = 0; bphase ^= 1; }
                if (++a_head == p.a_slots) { a_head = 0; a_phase ^= 1; }
            }
            // the last rg-1 segments of the item are still held: release them and move the ring head past them
            if (ptx::elect_one_sync()) {
                for (int k = 0; k < it.rg - 1; ++k) {
                    int sl = a_head + k;
                    if (sl >= p.a_slots) sl -= p.a_slots;
                    commit(a_empty(sl));
                }
                commit(acc_full);
            }
            __syncwarp();
            a_head += it.rg - 1;
            if (a_head >= p.a_slots) { a_head -= p.a_slots; a_phase ^= 1; }
            acc_phase ^= 1;
        }
        }
    } else if (warp >= 2 && warp <= 5) {
        // ===================== epilogue: TMEM -> red.add into ws[tap][ci][co] =====================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        uint32_t acc_phase = 0;
        WSched sched(p, unit, rank);
        WItem it;
        while (sched.next(it)) {
            ptx::mbar_wait(acc_full, acc_phase);
            ptx::tc_fence_after();
            const int n_full = it.rg * p.SPf, n_slots = n_full + (it.rg + 1) / 2;
            {
                for (int qs = 0; qs < n_slots; ++qs) {
                    int rr, s;
                    if (qs < n_full) { rr = qs / p.SPf; s = 2 * (qs - rr * p.SPf) + (row >> 6); }      // taps (rr, 2sp) | (rr, 2sp+1)
                    else { rr = 2 * (qs - n_full) + (row >> 6); s = p.S - 1; }                         // taps (2j, S-1) | (2j+1, S-1)
                    const bool valid = rr < it.rg && it.r0 + rr < p.R;
                    const int tap = (it.r0 + (valid ? rr : 0)) * p.S + s;
                    float* o = p.ws + ((long long)tap * p.Cin + it.cc * 64 + (row & 63)) * p.Cout + it.nt * p.BN;
                    const uint32_t t_addr = (uint32_t)(qs * p.BN) + ((uint32_t)(q * 32) << 16);
                    for (int c = 0; c < p.BN; c += 32) {
                        uint32_t v[32];
                        ptx::tmem_ld_32x32(t_addr + c, v);
                        ptx::tmem_ld_wait();
                        if (!valid) continue;
#pragma unroll
                        for (int j = 0; j < 32; j += 4)   // 16-byte vector reductions: 4x fewer L2 atomic instructions
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + c + j), "f"(__uint_as_float(v[j])),
                                         "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                                         : "memory");
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CTA2) ptx::mbar_arrive_cluster(ptx::mapa_shared(acc_empty, 0)); else ptx::mbar_arrive(acc_empty);
            }
            acc_phase ^= 1;
        }
    }

    ptx::tc_fence_before();
    if (CTA2) ptx::cluster_sync_all(); else __syncthreads();   // CTA2: the peer's shared memory / barriers stay alive until both are done
    if (warp == 1) {
        ptx::tc_fence_after();
        if (CTA2) ptx::tmem_dealloc2(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512);
    }
}

int encode_map_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box);
int encode_map_bf16_sw64(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box);

bool conv_wgradh_eligible(int W, int Cin, int Cout, int R, int S, int dtype) {
    static int disabled = -1;
    if (disabled < 0) { const char* e = getenv("STC_WGRADH"); disabled = (e && e[0] == '0') ? 1 : 0; }
    if (disabled) return false;
    return dtype == STC_BF16 && W >= 128 && Cin % 64 == 0 && Cout % 64 == 0 && R == S && (R == 3 || R == 5 || R == 7);
}

int check_cat(const ChanCat* c, int total, const char* what);

int conv_wgrad_wgradh(const void* x, const void* dy, float* ws, int N, int H, int W, int Cin, int Cout, int R, int S, cudaStream_t st,
                      const ChanCat* src) {
    WgradHParams p;
    memset(&p, 0, sizeof(p));
    if (src && src->n == 1) { x = src->ptr[0]; src = nullptr; }
    if (int rc = check_cat(src, Cin, "conv_wgrad_wgradh input")) return rc;
    // BN = 128 when possible (measured 5-40 % faster than 64 on the Cout >= 128 layers: fewer, longer MMAs per dy tile)
    { const char* e = getenv("STC_WGRADH_BN"); p.BN = (Cout % 128 == 0 && !(e && atoi(e) == 64)) ? 128 : 64; }
    p.SPf = S / 2;
    p.cin_chunks = Cin / 64;
    {   // CTA pairs (STC_WGRADH_CTA2=0 disables): over two ci chunks when there is an even number of them, else over two filter-row groups
        static int en = -1;
        if (en < 0) { const char* e = getenv("STC_WGRADH_CTA2"); en = (e && e[0] == '0') ? 0 : 1; }
        p.cta2 = en ? 1 : 0;
        p.pair_mode = (p.cin_chunks % 2 == 0) ? 0 : 1;
        if (p.cta2 && p.pair_mode == 1 && R == 1) p.cta2 = 0;   // one filter row: nothing to pair
    }
    // rows per group: a group of rg rows needs rg * (S/2) + ceil(rg / 2) accumulators of BN columns in the 512 TMEM columns.  Single CTAs and
    // ci-chunk pairs take as many rows as fit; group pairs take the row count with the fewest MMAs per pixel block over all pairs.
    p.RG = 0;
    int best_cost = 1 << 30;
    for (int rg = R; rg >= 1; --rg) {
        const int nacc = rg * p.SPf + (rg + 1) / 2;
        if (nacc * p.BN > 512) continue;
        if (!(p.cta2 && p.pair_mode == 1)) { p.RG = rg; break; }
        const int ng = (R + rg - 1) / rg, cost = ((ng + 1) / 2) * nacc;
        if (cost < best_cost) { best_cost = cost; p.RG = rg; }
    }
    STC_REQUIRE(p.RG >= 1, "conv_wgrad_wgradh: no plan");
    p.num_groups = (R + p.RG - 1) / p.RG;
    if (p.cta2 && p.pair_mode == 1) p.num_groups = (p.num_groups + 1) / 2;   // pairs of groups
    if (p.cta2 && p.pair_mode == 0) p.cin_chunks /= 2;                        // pairs of chunks
    p.b_kstep16 = (p.cta2 && p.BN == 64) ? 64 : 128;
    const int bwh = 128 + S - 1;
    p.n_src = src ? src->n : 1;
    for (int j = 0, acc = 0; j < p.n_src; ++j) {
        const int cj = src ? src->c[j] : Cin;
        uint64_t dims[4] = {(uint64_t)cj, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[4] = {2, (uint64_t)cj * 2, (uint64_t)W * cj * 2, (uint64_t)H * W * cj * 2};
        uint32_t box[4] = {64, (uint32_t)bwh, 1, 1};
        int rc = encode_map_bf16(j ? &p.tmX2[j - 1] : &p.tmX, src ? src->ptr[j] : x, 4, dims, str, box);
        if (rc) return rc;
        acc += cj / 64;
        p.src_chunk_end[j] = acc;
    }
    {
        uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t str[4] = {2, (uint64_t)Cout * 2, (uint64_t)W * Cout * 2, (uint64_t)H * W * Cout * 2};
        uint32_t box[4] = {64, 128, 1, 1};
        int rc;
        if (p.b_kstep16 == 64) { box[0] = 32; rc = encode_map_bf16_sw64(&p.tmDY, dy, 4, dims, str, box); }   // 32-channel half tiles, 64-byte swizzle
        else rc = encode_map_bf16(&p.tmDY, dy, 4, dims, str, box);
        if (rc) return rc;
    }
    p.H = H; p.W = W; p.R = R; p.S = S; p.Cin = Cin; p.Cout = Cout;
    p.num_n_tiles = Cout / p.BN;
    p.blocks_w = (W + 127) / 128;
    STC_REQUIRE(p.num_groups <= 8, "conv_wgrad_wgradh: too many filter-row groups");
    p.rows_total = (long long)p.num_n_tiles * p.cin_chunks * p.blocks_w * N * H;
    p.a_slot_bytes = 17408;
    p.a_box_bytes = (uint32_t)bwh * 128;
    p.b_stage_bytes = (uint32_t)(p.cta2 ? p.BN / 2 : p.BN) * 256;
    p.a_slots = p.RG + 3;
    p.b_stages = (p.BN == 64 || p.cta2) ? 4 : 3;
    p.idesc = make_idesc_bf16(p.cta2 ? 256 : 128, p.BN, 1, 1);
    p.ws = ws;
    size_t smem = (size_t)(p.a_slots + 1) * p.a_slot_bytes + (size_t)p.b_stages * p.b_stage_bytes + (2 * p.a_slots + 2 * p.b_stages + 2) * 8 + 16 + 1024;
    STC_REQUIRE(smem <= 227 * 1024, "conv_wgrad_wgradh: smem %zu", smem);
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        STC_CUDA(cudaFuncSetAttribute(umma_wgradh_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        STC_CUDA(cudaFuncSetAttribute(umma_wgradh_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[dev] = true;
    }
    static int row_overhead = -1;   // fixed per-row share (TMA issue, barrier round trips) in MMA units; STC_WGRADH_ROWCOST
    if (row_overhead < 0) { const char* e = getenv("STC_WGRADH_ROWCOST"); row_overhead = e ? atoi(e) : 16; }
    const int units = p.cta2 ? max_cta_pairs((const void*)umma_wgradh_kernel<true>, kWgradHThreads, smem) : num_sms();   // scheduling units: CTAs, or CTA pairs
    int grid = p.rows_total * p.num_groups < units ? (int)(p.rows_total * p.num_groups) : units;
    {   // CTAs per group in proportion to the group's cost per row; every group gets at least one, the counts sum to the grid
        int wgt[8], wsum = 0;
        for (int g = 0; g < p.num_groups; ++g) {
            const int gl = (p.cta2 && p.pair_mode == 1) ? 2 * g : g;   // the (leader) group whose row count the unit runs
            const int rg = (R - gl * p.RG) < p.RG ? (R - gl * p.RG) : p.RG;
            wgt[g] = (rg * p.SPf + (rg + 1) / 2) * 8 + row_overhead;
            wsum += wgt[g];
        }
        if (grid < p.num_groups) grid = p.num_groups;
        int cnt[8], used = 0;
        for (int g = 0; g < p.num_groups; ++g) {
            cnt[g] = (int)((long long)grid * wgt[g] / wsum);
            if (cnt[g] < 1) cnt[g] = 1;
            used += cnt[g];
        }
        for (int g = 0; used < grid; g = (g + 1) % p.num_groups) { ++cnt[g]; ++used; }          // leftovers: one each, heaviest groups first
        for (int g = p.num_groups - 1; used > grid; g = (g + p.num_groups - 1) % p.num_groups)
            if (cnt[g] > 1) { --cnt[g]; --used; }
        p.grp_cta_begin[0] = 0;
        for (int g = 0; g < p.num_groups; ++g) p.grp_cta_begin[g + 1] = p.grp_cta_begin[g] + cnt[g];
    }
    if (p.cta2) {   // `grid` counted pairs: clusters of two CTAs (the two SMs of a TPC)
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)(2 * grid)); cfg.blockDim = dim3(kWgradHThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        STC_CUDA(cudaLaunchKernelEx(&cfg, umma_wgradh_kernel<true>, p));
        return check_launch("umma_wgradh_kernel (cta_group::2)");
    }
    umma_wgradh_kernel<false><<<grid, kWgradHThreads, smem, st>>>(p);
    return check_launch("umma_wgradh_kernel");
}

}  // namespace stc
