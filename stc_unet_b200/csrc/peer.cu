// Data-parallel exchanges over NVLink PEER MEMORY (C1/C2 SyncBN statistics, C3 gradient all-reduce) without NCCL in the
// step: every rank maps every other rank's symmetric buffers (torch.distributed._symmetric_memory hands out the peer pointers) and
// the kernels below load / store them directly.  Being plain kernels they can be captured in the whole-step CUDA graph.
//
// Synchronisation = monotonically increasing 64-bit tickets written with st.release.sys into the CONSUMER's flag array and polled
// with ld.acquire.sys.  Every wait is bounded in WALL-CLOCK time (%globaltimer; stc_peer_configure: default 10 min, the order of NCCL's
// watchdog - rank skew from a checkpoint save, an evaluation pass or a loader stall is ordinary).  A wait that does run out does not
// trap (that would destroy the CUDA context of every waiting rank): it raises the error flag the host polls (pinned host memory, so no
// synchronisation is needed to read it), and the kernel returns; the host raises at its next check (PeerExchange.check).
//   small all-reduce (fp64, n <= max_n): one CTA; ticket t uses data slot t & 1 (a rank can be at most one call ahead of a peer, so
//       two slots suffice): publish own values -> ticket to every peer -> wait for every peer's ticket -> sum the W slots in rank order.
//   arena all-reduce (fp32, in place on the symmetric gradient arena): CTA c of every rank works on the c-th sub-range and only
//       synchronises with CTA c of the peers (no grid-wide barrier): barrier, reduce-scatter (rank r sums chunk r of all ranks in rank
//       order and scales), barrier, all-gather (copy chunk p from rank p), barrier (nobody may touch its arena before all peers are
//       done reading).  Sums are formed once, by the chunk's owner, in a fixed order: deterministic and identical on all ranks.
#include "common.cuh"

namespace stc {

struct PeerTable {
    unsigned long long base[STC_PEER_MAX];
    int rank, world;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
struct PeerWait {
    unsigned long long timeout_ns;   // wall-clock bound of one wait
    int* err;                        // host-visible error flag (may be null): set to 1 when a wait ran out
};
// bounded poll; false when the peer did not arrive in time (the flag is raised, the caller leaves the kernel)
__device__ __forceinline__ bool wait_ticket(const unsigned long long* flag, unsigned long long want, const PeerWait& w) {
    if (ld_acquire_sys(flag) >= want) return true;
    const unsigned long long t0 = global_ns();
    unsigned int spins = 0;
    while (ld_acquire_sys(flag) < want) {
        if ((++spins & 1023u) == 0 && global_ns() - t0 > w.timeout_ns) {
            if (w.err) { *reinterpret_cast<volatile int*>(w.err) = 1; __threadfence_system(); }
            return false;
        }
        __nanosleep(64);
    }
    return true;
}

// control block layout (bytes, identical on every rank): [0, 2*max_n*8) two fp64 data slots | then STC_PEER_MAX u64 flags (one per sender)
__global__ void __launch_bounds__(256) peer_allreduce_small_kernel(PeerTable t, int max_n, const double* __restrict__ in, double* __restrict__ out,
                                                                   int n, unsigned long long* __restrict__ seq, PeerWait pw) {
    __shared__ unsigned long long s_ticket;
    __shared__ int s_fail;
    if (threadIdx.x == 0) { s_ticket = ++(*seq); s_fail = 0; }
    __syncthreads();
    const unsigned long long ticket = s_ticket;
    const size_t slot = (size_t)(ticket & 1) * max_n;
    double* mine = reinterpret_cast<double*>(t.base[t.rank]) + slot;
    for (int i = threadIdx.x; i < n; i += blockDim.x) mine[i] = in[i];
    __threadfence_system();
    __syncthreads();
    const size_t flag_off = (size_t)2 * max_n * sizeof(double);
    if ((int)threadIdx.x < t.world && (int)threadIdx.x != t.rank) {
        unsigned long long* peer_flags = reinterpret_cast<unsigned long long*>(t.base[threadIdx.x] + flag_off);
        st_release_sys(peer_flags + t.rank, ticket);
        const unsigned long long* my_flags = reinterpret_cast<const unsigned long long*>(t.base[t.rank] + flag_off);
        if (!wait_ticket(my_flags + threadIdx.x, ticket, pw)) s_fail = 1;
    }
    __syncthreads();
    if (s_fail) return;   // the error flag is up: the host raises; `out` keeps the local values
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < t.world; ++r) s += __ldcv(reinterpret_cast<const double*>(t.base[r]) + slot + i);
        out[i] = s;
    }
}

// CTA-local cross-rank barrier: CTA c of this rank <-> CTA c of every peer.  flags: [ctas][STC_PEER_MAX] u64 in the control block
__device__ __forceinline__ bool cta_peer_barrier(const PeerTable& ctl, size_t flag_off, unsigned long long ticket, const PeerWait& pw, int* s_fail) {
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < ctl.world && (int)threadIdx.x != ctl.rank) {
        unsigned long long* peer_flags = reinterpret_cast<unsigned long long*>(ctl.base[threadIdx.x] + flag_off) + (size_t)blockIdx.x * STC_PEER_MAX;
        st_release_sys(peer_flags + ctl.rank, ticket);
        const unsigned long long* my_flags =
            reinterpret_cast<const unsigned long long*>(ctl.base[ctl.rank] + flag_off) + (size_t)blockIdx.x * STC_PEER_MAX;
        if (!wait_ticket(my_flags + threadIdx.x, ticket, pw)) *s_fail = 1;
    }
    __syncthreads();
    return *s_fail == 0;
}

// arena: symmetric fp32 buffers (one per rank); [elem_off, elem_off + n) is reduced in place on every rank; n % (4 * world * gridDim.x) == 0
// is NOT required: the range is cut into float4 units, the tail (< 4 floats) is handled by CTA 0 / rank chunk 0 as scalars.
__global__ void __launch_bounds__(256) peer_allreduce_arena_kernel(PeerTable arena, PeerTable ctl, size_t flag_off, long long elem_off, long long n,
                                                                   float scale, unsigned long long* __restrict__ seq, PeerWait pw) {
    __shared__ unsigned long long s_ticket;
    __shared__ int s_fail;
    if (threadIdx.x == 0) { s_ticket = (seq[blockIdx.x] += 3); s_fail = 0; }
    __syncthreads();
    const unsigned long long ticket = s_ticket;   // this call uses tickets ticket-2, ticket-1, ticket
    const int W = arena.world, R = arena.rank;
    // float4 units of the range, split first over CTAs, then over ranks
    const long long nv = n >> 2;
    const long long per_cta = (nv + gridDim.x - 1) / gridDim.x;
    const long long c0 = min(nv, (long long)blockIdx.x * per_cta), c1 = min(nv, c0 + per_cta);
    const long long per_rank = (c1 - c0 + W - 1) / W;
    float4* mine = reinterpret_cast<float4*>(reinterpret_cast<float*>(arena.base[R]) + elem_off);

    if (!cta_peer_barrier(ctl, flag_off, ticket - 2, pw, &s_fail)) return;   // every rank's gradients of this range are complete
    {   // reduce-scatter: my chunk
        const long long a = min(c1, c0 + R * per_rank), b = min(c1, a + per_rank);
        // NVLink loads have ~1-2 us of latency: kU independent 16-byte loads per rank are issued before any of them is consumed
        constexpr int kU = 4;
        long long i = a + threadIdx.x;
        for (; i + (kU - 1) * (long long)blockDim.x < b; i += kU * (long long)blockDim.x) {
            float4 s[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) s[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < W; ++r) {
                const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(arena.base[r]) + elem_off) + i;
                float4 v[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) v[u] = __ldcv(src + u * (long long)blockDim.x);
#pragma unroll
                for (int u = 0; u < kU; ++u) { s[u].x += v[u].x; s[u].y += v[u].y; s[u].z += v[u].z; s[u].w += v[u].w; }
            }
#pragma unroll
            for (int u = 0; u < kU; ++u)
                mine[i + u * (long long)blockDim.x] = make_float4(s[u].x * scale, s[u].y * scale, s[u].z * scale, s[u].w * scale);
        }
        for (; i < b; i += blockDim.x) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < W; ++r) {
                const float4 v = __ldcv(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(arena.base[r]) + elem_off) + i);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            mine[i] = make_float4(s.x * scale, s.y * scale, s.z * scale, s.w * scale);
        }
        if (blockIdx.x == 0 && R == 0) {   // scalar tail of the range
            for (long long i = (nv << 2) + threadIdx.x; i < n; i += blockDim.x) {
                float s = 0.f;
                for (int r = 0; r < W; ++r) s += __ldcv(reinterpret_cast<const float*>(arena.base[r]) + elem_off + i);
                reinterpret_cast<float*>(mine)[i] = s * scale;
            }
        }
    }
    if (!cta_peer_barrier(ctl, flag_off, ticket - 1, pw, &s_fail)) return;   // all chunks reduced
    for (int k = 1; k < W; ++k) {                         // all-gather: chunk p from rank p, starting with the next rank
        const int p = (R + k) % W;
        const long long a = min(c1, c0 + p * per_rank), b = min(c1, a + per_rank);
        const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(arena.base[p]) + elem_off);
        constexpr int kG = 8;
        long long i = a + threadIdx.x;
        for (; i + (kG - 1) * (long long)blockDim.x < b; i += kG * (long long)blockDim.x) {
            float4 v[kG];
#pragma unroll
            for (int u = 0; u < kG; ++u) v[u] = __ldcv(src + i + u * (long long)blockDim.x);
#pragma unroll
            for (int u = 0; u < kG; ++u) mine[i + u * (long long)blockDim.x] = v[u];
        }
        for (; i < b; i += blockDim.x) mine[i] = __ldcv(src + i);
    }
    if (blockIdx.x == 0 && R != 0) {
        const float* src = reinterpret_cast<const float*>(arena.base[0]) + elem_off;
        for (long long i = (nv << 2) + threadIdx.x; i < n; i += blockDim.x) reinterpret_cast<float*>(mine)[i] = __ldcv(src + i);
    }
    cta_peer_barrier(ctl, flag_off, ticket, pw, &s_fail);   // every peer has taken its copy: the arena may be overwritten again
}

}  // namespace stc

using namespace stc;

static int fill_table(PeerTable& t, const unsigned long long* ptrs, int rank, int world, const char* what) {
    STC_REQUIRE(ptrs && world >= 1 && world <= STC_PEER_MAX && rank >= 0 && rank < world, "%s: bad peer table (rank %d of %d)", what, rank, world);
    memset(&t, 0, sizeof(t));
    for (int i = 0; i < world; ++i) {
        STC_REQUIRE(ptrs[i] != 0 && (ptrs[i] & 15) == 0, "%s: peer pointer %d is null or unaligned", what, i);
        t.base[i] = ptrs[i];
    }
    t.rank = rank; t.world = world;
    return STC_OK;
}

static PeerWait g_wait = {600ull * 1000000000ull, nullptr};

extern "C" int stc_peer_configure(long long timeout_ms, int* err_flag) {
    STC_REQUIRE(timeout_ms > 0, "stc_peer_configure: timeout must be positive (got %lld ms)", timeout_ms);
    g_wait.timeout_ns = (unsigned long long)timeout_ms * 1000000ull;
    g_wait.err = err_flag;
    return STC_OK;
}

extern "C" long long stc_peer_ctrl_bytes(int max_n, int ctas) {
    return (long long)2 * max_n * (long long)sizeof(double) + (long long)STC_PEER_MAX * 8 + (long long)ctas * STC_PEER_MAX * 8;
}

extern "C" int stc_peer_allreduce_small_f64(const unsigned long long* ctrl_ptrs, int rank, int world, int max_n, const double* in, double* out,
                                            int n, unsigned long long* seq, void* stream) {
    STC_REQUIRE(n >= 0 && n <= max_n && in && out && seq, "peer_allreduce_small: n=%d exceeds the slot size %d (or null pointer)", n, max_n);
    PeerTable t;
    int rc = fill_table(t, ctrl_ptrs, rank, world, "peer_allreduce_small");
    if (rc) return rc;
    if (n == 0) return STC_OK;
    peer_allreduce_small_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(t, max_n, in, out, n, seq, g_wait);
    return check_launch("peer_allreduce_small");
}

extern "C" int stc_peer_allreduce_arena_f32(const unsigned long long* arena_ptrs, const unsigned long long* ctrl_ptrs, int rank, int world, int max_n,
                                            long long elem_off, long long n, float scale, unsigned long long* seq, int ctas, void* stream) {
    STC_REQUIRE(n >= 0 && elem_off >= 0 && elem_off % 4 == 0 && ctas >= 1 && ctas <= 1024 && seq, "peer_allreduce_arena: bad range / grid");
    PeerTable a, c;
    int rc = fill_table(a, arena_ptrs, rank, world, "peer_allreduce_arena");
    if (rc) return rc;
    rc = fill_table(c, ctrl_ptrs, rank, world, "peer_allreduce_arena");
    if (rc) return rc;
    if (n == 0) return STC_OK;
    const size_t flag_off = (size_t)2 * max_n * sizeof(double) + (size_t)STC_PEER_MAX * 8;
    peer_allreduce_arena_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(a, c, flag_off, elem_off, n, scale, seq, g_wait);
    return check_launch("peer_allreduce_arena");
}
