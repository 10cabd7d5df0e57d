// One-node all-reduce(sum) of a handful of float64 accumulators over NVLink / NVSwitch PEER MEMORY, done inside the kernel
// that produces them (the only collective of the path is the sum of {n_images, per-group metric sums} -- at
// most 1 + 16 x 7 doubles -- which a library all-reduce turns into a launch + a ~40 us protocol round trip, longer than
// the 15-image evaluation shard it follows).
//
// Every rank owns one PeerBlock in its own HBM and maps the blocks of all ranks (CUDA IPC).  A call, one CTA per rank:
//   1. seq = ++block.seq (device-side counter, so CUDA-graph replays advance it); parity = seq & 1;
//   2. thread k stores its value into slots[parity][my rank][k] of EVERY rank's block (8-byte stores through NVLink),
//      fences at system scope, and after a CTA barrier one thread per destination releases flags[parity][my rank] = seq;
//   3. thread r acquires its OWN block's flags[parity][r] >= seq (bounded spin), CTA barrier;
//   4. thread k adds slots[parity][0..world-1][k] of its own block IN RANK ORDER: every rank forms the identical sum,
//      bitwise reproducible (a ring / tree all-reduce does not promise that).
// Two parities are enough: a rank can pass the wait of call s + 1 only after every peer has entered call s + 1, i.e. has
// finished reading the slots of call s (calls on one rank are stream-ordered), so the writes of call s + 2 cannot overtake them.
// A peer that never arrives ends the spin after ~4 s: the results become NaN and block.failed_call records the call.
#pragma once
#include <cstdint>

namespace polcue {

constexpr int kPeerMaxRanks = 16;
constexpr int kPeerMaxValues = 128;
constexpr int kPeerThreads = 128;            // threads that take part in an exchange (>= kPeerMaxValues, >= kPeerMaxRanks)
constexpr long long kPeerSpinCycles = 8000000000ll;   // ~4 s at 1.9 GHz

struct PeerBlock {
    unsigned long long seq;                                  // exchanges started by the owning rank
    unsigned long long failed_call;                          // first call whose wait timed out (0 = none)
    unsigned long long pad[14];
    unsigned long long flags[2][kPeerMaxRanks];              // flags[parity][src] = src's call whose values are complete
    double slots[2][kPeerMaxRanks][kPeerMaxValues];
};

struct PeerParams {
    PeerBlock* block[kPeerMaxRanks];   // block[r] = rank r's block as mapped into this process (block[rank] is local)
    int world, rank;
};

__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_f64(double* p, double v) {
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// Called by ALL threads of a CTA with blockDim.x >= kPeerThreads (one CTA per rank).  Thread k < n contributes `v` and
// receives the sum over ranks of the k-th values; the other threads receive 0.
__device__ __forceinline__ double peer_allreduce_cta(const PeerParams& pp, double v, int n) {
    __shared__ unsigned long long s_seq;
    __shared__ int s_failed;
    PeerBlock* me = pp.block[pp.rank];
    const int k = threadIdx.x;
    if (k == 0) {
        s_seq = me->seq + 1;
        me->seq = s_seq;
        s_failed = 0;
    }
    __syncthreads();
    const unsigned long long seq = s_seq;
    const int par = (int)(seq & 1ull);
    if (k < n) {
        for (int r = 0; r < pp.world; ++r) st_relaxed_sys_f64(&pp.block[r]->slots[par][pp.rank][k], v);
        __threadfence_system();
    }
    __syncthreads();
    if (k < pp.world) {
        st_release_sys_u64(&pp.block[k]->flags[par][pp.rank], seq);
        const long long t0 = clock64();
        while (ld_acquire_sys_u64(&me->flags[par][k]) < seq) {
            if (clock64() - t0 > kPeerSpinCycles) {
                s_failed = 1;
                break;
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    double total = 0.0;
    if (k < n) {
        for (int r = 0; r < pp.world; ++r) total += ld_relaxed_sys_f64(&me->slots[par][r][k]);
        if (s_failed) total = __longlong_as_double(0x7ff8000000000000ll);
    }
    if (k == 0 && s_failed && me->failed_call == 0) me->failed_call = seq;
    return total;
}

}  // namespace polcue
