// One-node all-reduce(sum) of a handful of float64 accumulators over NVLink / NVSwitch PEER MEMORY, done inside the kernel
// that produces them (the only collective of the path is the sum of {n_images, per-group metric sums} -- at
// most 1 + 16 x 7 doubles -- which a library all-reduce turns into a launch + a protocol round trip that is longer than
// the 15-image evaluation shard it follows).
//
// Every rank owns one PeerBlock in its own HBM and maps the blocks of all ranks (CUDA IPC).  A call, one CTA per rank:
//   1. seq = ++block.seq (device-side counter, so CUDA-graph replays advance it); parity = seq & 1;
//   2. thread k splits its value into two 32-bit halves and stores each half TOGETHER WITH the 32-bit call number as one
//      64-bit scalar store into words[parity][my rank][k] of EVERY rank's block: a 64-bit store is single-copy atomic, so a
//      word whose tag equals the call number carries valid data -- no fence, no separate flag, one NVLink one-way trip;
//   3. thread k polls words[parity][r][k] of its OWN block for r = 0 .. world-1 until both tags match (bounded spin) and
//      adds the values IN RANK ORDER: every rank forms the identical sum, bitwise reproducible (a ring or tree
//      all-reduce does not promise that).
// Two parities are enough: a rank completes call s + 1 only after every peer's values of call s + 1 arrived, which a peer
// sends after it finished reading call s (the calls of one rank are stream-ordered), so the stores of call s + 2 cannot
// overwrite words a peer still has to read.  A peer that never arrives ends the spin after ~4 s: the results become NaN
// and block.failed_call records the call.
#pragma once
#include <cstdint>

namespace polcue {

constexpr int kPeerMaxRanks = 16;
constexpr int kPeerMaxValues = 128;
constexpr int kPeerThreads = 128;            // threads that take part in an exchange (>= kPeerMaxValues)
constexpr long long kPeerSpinCycles = 8000000000ll;   // ~4 s at 1.9 GHz

struct PeerBlock {
    unsigned long long seq;                                  // exchanges started by the owning rank
    unsigned long long failed_call;                          // first call whose wait timed out (0 = none)
    unsigned long long pad[14];
    unsigned long long words[2][kPeerMaxRanks][kPeerMaxValues][2];   // [parity][source rank][value]{low half, high half}: call number << 32 | 32 data bits
};

struct PeerParams {
    PeerBlock* block[kPeerMaxRanks];   // block[r] = rank r's block as mapped into this process (block[rank] is local)
    int world, rank;
};

__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ ulonglong2 ld_sys_u64x2(const unsigned long long* p) {      // both halves of one value (16-byte aligned)
    ulonglong2 v;
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
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
    const unsigned long long tag = seq << 32;          // never 0 within 2^32 calls of a zero-initialised block
    double total = 0.0;
    if (k < n) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
        for (int r = 0; r < pp.world; ++r) {
            unsigned long long* w = pp.block[r]->words[par][pp.rank][k];
            st_sys_u64(w, tag | (bits & 0xffffffffull));
            st_sys_u64(w + 1, tag | (bits >> 32));
        }
        const long long t0 = clock64();
        for (int r = 0; r < pp.world; ++r) {
            const unsigned long long* w = me->words[par][r][k];
            ulonglong2 got = ld_sys_u64x2(w);
            while (((got.x ^ tag) | (got.y ^ tag)) >> 32) {
                if (clock64() - t0 > kPeerSpinCycles) {
                    s_failed = 1;
                    break;
                }
                got = ld_sys_u64x2(w);
            }
            total += __longlong_as_double((long long)((got.y << 32) | (got.x & 0xffffffffull)));
        }
    }
    __syncthreads();
    if (k < n && s_failed) total = __longlong_as_double(0x7ff8000000000000ll);
    if (k == 0 && s_failed && me->failed_call == 0) me->failed_call = seq;
    return total;
}

}  // namespace polcue
