// Per-channel sum and sum of squares of a planar float tensor [B, C, HW], in float64.
//
// Replaces the reductions of polarisation/xolp_mean_and_std_dev.py:26-32 (mean / std of the DoLP and AoLP maps of a
// set of frames -> the constants hard-coded at manydepth/networks/pre_encoders.py:79) and serves as the output
// checksum of the sequence benchmark.  The two sums are additive across frames and ranks
// (one all-reduce of 2 C doubles).
//
// Roofline: HBM, 4 B per element.  One CTA per 64 KB slab of one (b, c) plane: float4 streaming loads, float32
// partial sums over 64 elements per thread, float64 from there on; warp shuffles -> shared -> one partial per CTA;
// the last CTA to finish folds the partials of every channel in a fixed order, so results are bitwise reproducible.
#include "polcue_device.cuh"
#include "polcue_host.h"

namespace polcue {
namespace {

constexpr int kStatThreads = 256, kStatVecPerThread = 16;               // 16 float4 per thread
constexpr int kStatSlab = kStatThreads * kStatVecPerThread * 4;         // 16384 floats = 64 KB per CTA

struct StatParams {
    const float* x;
    size_t hw;
    int planes;          // B * C
    int channels;        // C
    unsigned chunks;     // slabs per plane
    bool vec4;
    unsigned long long* ticket;
    double* partials;    // [planes * chunks][2]
    double* stats;       // [C][2]
};

__device__ __forceinline__ double block_sum(double v, double* scratch) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();                      // scratch may still be read from a previous call
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kStatThreads / 32; ++w) t += scratch[w];
    return t;                             // valid in thread 0
}

__global__ void __launch_bounds__(kStatThreads) channel_stats_kernel(const StatParams p) {
    __shared__ double scratch[kStatThreads / 32];
    __shared__ bool last;
    const int plane = blockIdx.y;
    const float* src = p.x + (size_t)plane * p.hw;
    const size_t lo = (size_t)blockIdx.x * kStatSlab;
    const size_t hi = lo + kStatSlab < p.hw ? lo + kStatSlab : p.hw;
    double s = 0.0, q = 0.0;
    if (p.vec4) {
        float fs = 0.0f, fq = 0.0f;
#pragma unroll 4
        for (int k = 0; k < kStatVecPerThread; ++k) {
            const size_t i = lo + ((size_t)k * kStatThreads + threadIdx.x) * 4;
            if (i < hi) {                       // hw % 4 == 0: a vector never straddles the end
                const float4 v = ld_stream_f32x4(src + i);
                fs += (v.x + v.y) + (v.z + v.w);
                fq = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, fq))));
            }
        }
        s = fs;
        q = fq;
    } else {
        for (size_t i = lo + threadIdx.x; i < hi; i += kStatThreads) {
            const float v = ld_stream_f32(src + i);
            s += v;
            q += (double)v * v;
        }
    }
    const double ts = block_sum(s, scratch);
    const double tq = block_sum(q, scratch);
    if (threadIdx.x == 0) {
        const size_t slot = ((size_t)plane * p.chunks + blockIdx.x) * 2;
        p.partials[slot] = ts;
        p.partials[slot + 1] = tq;
        __threadfence();
        last = atomicAdd(p.ticket, 1ull) == (unsigned long long)gridDim.x * gridDim.y - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    // fixed-order fold: channel c gathers planes c, c + C, c + 2C, ... and all their slabs
    const int batches = p.planes / p.channels;
    for (int c = 0; c < p.channels; ++c) {
        double fs = 0.0, fq = 0.0;
        const size_t per_channel = (size_t)batches * p.chunks;
        for (size_t j = threadIdx.x; j < per_channel; j += kStatThreads) {
            const size_t b = j / p.chunks, ch = j - b * p.chunks;
            const size_t slot = ((b * p.channels + c) * p.chunks + ch) * 2;
            fs += __ldcg(p.partials + slot);
            fq += __ldcg(p.partials + slot + 1);
        }
        const double a = block_sum(fs, scratch);
        const double b2 = block_sum(fq, scratch);
        if (threadIdx.x == 0) {
            p.stats[2 * c] = a;
            p.stats[2 * c + 1] = b2;
        }
    }
    if (threadIdx.x == 0) *p.ticket = 0ull;
}

}  // namespace
}  // namespace polcue

using namespace polcue;

extern "C" {

size_t polcue_channel_stats_workspace_bytes(int B, int C, size_t hw) {
    if (B <= 0 || C <= 0) return 64;
    const size_t chunks = (hw + kStatSlab - 1) / kStatSlab;
    return 64 + (size_t)B * C * chunks * 2 * sizeof(double);
}

int polcue_channel_stats_f32(const float* x, int B, int C, size_t hw, void* workspace, double* stats, polcue_stream_t stream) {
    if (!x || !workspace || !stats || B <= 0 || C <= 0 || hw == 0) return POLCUE_EINVAL;
    if ((reinterpret_cast<uintptr_t>(workspace) & 63) || (reinterpret_cast<uintptr_t>(stats) & 7) ||
        (reinterpret_cast<uintptr_t>(x) & 3))
        return POLCUE_EINVAL;
    const size_t chunks = (hw + kStatSlab - 1) / kStatSlab;
    if ((size_t)B * C > 65535 || chunks >= (1ull << 31)) return POLCUE_E2BIG;
    StatParams p;
    p.x = x;
    p.hw = hw;
    p.planes = B * C;
    p.channels = C;
    p.chunks = (unsigned)chunks;
    p.vec4 = hw % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    p.ticket = static_cast<unsigned long long*>(workspace);
    p.partials = reinterpret_cast<double*>(static_cast<char*>(workspace) + 64);
    p.stats = stats;
    channel_stats_kernel<<<dim3((unsigned)chunks, (unsigned)(B * C)), kStatThreads, 0, (cudaStream_t)stream>>>(p);
    return launch_status();
}

}  // extern "C"
