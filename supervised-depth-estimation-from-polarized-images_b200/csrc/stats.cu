// Per-channel sum and sum of squares of a planar float tensor [B, C, HW], in float64.
//
// Replaces the reductions of polarisation/xolp_mean_and_std_dev.py:26-32 (mean / std of the DoLP and AoLP maps of a
// set of frames -> the constants hard-coded at manydepth/networks/pre_encoders.py:79) and serves as the output
// checksum of the sequence benchmark.  The two sums are additive across frames and ranks
// (one all-reduce of 2 C doubles).
//
// Roofline: HBM, 4 B per element.  The sums follow the library's canonical order (polcue_device.cuh: group -> tile ->
// segment -> total), the same order the statistics by-product of the fused kernel uses, so
// polcue_fused_mosaic_stats_u8 and this function return the same bits for the same data, and every run does.
// One WARP per tile of 256 groups: lane l loads groups l, l + 32, ... (eight 16-byte loads in flight), the eight
// butterfly trees of the tile run in registers, lane 0 writes the tile record; fold_tiles_kernel then adds the
// records in float64 (one CTA per segment of 1024 tiles, the last CTA to finish adds the segments).
#include "polcue_device.cuh"
#include "polcue_host.h"

namespace polcue {
namespace {

constexpr int kStatThreads = 256;
constexpr int kFoldValues = 32;          // values per tile record the fold kernel can carry

struct StatParams {
    const float* x;
    size_t hw;
    int channels;              // C of the tensor
    int c0;                    // first channel of this launch (records hold channels c0 .. c0 + gridDim.y - 1)
    int vec;                   // elements per group: 4 (hw % 4 == 0) or 1
    uint32_t groups;           // B * hw / vec, per channel (< 2^31)
    FastDiv groups_per_image;  // hw / vec
    uint32_t n_tiles;
    int stride;                // floats per tile record
    bool aligned16;
    float* records;            // [n_tiles][stride]: (sum, squares) of channel c at [2 (c - c0)], [2 (c - c0) + 1]
};

__global__ void __launch_bounds__(kStatThreads) channel_tiles_kernel(const StatParams p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tile = blockIdx.x * 8 + warp;
    if (tile >= p.n_tiles) return;
    const int c = p.c0 + blockIdx.y;
    float gs[8], gq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t g = tile * kSumTileGroups + 32 * j + lane;
        gs[j] = gq[j] = 0.0f;
        if (g < p.groups) {
            const uint32_t b = fastdiv(g, p.groups_per_image), r = g - b * p.groups_per_image.div;
            const float* src = p.x + ((size_t)(b * (uint32_t)p.channels + c) * p.hw + (size_t)r * p.vec);
            if (p.vec == 4) {
                float4 v;
                if (p.aligned16) v = ld_stream_f32x4(src);
                else v = make_float4(ld_stream_f32(src), ld_stream_f32(src + 1), ld_stream_f32(src + 2), ld_stream_f32(src + 3));
                gs[j] = group_sum4(v.x, v.y, v.z, v.w);
                gq[j] = group_squares4(v.x, v.y, v.z, v.w);
            } else {
                const float v = ld_stream_f32(src);
                gs[j] = v;
                gq[j] = __fmul_rn(v, v);
            }
        }
    }
    // The eight butterfly trees of the sums and the eight of the squares, all at once: at every level (xor 16, 8, 4, 2, 1)
    // a lane keeps half of its values and hands the other half to its partner, so 16 values x 5 levels cost 16 shuffles
    // instead of 80.  Each value still sees the canonical pairing order (float add commutes), so the bits are the canonical
    // ones.  Afterwards lane l holds the warp-tree sum of (squares if l & 16 else sums) of "warp" j = bits 3, 2, 1 of l.
    float v8[8];
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float keep = hi ? gq[k] : gs[k], send = hi ? gs[k] : gq[k];
            v8[k] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 16));
        }
    }
    float v4[4];
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float keep = hi ? v8[4 + k] : v8[k], send = hi ? v8[k] : v8[4 + k];
            v4[k] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 8));
        }
    }
    float v2[2];
    {
        const bool hi = lane & 4;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float keep = hi ? v4[2 + k] : v4[k], send = hi ? v4[k] : v4[2 + k];
            v2[k] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 4));
        }
    }
    float v1;
    {
        const bool hi = lane & 2;
        const float keep = hi ? v2[1] : v2[0], send = hi ? v2[0] : v2[1];
        v1 = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 2));
    }
    v1 = __fadd_rn(v1, __shfl_xor_sync(0xffffffffu, v1, 1));
    // the eight warp sums of the tile in order 0..7: "warp" j sits in lane 2 j (sums) and 16 + 2 j (squares)
    float ts = __shfl_sync(0xffffffffu, v1, 0), tq = __shfl_sync(0xffffffffu, v1, 16);
#pragma unroll
    for (int j = 1; j < 8; ++j) {
        ts = __fadd_rn(ts, __shfl_sync(0xffffffffu, v1, 2 * j));
        tq = __fadd_rn(tq, __shfl_sync(0xffffffffu, v1, 16 + 2 * j));
    }
    if (lane == 0) {
        float* rec = p.records + (size_t)tile * p.stride + 2 * blockIdx.y;
        rec[0] = ts;
        rec[1] = tq;
    }
}

// records [n_tiles][stride] (float) -> out[n_values] (double), canonical segment / total order.
__global__ void __launch_bounds__(kStatThreads) fold_tiles_kernel(const float* __restrict__ records, uint32_t n_tiles, int stride,
                                                                  int n_values, double* __restrict__ seg_part,
                                                                  unsigned long long* ticket, double* __restrict__ out) {
    __shared__ double scratch[8];
    __shared__ bool last;
    const uint32_t seg = blockIdx.x;
    double acc[kFoldValues];
#pragma unroll
    for (int v = 0; v < kFoldValues; ++v) acc[v] = 0.0;
    for (int m = 0; m < kSumSegmentTiles / kStatThreads; ++m) {
        const uint32_t t = seg * kSumSegmentTiles + m * kStatThreads + threadIdx.x;
        if (t < n_tiles) {
            const float* rec = records + (size_t)t * stride;
#pragma unroll
            for (int v = 0; v < kFoldValues; ++v)
                if (v < n_values) acc[v] += (double)__ldcg(rec + v);
        }
    }
#pragma unroll
    for (int v = 0; v < kFoldValues; ++v) {
        if (v < n_values) {                                  // uniform
            const double s = block_sum_f64(acc[v], scratch);
            if (threadIdx.x == 0) seg_part[(size_t)seg * kFoldValues + v] = s;
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(ticket, 1ull) == (unsigned long long)gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    for (int v = 0; v < n_values; ++v) {
        double a = 0.0;
        for (uint32_t s = threadIdx.x; s < gridDim.x; s += kStatThreads) a += __ldcg(seg_part + (size_t)s * kFoldValues + v);
        const double t = block_sum_f64(a, scratch);
        if (threadIdx.x == 0) out[v] = t;
    }
    if (threadIdx.x == 0) *ticket = 0ull;
}

}  // namespace

size_t fold_workspace_bytes(uint32_t n_tiles, int stride) {
    const size_t segs = ((size_t)n_tiles + kSumSegmentTiles - 1) / kSumSegmentTiles;
    return 64 + segs * kFoldValues * sizeof(double) + (size_t)n_tiles * stride * sizeof(float);
}

float* fold_records(void* workspace, uint32_t n_tiles) {
    const size_t segs = ((size_t)n_tiles + kSumSegmentTiles - 1) / kSumSegmentTiles;
    return reinterpret_cast<float*>(static_cast<char*>(workspace) + 64 + segs * kFoldValues * sizeof(double));
}

int launch_fold_tiles(void* workspace, uint32_t n_tiles, int stride, int n_values, double* out, cudaStream_t stream) {
    if (n_values > kFoldValues || n_values > stride) return POLCUE_EINVAL;
    const uint32_t segs = (n_tiles + kSumSegmentTiles - 1) / kSumSegmentTiles;
    fold_tiles_kernel<<<segs ? segs : 1, kStatThreads, 0, stream>>>(fold_records(workspace, n_tiles), n_tiles, stride, n_values,
                                                                    reinterpret_cast<double*>(static_cast<char*>(workspace) + 64),
                                                                    static_cast<unsigned long long*>(workspace), out);
    return launch_status();
}

}  // namespace polcue

using namespace polcue;

namespace {
constexpr int kChannelsPerPass = 16;     // 32 values per tile record
unsigned long long tiles_of(int B, size_t hw) {
    const int vec = hw % 4 == 0 ? 4 : 1;
    const unsigned long long groups = (unsigned long long)B * (hw / vec);
    return (groups + kSumTileGroups - 1) / kSumTileGroups;
}
}  // namespace

extern "C" {

size_t polcue_channel_stats_workspace_bytes(int B, int C, size_t hw) {
    if (B <= 0 || C <= 0 || hw == 0) return 64;
    const unsigned long long tiles = tiles_of(B, hw);
    if (tiles >= (1ull << 31)) return 64;
    return fold_workspace_bytes((uint32_t)tiles, 2 * (C < kChannelsPerPass ? C : kChannelsPerPass));
}

int polcue_channel_stats_f32(const float* x, int B, int C, size_t hw, void* workspace, double* stats, polcue_stream_t stream) {
    if (!x || !workspace || !stats || B <= 0 || C <= 0 || hw == 0) return POLCUE_EINVAL;
    if ((reinterpret_cast<uintptr_t>(workspace) & 63) || (reinterpret_cast<uintptr_t>(stats) & 7) ||
        (reinterpret_cast<uintptr_t>(x) & 3))
        return POLCUE_EINVAL;
    const unsigned long long tiles = tiles_of(B, hw);
    if (tiles >= (1ull << 31) - 8 || hw >= (1ull << 32)) return POLCUE_E2BIG;
    StatParams p;
    p.x = x;
    p.hw = hw;
    p.channels = C;
    p.vec = hw % 4 == 0 ? 4 : 1;
    if ((unsigned long long)B * (hw / p.vec) >= (1ull << 31)) return POLCUE_E2BIG;
    p.groups = (uint32_t)((unsigned long long)B * (hw / p.vec));
    p.groups_per_image.div = (uint32_t)(hw / p.vec);
    make_fastdiv(p.groups_per_image.div, p.groups_per_image.mul, p.groups_per_image.shift);
    p.n_tiles = (uint32_t)tiles;
    p.aligned16 = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    p.records = fold_records(workspace, p.n_tiles);
    for (int c0 = 0; c0 < C; c0 += kChannelsPerPass) {       // 16 channels per pass share the workspace (stream order)
        const int nc = C - c0 < kChannelsPerPass ? C - c0 : kChannelsPerPass;
        p.c0 = c0;
        p.stride = 2 * (C < kChannelsPerPass ? C : kChannelsPerPass);
        channel_tiles_kernel<<<dim3((p.n_tiles + 7) / 8, (unsigned)nc), kStatThreads, 0, (cudaStream_t)stream>>>(p);
        int rc = launch_status();
        if (rc != POLCUE_OK) return rc;
        rc = launch_fold_tiles(workspace, p.n_tiles, p.stride, 2 * nc, stats + 2 * c0, (cudaStream_t)stream);
        if (rc != POLCUE_OK) return rc;
    }
    return POLCUE_OK;
}

}  // extern "C"
