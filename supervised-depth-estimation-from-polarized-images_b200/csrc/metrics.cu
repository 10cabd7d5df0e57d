// Depth error metrics: one pass, eight additive accumulators.
//
// Replaces manydepth/layers.py:539-557 (compute_depth_errors) / :559-577 (compute_depth_errors_numpy),
// which make seven full passes with six temporaries:
//     thresh = max(gt/pred, pred/gt);  a_k = mean(thresh < 1.25^k)
//     rmse = sqrt(mean((gt-pred)^2));  rmse_log = sqrt(mean((log gt - log pred)^2))
//     abs_rel = mean(|gt-pred|/gt);    sq_rel = mean((gt-pred)^2/gt)
// and the per-image masked loop of Trainer.compute_depth_losses_from_list (manydepth/trainer.py:1376-1428).
//
// Accumulators: count, n[t<1.25], n[t<1.25^2], n[t<1.25^3], S d^2, S dlog^2, S |d|/gt, S d^2/gt.
// The threshold ratio is ONE IEEE-rounded division max(gt,pred)/min(gt,pred): for positive inputs it equals
// the reference's max(fl(gt/pred), fl(pred/gt)) bit for bit, so the three counts are exact integers.
// Per-thread partial sums are float32 over at most 64 elements, then carried in float64; the cross-thread
// reduction is warp shuffles -> shared memory -> (flat) a fixed-order pass by the last CTA, or
// (per image) a thread-block cluster reducing through distributed shared memory.  Both are deterministic.
// Roofline: HBM, 8 B per element (+1 B with an instance-id filter).
#include <cooperative_groups.h>

#include "polcue_device.cuh"
#include "polcue_host.h"

namespace cg = cooperative_groups;

namespace polcue {
namespace {

constexpr int kMetricThreads = 256;
constexpr int kMaxMetricBlocks = 148 * 16;
constexpr int kCluster = 8;  // CTAs per image in the per-image kernel (portable cluster size)

struct Acc {
    float f[4];  // d^2, (log2 hi - log2 lo)^2, |d|/gt, d^2/gt
    int n, c1, c2, c3;   // kept elements; raw threshold counts (kept + rejected, see acc_add)
    int seen;            // elements passed to acc_add since the last flush
};

__device__ __forceinline__ void acc_clear(Acc& a) {
    a.f[0] = a.f[1] = a.f[2] = a.f[3] = 0.0f;
    a.n = a.c1 = a.c2 = a.c3 = a.seen = 0;
}

// fl(hi / lo) < c, decided WITHOUT the division, exactly as IEEE round-to-nearest-even would decide it.
// For c in [1, 2) with an even significand (1.25, 1.5625 and 1.953125 all are), fl(x) < c  <=>  x < c - 2^-24
// (the midpoint below c rounds up to c).  So the test is hi - c lo < -2^-24 lo; near the threshold hi - c lo is
// a small multiple of ulp(lo)/8 and the FMA returns it exactly, far from it the sign is unambiguous.
__device__ __forceinline__ bool ratio_below(float hi, float lo, float c, float neg_half_ulp_lo) {
    return fmaf(-c, lo, hi) < neg_half_ulp_lo;
}

// One element.  The masked kernels stay branch-free: a rejected element is replaced by the pair (1, 1), which adds
// exactly zero to the four float sums (d = 0, log ratio = 0) and exactly one to each threshold count; `flush`
// removes those spurious counts again (they equal the number of rejected elements).
__device__ __forceinline__ void acc_add(Acc& a, float gt, float pred) {
    const float hi = fmaxf(gt, pred), lo = fminf(gt, pred);
    const float thr = -5.9604644775390625e-08f * lo;     // -2^-24 lo
    // == (max(gt/pred, pred/gt) < 1.25 ** k) of layers.py:542-545, bit for bit, for positive inputs
    a.c1 += ratio_below(hi, lo, 1.25f, thr);
    a.c2 += ratio_below(hi, lo, 1.5625f, thr);
    a.c3 += ratio_below(hi, lo, 1.953125f, thr);
    const float d = gt - pred, d2 = d * d;
    const float inv_gt = rcp_approx(gt);
    const float dl = lg2_approx(hi) - lg2_approx(lo);    // |log gt - log pred| / ln 2; ln^2 2 is applied at the flush
    a.f[0] += d2;
    a.f[1] = fmaf(dl, dl, a.f[1]);
    a.f[2] = fmaf(fabsf(d), inv_gt, a.f[2]);
    a.f[3] = fmaf(d2, inv_gt, a.f[3]);
}

struct Acc64 {
    double v[8];
};

__device__ __forceinline__ void flush(Acc64& s, Acc& a) {
    const int rejected = a.seen - a.n;
    s.v[0] += a.n;
    s.v[1] += a.c1 - rejected;
    s.v[2] += a.c2 - rejected;
    s.v[3] += a.c3 - rejected;
    s.v[4] += a.f[0];
    s.v[5] += 0.4804530139182014 * (double)a.f[1];   // ln(2)^2
    s.v[6] += a.f[2];
    s.v[7] += a.f[3];
    acc_clear(a);
}

// Block reduction: shuffles inside a warp, then warp 0 folds the per-warp rows in a fixed order.
__device__ __forceinline__ void block_reduce(Acc64& s, double (*warp_rows)[8]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        double v = s.v[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        s.v[k] = v;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < 8; ++k) warp_rows[warp][k] = s.v[k];
    __syncthreads();
    if (threadIdx.x < 8) {
        double v = 0.0;
        for (int w = 0; w < kMetricThreads / 32; ++w) v += warp_rows[w][threadIdx.x];
        warp_rows[0][threadIdx.x] = v;  // thread k owns accumulator k; row 0 now holds the block total
    }
    __syncthreads();
}

// sums -> the seven metrics in the reference's return order (layers.py:557).  Empty -> NaN like mean([]).
__device__ __forceinline__ void finalize(const double* s, float* m) {
    const double n = s[0];
    m[0] = (float)(s[6] / n);
    m[1] = (float)(s[7] / n);
    m[2] = (float)sqrt(s[4] / n);
    m[3] = (float)sqrt(s[5] / n);
    m[4] = (float)(s[1] / n);
    m[5] = (float)(s[2] / n);
    m[6] = (float)(s[3] / n);
}

// ------------------------------------------------------------------------------------------
// flat: already masked / compacted arrays (the signature of compute_depth_errors)
// workspace: [0] uint64 ticket (zero between launches), [64...] partials[grid][8] doubles
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMetricThreads) depth_errors_kernel(const float* __restrict__ gt, const float* __restrict__ pred,
                                                                      size_t count, bool vec4, unsigned long long* ticket,
                                                                      double* partials, double* sums8, float* metrics7) {
    __shared__ double warp_rows[kMetricThreads / 32][8];
    __shared__ bool last;
    Acc a;
    acc_clear(a);
    Acc64 s{};
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    if (vec4) {
        const size_t n4 = count >> 2;
        int since = 0;
        for (size_t i = tid; i < n4; i += stride) {
            const float4 g = ld_stream_f32x4(gt + 4 * i), q = ld_stream_f32x4(pred + 4 * i);
            acc_add(a, g.x, q.x);
            acc_add(a, g.y, q.y);
            acc_add(a, g.z, q.z);
            acc_add(a, g.w, q.w);
            a.n += 4;
            a.seen += 4;
            if (++since == 16) {
                flush(s, a);
                since = 0;
            }
        }
        for (size_t i = (n4 << 2) + tid; i < count; i += stride) {
            acc_add(a, gt[i], pred[i]);
            a.n += 1;
            a.seen += 1;
        }
    } else {
        int since = 0;
        for (size_t i = tid; i < count; i += stride) {
            acc_add(a, ld_stream_f32(gt + i), ld_stream_f32(pred + i));
            a.n += 1;
            a.seen += 1;
            if (++since == 64) {
                flush(s, a);
                since = 0;
            }
        }
    }
    flush(s, a);
    block_reduce(s, warp_rows);
    if (threadIdx.x < 8) partials[(size_t)blockIdx.x * 8 + threadIdx.x] = warp_rows[0][threadIdx.x];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1ull) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    {   // fixed summation order over CTAs (bitwise reproducible): 32 lanes x 8 accumulators, then lanes in order
        __shared__ double lane_rows[32][8];
        const unsigned k = threadIdx.x & 7, lane = threadIdx.x >> 3;
        double v = 0.0;
        for (unsigned blk = lane; blk < gridDim.x; blk += 32) v += __ldcg(partials + (size_t)blk * 8 + k);
        lane_rows[lane][k] = v;
        __syncthreads();
        if (threadIdx.x < 8) {
            double t = 0.0;
            for (int l = 0; l < 32; ++l) t += lane_rows[l][threadIdx.x];
            warp_rows[0][threadIdx.x] = t;
            sums8[threadIdx.x] = t;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (metrics7) finalize(warp_rows[0], metrics7);
        *ticket = 0ull;  // ready for the next launch on this workspace
    }
}

// ------------------------------------------------------------------------------------------
// per image, masked: one cluster of kCluster CTAs per image, DSMEM reduction into CTA rank 0
// ------------------------------------------------------------------------------------------
struct ImageParams {
    const float* gt;
    const float* pred;
    const uint8_t* inst;  // may be null
    size_t px;
    float min_d, max_d;
    int inst_id;
    double* sums;
    float* metrics;  // may be null
    bool vec4;
};

__device__ __forceinline__ void acc_masked(Acc& a, const ImageParams& p, float g, float q, int id) {
    // trainer.py:1380 mask, :1410-1411 material filter, :1417-1418 clamp of the prediction
    const bool keep = (g > p.min_d) & (g < p.max_d) & ((p.inst == nullptr) | (id == p.inst_id));
    acc_add(a, keep ? g : 1.0f, keep ? fminf(fmaxf(q, p.min_d), p.max_d) : 1.0f);   // rejected -> the neutral pair (1, 1)
    a.n += keep;
    a.seen += 1;
}

__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kMetricThreads)
    depth_errors_images_kernel(const ImageParams p) {
    __shared__ double warp_rows[kMetricThreads / 32][8];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    const size_t b = blockIdx.y;
    const float* gt = p.gt + b * p.px;
    const float* pred = p.pred + b * p.px;
    const uint8_t* inst = p.inst ? p.inst + b * p.px : nullptr;

    Acc a;
    acc_clear(a);
    Acc64 s{};
    const size_t tid = (size_t)rank * kMetricThreads + threadIdx.x, stride = (size_t)kCluster * kMetricThreads;
    int since = 0;
    if (p.vec4) {
        const size_t n4 = p.px >> 2;
        for (size_t i = tid; i < n4; i += stride) {
            const float4 g = ld_stream_f32x4(gt + 4 * i), q = ld_stream_f32x4(pred + 4 * i);
            uint32_t ids = 0;
            if (inst) ids = ld_stream_u32(inst + 4 * i);
            acc_masked(a, p, g.x, q.x, ids & 0xff);
            acc_masked(a, p, g.y, q.y, (ids >> 8) & 0xff);
            acc_masked(a, p, g.z, q.z, (ids >> 16) & 0xff);
            acc_masked(a, p, g.w, q.w, ids >> 24);
            if (++since == 16) {
                flush(s, a);
                since = 0;
            }
        }
    } else {
        for (size_t i = tid; i < p.px; i += stride) {
            acc_masked(a, p, gt[i], pred[i], inst ? inst[i] : 0);
            if (++since == 64) {
                flush(s, a);
                since = 0;
            }
        }
    }
    flush(s, a);
    block_reduce(s, warp_rows);
    cluster.sync();  // every CTA's row 0 is final and visible cluster-wide
    if (rank == 0 && threadIdx.x < 8) {
        double v = 0.0;
        for (unsigned r = 0; r < kCluster; ++r) {
            const double* remote = cluster.map_shared_rank(&warp_rows[0][0], r);
            v += remote[threadIdx.x];
        }
        p.sums[b * 8 + threadIdx.x] = v;
        warp_rows[1][threadIdx.x] = v;
    }
    cluster.sync();  // keep remote shared memory alive until rank 0 has read it
    if (rank == 0 && threadIdx.x == 0 && p.metrics) finalize(warp_rows[1], p.metrics + b * 7);
}

}  // namespace
}  // namespace polcue

using namespace polcue;

extern "C" {

size_t polcue_depth_errors_workspace_bytes(void) { return 64 + (size_t)kMaxMetricBlocks * 8 * sizeof(double); }

int polcue_depth_errors_f32(const float* gt, const float* pred, size_t count, void* workspace, double* sums8, float* metrics7,
                            polcue_stream_t stream) {
    if ((!gt || !pred) && count) return POLCUE_EINVAL;
    if (!workspace || !sums8) return POLCUE_EINVAL;
    if ((reinterpret_cast<uintptr_t>(workspace) & 63) || (reinterpret_cast<uintptr_t>(sums8) & 7)) return POLCUE_EINVAL;
    const bool vec4 = ((reinterpret_cast<uintptr_t>(gt) | reinterpret_cast<uintptr_t>(pred)) & 15) == 0;
    const size_t items = vec4 ? (count + 3) / 4 : count;
    size_t blocks = (items + kMetricThreads - 1) / kMetricThreads;
    const size_t cap = (size_t)device_info().sms * 16;   // two waves of 8 resident CTAs: evens out per-SM bandwidth
    blocks = blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
    if (blocks > (size_t)kMaxMetricBlocks) blocks = kMaxMetricBlocks;
    auto* ticket = static_cast<unsigned long long*>(workspace);
    auto* partials = reinterpret_cast<double*>(static_cast<char*>(workspace) + 64);
    depth_errors_kernel<<<(unsigned)blocks, kMetricThreads, 0, (cudaStream_t)stream>>>(gt, pred, count, vec4, ticket, partials,
                                                                                      sums8, metrics7);
    return launch_status();
}

int polcue_depth_errors_images_f32(const float* gt, const float* pred, const uint8_t* inst, int B, size_t px, float min_d,
                                   float max_d, int inst_id, double* sums, float* metrics, polcue_stream_t stream) {
    if (!gt || !pred || !sums || B < 0 || B > 65535) return POLCUE_EINVAL;
    if (reinterpret_cast<uintptr_t>(sums) & 7) return POLCUE_EINVAL;
    if (B == 0) return POLCUE_OK;
    ImageParams p;
    p.gt = gt;
    p.pred = pred;
    p.inst = inst;
    p.px = px;
    p.min_d = min_d;
    p.max_d = max_d;
    p.inst_id = inst_id;
    p.sums = sums;
    p.metrics = metrics;
    p.vec4 = (px % 4 == 0) && (((reinterpret_cast<uintptr_t>(gt) | reinterpret_cast<uintptr_t>(pred)) & 15) == 0) &&
             ((reinterpret_cast<uintptr_t>(inst) & 3) == 0);
    depth_errors_images_kernel<<<dim3(kCluster, B, 1), kMetricThreads, 0, (cudaStream_t)stream>>>(p);
    return launch_status();
}

}  // extern "C"
