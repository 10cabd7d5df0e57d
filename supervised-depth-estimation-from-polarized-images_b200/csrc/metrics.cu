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

#include <mutex>

#include "polcue_device.cuh"
#include "peer.cuh"
#include "polcue_host.h"

namespace cg = cooperative_groups;

namespace polcue {
namespace {

constexpr int kMetricThreads = 256;
constexpr int kMaxMetricBlocks = 148 * 16;
constexpr int kMaxCluster = 8;  // most CTAs per image (portable cluster size); every launch picks its size, see pick_cluster

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
    const float dl = lg2_approx(pred * inv_gt);          // (log pred - log gt) / ln 2 (only its square is used); ln^2 2 is
                                                         // applied at the flush.  One MUFU less than lg2(hi) - lg2(lo); the
                                                         // neutral pair gives lg2(1 * rcp(1)) = 0 exactly
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
        unsigned blk = lane;
        for (; blk + 7 * 32 < gridDim.x; blk += 8 * 32) {      // eight loads in flight, added in the same fixed order
            double t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = __ldcg(partials + (size_t)(blk + 32 * j) * 8 + k);
#pragma unroll
            for (int j = 0; j < 8; ++j) v += t[j];
        }
        for (; blk < gridDim.x; blk += 32) v += __ldcg(partials + (size_t)blk * 8 + k);
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
// per image, masked: one cluster of 1..8 CTAs per image (pick_cluster), DSMEM reduction into CTA rank 0
// ------------------------------------------------------------------------------------------
struct ImageParams {
    const float* gt;
    const float* pred;
    const uint8_t* inst;  // may be null
    const float* pred_scale;   // may be null: per-image factor applied to pred before the clamp (median scaling, trainer.py:1413-1414)
    int want_hi;               // material filter is group_ids[0] <= id <= want_hi (evaluation.py:259-262: "objects" = 20..160)
    int clamp_first;           // clamp pred to [min_d, max_d] BEFORE the scaling as well (the batch-level clamp of trainer.py:1368-1370)
    size_t px;
    float min_d, max_d;
    int n_groups;         // mask groups evaluated in this launch (grid.y); outputs are [B][n_groups][...]
    int group_ids[16];    // instance id of each group, or -1 for "all" (range mask only)
    double* sums;
    float* metrics;  // may be null
    bool vec4;
};

// EXT = false: the plain loop (one material id, no scaling): this kernel is bound by instruction issue, so the extras
// of the full evaluation loops (id range, batch-level clamp, median scaling) are compiled in only when asked for.
template <bool EXT, bool INST = true>   // INST = false: no material filter at all (object == "all")
__device__ __forceinline__ void acc_masked(Acc& a, const ImageParams& p, int want_id, float g, float q, int id, float scale) {
    // trainer.py:1380 mask, :1410-1411 material filter, :1413-1414 median scaling, :1417-1418 clamp of the prediction
    bool keep = (g > p.min_d) & (g < p.max_d);
    if constexpr (EXT) {
        keep &= (want_id < 0) | ((id >= want_id) & (id <= p.want_hi));
        if (p.clamp_first) q = fminf(fmaxf(q, p.min_d), p.max_d);
        q = __fmul_rn(q, scale);
    } else if constexpr (INST) {
        keep &= (want_id < 0) | (id == want_id);
    }
    acc_add(a, keep ? g : 1.0f, keep ? fminf(fmaxf(q, p.min_d), p.max_d) : 1.0f);   // rejected -> the neutral pair (1, 1)
    a.n += keep;
    a.seen += 1;
}

template <bool EXT, bool INST>
__global__ void __launch_bounds__(kMetricThreads) depth_errors_images_kernel(const ImageParams p) {   // cluster (c, 1, 1): a launch attribute
    __shared__ double warp_rows[kMetricThreads / 32][8];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank(), n_rank = cluster.num_blocks();
    const size_t b = blockIdx.y;
    const int want_id = p.inst ? p.group_ids[0] : -1;
    const float* gt = p.gt + b * p.px;
    const float* pred = p.pred + b * p.px;
    const uint8_t* inst = (p.inst && want_id >= 0) ? p.inst + b * p.px : nullptr;
    const float scale = p.pred_scale ? __ldg(p.pred_scale + b) : 1.0f;

    Acc a;
    acc_clear(a);
    Acc64 s{};
    const size_t tid = (size_t)rank * kMetricThreads + threadIdx.x, stride = (size_t)n_rank * kMetricThreads;
    int since = 0;
    if (p.vec4) {
        // the loads of the next group are issued before the current one is reduced (the top stall of this kernel was the
        // long scoreboard: nothing else in a thread overlaps its own loads)
        const size_t n4 = p.px >> 2;
        float4 g = make_float4(1.f, 1.f, 1.f, 1.f), q = g;
        uint32_t ids = 0;
        size_t i = tid;
        if (i < n4) {
            g = ld_stream_f32x4(gt + 4 * i);
            q = ld_stream_f32x4(pred + 4 * i);
            if (inst) ids = ld_stream_u32(inst + 4 * i);
        }
        while (i < n4) {
            const size_t nx = i + stride;
            float4 g2 = g, q2 = q;
            uint32_t ids2 = 0;
            if (nx < n4) {
                g2 = ld_stream_f32x4(gt + 4 * nx);
                q2 = ld_stream_f32x4(pred + 4 * nx);
                if (inst) ids2 = ld_stream_u32(inst + 4 * nx);
            }
            acc_masked<EXT, INST>(a, p, want_id, g.x, q.x, ids & 0xff, scale);
            acc_masked<EXT, INST>(a, p, want_id, g.y, q.y, (ids >> 8) & 0xff, scale);
            acc_masked<EXT, INST>(a, p, want_id, g.z, q.z, (ids >> 16) & 0xff, scale);
            acc_masked<EXT, INST>(a, p, want_id, g.w, q.w, ids >> 24, scale);
            if (++since == 16) {
                flush(s, a);
                since = 0;
            }
            g = g2; q = q2; ids = ids2;
            i = nx;
        }
    } else {
        for (size_t i = tid; i < p.px; i += stride) {
            acc_masked<EXT, INST>(a, p, want_id, gt[i], pred[i], inst ? inst[i] : 0, scale);
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
        for (unsigned r = 0; r < n_rank; ++r) {
            const double* remote = cluster.map_shared_rank(&warp_rows[0][0], r);
            v += remote[threadIdx.x];
        }
        p.sums[b * 8 + threadIdx.x] = v;
        warp_rows[1][threadIdx.x] = v;
    }
    cluster.sync();  // keep remote shared memory alive until rank 0 has read it
    if (rank == 0 && threadIdx.x == 0 && p.metrics) finalize(warp_rows[1], p.metrics + b * 7);
}

// ------------------------------------------------------------------------------------------
// median scaling (trainer.py:1413-1414, self-supervised configurations only):
//     depth_pred *= np.median(depth_gt[mask]) / np.median(depth_pred[mask])
// One CTA per (image, array).  The k-th smallest masked value is found by a three-pass radix select on the order-
// preserving 32-bit key of the float (11 + 11 + 10 bits, shared-memory histogram with integer atomics: exact and
// order-independent); an even count averages the two middle values in float32 as np.median does.
// ------------------------------------------------------------------------------------------
constexpr int kMedianThreads = 1024;

struct MedianParams {
    const float* gt;
    const float* pred;
    const uint8_t* inst;   // may be null
    size_t px;
    float min_d, max_d;
    int want_id, want_hi;  // want_id = -1: range mask only, else want_id <= id <= want_hi
    int clamp_first;       // the prediction is clamped to [min_d, max_d] before its median is taken
    float* medians;        // [B][2]: median of gt[mask], of pred[mask]
};

__device__ __forceinline__ uint32_t order_key(float v) {
    const uint32_t u = __float_as_uint(v);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float key_value(uint32_t k) {
    return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}

__global__ void __launch_bounds__(kMedianThreads) masked_median_kernel(const MedianParams p) {
    __shared__ unsigned hist[2048];
    __shared__ unsigned warp_tot[kMedianThreads / 32];
    __shared__ unsigned sel_bin, sel_rank, total, next_key;
    const size_t b = blockIdx.y;
    const float* gt = p.gt + b * p.px;
    const float* key_src = (blockIdx.x == 0 ? p.gt : p.pred) + b * p.px;
    const uint8_t* inst = (p.inst && p.want_id >= 0) ? p.inst + b * p.px : nullptr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool clamp_key = p.clamp_first && blockIdx.x == 1;

    // One sweep over the image: `visit(keep, key)` for every pixel; four pixels' loads are in flight per thread.
    auto sweep = [&](auto&& visit) {
        constexpr int kBatch = 4;
        for (size_t base = 0; base < p.px; base += (size_t)kBatch * kMedianThreads) {
            float g[kBatch], v[kBatch];
            int id[kBatch];
#pragma unroll
            for (int j = 0; j < kBatch; ++j) {
                const size_t i = base + (size_t)j * kMedianThreads + threadIdx.x;
                g[j] = 0.0f;                    // outside [min_d, max_d]: never kept
                v[j] = 0.0f;
                id[j] = p.want_id;
                if (i < p.px) {
                    g[j] = gt[i];
                    v[j] = key_src[i];
                    if (inst) id[j] = inst[i];
                }
            }
#pragma unroll
            for (int j = 0; j < kBatch; ++j) {
                const bool keep = (g[j] > p.min_d) & (g[j] < p.max_d) & (!inst || ((id[j] >= p.want_id) & (id[j] <= p.want_hi)));
                const float val = clamp_key ? fminf(fmaxf(v[j], p.min_d), p.max_d) : v[j];
                visit(keep, order_key(val));
            }
        }
    };

    // lower middle element: three-pass radix select
    uint32_t prefix = 0;
    unsigned rank = 0, n = 0, dup = 0;
    for (int pass = 0; pass < 3; ++pass) {
        const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
        const int bits = pass == 2 ? 10 : 11;
        for (int i = threadIdx.x; i < 2048; i += kMedianThreads) hist[i] = 0;
        __syncthreads();
        // a thread counts runs of the same bin in a register and issues one shared-memory atomic per run (the depths of
        // an image share a handful of exponents: the first pass would otherwise hammer a few addresses)
        uint32_t run_bin = 0xFFFFFFFFu;
        unsigned run_len = 0;
        sweep([&](bool keep, uint32_t k) {
            if (pass > 0 && (k >> (shift + bits)) != prefix) keep = false;
            if (!keep) return;
            const uint32_t bin = (k >> shift) & ((1u << bits) - 1);
            if (bin == run_bin) {
                ++run_len;
            } else {
                if (run_len) atomicAdd(&hist[run_bin], run_len);
                run_bin = bin;
                run_len = 1;
            }
        });
        if (run_len) atomicAdd(&hist[run_bin], run_len);
        __syncthreads();
        // exclusive scan over the 2048 bins (two per thread) to find the bin that holds `rank`
        const unsigned h0 = hist[2 * threadIdx.x], h1 = hist[2 * threadIdx.x + 1];
        unsigned incl = h0 + h1;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += v;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        unsigned before = incl - (h0 + h1);
        for (int w = 0; w < warp; ++w) before += warp_tot[w];
        if (pass == 0 && threadIdx.x == kMedianThreads - 1) total = before + h0 + h1;
        __syncthreads();
        if (pass == 0) {
            n = total;
            if (n == 0) break;
            rank = (n - 1) / 2;
        }
        if (rank >= before && rank < before + h0) {
            sel_bin = 2 * threadIdx.x;
            sel_rank = rank - before;
        } else if (rank >= before + h0 && rank < before + h0 + h1) {
            sel_bin = 2 * threadIdx.x + 1;
            sel_rank = rank - before - h0;
        }
        __syncthreads();
        prefix = (prefix << bits) | sel_bin;
        rank = sel_rank;
        dup = hist[sel_bin];          // after the last pass: how many masked values equal the selected one
        __syncthreads();
    }
    float lower = 0.0f, upper = 0.0f;
    if (n) {
        lower = upper = key_value(prefix);
        // even count: the upper middle element is the next one in sorted order -- the same value if it occurs again
        // beyond the selected rank, else the smallest masked value above it (one more sweep)
        if (!(n & 1) && rank + 1 >= dup) {
            if (threadIdx.x == 0) next_key = 0xFFFFFFFFu;
            __syncthreads();
            unsigned best = 0xFFFFFFFFu;
            sweep([&](bool keep, uint32_t k) {
                if (keep && k > prefix) best = min(best, k);
            });
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, off));
            if (lane == 0) atomicMin(&next_key, best);
            __syncthreads();
            upper = key_value(next_key);
        }
    }
    if (threadIdx.x == 0)
        p.medians[b * 2 + blockIdx.x] = n ? __fmul_rn(__fadd_rn(lower, upper), 0.5f) : __int_as_float(0x7fc00000);
}

__global__ void median_ratio_kernel(const float* medians, int B, float* scale) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) scale[b] = __fdiv_rn(medians[2 * b], medians[2 * b + 1]);     // np.float32 / np.float32
}

// ------------------------------------------------------------------------------------------
// every mask group of the evaluation loop in ONE pass over (gt, pred, inst)
// ------------------------------------------------------------------------------------------
// Each pixel's eight contributions are computed once.  A thread keeps the "all" group and the current material
// run in registers; when the material under its pixels changes it adds the finished run into its PRIVATE column of
// a shared-memory table [group][accumulator][thread] (no atomics, so the result is bitwise reproducible), then the
// CTA folds the table over threads in a fixed order and the cluster folds the CTAs through distributed shared memory.
struct Contrib {
    float f[4];
    int n, c1, c2, c3;
};

__device__ __forceinline__ Contrib make_contrib(float gt, float pred, bool keep) {
    gt = keep ? gt : 1.0f;          // rejected -> the neutral pair (1, 1): every float term is exactly 0
    pred = keep ? pred : 1.0f;
    const float hi = fmaxf(gt, pred), lo = fminf(gt, pred);
    const float thr = -5.9604644775390625e-08f * lo;
    Contrib c;
    c.n = keep;
    c.c1 = (int)keep & (int)ratio_below(hi, lo, 1.25f, thr);
    c.c2 = (int)keep & (int)ratio_below(hi, lo, 1.5625f, thr);
    c.c3 = (int)keep & (int)ratio_below(hi, lo, 1.953125f, thr);
    const float d = gt - pred, d2 = d * d, inv_gt = rcp_approx(gt);
    const float dl = lg2_approx(pred * inv_gt);
    c.f[0] = d2;
    c.f[1] = dl * dl;
    c.f[2] = fabsf(d) * inv_gt;
    c.f[3] = d2 * inv_gt;
    return c;
}

struct Run {
    float f[4];
    int n, c1, c2, c3;
};
__device__ __forceinline__ void run_clear(Run& r) {
    r.f[0] = r.f[1] = r.f[2] = r.f[3] = 0.0f;
    r.n = r.c1 = r.c2 = r.c3 = 0;
}
__device__ __forceinline__ void run_add(Run& r, const Contrib& c, bool on) {
    const float k = on ? 1.0f : 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) r.f[j] = fmaf(k, c.f[j], r.f[j]);
    r.n += on ? c.n : 0;
    r.c1 += on ? c.c1 : 0;
    r.c2 += on ? c.c2 : 0;
    r.c3 += on ? c.c3 : 0;
}
// table layout: [slot][kGroupRows][kMetricThreads] 32-bit words.  Rows 0-1 hold the four counts as 16-bit pairs
// (n | c1 << 16, c2 | c3 << 16: a thread sees at most px / 2048 <= 65535 pixels, checked on the host), rows 2-5 the four
// float sums: 6 words per thread and group instead of 8, which is what lets three CTAs instead of two share an SM.
constexpr int kGroupRows = 6;
__device__ __forceinline__ void run_flush(float* table, int slot, Run& r) {
    float* col = table + (size_t)slot * kGroupRows * kMetricThreads + threadIdx.x;
    unsigned* cnt = reinterpret_cast<unsigned*>(col);
    cnt[0 * kMetricThreads] += (unsigned)r.n | ((unsigned)r.c1 << 16);
    cnt[1 * kMetricThreads] += (unsigned)r.c2 | ((unsigned)r.c3 << 16);
    col[2 * kMetricThreads] += r.f[0];
    col[3 * kMetricThreads] += r.f[1];
    col[4 * kMetricThreads] += r.f[2];
    col[5 * kMetricThreads] += r.f[3];
    run_clear(r);
}

struct GroupParams {
    const float* gt;
    const float* pred;
    const uint8_t* inst;   // may be null when no material group is requested
    size_t px;
    float min_d, max_d;
    int n_groups;
    int all_slot;          // slot of the "all" group or -1
    uint8_t slot_of[256];  // instance id -> slot, 0xFF = not requested
    double* sums;          // [B][n_groups][8]
    float* metrics;        // [B][n_groups][7] or null
    bool vec4;
};

template <bool ALL>   // ALL: the "all" group (range mask only) is among the requested groups
__global__ void __launch_bounds__(kMetricThreads) depth_errors_groups_kernel(const __grid_constant__ GroupParams p) {   // cluster (c, 1, 1): a launch attribute
    extern __shared__ __align__(16) float table[];          // [n_groups][kGroupRows][kMetricThreads], see run_flush
    __shared__ double cta_tot[16][8];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank(), n_rank = cluster.num_blocks();
    const size_t b = blockIdx.y;
    const float* gt = p.gt + b * p.px;
    const float* pred = p.pred + b * p.px;
    const uint8_t* inst = p.inst ? p.inst + b * p.px : nullptr;
    for (int i = threadIdx.x; i < p.n_groups * kGroupRows * kMetricThreads; i += kMetricThreads) table[i] = 0.0f;   // 0.0f == 0u
    // (each thread only ever touches its own column, no barrier needed before use)

    // Every pixel is routed to exactly ONE slot: its material's, or -- when "all" is requested -- the slot of "all", which
    // collects the pixels of no requested material; "all" itself is the sum of every slot, formed after the cluster fold.
    // (One accumulate per pixel instead of two.)
    Run mat;
    run_clear(mat);
    int cur = 0xFF;
    const int other = ALL ? p.all_slot : 0xFF;
    auto pixel = [&](float g, float q, int id) {
        const bool keep = (g > p.min_d) & (g < p.max_d);
        const Contrib c = make_contrib(g, fminf(fmaxf(q, p.min_d), p.max_d), keep);
        int slot = inst ? p.slot_of[id] : 0xFF;
        slot = (slot == 0xFF) ? other : slot;
        if (slot != cur) {                                   // the material under this thread's pixels changed
            if (cur != 0xFF) run_flush(table, cur, mat);
            cur = slot;
        }
        run_add(mat, c, ALL ? true : slot != 0xFF);
    };
    const size_t tid = (size_t)rank * kMetricThreads + threadIdx.x, stride = (size_t)n_rank * kMetricThreads;
    if (p.vec4) {
        // The per-thread table limits this kernel to ~2 CTAs per SM, so memory latency is hidden inside the thread:
        // the loads of the next group are issued before the current one is reduced (1.70 -> 2.25 TB/s; a deeper
        // look-ahead gains nothing more, the dependent accumulation chain is what remains).
        const size_t n4 = p.px >> 2;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f), q = g;
        uint32_t ids = 0;
        size_t i = tid;
        if (i < n4) {
            g = ld_stream_f32x4(gt + 4 * i);
            q = ld_stream_f32x4(pred + 4 * i);
            if (inst) ids = ld_stream_u32(inst + 4 * i);
        }
        while (i < n4) {
            const size_t nx = i + stride;
            float4 g2 = g, q2 = q;
            uint32_t ids2 = 0;
            if (nx < n4) {
                g2 = ld_stream_f32x4(gt + 4 * nx);
                q2 = ld_stream_f32x4(pred + 4 * nx);
                if (inst) ids2 = ld_stream_u32(inst + 4 * nx);
            }
            pixel(g.x, q.x, ids & 0xff);
            pixel(g.y, q.y, (ids >> 8) & 0xff);
            pixel(g.z, q.z, (ids >> 16) & 0xff);
            pixel(g.w, q.w, ids >> 24);
            g = g2; q = q2; ids = ids2;
            i = nx;
        }
    } else {
        for (size_t i = tid; i < p.px; i += stride) pixel(gt[i], pred[i], inst ? inst[i] : 0);
    }
    if (cur != 0xFF) run_flush(table, cur, mat);
    __syncthreads();
    // CTA fold over threads in a fixed order: a warp takes (slot, row) pairs warp, warp + 8, ...; each lane adds its
    // eight columns (bank-conflict free), then a fixed shuffle tree.  Count rows are summed as exact integers.
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int pair = warp; pair < p.n_groups * kGroupRows; pair += kMetricThreads / 32) {
            const int slot = pair / kGroupRows, row = pair - slot * kGroupRows;
            const float* col = table + (size_t)pair * kMetricThreads;
            if (row < 2) {
                const unsigned* cnt = reinterpret_cast<const unsigned*>(col);
                unsigned lo = 0, hi = 0;
#pragma unroll
                for (int j = 0; j < kMetricThreads / 32; ++j) {
                    const unsigned w = cnt[lane + 32 * j];
                    lo += w & 0xffffu;
                    hi += w >> 16;
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    lo += __shfl_down_sync(0xffffffffu, lo, off);
                    hi += __shfl_down_sync(0xffffffffu, hi, off);
                }
                if (lane == 0) {
                    cta_tot[slot][2 * row] = (double)lo;
                    cta_tot[slot][2 * row + 1] = (double)hi;
                }
            } else {
                double v = 0.0;
#pragma unroll
                for (int j = 0; j < kMetricThreads / 32; ++j) v += (double)col[lane + 32 * j];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
                if (lane == 0) cta_tot[slot][row + 2] = (row == 3) ? v * 0.4804530139182014 : v;   // ln(2)^2 on the log2 term
            }
        }
    }
    cluster.sync();
    if (rank == 0 && threadIdx.x < p.n_groups * 8) {
        const int slot = threadIdx.x >> 3, k = threadIdx.x & 7;
        double v = 0.0;
        for (unsigned r = 0; r < n_rank; ++r) v += cluster.map_shared_rank(&cta_tot[0][0], r)[slot * 8 + k];
        if (!ALL || slot != p.all_slot) p.sums[(b * p.n_groups + slot) * 8 + k] = v;
        cta_tot[slot][k] = v;       // rank 0's own copy has been read by this same thread already
    }
    __syncthreads();
    if constexpr (ALL) {            // "all" = the pixels of no requested material + every material, in slot order
        double total = 0.0;
        if (rank == 0 && threadIdx.x < 8)
            for (int g = 0; g < p.n_groups; ++g) total += cta_tot[g][threadIdx.x];
        __syncthreads();
        if (rank == 0 && threadIdx.x < 8) {
            cta_tot[p.all_slot][threadIdx.x] = total;
            p.sums[(b * p.n_groups + p.all_slot) * 8 + threadIdx.x] = total;
        }
    }
    cluster.sync();
    if (rank == 0 && p.metrics && threadIdx.x < p.n_groups) finalize(cta_tot[threadIdx.x], p.metrics + (b * p.n_groups + threadIdx.x) * 7);
}

// {n_images, sum over images of the per-image metric rows} -- the accumulators of the reference's mean over images
// (np.array(errors).mean(0), trainer.py:1426; evaluation.py:283-285), additive across ranks.  Fixed summation order,
// so the result is bitwise reproducible.  NaN rows (empty masks)
// poison the mean exactly as they do in the reference.
__global__ void __launch_bounds__(128) image_mean_acc_kernel(const float* __restrict__ metrics, int B, int n_values, double* __restrict__ acc) {
    // one warp per (group, metric): lane l adds images l, l + 32, ... in order, then a fixed shuffle tree
    const int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (v == 0 && lane == 0) acc[0] = (double)B;
    if (v >= n_values) return;
    double t = 0.0;
    for (int b = lane; b < B; b += 32) t += (double)__ldg(metrics + (size_t)b * n_values + v);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) t += __shfl_down_sync(0xffffffffu, t, off);
    if (lane == 0) acc[1 + v] = t;
}

// image_mean_acc_kernel and the sum over ranks in ONE CTA (peer.cuh): warp w adds the images of values w, w + 32, ..., the
// accumulators go to shared memory, thread k publishes the k-th one to every rank and adds what the ranks sent, in rank order.
static_assert(1 + 16 * 7 <= kPeerMaxValues, "accumulators of the largest group list fit one exchange");
__global__ void __launch_bounds__(1024) image_mean_acc_peer_kernel(const __grid_constant__ PeerParams pp, const float* __restrict__ metrics,
                                                                   int B, int n_values, double* __restrict__ acc,
                                                                   double* __restrict__ acc_all) {
    __shared__ double local[kPeerMaxValues];
    const int lane = threadIdx.x & 31;
    for (int v = threadIdx.x >> 5; v < n_values; v += blockDim.x >> 5) {
        double t = 0.0;
        for (int b = lane; b < B; b += 32) t += (double)__ldg(metrics + (size_t)b * n_values + v);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_down_sync(0xffffffffu, t, off);
        if (lane == 0) local[1 + v] = t;
    }
    if (threadIdx.x == 0) local[0] = (double)B;
    __syncthreads();
    const int n = 1 + n_values, k = threadIdx.x;
    const double mine = k < n ? local[k] : 0.0;
    if (k < n) acc[k] = mine;
    const double total = peer_allreduce_cta(pp, mine, n);
    if (k < n) acc_all[k] = total;
}


// ------------------------------------------------------------------------------------------
// CTAs per image.  Both per-image kernels pay a fixed price per CTA (the table / accumulator fold and two cluster
// barriers), so the best cluster size depends on how the batch fills the machine (tools/probes/cluster_sweep.py, B200,
// 320 x 480 images, 11 groups: 8 CTAs per image 69 us at 120 images and 856 us at 1920; 3 CTAs 57 us; 2 CTAs 618 us):
//   * the batch fits one wave at some size: the LARGEST such size (shortest critical path, no second wave);
//   * it does not: 4 per image while the images alone still fit one wave (a 1.1-wave launch is the worst case),
//     2 per image beyond that (many waves: the fixed price per CTA is what is left to save).
// The per-thread summation order follows the size, so the float sums of an image are reproducible run to run for the
// same launch geometry (batch size, image size, GPU model) and agree to rounding (1e-7) across geometries; the counts
// are exact integers in every case.
// ------------------------------------------------------------------------------------------
int cta_capacity(const void* kernel, size_t smem) {       // CTAs of this kernel resident on the current device (cached)
    struct Entry {
        const void* kernel;
        size_t smem;
        int dev, capacity;
    };
    static std::mutex mutex;
    static Entry cache[32];
    static int used = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    std::lock_guard<std::mutex> lock(mutex);
    for (int i = 0; i < used; ++i)
        if (cache[i].kernel == kernel && cache[i].smem == smem && cache[i].dev == dev) return cache[i].capacity;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kMetricThreads, smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        return device_info().sms;
    }
    const int capacity = device_info().sms * per_sm;
    if (used < 32) cache[used++] = Entry{kernel, smem, dev, capacity};
    return capacity;
}
int pick_cluster(int B, int capacity, bool fine, size_t px) {
    int c = 2;
    if ((long long)B * 8 <= capacity) c = 8;
    else if (fine && (long long)B * 6 <= capacity) c = 6;
    else if ((long long)B * 4 <= capacity) c = 4;
    else if (fine && (long long)B * 3 <= capacity) c = 3;
    else if (B <= capacity) c = 4;
    while (c < kMaxCluster && px / ((size_t)c * kMetricThreads) + 8 > 65535) c = c < 4 ? c + 1 : (c == 4 ? 6 : 8);   // 16-bit per-thread counts (run_flush)
    return c;
}
template <typename Params>
int launch_clustered(void (*kernel)(const Params), const Params& p, int cluster, int B, size_t smem, cudaStream_t stream) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)cluster, (unsigned)B, 1);
    cfg.blockDim = dim3(kMetricThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = (unsigned)cluster;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return (int)e;
    }
    return launch_status();
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

static int launch_images(const float* gt, const float* pred, const uint8_t* inst, int B, size_t px, float min_d, float max_d,
                         const int* group_ids, int n_groups, double* sums, float* metrics, polcue_stream_t stream,
                         const float* pred_scale = nullptr, int inst_hi = -1, int clamp_first = 0) {
    if (!gt || !pred || !sums || B < 0 || B > 65535 || n_groups < 1 || n_groups > 16) return POLCUE_EINVAL;
    if (reinterpret_cast<uintptr_t>(sums) & 7) return POLCUE_EINVAL;
    if (B == 0) return POLCUE_OK;
    ImageParams p;
    p.gt = gt;
    p.pred = pred;
    p.inst = inst;
    p.pred_scale = pred_scale;
    p.want_hi = (inst && inst_hi >= 0) ? inst_hi : ((inst && n_groups >= 1) ? group_ids[0] : -1);
    p.clamp_first = clamp_first;
    p.px = px;
    p.min_d = min_d;
    p.max_d = max_d;
    p.n_groups = n_groups;
    for (int g = 0; g < 16; ++g) p.group_ids[g] = (g < n_groups && inst) ? group_ids[g] : -1;
    p.sums = sums;
    p.metrics = metrics;
    p.vec4 = (px % 4 == 0) && (((reinterpret_cast<uintptr_t>(gt) | reinterpret_cast<uintptr_t>(pred)) & 15) == 0) &&
             ((reinterpret_cast<uintptr_t>(inst) & 3) == 0);
    const bool ext = pred_scale || clamp_first || (inst && p.want_hi != p.group_ids[0]);
    auto kern = ext ? depth_errors_images_kernel<true, true>
                    : ((inst && p.group_ids[0] >= 0) ? depth_errors_images_kernel<false, true> : depth_errors_images_kernel<false, false>);
    const int cluster = pick_cluster(B, cta_capacity(reinterpret_cast<const void*>(kern), 0), false, 0);
    return launch_clustered(kern, p, cluster, B, 0, (cudaStream_t)stream);
}

int polcue_depth_errors_images_f32(const float* gt, const float* pred, const uint8_t* inst, int B, size_t px, float min_d,
                                   float max_d, int inst_id, double* sums, float* metrics, polcue_stream_t stream) {
    const int id = inst ? inst_id : -1;
    return launch_images(gt, pred, inst, B, px, min_d, max_d, &id, 1, sums, metrics, stream);
}

int polcue_depth_errors_images_scaled_f32(const float* gt, const float* pred, const uint8_t* inst, int B, size_t px, float min_d,
                                          float max_d, int inst_lo, int inst_hi, int clamp_first, const float* pred_scale,
                                          double* sums, float* metrics, polcue_stream_t stream) {
    if (inst && (inst_lo < 0 || inst_hi < inst_lo)) return POLCUE_EINVAL;
    const int id = inst ? inst_lo : -1;
    return launch_images(gt, pred, inst, B, px, min_d, max_d, &id, 1, sums, metrics, stream, pred_scale, inst_hi, clamp_first);
}

int polcue_masked_median_scale_f32(const float* gt, const float* pred, const uint8_t* inst, int B, size_t px, float min_d,
                                   float max_d, int inst_lo, int inst_hi, int clamp_first, float* medians, float* scale,
                                   polcue_stream_t stream) {
    if (!gt || !pred || !medians || B < 0 || B > 65535) return POLCUE_EINVAL;
    if (inst && (inst_lo < 0 || inst_hi < inst_lo)) return POLCUE_EINVAL;
    if (B == 0) return POLCUE_OK;
    MedianParams p;
    p.gt = gt;
    p.pred = pred;
    p.inst = inst;
    p.px = px;
    p.min_d = min_d;
    p.max_d = max_d;
    p.want_id = inst ? inst_lo : -1;
    p.want_hi = inst_hi;
    p.clamp_first = clamp_first;
    p.medians = medians;
    masked_median_kernel<<<dim3(2, B, 1), kMedianThreads, 0, (cudaStream_t)stream>>>(p);
    int rc = launch_status();
    if (rc != POLCUE_OK || !scale) return rc;
    median_ratio_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(medians, B, scale);
    return launch_status();
}

int polcue_depth_errors_groups_f32(const float* gt, const float* pred, const uint8_t* inst, int B, size_t px, float min_d,
                                   float max_d, const int* group_ids, int n_groups, double* sums, float* metrics,
                                   polcue_stream_t stream) {
    if (!gt || !pred || !sums || !group_ids || B < 0 || B > 65535 || n_groups < 1 || n_groups > 16) return POLCUE_EINVAL;
    if (reinterpret_cast<uintptr_t>(sums) & 7) return POLCUE_EINVAL;
    GroupParams p;
    p.all_slot = -1;
    for (int i = 0; i < 256; ++i) p.slot_of[i] = 0xFF;
    bool any_material = false;
    for (int g = 0; g < n_groups; ++g) {
        const int id = group_ids[g];
        if (id < 0) {
            if (p.all_slot >= 0) return POLCUE_EINVAL;            // "all" listed twice
            p.all_slot = g;
        } else {
            if (id > 255 || p.slot_of[id] != 0xFF || !inst) return POLCUE_EINVAL;   // out of range, duplicate, or no map
            p.slot_of[id] = (uint8_t)g;
            any_material = true;
        }
    }
    if (B == 0) return POLCUE_OK;
    p.gt = gt;
    p.pred = pred;
    p.inst = any_material ? inst : nullptr;
    p.px = px;
    p.min_d = min_d;
    p.max_d = max_d;
    p.n_groups = n_groups;
    p.sums = sums;
    p.metrics = metrics;
    p.vec4 = (px % 4 == 0) && (((reinterpret_cast<uintptr_t>(gt) | reinterpret_cast<uintptr_t>(pred)) & 15) == 0) &&
             ((reinterpret_cast<uintptr_t>(inst) & 3) == 0);
    if (px / ((size_t)kMaxCluster * kMetricThreads) + 8 > 65535) return POLCUE_E2BIG;   // 16-bit per-thread counts (run_flush)
    const size_t smem = (size_t)n_groups * kGroupRows * kMetricThreads * sizeof(float);
    auto kern = p.all_slot >= 0 ? depth_errors_groups_kernel<true> : depth_errors_groups_kernel<false>;
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int cluster = pick_cluster(B, cta_capacity(reinterpret_cast<const void*>(kern), smem), true, px);
    return launch_clustered(kern, p, cluster, B, smem, (cudaStream_t)stream);
}

namespace {
// A side stream and two events per device for the fork / join inside polcue_eval_pass_f32.  The mutex is held only while a
// call ENQUEUES its work, which fixes the order of the event records; it is never held while the GPU runs.
struct EvalLanes {
    cudaStream_t side = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
std::mutex g_eval_mutex;
EvalLanes g_eval_lanes[64];

cudaError_t eval_lanes(int dev, EvalLanes*& out) {
    EvalLanes& l = g_eval_lanes[dev];
    if (!l.side) {
        cudaError_t e = cudaStreamCreateWithFlags(&l.side, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&l.fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&l.join, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    out = &l;
    return cudaSuccess;
}
}  // namespace

static int eval_pass(const float* gt, const float* pred, const uint8_t* inst, const float* K, int B, int H, int W, float min_d,
                     float max_d, const int* group_ids, int n_groups, float* normals, double* sums, float* metrics,
                     double* mean_acc, polcue_peer* peer, double* mean_acc_all, polcue_stream_t stream) {
    if (!metrics || !mean_acc || H <= 0 || W <= 0 || B <= 0) return POLCUE_EINVAL;
    if (reinterpret_cast<uintptr_t>(mean_acc) & 7) return POLCUE_EINVAL;
    PeerParams pp;
    if (peer && (!peer_params(peer, pp) || !mean_acc_all || (reinterpret_cast<uintptr_t>(mean_acc_all) & 7))) return POLCUE_EINVAL;
    if (normals && !K) return POLCUE_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    int rc = POLCUE_OK;
    // The stencil and the metric pass only share their INPUT, and at the evaluation split's size (120 images) neither
    // fills the machine on its own (2-3 partial waves each): the stencil runs on a side stream, forked from and joined
    // back into `stream` with events, so the two overlap.  (Works under stream capture: the side stream joins the capture.)
    EvalLanes* lanes = nullptr;
    std::unique_lock<std::mutex> guard(g_eval_mutex, std::defer_lock);
    if (normals) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess || dev < 0 || dev >= 64) return e != cudaSuccess ? (int)e : POLCUE_EINVAL;
        guard.lock();
        e = eval_lanes(dev, lanes);
        if (e == cudaSuccess) e = cudaEventRecord(lanes->fork, s);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(lanes->side, lanes->fork, 0);
        if (e != cudaSuccess) return (int)e;
        rc = polcue_depth_to_normals_f32(gt, K, B, H, W, normals, lanes->side);
        e = cudaEventRecord(lanes->join, lanes->side);           // joined below even if the launch was refused
        if (e != cudaSuccess) return (int)e;
    }
    int rc2 = POLCUE_OK;
    if (rc == POLCUE_OK) rc2 = polcue_depth_errors_groups_f32(gt, pred, inst, B, (size_t)H * W, min_d, max_d, group_ids, n_groups, sums, metrics, s);
    if (rc == POLCUE_OK && rc2 == POLCUE_OK) {
        const int n_values = n_groups * 7;
        if (peer) image_mean_acc_peer_kernel<<<1, 1024, 0, s>>>(pp, metrics, B, n_values, mean_acc, mean_acc_all);
        else image_mean_acc_kernel<<<(n_values + 3) / 4, 128, 0, s>>>(metrics, B, n_values, mean_acc);
        rc2 = launch_status();
    }
    if (normals) {
        const cudaError_t e = cudaStreamWaitEvent(s, lanes->join, 0);
        if (e != cudaSuccess && rc == POLCUE_OK && rc2 == POLCUE_OK) return (int)e;
    }
    return rc != POLCUE_OK ? rc : rc2;
}

int polcue_eval_pass_f32(const float* gt, const float* pred, const uint8_t* inst, const float* K, int B, int H, int W, float min_d,
                         float max_d, const int* group_ids, int n_groups, float* normals, double* sums, float* metrics,
                         double* mean_acc, polcue_stream_t stream) {
    return eval_pass(gt, pred, inst, K, B, H, W, min_d, max_d, group_ids, n_groups, normals, sums, metrics, mean_acc, nullptr, nullptr, stream);
}

int polcue_eval_pass_peer_f32(const float* gt, const float* pred, const uint8_t* inst, const float* K, int B, int H, int W, float min_d,
                              float max_d, const int* group_ids, int n_groups, float* normals, double* sums, float* metrics,
                              double* mean_acc, polcue_peer* peer, double* mean_acc_all, polcue_stream_t stream) {
    if (!peer) return POLCUE_EINVAL;
    return eval_pass(gt, pred, inst, K, B, H, W, min_d, max_d, group_ids, n_groups, normals, sums, metrics, mean_acc, peer, mean_acc_all, stream);
}

}  // extern "C"
