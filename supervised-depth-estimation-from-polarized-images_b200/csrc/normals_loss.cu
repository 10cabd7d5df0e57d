// Supervised normals loss, forward and backward.
//
// Replaces Trainer.compute_supervised_normals_losses, manydepth/trainer.py:1298-1309 (called once per scale at :1248):
//     n_gt = depth_to_normals(depth_gt, K); n_pred = depth_to_normals(depth_pred, K)        (kornia 0.5.11)
//     cos  = F.cosine_similarity(n_gt, n_pred, dim=1)      = <a,b> / max(|a| |b|, 1e-8)      (torch 1.7.1)
//     loss = sum((2 - cos) * mask) / sum(mask)
// The reference runs ~25 launches per scale (two pads, two conv3d, crosses, normalisations, the similarity, the
// masked mean) and autograd stores every intermediate.  Here the forward is ONE kernel (both stencils, the cosine
// and the masked sums; 12 B read per pixel) and the backward is ONE kernel that recomputes the window quantities
// instead of storing them (12 B read + 4 B written per pixel).
//
// Backward, per image (depth_to_normals as in stencil.cu, with the unnormalised Sobel taps s = (1,2,1), d = (-1,0,1):
// gu = sum_ab s[a] d[b] P(p+(a,b)), gv = sum_ab d[a] s[b] P(p+(a,b)), P = (fx(u) Z, fy(v) Z, Z), n = gu x gv, b = n/|n|):
//     k_p   = -grad_out * m_p / sum(m)
//     n_bar = k_p (g - <g,b> b) / |n|,   g = dcos/db = a/D - <a,b> b / (D |b|^2),  D = max(|a||b|, 1e-8)
//     gu_bar = gv x n_bar,  gv_bar = n_bar x gu
//     Zbar(q) = fx(q) A_0(q) + fy(q) A_1(q) + A_2(q),
//     A_c(q)  = sum_{i,j in {-1,0,1}} Sv[i] Du[j] gu_bar_c(q+(i,j)) + Dv[i] Su[j] gv_bar_c(q+(i,j))
// where the per-axis weights fold the replicate padding back onto border pixels:
//     S[-1] = S[+1] = 1, D[-1] = +1, D[+1] = -1, S[0] = 2 + [q == 0] + [q == last], D[0] = [q == last] - [q == 0].
// Phase 1 of the backward kernel writes gu_bar, gv_bar of the tile + 1-pixel ring to shared memory, phase 2 gathers.
// Depth tiles (halo 2) are staged by TMA tensor copies when rows are 16-byte multiples, else by hand.
#include <cuda.h>

#include <mutex>

#include "polcue_device.cuh"
#include "polcue_host.h"

namespace polcue {
namespace {

constexpr int kLW = 128, kLossThreads = 256;
constexpr int kFwdH = 32, kBwdH = 14;                        // tile heights: forward keeps two depth tiles, backward also six adjoint fields;
                                                             // 14 + 2 ring rows = 16 rows x 32 pixel groups = two full rounds of 256 threads in phase 1
constexpr int kHalo = 2;
constexpr int kLBoxW = kLW + 8;                              // interior at column 4, row 2
constexpr int kLCol = 4, kLRow = kHalo;
__host__ __device__ constexpr int box_rows(int th) { return th + 2 * kHalo; }
__host__ __device__ constexpr uint32_t tile_bytes(int th) { return kLBoxW * box_rows(th) * sizeof(float); }
constexpr int kGH = kBwdH + 2, kGPitch = kLW + 8;                // adjoint fields of the tile + 1-pixel ring
constexpr int kGCol = 4;                                          // G column of tile column 0 (16-byte aligned rows)
constexpr int kMaxLossBlocks = 1 << 18;
constexpr uint32_t kLTilePad = (tile_bytes(kBwdH) + 127) / 128 * 128;           // TMA destinations are 128-byte aligned
constexpr size_t kBwdSmem = 2 * kLTilePad + 6 * kGH * kGPitch * sizeof(float);

struct LossParams {
    const float* gt;
    const float* pred;
    const float* K;
    const float* mask;     // may be null when range_mask: mask = (min_d <= gt <= max_d), trainer.py:1241-1242
    int range_mask;
    float min_d, max_d;
    int H, W;
    // forward
    unsigned long long* ticket;
    double* partials;      // [ctas][3]
    double* sums2;         // S = sum (2 - cos) m, M = sum m (and, with the L1 term, [2] = sum |gt - pred| m)
    float* loss;           // S / M
    // backward
    const float* grad_out; // device scalar
    const float* grad_l1;  // device scalar: upstream gradient of the supervised depth loss (L1 variant), may be null
    float* grad_pred;
};

// Stage a halo'd depth tile: TMA (zero-filled outside the image) or manual with clamped coordinates.
template <bool TMA, int TH>
__device__ __forceinline__ void stage_begin(float (*tile)[kLBoxW], const CUtensorMap* tmap, const float* plane, int H, int W, int x0,
                                            int y0, int b, uint64_t* bar) {
    if constexpr (TMA) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(tile_bytes(TH)) : "memory");
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                    smem_u32(&tile[0][0])),
                "l"(tmap), "r"(x0 - kLCol), "r"(y0 - kLRow), "r"(b), "r"(smem_u32(bar))
                : "memory");
        }
    } else {
        for (int i = threadIdx.x; i < box_rows(TH) * (kLW + 2 * kHalo); i += kLossThreads) {
            const int r = i / (kLW + 2 * kHalo), c = i - r * (kLW + 2 * kHalo);
            const int yy = min(max(y0 + r - kLRow, 0), H - 1);
            const int xx = min(max(x0 + c - kHalo, 0), W - 1);
            tile[r][kLCol - kHalo + c] = __ldg(plane + (size_t)yy * W + xx);
        }
    }
}

// After a TMA load: replicate the image border into the first ring outside the image (the only out-of-image cells
// any in-image pixel's window touches).
template <int TH>
__device__ __forceinline__ void patch_replicate(float (*tile)[kLBoxW], int H, int W, int x0, int y0) {
    const int last_x = W - 1 - x0, last_y = H - 1 - y0;
    const bool left = x0 == 0, right = last_x < kLW + 1, top = y0 == 0, bottom = last_y < TH + 1;
    if (left | right) {
        for (int r = threadIdx.x; r < box_rows(TH); r += kLossThreads) {
            if (left) tile[r][kLCol - 1] = tile[r][kLCol];
            if (right && last_x >= -1) tile[r][kLCol + last_x + 1] = tile[r][kLCol + last_x];
        }
    }
    __syncthreads();
    if (top | bottom) {
        for (int c = threadIdx.x; c < kLBoxW; c += kLossThreads) {
            if (top) tile[kLRow - 1][c] = tile[kLRow][c];
            if (bottom && last_y >= -1) tile[kLRow + last_y + 1][c] = tile[kLRow + last_y][c];
        }
    }
    __syncthreads();
}

struct Cam {
    float inv_fx, cx, inv_fy, cy;
};
__device__ __forceinline__ Cam load_cam(const float* K, int b) {
    const float* k = K + (size_t)b * 9;
    Cam c;
    c.inv_fx = 1.0f / __ldg(k + 0);
    c.cx = __ldg(k + 2);
    c.inv_fy = 1.0f / __ldg(k + 4);
    c.cy = __ldg(k + 5);
    return c;
}

// 8 x gradients of xyz at tile-local pixel (ty, tx) (image pixel (y0+ty, x0+tx)), same arithmetic as stencil.cu.
__device__ __forceinline__ void gradients(const float (*tile)[kLBoxW], int ty, int tx, int x, int y, int H, int W, const Cam& cam,
                                          float (&gu)[3], float (&gv)[3]) {
    float fx3[3], fy3[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        fx3[k] = ((float)min(max(x + k - 1, 0), W - 1) - cam.cx) * cam.inv_fx;
        fy3[k] = ((float)min(max(y + k - 1, 0), H - 1) - cam.cy) * cam.inv_fy;
    }
    float Z[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Z[r][c] = tile[kLRow + ty + r - 1][kLCol + tx + c - 1];
    float Su[3][3], Dv[3][3];
    const float fy1x2 = 2.0f * fy3[1];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        Su[2][c] = fmaf(2.0f, Z[1][c], Z[0][c] + Z[2][c]);
        Dv[2][c] = Z[2][c] - Z[0][c];
        Su[0][c] = fx3[c] * Su[2][c];
        Dv[0][c] = fx3[c] * Dv[2][c];
        Su[1][c] = fmaf(fy3[2], Z[2][c], fmaf(fy1x2, Z[1][c], fy3[0] * Z[0][c]));
        Dv[1][c] = __fsub_rn(__fmul_rn(fy3[2], Z[2][c]), __fmul_rn(fy3[0], Z[0][c]));
    }
#pragma unroll
    for (int comp = 0; comp < 3; ++comp) {
        gu[comp] = Su[comp][2] - Su[comp][0];
        gv[comp] = fmaf(2.0f, Dv[comp][1], Dv[comp][0] + Dv[comp][2]);
    }
}

// The same for four consecutive pixels starting at tile-local column tx0 (a multiple of 4): the 3 x 6 window is read
// as one LDS.128 + two LDS.32 per row and its column sums are shared by the four pixels (as stencil.cu does).
__device__ __forceinline__ void gradients4(const float (*tile)[kLBoxW], int ty, int tx0, int x, int y, int H, int W, const Cam& cam,
                                           float (&gu)[3][4], float (&gv)[3][4]) {
    float fx6[6], fy3[3];
#pragma unroll
    for (int c = 0; c < 6; ++c) fx6[c] = ((float)min(max(x + c - 1, 0), W - 1) - cam.cx) * cam.inv_fx;
#pragma unroll
    for (int r = 0; r < 3; ++r) fy3[r] = ((float)min(max(y + r - 1, 0), H - 1) - cam.cy) * cam.inv_fy;
    float Z[3][6];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float* row = &tile[kLRow + ty + r - 1][kLCol + tx0];
        const float4 mid = *reinterpret_cast<const float4*>(row);
        Z[r][0] = row[-1];
        Z[r][1] = mid.x; Z[r][2] = mid.y; Z[r][3] = mid.z; Z[r][4] = mid.w;
        Z[r][5] = row[4];
    }
    float Su[3][6], Dv[3][6];
    const float fy1x2 = 2.0f * fy3[1];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        Su[2][c] = fmaf(2.0f, Z[1][c], Z[0][c] + Z[2][c]);
        Dv[2][c] = Z[2][c] - Z[0][c];
        Su[0][c] = fx6[c] * Su[2][c];
        Dv[0][c] = fx6[c] * Dv[2][c];
        Su[1][c] = fmaf(fy3[2], Z[2][c], fmaf(fy1x2, Z[1][c], fy3[0] * Z[0][c]));
        Dv[1][c] = __fsub_rn(__fmul_rn(fy3[2], Z[2][c]), __fmul_rn(fy3[0], Z[0][c]));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int comp = 0; comp < 3; ++comp) {
            gu[comp][j] = Su[comp][j + 2] - Su[comp][j];
            gv[comp][j] = fmaf(2.0f, Dv[comp][j + 1], Dv[comp][j] + Dv[comp][j + 2]);
        }
}

__device__ __forceinline__ void cross_rn(const float (&a)[3], const float (&b)[3], float (&n)[3]) {
    n[0] = __fsub_rn(__fmul_rn(a[1], b[2]), __fmul_rn(a[2], b[1]));
    n[1] = __fsub_rn(__fmul_rn(a[2], b[0]), __fmul_rn(a[0], b[2]));
    n[2] = __fsub_rn(__fmul_rn(a[0], b[1]), __fmul_rn(a[1], b[0]));
}

constexpr float kInvCap = 1.0f / (64.0f * 1e-12f);   // n here is 64 x the reference's cross product (see stencil.cu)

// unit (or capped) normal and the factor it was scaled by
__device__ __forceinline__ float normalize3(const float (&n)[3], float (&u)[3]) {
    const float inv = fminf(rsqrt_approx(fmaf(n[0], n[0], fmaf(n[1], n[1], n[2] * n[2]))), kInvCap);
    u[0] = n[0] * inv;
    u[1] = n[1] * inv;
    u[2] = n[2] * inv;
    return inv;
}

// cos = <a,b> / max(|a||b|, 1e-8)   (torch 1.7.1 F.cosine_similarity: w12 * rsqrt(clamp_min(w1 w2, eps^2)))
__device__ __forceinline__ float cosine(const float (&a)[3], const float (&b)[3], float& inv_den, float& ab, float& bb,
                                        bool& clamped) {
    ab = fmaf(a[0], b[0], fmaf(a[1], b[1], a[2] * b[2]));
    const float aa = fmaf(a[0], a[0], fmaf(a[1], a[1], a[2] * a[2]));
    bb = fmaf(b[0], b[0], fmaf(b[1], b[1], b[2] * b[2]));
    clamped = aa * bb <= 1e-16f;
    inv_den = clamped ? 1e8f : rsqrtf(aa * bb);
    return ab * inv_den;
}

// Adjoints of the predicted-depth gradients of one pixel: (gu, gv) of the GT tile -> a; of the prediction -> b, n;
// k = -grad_out m / sum(m).  See the derivation in the file header.
__device__ __forceinline__ void adjoint_px(const float (&ug)[3], const float (&vg)[3], const float (&up)[3], const float (&vp)[3],
                                           float k, float (&gub)[3], float (&gvb)[3]) {
    gub[0] = gub[1] = gub[2] = gvb[0] = gvb[1] = gvb[2] = 0.0f;
    if (k == 0.0f) return;
    float n[3], a[3], bn[3];
    cross_rn(ug, vg, n);
    normalize3(n, a);
    cross_rn(up, vp, n);
    const float inv = normalize3(n, bn);
    float inv_den, ab, bb;
    bool clamped;
    cosine(a, bn, inv_den, ab, bb, clamped);
    // g = dcos/db: above the eps clamp a/D - <a,b> b / (D |b|^2), inside it a / eps
    const float w_b = clamped ? 0.0f : ab * inv_den / fmaxf(bb, 1e-30f);
    float g[3], nb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c] = a[c] * inv_den - w_b * bn[c];
    // b = n * inv: capped normalisation is linear, otherwise project out b and divide by |n| (= multiply by inv)
    const float gb = (inv >= kInvCap) ? 0.0f : fmaf(g[0], bn[0], fmaf(g[1], bn[1], g[2] * bn[2]));
#pragma unroll
    for (int c = 0; c < 3; ++c) nb[c] = k * inv * (g[c] - gb * bn[c]);
    // n = gu x gv  =>  gu_bar = gv x n_bar, gv_bar = n_bar x gu
    gub[0] = vp[1] * nb[2] - vp[2] * nb[1];
    gub[1] = vp[2] * nb[0] - vp[0] * nb[2];
    gub[2] = vp[0] * nb[1] - vp[1] * nb[0];
    gvb[0] = nb[1] * up[2] - nb[2] * up[1];
    gvb[1] = nb[2] * up[0] - nb[0] * up[2];
    gvb[2] = nb[0] * up[1] - nb[1] * up[0];
}

// mask of one pixel: the caller's float mask, or the supervised range test on the GT depth (trainer.py:1241-1242)
template <bool RANGE>   // RANGE is what the supervised (L1) entry points use; the plain ones read the caller's mask
__device__ __forceinline__ float mask_value(const LossParams& p, size_t idx, float gt_depth) {
    if constexpr (RANGE) return (gt_depth >= p.min_d && gt_depth <= p.max_d) ? 1.0f : 0.0f;
    else return __ldg(p.mask + idx);
}

// L1 = true adds the supervised depth loss of the same block of the trainer (trainer.py:1246):
//     supervised_depth_loss = (|gt - pred| * mask).sum() / mask.sum()
template <bool TMA, bool L1>
__global__ void __launch_bounds__(kLossThreads, 4) normals_loss_fwd_kernel(const __grid_constant__ CUtensorMap tm_gt,
                                                                        const __grid_constant__ CUtensorMap tm_pred, const LossParams p) {
    __shared__ __align__(128) float tg[box_rows(kFwdH)][kLBoxW];
    __shared__ __align__(128) float tp[box_rows(kFwdH)][kLBoxW];
    __shared__ uint64_t bar;
    __shared__ double red[kLossThreads / 32][3];
    __shared__ bool last;
    const int b = blockIdx.z, x0 = blockIdx.x * kLW, y0 = blockIdx.y * kFwdH;
    const size_t hw = (size_t)p.H * p.W;
    if (TMA && threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (TMA) __syncthreads();
    stage_begin<TMA, kFwdH>(tg, &tm_gt, p.gt + b * hw, p.H, p.W, x0, y0, b, &bar);
    stage_begin<TMA, kFwdH>(tp, &tm_pred, p.pred + b * hw, p.H, p.W, x0, y0, b, &bar);
    if constexpr (TMA) {
        lut_stage_wait(&bar);
        patch_replicate<kFwdH>(tg, p.H, p.W, x0, y0);
        patch_replicate<kFwdH>(tp, p.H, p.W, x0, y0);
    } else {
        __syncthreads();
    }
    const Cam cam = load_cam(p.K, b);
    float s = 0.0f, m = 0.0f, l1 = 0.0f;   // at most 16 pixels per thread: float32 partials, float64 from the warp level on
    for (int i = threadIdx.x; i < (kLW / 4) * kFwdH; i += kLossThreads) {
        const int ty = i / (kLW / 4), tx0 = 4 * (i - ty * (kLW / 4));
        const int x = x0 + tx0, y = y0 + ty;
        if (x >= p.W || y >= p.H) continue;
        float gu[3][4], gv[3][4], a4[3][4], b4[3][4];
        gradients4(tg, ty, tx0, x, y, p.H, p.W, cam, gu, gv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float u[3] = {gu[0][j], gu[1][j], gu[2][j]}, v[3] = {gv[0][j], gv[1][j], gv[2][j]};
            float n[3], un[3];
            cross_rn(u, v, n);
            normalize3(n, un);
            a4[0][j] = un[0]; a4[1][j] = un[1]; a4[2][j] = un[2];
        }
        gradients4(tp, ty, tx0, x, y, p.H, p.W, cam, gu, gv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float u[3] = {gu[0][j], gu[1][j], gu[2][j]}, v[3] = {gv[0][j], gv[1][j], gv[2][j]};
            float n[3], un[3];
            cross_rn(u, v, n);
            normalize3(n, un);
            b4[0][j] = un[0]; b4[1][j] = un[1]; b4[2][j] = un[2];
        }
        float fs = 0.0f, fm = 0.0f, fl = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (x + j < p.W) {
                const float zg = tg[kLRow + ty][kLCol + tx0 + j];
                const float mk = mask_value<L1>(p, b * hw + (size_t)y * p.W + x + j, zg);
                if constexpr (L1) fl = fmaf(fabsf(zg - tp[kLRow + ty][kLCol + tx0 + j]), mk, fl);
                const float a[3] = {a4[0][j], a4[1][j], a4[2][j]}, bb3[3] = {b4[0][j], b4[1][j], b4[2][j]};
                float inv_den, ab, bb;
                bool clamped;
                const float c = cosine(a, bb3, inv_den, ab, bb, clamped);
                fs = fmaf(2.0f - c, mk, fs);
                fm += mk;
            }
        }
        s += fs;
        m += fm;
        l1 += fl;
    }
    // deterministic reduction: warp shuffle -> shared -> per-CTA partial -> last CTA folds in fixed order
    double sd = s, md = m, ld = l1;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        sd += __shfl_down_sync(0xffffffffu, sd, off);
        md += __shfl_down_sync(0xffffffffu, md, off);
        if constexpr (L1) ld += __shfl_down_sync(0xffffffffu, ld, off);
    }
    if ((threadIdx.x & 31) == 0) {
        red[threadIdx.x >> 5][0] = sd;
        red[threadIdx.x >> 5][1] = md;
        red[threadIdx.x >> 5][2] = ld;
    }
    __syncthreads();
    const unsigned cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const unsigned ctas = gridDim.x * gridDim.y * gridDim.z;
    if (threadIdx.x == 0) {
        double ts = 0.0, tmk = 0.0, tl = 0.0;
        for (int w = 0; w < kLossThreads / 32; ++w) {
            ts += red[w][0];
            tmk += red[w][1];
            tl += red[w][2];
        }
        p.partials[3 * (size_t)cta] = ts;
        p.partials[3 * (size_t)cta + 1] = tmk;
        p.partials[3 * (size_t)cta + 2] = tl;
        __threadfence();
        last = atomicAdd(p.ticket, 1ull) == ctas - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    double fs = 0.0, fm = 0.0, fl1 = 0.0;
    for (unsigned j = threadIdx.x; j < ctas; j += kLossThreads) {
        fs += __ldcg(p.partials + 3 * (size_t)j);
        fm += __ldcg(p.partials + 3 * (size_t)j + 1);
        if constexpr (L1) fl1 += __ldcg(p.partials + 3 * (size_t)j + 2);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        fs += __shfl_down_sync(0xffffffffu, fs, off);
        fm += __shfl_down_sync(0xffffffffu, fm, off);
        if constexpr (L1) fl1 += __shfl_down_sync(0xffffffffu, fl1, off);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        red[threadIdx.x >> 5][0] = fs;
        red[threadIdx.x >> 5][1] = fm;
        red[threadIdx.x >> 5][2] = fl1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0.0, tmk = 0.0, tl = 0.0;
        for (int w = 0; w < kLossThreads / 32; ++w) {
            ts += red[w][0];
            tmk += red[w][1];
            tl += red[w][2];
        }
        p.sums2[0] = ts;
        p.sums2[1] = tmk;
        if constexpr (L1) p.sums2[2] = tl;
        if (p.loss) {
            p.loss[0] = (float)(ts / tmk);
            if constexpr (L1) p.loss[1] = (float)(tl / tmk);
        }
        *p.ticket = 0ull;
    }
}

// 72 KB of shared memory per CTA: three CTAs per SM, so up to 85 registers per thread (no spills)
template <bool TMA, bool L1>
__global__ void __launch_bounds__(kLossThreads, 3) normals_loss_bwd_kernel(const __grid_constant__ CUtensorMap tm_gt,
                                                                        const __grid_constant__ CUtensorMap tm_pred, const LossParams p) {
    // dynamic shared memory: two halo'd depth tiles and the six adjoint fields of the tile + 1-pixel ring (46 KB)
    extern __shared__ __align__(128) unsigned char bwd_smem[];
    float (*tg)[kLBoxW] = reinterpret_cast<float (*)[kLBoxW]>(bwd_smem);
    float (*tp)[kLBoxW] = reinterpret_cast<float (*)[kLBoxW]>(bwd_smem + kLTilePad);
    float (*G)[kGH][kGPitch] = reinterpret_cast<float (*)[kGH][kGPitch]>(bwd_smem + 2 * kLTilePad);   // gu_bar xyz, gv_bar xyz
    __shared__ uint64_t bar;
    const int b = blockIdx.z, x0 = blockIdx.x * kLW, y0 = blockIdx.y * kBwdH;
    const size_t hw = (size_t)p.H * p.W;
    if (TMA && threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (TMA) __syncthreads();
    stage_begin<TMA, kBwdH>(tg, &tm_gt, p.gt + b * hw, p.H, p.W, x0, y0, b, &bar);
    stage_begin<TMA, kBwdH>(tp, &tm_pred, p.pred + b * hw, p.H, p.W, x0, y0, b, &bar);
    if constexpr (TMA) {
        lut_stage_wait(&bar);
        patch_replicate<kBwdH>(tg, p.H, p.W, x0, y0);
        patch_replicate<kBwdH>(tp, p.H, p.W, x0, y0);
    } else {
        __syncthreads();
    }
    const Cam cam = load_cam(p.K, b);
    const float scale = -__ldg(p.grad_out) / (float)p.sums2[1];    // -grad_out / sum(mask)
    const float l1_scale = (L1 && p.grad_l1) ? __ldg(p.grad_l1) / (float)p.sums2[1] : 0.0f;

    // phase 1: adjoint of the two gradients at every pixel of the tile and its 1-pixel ring.
    // Core columns go four pixels at a time (shared 3 x 6 windows), the two ring columns one pixel at a time.
    for (int i = threadIdx.x; i < kGH * (kLW / 4); i += kLossThreads) {
        const int gy = i / (kLW / 4), tx0 = 4 * (i - gy * (kLW / 4));
        const int ty = gy - 1, x = x0 + tx0, y = y0 + ty;
        float out[6][4];
#pragma unroll
        for (int c = 0; c < 6; ++c)
#pragma unroll
            for (int j = 0; j < 4; ++j) out[c][j] = 0.0f;
        if (y >= 0 && y < p.H && x < p.W) {
            float gug[3][4], gvg[3][4], gup[3][4], gvp[3][4];
            gradients4(tg, ty, tx0, x, y, p.H, p.W, cam, gug, gvg);
            gradients4(tp, ty, tx0, x, y, p.H, p.W, cam, gup, gvp);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (x + j < p.W) {
                    const float k = scale * mask_value<L1>(p, b * hw + (size_t)y * p.W + x + j, tg[kLRow + ty][kLCol + tx0 + j]);
                    const float ug[3] = {gug[0][j], gug[1][j], gug[2][j]}, vg[3] = {gvg[0][j], gvg[1][j], gvg[2][j]};
                    const float up[3] = {gup[0][j], gup[1][j], gup[2][j]}, vp[3] = {gvp[0][j], gvp[1][j], gvp[2][j]};
                    float gub[3], gvb[3];
                    adjoint_px(ug, vg, up, vp, k, gub, gvb);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        out[c][j] = gub[c];
                        out[3 + c][j] = gvb[c];
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 6; ++c)
#pragma unroll
            for (int j = 0; j < 4; ++j) G[c][gy][kGCol + tx0 + j] = out[c][j];
    }
    for (int i = threadIdx.x; i < kGH * 2; i += kLossThreads) {
        const int gy = i >> 1, tx = (i & 1) ? kLW : -1;
        const int ty = gy - 1, gx = kGCol + tx;
        const int x = x0 + tx, y = y0 + ty;
        float gub[3] = {0.f, 0.f, 0.f}, gvb[3] = {0.f, 0.f, 0.f};
        if (x >= 0 && x < p.W && y >= 0 && y < p.H) {
            const float k = scale * mask_value<L1>(p, b * hw + (size_t)y * p.W + x, tg[kLRow + ty][kLCol + tx]);
            float ug[3], vg[3], up[3], vp[3];
            gradients(tg, ty, tx, x, y, p.H, p.W, cam, ug, vg);
            gradients(tp, ty, tx, x, y, p.H, p.W, cam, up, vp);
            adjoint_px(ug, vg, up, vp, k, gub, gvb);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            G[c][gy][gx] = gub[c];
            G[3 + c][gy][gx] = gvb[c];
        }
    }
    __syncthreads();

    // phase 2: gather the adjoint stencil four pixels at a time; the replicate padding is folded onto border pixels
    // by the per-axis weights (file header).  Vertical combinations first (shared by the four pixels), then horizontal.
    for (int i = threadIdx.x; i < (kLW / 4) * kBwdH; i += kLossThreads) {
        const int ty = i / (kLW / 4), tx0 = 4 * (i - ty * (kLW / 4));
        const int x = x0 + tx0, y = y0 + ty;
        if (x >= p.W || y >= p.H) continue;
        const float Sv[3] = {1.0f, 2.0f + (y == 0) + (y == p.H - 1), 1.0f};
        const float Dv[3] = {1.0f, (float)(y == p.H - 1) - (float)(y == 0), -1.0f};
        float VU[3][6], VV[3][6];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int k6 = 0; k6 < 6; ++k6) VU[c][k6] = VV[c][k6] = 0.0f;
#pragma unroll
        for (int di = 0; di < 3; ++di)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float* ru = &G[c][ty + di][kGCol + tx0];
                const float* rv = &G[3 + c][ty + di][kGCol + tx0];
                const float4 mu = *reinterpret_cast<const float4*>(ru), mv = *reinterpret_cast<const float4*>(rv);
                const float wu[6] = {ru[-1], mu.x, mu.y, mu.z, mu.w, ru[4]};
                const float wv[6] = {rv[-1], mv.x, mv.y, mv.z, mv.w, rv[4]};
#pragma unroll
                for (int k6 = 0; k6 < 6; ++k6) {
                    VU[c][k6] = fmaf(Sv[di], wu[k6], VU[c][k6]);
                    VV[c][k6] = fmaf(Dv[di], wv[k6], VV[c][k6]);
                }
            }
        const float fyq = ((float)y - cam.cy) * cam.inv_fy;
        float res[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int xj = x + j;
            const float Sh[3] = {1.0f, 2.0f + (xj == 0) + (xj == p.W - 1), 1.0f};
            const float Dh[3] = {1.0f, (float)(xj == p.W - 1) - (float)(xj == 0), -1.0f};
            float A[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float acc = 0.0f;
#pragma unroll
                for (int dj = 0; dj < 3; ++dj) acc = fmaf(Dh[dj], VU[c][j + dj], fmaf(Sh[dj], VV[c][j + dj], acc));
                A[c] = acc;
            }
            const float fxq = ((float)xj - cam.cx) * cam.inv_fx;
            res[j] = fmaf(fxq, A[0], fmaf(fyq, A[1], A[2]));
            if constexpr (L1) {     // d/dpred of sum |gt - pred| m / sum m: sign(pred - gt) m / sum m (sign(0) = 0, as torch.abs)
                if (xj < p.W) {
                    const float zg = tg[kLRow + ty][kLCol + tx0 + j], zp = tp[kLRow + ty][kLCol + tx0 + j];
                    const float mk = mask_value<L1>(p, b * hw + (size_t)y * p.W + xj, zg);
                    const float sgn = (zp > zg) ? 1.0f : ((zp < zg) ? -1.0f : 0.0f);
                    res[j] = fmaf(l1_scale * mk, sgn, res[j]);
                }
            }
        }
        float* o = p.grad_pred + b * hw + (size_t)y * p.W + x;
        if (x + 3 < p.W && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
            *reinterpret_cast<float4*>(o) = make_float4(res[0], res[1], res[2], res[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (x + j < p.W) o[j] = res[j];
        }
    }
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn loss_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

bool make_map(CUtensorMap* tmap, const float* base, int B, int H, int W, int tile_h) {
    EncodeTiledFn encode = loss_encode_tiled();
    if (!encode || W % 4 != 0 || (reinterpret_cast<uintptr_t>(base) & 15)) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)W * sizeof(float), (cuuint64_t)W * H * sizeof(float)};
    const cuuint32_t box[3] = {kLBoxW, (cuuint32_t)box_rows(tile_h), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return encode(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int check_common(const float* gt, const float* pred, const float* K, const float* mask, int B, int H, int W, int tile_h, bool cap,
                 dim3& grid) {
    if (!gt || !pred || !K || !mask || B < 0 || H <= 0 || W <= 0 || B > 65535) return POLCUE_EINVAL;
    if ((reinterpret_cast<uintptr_t>(gt) | reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(K) |
         reinterpret_cast<uintptr_t>(mask)) & 3)
        return POLCUE_EINVAL;
    grid = dim3((W + kLW - 1) / kLW, (H + tile_h - 1) / tile_h, B);
    if (grid.y > 65535) return POLCUE_E2BIG;
    if (cap && (unsigned long long)grid.x * grid.y * grid.z > (unsigned long long)kMaxLossBlocks) return POLCUE_E2BIG;   // one partial per CTA
    return POLCUE_OK;
}

}  // namespace
}  // namespace polcue

using namespace polcue;

extern "C" {

size_t polcue_normals_loss_workspace_bytes(void) { return 64 + (size_t)kMaxLossBlocks * 3 * sizeof(double); }

static int loss_forward(const float* depth_gt, const float* depth_pred, const float* K, const float* mask, bool range_mask, float min_d,
                        float max_d, bool with_l1, int B, int H, int W, void* workspace, double* sums, float* loss,
                        polcue_stream_t stream) {
    dim3 grid;
    const int rc = check_common(depth_gt, depth_pred, K, range_mask ? depth_gt : mask, B, H, W, kFwdH, true, grid);
    if (rc != POLCUE_OK) return rc;
    if (!workspace || !sums || (reinterpret_cast<uintptr_t>(workspace) & 63) || (reinterpret_cast<uintptr_t>(sums) & 7))
        return POLCUE_EINVAL;
    if (B == 0) return POLCUE_EINVAL;   // the loss of an empty batch is 0/0; let the caller decide
    LossParams p{};
    p.gt = depth_gt;
    p.pred = depth_pred;
    p.K = K;
    p.mask = mask;
    p.range_mask = range_mask ? 1 : 0;
    p.min_d = min_d;
    p.max_d = max_d;
    p.H = H;
    p.W = W;
    p.ticket = static_cast<unsigned long long*>(workspace);
    p.partials = reinterpret_cast<double*>(static_cast<char*>(workspace) + 64);
    p.sums2 = sums;
    p.loss = loss;
    CUtensorMap mg, mp;
    cudaStream_t s = (cudaStream_t)stream;
    const bool tma = make_map(&mg, depth_gt, B, H, W, kFwdH) && make_map(&mp, depth_pred, B, H, W, kFwdH);
    if (with_l1) {
        if (tma) normals_loss_fwd_kernel<true, true><<<grid, kLossThreads, 0, s>>>(mg, mp, p);
        else normals_loss_fwd_kernel<false, true><<<grid, kLossThreads, 0, s>>>(mg, mp, p);
    } else {
        if (tma) normals_loss_fwd_kernel<true, false><<<grid, kLossThreads, 0, s>>>(mg, mp, p);
        else normals_loss_fwd_kernel<false, false><<<grid, kLossThreads, 0, s>>>(mg, mp, p);
    }
    return launch_status();
}

static int loss_backward(const float* depth_gt, const float* depth_pred, const float* K, const float* mask, bool range_mask, float min_d,
                         float max_d, bool with_l1, int B, int H, int W, const double* sums, const float* grad_out,
                         const float* grad_l1, float* grad_pred, polcue_stream_t stream) {
    dim3 grid;
    const int rc = check_common(depth_gt, depth_pred, K, range_mask ? depth_gt : mask, B, H, W, kBwdH, false, grid);
    if (rc != POLCUE_OK) return rc;
    if (!sums || !grad_out || !grad_pred || (reinterpret_cast<uintptr_t>(grad_pred) & 3)) return POLCUE_EINVAL;
    if (B == 0) return POLCUE_OK;
    LossParams p{};
    p.gt = depth_gt;
    p.pred = depth_pred;
    p.K = K;
    p.mask = mask;
    p.range_mask = range_mask ? 1 : 0;
    p.min_d = min_d;
    p.max_d = max_d;
    p.H = H;
    p.W = W;
    p.sums2 = const_cast<double*>(sums);
    p.grad_out = grad_out;
    p.grad_l1 = grad_l1;
    p.grad_pred = grad_pred;
    CUtensorMap mg, mp;
    cudaStream_t s = (cudaStream_t)stream;
    const bool tma = make_map(&mg, depth_gt, B, H, W, kBwdH) && make_map(&mp, depth_pred, B, H, W, kBwdH);
    auto kern = with_l1 ? (tma ? normals_loss_bwd_kernel<true, true> : normals_loss_bwd_kernel<false, true>)
                        : (tma ? normals_loss_bwd_kernel<true, false> : normals_loss_bwd_kernel<false, false>);
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, kLossThreads, kBwdSmem, s>>>(mg, mp, p);
    return launch_status();
}

int polcue_normals_loss_fwd_f32(const float* depth_gt, const float* depth_pred, const float* K, const float* mask, int B, int H,
                                int W, void* workspace, double* sums2, float* loss, polcue_stream_t stream) {
    return loss_forward(depth_gt, depth_pred, K, mask, false, 0.0f, 0.0f, false, B, H, W, workspace, sums2, loss, stream);
}

int polcue_normals_loss_bwd_f32(const float* depth_gt, const float* depth_pred, const float* K, const float* mask, int B, int H,
                                int W, const double* sums2, const float* grad_out, float* grad_pred, polcue_stream_t stream) {
    return loss_backward(depth_gt, depth_pred, K, mask, false, 0.0f, 0.0f, false, B, H, W, sums2, grad_out, nullptr, grad_pred, stream);
}

int polcue_supervised_losses_fwd_f32(const float* depth_gt, const float* depth_pred, const float* K, float min_d, float max_d, int B,
                                     int H, int W, void* workspace, double* sums3, float* losses2, polcue_stream_t stream) {
    return loss_forward(depth_gt, depth_pred, K, nullptr, true, min_d, max_d, true, B, H, W, workspace, sums3, losses2, stream);
}

int polcue_supervised_losses_bwd_f32(const float* depth_gt, const float* depth_pred, const float* K, float min_d, float max_d, int B,
                                     int H, int W, const double* sums3, const float* grad_normals, const float* grad_depth,
                                     float* grad_pred, polcue_stream_t stream) {
    if (!grad_depth) return POLCUE_EINVAL;
    return loss_backward(depth_gt, depth_pred, K, nullptr, true, min_d, max_d, true, B, H, W, sums3, grad_normals, grad_depth, grad_pred,
                         stream);
}

}  // extern "C"
