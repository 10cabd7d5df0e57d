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
// Forward: the two cross products in a cancellation-free form (scaled_normals4 below: half the FP operations of the
// reference's operation order, which the stencil kernel and the backward keep), the cosine from the unnormalised vectors.
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
// Two kernel families: packed (rows that are 16-byte multiples: GT and prediction in the two lanes of packed FP32
// instructions, interleaved shared tile) and scalar (any width / alignment); both stage their tiles by hand.
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
constexpr uint32_t kLTilePad = (tile_bytes(kBwdH) + 127) / 128 * 128;
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

// Stage a halo'd depth tile by hand with clamped coordinates (replicate padding).
template <int TH>
__device__ __forceinline__ void stage_tile(float (*tile)[kLBoxW], const float* plane, int H, int W, int x0, int y0) {
    for (int i = threadIdx.x; i < box_rows(TH) * (kLW + 2 * kHalo); i += kLossThreads) {
        const int r = i / (kLW + 2 * kHalo), c = i - r * (kLW + 2 * kHalo);
        const int yy = min(max(y0 + r - kLRow, 0), H - 1);
        const int xx = min(max(x0 + c - kHalo, 0), W - 1);
        tile[r][kLCol - kHalo + c] = __ldg(plane + (size_t)yy * W + xx);
    }
}

struct Cam {
    float inv_fx, cx, inv_fy, cy;
};
// The four intrinsics are REQUESTED at the top of a kernel (cam_fetch) and turned into reciprocals after the tile's barrier
// (cam_finish): fetched after the barrier, every warp of the CTA sat out an L2 round trip there (7 % of the backward's stall samples).
struct CamRaw {
    float fx, cx, fy, cy;
};
__device__ __forceinline__ CamRaw cam_fetch(const float* K, int b) {
    const float* k = K + (size_t)b * 9;
    CamRaw r;
    r.fx = __ldg(k + 0);
    r.cx = __ldg(k + 2);
    r.fy = __ldg(k + 4);
    r.cy = __ldg(k + 5);
    return r;
}
__device__ __forceinline__ Cam cam_finish(const CamRaw& r) {
    Cam c;
    c.inv_fx = 1.0f / r.fx;
    c.cx = r.cx;
    c.inv_fy = 1.0f / r.fy;
    c.cy = r.cy;
    return c;
}
__device__ __forceinline__ Cam load_cam(const float* K, int b) { return cam_finish(cam_fetch(K, b)); }

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
    // The stencil arithmetic is DEFINED here, rounding by rounding (no compiler-chosen contraction), and repeated verbatim by
    // gradients4, the packed twins below and stencil.cu, so every kernel gives the same bits:
    //   S2 = fma(2, Z1, Z0 + Z2), D2 = Z2 - Z0                     vertical smoothing / difference of Z per column
    //   X: P = fl(fx S2), Q = fl(fx D2);  gu_x = P[+1] - P[-1];  gv_x = fma(2, Q[0], Q[-1] + Q[+1])
    //   Y: lo = fl(fy[-1] Z0), hi = fl(fy[+1] Z2), S1 = fma(fy[+1], Z2, fma(2 fy[0], Z1, lo)), D1 = hi - lo;
    //      gu_y = S1[+1] - S1[-1];  gv_y = fma(2, D1[0], D1[-1] + D1[+1])
    //   Z: gu_z = S2[+1] - S2[-1];  gv_z = fma(2, D2[0], D2[-1] + D2[+1])
    // Products that feed a difference are rounded separately, so replicated rows / columns (image borders, one-pixel-wide
    // images) and parallel gradients cancel to exact zeros.
    float S2[3], D2[3], S1[3], D1[3], P[3], Q[3];
    const float fy1x2 = 2.0f * fy3[1];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        S2[c] = fmaf(2.0f, Z[1][c], __fadd_rn(Z[0][c], Z[2][c]));
        D2[c] = __fsub_rn(Z[2][c], Z[0][c]);
        P[c] = __fmul_rn(fx3[c], S2[c]);
        Q[c] = __fmul_rn(fx3[c], D2[c]);
        const float lo = __fmul_rn(fy3[0], Z[0][c]);
        S1[c] = fmaf(fy3[2], Z[2][c], fmaf(fy1x2, Z[1][c], lo));
        D1[c] = __fsub_rn(__fmul_rn(fy3[2], Z[2][c]), lo);
    }
    gu[0] = __fsub_rn(P[2], P[0]);
    gv[0] = fmaf(2.0f, Q[1], __fadd_rn(Q[0], Q[2]));
    gu[1] = __fsub_rn(S1[2], S1[0]);
    gv[1] = fmaf(2.0f, D1[1], __fadd_rn(D1[0], D1[2]));
    gu[2] = __fsub_rn(S2[2], S2[0]);
    gv[2] = fmaf(2.0f, D2[1], __fadd_rn(D2[0], D2[2]));
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
    float S2[6], D2[6], S1[6], D1[6], P[6], Q[6];
    const float fy1x2 = 2.0f * fy3[1];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        S2[c] = fmaf(2.0f, Z[1][c], __fadd_rn(Z[0][c], Z[2][c]));
        D2[c] = __fsub_rn(Z[2][c], Z[0][c]);
        P[c] = __fmul_rn(fx6[c], S2[c]);
        Q[c] = __fmul_rn(fx6[c], D2[c]);
        const float lo = __fmul_rn(fy3[0], Z[0][c]);
        S1[c] = fmaf(fy3[2], Z[2][c], fmaf(fy1x2, Z[1][c], lo));
        D1[c] = __fsub_rn(__fmul_rn(fy3[2], Z[2][c]), lo);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        gu[0][j] = __fsub_rn(P[j + 2], P[j]);
        gv[0][j] = fmaf(2.0f, Q[j + 1], __fadd_rn(Q[j], Q[j + 2]));
        gu[1][j] = __fsub_rn(S1[j + 2], S1[j]);
        gv[1][j] = fmaf(2.0f, D1[j + 1], __fadd_rn(D1[j], D1[j + 2]));
        gu[2][j] = __fsub_rn(S2[j + 2], S2[j]);
        gv[2][j] = fmaf(2.0f, D2[j + 1], __fadd_rn(D2[j], D2[j + 2]));
    }
}

// products rounded separately (never contracted): parallel vectors cancel to an exact zero vector
__device__ __forceinline__ void cross_rn(const float (&a)[3], const float (&b)[3], float (&n)[3]) {
    n[0] = __fsub_rn(__fmul_rn(a[1], b[2]), __fmul_rn(a[2], b[1]));
    n[1] = __fsub_rn(__fmul_rn(a[2], b[0]), __fmul_rn(a[0], b[2]));
    n[2] = __fsub_rn(__fmul_rn(a[0], b[1]), __fmul_rn(a[1], b[0]));
}

constexpr float kInvCap = 1.0f / (64.0f * 1e-12f);   // n here is 64 x the reference's cross product (see stencil.cu)

// unit (or capped) normal and the factor it was scaled by
__device__ __forceinline__ float normalize3(const float (&n)[3], float (&u)[3]) {
    const float inv = fminf(rsqrt_approx(fmaf(n[0], n[0], fmaf(n[1], n[1], __fmul_rn(n[2], n[2])))), kInvCap);
    u[0] = __fmul_rn(n[0], inv);
    u[1] = __fmul_rn(n[1], inv);
    u[2] = __fmul_rn(n[2], inv);
    return inv;
}

// cos = <a,b> / max(|a||b|, 1e-8)   (torch 1.7.1 F.cosine_similarity: w12 * rsqrt(clamp_min(w1 w2, eps^2)))
// the part of `cosine` after the three dot products (the packed kernels form aa and bb in the two lanes of one register pair)
__device__ __forceinline__ float cosine_from_dots(float ab, float aa, float bb, float& inv_den, bool& clamped) {
    const float den2 = __fmul_rn(aa, bb);
    clamped = den2 <= 1e-16f;
    inv_den = clamped ? 1e8f : rsqrt_approx(den2);   // den2 > 1e-16 here: a normal number, the same MUFU.RSQ result rsqrtf() returns without its denormal fix-up
    return __fmul_rn(ab, inv_den);
}
__device__ __forceinline__ float cosine(const float (&a)[3], const float (&b)[3], float& inv_den, float& ab, float& bb,
                                        bool& clamped) {
    ab = fmaf(a[0], b[0], fmaf(a[1], b[1], __fmul_rn(a[2], b[2])));
    const float aa = fmaf(a[0], a[0], fmaf(a[1], a[1], __fmul_rn(a[2], a[2])));
    bb = fmaf(b[0], b[0], fmaf(b[1], b[1], __fmul_rn(b[2], b[2])));
    return cosine_from_dots(ab, aa, bb, inv_den, clamped);
}
// the same from the packed unit normals n (lane 0 = a, lane 1 = b): (aa, bb) cost three packed instructions instead of six
__device__ __forceinline__ float cosine_pairs(const f32x2 (&n)[3], const float (&a)[3], const float (&b)[3], float& inv_den, float& ab,
                                              float& bb, bool& clamped) {
    ab = fmaf(a[0], b[0], fmaf(a[1], b[1], __fmul_rn(a[2], b[2])));
    float aa;
    unpk2(fma2(n[0], n[0], fma2(n[1], n[1], mul2(n[2], n[2]))), aa, bb);
    return cosine_from_dots(ab, aa, bb, inv_den, clamped);
}

// The adjoint arithmetic after the cosine, with every rounding spelled out (no compiler-chosen contraction) so that the
// scalar and the packed kernel produce the same bits:
//   g = dcos/db: above the eps clamp a/D - <a,b> b / (D |b|^2), inside it a / eps;   b = n * inv: the capped normalisation is
//   linear, otherwise project out b and divide by |n| (= multiply by inv);   n = gu x gv  =>  gu_bar = gv x n_bar, gv_bar = n_bar x gu.
__device__ __forceinline__ void adjoint_tail(const float (&a)[3], const float (&bn)[3], float inv, const float (&up)[3], const float (&vp)[3],
                                             float k, float inv_den, float ab, float bb, bool clamped, float (&gub)[3], float (&gvb)[3]) {
    // (MUFU reciprocal, 1 ulp: |b|^2 is 1 to rounding unless the normalisation was capped; the IEEE division was 7 % of the backward's instructions)
    const float w_b = clamped ? 0.0f : __fmul_rn(__fmul_rn(ab, inv_den), rcp_approx(fmaxf(bb, 1e-30f)));
    float g[3], nb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c] = fmaf(-w_b, bn[c], __fmul_rn(a[c], inv_den));
    const float gb = (inv >= kInvCap) ? 0.0f : fmaf(g[0], bn[0], fmaf(g[1], bn[1], __fmul_rn(g[2], bn[2])));
    const float ki = __fmul_rn(k, inv);
#pragma unroll
    for (int c = 0; c < 3; ++c) nb[c] = __fmul_rn(ki, fmaf(-gb, bn[c], g[c]));
    gub[0] = fmaf(vp[1], nb[2], -__fmul_rn(vp[2], nb[1]));
    gub[1] = fmaf(vp[2], nb[0], -__fmul_rn(vp[0], nb[2]));
    gub[2] = fmaf(vp[0], nb[1], -__fmul_rn(vp[1], nb[0]));
    gvb[0] = fmaf(nb[1], up[2], -__fmul_rn(nb[2], up[1]));
    gvb[1] = fmaf(nb[2], up[0], -__fmul_rn(nb[0], up[2]));
    gvb[2] = fmaf(nb[0], up[1], -__fmul_rn(nb[1], up[0]));
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
    adjoint_tail(a, bn, inv, up, vp, k, inv_den, ab, bb, clamped, gub, gvb);
}

// mask of one pixel: the caller's float mask, or the supervised range test on the GT depth (trainer.py:1241-1242)
template <bool RANGE>   // RANGE is what the supervised (L1) entry points use; the plain ones read the caller's mask
__device__ __forceinline__ float mask_value(const LossParams& p, size_t idx, float gt_depth) {
    if constexpr (RANGE) return (gt_depth >= p.min_d && gt_depth <= p.max_d) ? 1.0f : 0.0f;
    else return __ldg(p.mask + idx);
}

// Deterministic reduction shared by the forward kernels: float32 inside a warp (512 pixels; the mask sum is an exact
// integer), the eight warp sums added in order in float64 -> one partial per CTA.  loss_fold_kernel (a second, tiny
// launch) adds the partials in a fixed order -- bitwise reproducible, and no CTA of the main kernel waits on a ticket.
template <bool L1>
__device__ __forceinline__ void loss_block_reduce(const LossParams& p, float s, float m, float l1, double (*red)[3]) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s = __fadd_rn(s, __shfl_down_sync(0xffffffffu, s, off));
        m = __fadd_rn(m, __shfl_down_sync(0xffffffffu, m, off));
        if constexpr (L1) l1 = __fadd_rn(l1, __shfl_down_sync(0xffffffffu, l1, off));
    }
    if ((threadIdx.x & 31) == 0) {
        red[threadIdx.x >> 5][0] = (double)s;
        red[threadIdx.x >> 5][1] = (double)m;
        red[threadIdx.x >> 5][2] = (double)l1;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int w = 0; w < kLossThreads / 32; ++w) t += red[w][threadIdx.x];
        const unsigned cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        p.partials[3 * (size_t)cta + threadIdx.x] = t;
    }
}

constexpr int kFoldThreads = 1024;
__global__ void __launch_bounds__(kFoldThreads) loss_fold_kernel(const double* __restrict__ partials, unsigned ctas, int with_l1,
                                                                 double* __restrict__ sums, float* __restrict__ loss) {
    __shared__ double red[kFoldThreads / 32][3];
    double f[3] = {0.0, 0.0, 0.0};
    for (unsigned j = threadIdx.x; j < ctas; j += kFoldThreads) {
        f[0] += partials[3 * (size_t)j];
        f[1] += partials[3 * (size_t)j + 1];
        f[2] += partials[3 * (size_t)j + 2];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int k = 0; k < 3; ++k) f[k] += __shfl_down_sync(0xffffffffu, f[k], off);
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < 3; ++k) red[threadIdx.x >> 5][k] = f[k];
    __syncthreads();
    if (threadIdx.x == 0) {
        double t[3] = {0.0, 0.0, 0.0};
        for (int w = 0; w < kFoldThreads / 32; ++w)
            for (int k = 0; k < 3; ++k) t[k] += red[w][k];
        sums[0] = t[0];
        sums[1] = t[1];
        if (with_l1) sums[2] = t[2];
        if (loss) {
            loss[0] = (float)(t[0] / t[1]);
            if (with_l1) loss[1] = (float)(t[2] / t[1]);
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// FORWARD stencil in cancellation-free form.  The loss only needs the DIRECTION of the two cross products, and both
// kernels above (stencil.cu's arithmetic) spend most of their FP work forming fx(u) Z and fy(v) Z per window column
// only to difference them again.  With fx(u + b) = fx(u) + b / f_x (clamped coordinates: b in {-1, 0, 1} becomes the
// weights w below) the gradients of P = (fx Z, fy Z, Z) are
//     gu = G r + (A / f_x, Cu / f_y, 0),   gv = V r + (Cv / f_x, B / f_y, 0),   r = (fx(u), fy(v), 1)
// with, from the vertical column combinations S2 = Z0 + 2 Z1 + Z2, D2 = Z2 - Z0 (rows y-1, y, y+1) and the row-weighted
// T = wp Z2 + wm Z0, D = wp Z2 - wm Z0 (wm = [y > 0], wp = [y < H-1]: replicated rows carry no fy step):
//     G  = S2[x+1] - S2[x-1]                     V  = D2[x-1] + 2 D2[x] + D2[x+1]          (the z components of gu, gv)
//     A  = wr S2[x+1] + wl S2[x-1]               B  = T[x-1] + 2 T[x] + T[x+1]
//     Cu = D[x+1] - D[x-1]                       Cv = wr D2[x+1] - wl D2[x-1]              (wl = [x > 0], wr = [x < W-1])
// and the cross product, scaled by the positive constant f_x f_y (which the cosine does not see), is
//     m = f_x f_y (gu x gv) = ( f_x (V Cu - G B),  f_y (G Cv - V A),  -fx(u) m_x - fy(v) m_y + A B - Cu Cv ).
// 35 instead of 66 FP operations per pixel and field, and none of the differences of nearly equal products that make
// the reference's float32 normals noisy: the result is closer to the float64 oracle than the reference's own float32
// (tests: the loss against the float64 oracle; the stencil KERNEL keeps the reference's operation order).
// The clamps of F.normalize (|n| >= 1e-12) and of cosine_similarity (|a||b| >= 1e-8) only matter for |n| < 1e-12, i.e.
// q = |m|^2 < qthr: those pixels (zero-depth holes) take the explicit path of cos_scaled.
// Templated on the lane type: float (scalar kernel) and f32x2 (GT and prediction in the two lanes) run the identical
// operation sequence, so the two kernels agree bit for bit.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float vadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float vsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float vmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float vfma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ float vneg(float a) { return -a; }
__device__ __forceinline__ f32x2 vadd(f32x2 a, f32x2 b) { return add2(a, b); }
__device__ __forceinline__ f32x2 vsub(f32x2 a, f32x2 b) { return sub2(a, b); }
__device__ __forceinline__ f32x2 vmul(f32x2 a, f32x2 b) { return mul2(a, b); }
__device__ __forceinline__ f32x2 vfma(f32x2 a, f32x2 b, f32x2 c) { return fma2(a, b, c); }
__device__ __forceinline__ f32x2 vneg(f32x2 a) { return neg2(a); }
template <typename V> __device__ __forceinline__ V vdup(float c);
template <> __device__ __forceinline__ float vdup<float>(float c) { return c; }
template <> __device__ __forceinline__ f32x2 vdup<f32x2>(float c) { return dup2(c); }

struct FwdCam {
    float nfX, nfY;            // -f_x, -f_y
    float cx, inv_fx, cy, inv_fy;
    float kap, qthr;           // |n_ref| 1e12 = |m| kap (n_ref = the reference's cross product = m / (64 f_x f_y)); qthr = kap^-2
};
__device__ __forceinline__ FwdCam fwd_cam(const CamRaw& r) {
    FwdCam c;
    c.nfX = -r.fx;
    c.nfY = -r.fy;
    c.cx = r.cx;
    c.cy = r.cy;
    c.inv_fx = 1.0f / r.fx;
    c.inv_fy = 1.0f / r.fy;
    c.kap = (1e12f / 64.0f) * c.inv_fx * c.inv_fy;
    c.qthr = 1.0f / (c.kap * c.kap);
    return c;
}

// m of the four pixels at window columns 1..4 (Z[r][c] = depth at row y - 1 + r, column x - 1 + c, replicate padding applied).
// GENERAL = false: x is a multiple of 4 and so is W, so only pixel 0 can sit on the left image border and only pixel 3 on
// the right one (wl = weight of pixel 0's left neighbour, wr = weight of pixel 3's right neighbour); GENERAL = true: any
// pixel may (wls[j], wrs[j]).  With all weights 1 both forms round identically.
// BORDER_ROW (first / last image row, a warp-uniform case the callers branch on): T and D take the row weights; elsewhere
// they are the plain column sums, which keeps the common path free of the extra live values.
template <typename V, bool GENERAL, bool BORDER_ROW>
__device__ __forceinline__ void scaled_normals4(const V (&Z)[3][6], V wm, V wp, const V (&wls)[4], const V (&wrs)[4],
                                                V nfX, V nfY, const V (&nfx0)[4], V nfy0, V (&m)[3][4]) {
    const V two = vdup<V>(2.0f);
    const V pfX = vneg(nfX), nwl0 = vneg(wls[0]);      // loop-invariant for the callers (hoisted)
    V S2[6], D2[6], T[6], D[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        const V t = vadd(Z[0][c], Z[2][c]);
        S2[c] = vfma(two, Z[1][c], t);
        D2[c] = vsub(Z[2][c], Z[0][c]);
        if constexpr (BORDER_ROW) {
            const V lo = vmul(wm, Z[0][c]);
            T[c] = vfma(wp, Z[2][c], lo);
            D[c] = vfma(wp, Z[2][c], vneg(lo));
        } else {
            T[c] = t;
            D[c] = D2[c];
        }
    }
    // (differences are formed in the sign each product needs: a packed negation is two LOP3 on the half-rate ALU pipe)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const V nG = vsub(S2[j], S2[j + 2]);
        const V Vz = vfma(two, D2[j + 1], vadd(D2[j], D2[j + 2]));
        const V B = vfma(two, T[j + 1], vadd(T[j], T[j + 2]));
        const V Cu = vsub(D[j + 2], D[j]), nCu = vsub(D[j], D[j + 2]);
        V A, Cv;
        if constexpr (GENERAL) {
            A = vfma(wrs[j], S2[j + 2], vmul(wls[j], S2[j]));
            Cv = vfma(wrs[j], D2[j + 2], vmul(vneg(wls[j]), D2[j]));
        } else if (j == 0) {
            A = vfma(wls[0], S2[0], S2[2]);
            Cv = vfma(nwl0, D2[0], D2[2]);
        } else if (j == 3) {
            A = vfma(wrs[3], S2[5], S2[3]);
            Cv = vfma(wrs[3], D2[5], vmul(vdup<V>(-1.0f), D2[3]));
        } else {
            A = vadd(S2[j + 2], S2[j]);
            Cv = BORDER_ROW ? vsub(D2[j + 2], D2[j]) : Cu;
        }
        const V mx = vmul(pfX, vfma(Vz, Cu, vmul(nG, B)));              //  f_x (V Cu - G B)
        const V my = vmul(nfY, vfma(nG, Cv, vmul(Vz, A)));              // -f_y (V A - G Cv)
        m[0][j] = mx;
        m[1][j] = my;
        m[2][j] = vfma(nfx0[j], mx, vfma(nfy0, my, vfma(nCu, Cv, vmul(A, B))));
    }
}

// cos(n_gt, n_pred) as the reference forms it -- F.normalize(eps 1e-12), then cosine_similarity(eps 1e-8) -- from the scaled
// cross products: ab = <m_a, m_b>, qa = |m_a|^2, qb = |m_b|^2.
__device__ __forceinline__ float cos_scaled(float ab, float qa, float qb, const FwdCam& fc) {
    const float ra = rsqrt_approx(qa), rb = rsqrt_approx(qb);
    float c = __fmul_rn(ab, __fmul_rn(ra, rb));
    if (fminf(qa, qb) < fc.qthr) {     // |n| < 1e-12 for one of the two (a zero-depth hole): a^ = n 1e12, |a^| < 1
        const float sa = qa > 0.0f ? ra : 0.0f, sb = qb > 0.0f ? rb : 0.0f;
        const float na = fminf(1.0f, __fmul_rn(__fmul_rn(qa, sa), fc.kap)), nb = fminf(1.0f, __fmul_rn(__fmul_rn(qb, sb), fc.kap));
        c = __fmul_rn(__fmul_rn(ab, __fmul_rn(sa, sb)), fminf(1.0f, __fmul_rn(__fmul_rn(na, nb), 1e8f)));
    }
    return c;
}

// L1 = true adds the supervised depth loss of the same block of the trainer (trainer.py:1246):
//     supervised_depth_loss = (|gt - pred| * mask).sum() / mask.sum()
// Scalar kernel: any width / alignment (the packed kernel below serves rows that are 16-byte multiples).
template <bool L1>
__global__ void __launch_bounds__(kLossThreads, 4) normals_loss_fwd_kernel(const LossParams p) {
    __shared__ __align__(16) float tg[box_rows(kFwdH)][kLBoxW];
    __shared__ __align__(16) float tp[box_rows(kFwdH)][kLBoxW];
    __shared__ double red[kLossThreads / 32][3];
    const int b = blockIdx.z, x0 = blockIdx.x * kLW, y0 = blockIdx.y * kFwdH;
    const size_t hw = (size_t)p.H * p.W;
    stage_tile<kFwdH>(tg, p.gt + b * hw, p.H, p.W, x0, y0);
    stage_tile<kFwdH>(tp, p.pred + b * hw, p.H, p.W, x0, y0);
    __syncthreads();
    const FwdCam fc = fwd_cam(cam_fetch(p.K, b));
    float s = 0.0f, m = 0.0f, l1 = 0.0f;   // at most 16 pixels per thread: float32 partials, float64 from the warp level on
    for (int i = threadIdx.x; i < (kLW / 4) * kFwdH; i += kLossThreads) {
        const int ty = i / (kLW / 4), tx0 = 4 * (i - ty * (kLW / 4));
        const int x = x0 + tx0, y = y0 + ty;
        if (x >= p.W || y >= p.H) continue;
        float wls[4], wrs[4], nfx0[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            wls[j] = (x + j == 0) ? 0.0f : 1.0f;
            wrs[j] = (x + j == p.W - 1) ? 0.0f : 1.0f;
            nfx0[j] = -(((float)(x + j) - fc.cx) * fc.inv_fx);
        }
        const bool border_row = (y == 0) | (y == p.H - 1);
        const float wm = y > 0 ? 1.0f : 0.0f, wp = y < p.H - 1 ? 1.0f : 0.0f, nfy0 = -(((float)y - fc.cy) * fc.inv_fy);
        float Zg[3][6], Zp[3][6], mg[3][4], mp[3][4];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                Zg[r][c] = tg[kLRow + ty + r - 1][kLCol + tx0 + c - 1];
                Zp[r][c] = tp[kLRow + ty + r - 1][kLCol + tx0 + c - 1];
            }
        if (border_row) {
            scaled_normals4<float, true, true>(Zg, wm, wp, wls, wrs, fc.nfX, fc.nfY, nfx0, nfy0, mg);
            scaled_normals4<float, true, true>(Zp, wm, wp, wls, wrs, fc.nfX, fc.nfY, nfx0, nfy0, mp);
        } else {
            scaled_normals4<float, true, false>(Zg, wm, wp, wls, wrs, fc.nfX, fc.nfY, nfx0, nfy0, mg);
            scaled_normals4<float, true, false>(Zp, wm, wp, wls, wrs, fc.nfX, fc.nfY, nfx0, nfy0, mp);
        }
        float fs = 0.0f, fm = 0.0f, fl = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (x + j < p.W) {
                const float zg = Zg[1][1 + j];
                const float mk = mask_value<L1>(p, b * hw + (size_t)y * p.W + x + j, zg);
                if constexpr (L1) fl = fmaf(fabsf(__fsub_rn(zg, Zp[1][1 + j])), mk, fl);
                const float qa = fmaf(mg[0][j], mg[0][j], fmaf(mg[1][j], mg[1][j], __fmul_rn(mg[2][j], mg[2][j])));
                const float qb = fmaf(mp[0][j], mp[0][j], fmaf(mp[1][j], mp[1][j], __fmul_rn(mp[2][j], mp[2][j])));
                const float ab = fmaf(mg[0][j], mp[0][j], fmaf(mg[1][j], mp[1][j], __fmul_rn(mg[2][j], mp[2][j])));
                const float c = cos_scaled(ab, qa, qb, fc);
                fs = fmaf(__fsub_rn(2.0f, c), mk, fs);
                fm += mk;
            }
        }
        s += fs;
        m += fm;
        l1 += fl;
    }
    loss_block_reduce<L1>(p, s, m, l1, red);
}

// ---------------------------------------------------------------------------------------------------------------
// Packed forward (rows that are 16-byte multiples).  The GT and the predicted depth go through the IDENTICAL stencil,
// so they ride in the two lanes of Blackwell's packed FP32 instructions (FFMA2 / FADD2 / FMUL2: the scalar FLOP rate
// in half the issue slots, tools/probes/int_pipe_probe.cu): every lane runs exactly the operation sequence of
// gradients4 / cross_rn / normalize3 above, so losses are bit-identical to the scalar kernel's.  The tile is staged
// INTERLEAVED -- shared element = (gt, pred) of one pixel -- so one LDS.128 delivers two ready-made lane pairs and a
// thread's 3 x 6 window of both fields costs 9 loads (18 in the planar layout).  Staging is by hand (coalesced
// 16-byte global loads, replicate padding by clamped coordinates); TMA cannot interleave two tensors.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kPW = 128, kPH = 32;                    // tile

// Shared rows of (gt, pred) pairs are swizzled in 16-byte units (two pairs): unit u lives at u ^ ((u >> 3) & 1).  A thread's
// window is three consecutive units starting at an even or odd unit 2 * lane (+1); without the swizzle lanes l and l + 4 of
// every quarter-warp hit the same banks (32-byte lane stride: two-way conflicts on every LDS.128, four-way on the 8-byte
// stores); with it the eight lanes of a quarter-warp cover the eight bank groups exactly once for each of the three loads.
__device__ __forceinline__ int swz_pair(int i) { return i ^ (((i >> 4) & 1) << 1); }
constexpr int kPPitch = kPW + 4;                      // (gt, pred) pairs per shared row: column 0 = image column x0 - 1, 129 = x0 + 128
constexpr size_t kFwdPairsTileBytes = (size_t)(kPH + 2) * kPPitch * sizeof(float2);
constexpr size_t kFwdPairsMaskBytes = (size_t)kPH * kPW * sizeof(float);

// Stage rows y0 - HALO .. y0 + TH + HALO - 1 of both fields, interleaved; tile column c lives at pair index c + OFF.
// All global loads of a thread are issued before the first shared store (the loop is fully unrolled), so a tile costs
// one memory round trip, not one per row.  Replicate padding by clamped coordinates.
template <int TH, int HALO, int OFF, int PITCH>
__device__ __forceinline__ void stage_pairs(float2 (*T)[PITCH], const float* __restrict__ G, const float* __restrict__ P, int H, int W,
                                            int x0, int y0) {
    constexpr int kRows = TH + 2 * HALO, kIter = (kRows + 7) / 8, kFull = kRows / 8;   // iterations whose row always exists
    const int q = threadIdx.x & 31, r0 = threadIdx.x >> 5;
    const int x = x0 + 4 * q;
    const bool in_x = x < W;                           // W % 4 == 0: a group is inside the image or outside it
    float4 g[kIter], d[kIter];
    if (in_x) {
#pragma unroll
        for (int k = 0; k < kIter; ++k) {
            const int r = r0 + 8 * k;
            if (k < kFull || r < kRows) {
                const unsigned off = (unsigned)min(max(y0 + r - HALO, 0), H - 1) * (unsigned)W + (unsigned)x;   // H * W < 2^31 (checked on the host)
                g[k] = __ldg(reinterpret_cast<const float4*>(G + off));
                d[k] = __ldg(reinterpret_cast<const float4*>(P + off));
            }
        }
    }
    // the HALO columns left of the tile and the first HALO columns right of it (or of the image)
    const int cr = min(W - x0, kPW);                  // tile-relative index of the first column beyond the tile / image
    float2 edge[(kRows * 2 * HALO + kLossThreads - 1) / kLossThreads];
#pragma unroll
    for (int k = 0; k < (kRows * 2 * HALO + kLossThreads - 1) / kLossThreads; ++k) {
        const int i = threadIdx.x + k * kLossThreads;
        if (i < kRows * 2 * HALO) {
            const int r = i / (2 * HALO), e = i - r * (2 * HALO);
            const int c = (e < HALO) ? e - HALO : cr + (e - HALO);
            const int yy = min(max(y0 + r - HALO, 0), H - 1);
            const int xx = min(max(x0 + c, 0), W - 1);
            edge[k] = make_float2(__ldg(G + (size_t)yy * W + xx), __ldg(P + (size_t)yy * W + xx));
        }
    }
    if (in_x) {
        static_assert(OFF % 2 == 1, "the four pairs of a group are: upper half of a unit, a whole unit, lower half of the next");
        const int pa = swz_pair(OFF + 4 * q), pb = swz_pair(OFF + 4 * q + 1), pc = swz_pair(OFF + 4 * q + 3);
#pragma unroll
        for (int k = 0; k < kIter; ++k) {
            const int r = r0 + 8 * k;
            if (k < kFull || r < kRows) {
                float2* row = &T[r][0];
                row[pa] = make_float2(g[k].x, d[k].x);
                *reinterpret_cast<float4*>(&row[pb]) = make_float4(g[k].y, d[k].y, g[k].z, d[k].z);
                row[pc] = make_float2(g[k].w, d[k].w);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < (kRows * 2 * HALO + kLossThreads - 1) / kLossThreads; ++k) {
        const int i = threadIdx.x + k * kLossThreads;
        if (i < kRows * 2 * HALO) {
            const int r = i / (2 * HALO), e = i - r * (2 * HALO);
            const int c = (e < HALO) ? e - HALO : cr + (e - HALO);
            T[r][swz_pair(OFF + c)] = edge[k];
        }
    }
}

// 8 x gradients of xyz of both fields for the four pixels starting at tile column tx0 of tile row ty (shared row ty + 1):
// lane 0 = GT, lane 1 = prediction.  Same arithmetic, per lane, as gradients4.
// `top` = the shared row of image row y - 1 (`pitch` pairs per row), `p0` = the (even) pair index of column x - 1 in it.
__device__ __forceinline__ void gradients4_pairs(const float2* top, int p0, int pitch, const f32x2 (&fx6)[6], const f32x2 (&fy3)[3],
                                                 f32x2 (&gu)[3][4], f32x2 (&gv)[3][4], f32x2 (&centre)[4]) {
    f32x2 Z[3][6];
    const int u0 = swz_pair(p0), u1 = swz_pair(p0 + 2), u2 = swz_pair(p0 + 4);    // the window's three 16-byte units
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float2* row = top + r * pitch;                                          // pairs of columns x - 1 .. x + 4
        const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(row + u0), b = *reinterpret_cast<const ulonglong2*>(row + u1),
                         c = *reinterpret_cast<const ulonglong2*>(row + u2);
        Z[r][0] = a.x; Z[r][1] = a.y; Z[r][2] = b.x; Z[r][3] = b.y; Z[r][4] = c.x; Z[r][5] = c.y;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) centre[j] = Z[1][1 + j];
    f32x2 S2[6], D2[6], S1[6], D1[6], P[6], Q[6];
    const f32x2 two = dup2(2.0f);
    const f32x2 fy1x2 = mul2(two, fy3[1]);
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        S2[c] = fma2(two, Z[1][c], add2(Z[0][c], Z[2][c]));
        D2[c] = sub2(Z[2][c], Z[0][c]);
        P[c] = mul2(fx6[c], S2[c]);
        Q[c] = mul2(fx6[c], D2[c]);
        const f32x2 lo = mul2(fy3[0], Z[0][c]);
        S1[c] = fma2(fy3[2], Z[2][c], fma2(fy1x2, Z[1][c], lo));
        D1[c] = sub2_unfused(mul2(fy3[2], Z[2][c]), lo);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        gu[0][j] = sub2_unfused(P[j + 2], P[j]);
        gv[0][j] = fma2(two, Q[j + 1], add2_unfused(Q[j], Q[j + 2]));
        gu[1][j] = sub2(S1[j + 2], S1[j]);
        gv[1][j] = fma2(two, D1[j + 1], add2(D1[j], D1[j + 2]));
        gu[2][j] = sub2(S2[j + 2], S2[j]);
        gv[2][j] = fma2(two, D2[j + 1], add2(D2[j], D2[j + 2]));
    }
}

// cross_rn + normalize3 on both lanes; returns the scale factors (inv) of both lanes.
__device__ __forceinline__ f32x2 unit_normals_pairs(const f32x2 (&u)[3], const f32x2 (&v)[3], f32x2 (&n)[3]) {
    f32x2 c[3];
    c[0] = sub2_unfused(mul2(u[1], v[2]), mul2(u[2], v[1]));
    c[1] = sub2_unfused(mul2(u[2], v[0]), mul2(u[0], v[2]));
    c[2] = sub2_unfused(mul2(u[0], v[1]), mul2(u[1], v[0]));
    float qa, qb;
    unpk2(fma2(c[0], c[0], fma2(c[1], c[1], mul2(c[2], c[2]))), qa, qb);
    const f32x2 inv = pk2(fminf(rsqrt_approx(qa), kInvCap), fminf(rsqrt_approx(qb), kInvCap));
    n[0] = mul2(c[0], inv);
    n[1] = mul2(c[1], inv);
    n[2] = mul2(c[2], inv);
    return inv;
}

template <bool L1>
__global__ void __launch_bounds__(kLossThreads, 3) normals_loss_fwd_pairs_kernel(const LossParams p) {
    extern __shared__ __align__(128) unsigned char fwd_smem[];          // the (gt, pred) tile, then the mask tile (plain variant)
    float2 (*T)[kPPitch] = reinterpret_cast<float2 (*)[kPPitch]>(fwd_smem);
    float (*M)[kPW] = reinterpret_cast<float (*)[kPW]>(fwd_smem + kFwdPairsTileBytes);
    __shared__ double red[kLossThreads / 32][3];
    const int b = blockIdx.z, x0 = blockIdx.x * kPW, y0 = blockIdx.y * kPH;
    const size_t hw = (size_t)p.H * p.W;
    const CamRaw cam_raw = cam_fetch(p.K, b);
    // the caller's mask of the tile rides along with the depth tile (requested before the tile's loads are consumed), so
    // no global load is left inside the arithmetic loop
    float4 mreg[kPH / 8];
    if constexpr (!L1) {
#pragma unroll
        for (int k = 0; k < kPH / 8; ++k) {
            const int yy = y0 + (int)(threadIdx.x >> 5) + 8 * k, xx = x0 + 4 * (int)(threadIdx.x & 31);
            mreg[k] = (yy < p.H && xx < p.W) ? __ldg(reinterpret_cast<const float4*>(p.mask + b * hw + (size_t)yy * p.W + xx))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    stage_pairs<kPH, 1, 1, kPPitch>(T, p.gt + b * hw, p.pred + b * hw, p.H, p.W, x0, y0);
    if constexpr (!L1) {
#pragma unroll
        for (int k = 0; k < kPH / 8; ++k) *reinterpret_cast<float4*>(&M[(threadIdx.x >> 5) + 8 * k][4 * (threadIdx.x & 31)]) = mreg[k];
    }
    __syncthreads();
    const FwdCam fc = fwd_cam(cam_raw);
    const int tx0 = 4 * (threadIdx.x & 31), x = x0 + tx0;
    const f32x2 one = dup2(1.0f);
    const f32x2 wls[4] = {dup2(x == 0 ? 0.0f : 1.0f), one, one, one};                 // W % 4 == 0: only pixel 0 / pixel 3 of a group
    const f32x2 wrs[4] = {one, one, one, dup2(x + 3 == p.W - 1 ? 0.0f : 1.0f)};       // can sit on the left / right image border
    const f32x2 nfX = dup2(fc.nfX), nfY = dup2(fc.nfY);
    f32x2 nfx0[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) nfx0[j] = dup2(-(((float)(x + j) - fc.cx) * fc.inv_fx));
    float s = 0.0f, m = 0.0f, l1 = 0.0f;   // at most 16 pixels per thread
    if (x < p.W) {
#pragma unroll 1
        for (int ty = threadIdx.x >> 5; ty < kPH; ty += kLossThreads / 32) {
            const int y = y0 + ty;
            if (y >= p.H) break;
            float mk4[4] = {0.f, 0.f, 0.f, 0.f};
            if constexpr (!L1) {
                const float4 mv = *reinterpret_cast<const float4*>(&M[ty][tx0]);
                mk4[0] = mv.x; mk4[1] = mv.y; mk4[2] = mv.z; mk4[3] = mv.w;
            }
            const bool border_row = (y == 0) | (y == p.H - 1);
            const f32x2 wm = dup2(y > 0 ? 1.0f : 0.0f), wp = dup2(y < p.H - 1 ? 1.0f : 0.0f);
            const f32x2 nfy0 = dup2(-(((float)y - fc.cy) * fc.inv_fy));
            f32x2 Z[3][6], mm[3][4];
            {   // the 3 x 6 window of (gt, pred) pairs: three swizzled 16-byte units per row (tile column c lives at pair index c + 1)
                const float2* top = &T[ty][0];
                const int u0 = swz_pair(tx0), u1 = swz_pair(tx0 + 2), u2 = swz_pair(tx0 + 4);
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float2* row = top + r * kPPitch;
                    const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(row + u0), bq = *reinterpret_cast<const ulonglong2*>(row + u1),
                                     c = *reinterpret_cast<const ulonglong2*>(row + u2);
                    Z[r][0] = a.x; Z[r][1] = a.y; Z[r][2] = bq.x; Z[r][3] = bq.y; Z[r][4] = c.x; Z[r][5] = c.y;
                }
            }
            if (border_row) scaled_normals4<f32x2, false, true>(Z, wm, wp, wls, wrs, nfX, nfY, nfx0, nfy0, mm);
            else scaled_normals4<f32x2, false, false>(Z, wm, wp, wls, wrs, nfX, nfY, nfx0, nfy0, mm);
            float fs = 0.0f, fm = 0.0f, fl = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float zg, zp;
                unpk2(Z[1][1 + j], zg, zp);
                float mk;
                if constexpr (L1) mk = (zg >= p.min_d && zg <= p.max_d) ? 1.0f : 0.0f;
                else mk = mk4[j];
                if constexpr (L1) fl = fmaf(fabsf(__fsub_rn(zg, zp)), mk, fl);
                float qa, qb, a[3], bb3[3];
                unpk2(fma2(mm[0][j], mm[0][j], fma2(mm[1][j], mm[1][j], mul2(mm[2][j], mm[2][j]))), qa, qb);
                unpk2(mm[0][j], a[0], bb3[0]);
                unpk2(mm[1][j], a[1], bb3[1]);
                unpk2(mm[2][j], a[2], bb3[2]);
                const float ab = fmaf(a[0], bb3[0], fmaf(a[1], bb3[1], __fmul_rn(a[2], bb3[2])));
                const float c = cos_scaled(ab, qa, qb, fc);
                fs = fmaf(__fsub_rn(2.0f, c), mk, fs);
                fm += mk;
            }
            s += fs;
            m += fm;
            l1 += fl;
        }
    }
    loss_block_reduce<L1>(p, s, m, l1, red);
}

// ---------------------------------------------------------------------------------------------------------------
// Packed backward (rows that are 16-byte multiples).  Phase 1 runs the two stencils of every pixel of the tile + ring
// in the two lanes of packed FP32 instructions (as the packed forward does) and stores the six adjoint fields as three
// PAIR fields A = (gu_bar_x, gu_bar_y), B = (gv_bar_x, gv_bar_y), C = (gu_bar_z, gv_bar_z); phase 2 gathers them with
// packed instructions too: the vertical combinations of A / B / C are independent chains per lane, and the horizontal
// chain  acc = fma(Dh, VU, fma(Sh, VV, acc))  of components x and y runs in the two lanes of one register pair.  Every
// lane performs exactly the operation sequence of the scalar kernel below: gradients are bit-identical to it.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kBPitch = 136;                          // pairs per shared row; tile column c lives at index c + 3 (c = -2 .. 129)
constexpr int kBTileRows = kBwdH + 4;                 // rows y0 - 2 .. y0 + kBwdH + 1
constexpr size_t kBwdPairsSmem = (size_t)(kBTileRows + 3 * kGH) * kBPitch * sizeof(float2);

// 3 x 3 window, one pixel (the two ring columns): same arithmetic as `gradients`, both fields in the two lanes.
__device__ __forceinline__ void gradients1_pairs(const float2* top, int p0, int pitch, int x, int y, int H, int W, const Cam& cam,
                                                 f32x2 (&gu)[3], f32x2 (&gv)[3], f32x2& centre) {
    f32x2 fx3[3], fy3[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        fx3[k] = dup2(((float)min(max(x + k - 1, 0), W - 1) - cam.cx) * cam.inv_fx);
        fy3[k] = dup2(((float)min(max(y + k - 1, 0), H - 1) - cam.cy) * cam.inv_fy);
    }
    f32x2 Z[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Z[r][c] = *reinterpret_cast<const f32x2*>(top + r * pitch + swz_pair(p0 + c));
    centre = Z[1][1];
    f32x2 S2[3], D2[3], S1[3], D1[3];
    const f32x2 two = dup2(2.0f);
    const f32x2 fy1x2 = mul2(two, fy3[1]);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        S2[c] = fma2(two, Z[1][c], add2(Z[0][c], Z[2][c]));
        D2[c] = sub2(Z[2][c], Z[0][c]);
        const f32x2 lo = mul2(fy3[0], Z[0][c]);
        S1[c] = fma2(fy3[2], Z[2][c], fma2(fy1x2, Z[1][c], lo));
        D1[c] = sub2_unfused(mul2(fy3[2], Z[2][c]), lo);
    }
    gu[0] = sub2_unfused(mul2(fx3[2], S2[2]), mul2(fx3[0], S2[0]));
    gv[0] = fma2(two, mul2(fx3[1], D2[1]), add2_unfused(mul2(fx3[0], D2[0]), mul2(fx3[2], D2[2])));
    gu[1] = sub2(S1[2], S1[0]);
    gv[1] = fma2(two, D1[1], add2(D1[0], D1[2]));
    gu[2] = sub2(S2[2], S2[0]);
    gv[2] = fma2(two, D2[1], add2(D2[0], D2[2]));
}

// adjoint_px from the packed gradients of one pixel (lane 0 = GT, lane 1 = prediction).
__device__ __forceinline__ void adjoint_pairs(const f32x2 (&gu)[3], const f32x2 (&gv)[3], float k, f32x2& A, f32x2& B, f32x2& C) {
    A = B = C = 0ull;
    if (k == 0.0f) return;
    f32x2 n[3];
    const f32x2 inv2 = unit_normals_pairs(gu, gv, n);
    float a[3], bn[3], up[3], vp[3], tmp, inv;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        unpk2(n[c], a[c], bn[c]);
        unpk2(gu[c], tmp, up[c]);
        unpk2(gv[c], tmp, vp[c]);
    }
    unpk2(inv2, tmp, inv);
    float inv_den, ab, bb;
    bool clamped;
    cosine_pairs(n, a, bn, inv_den, ab, bb, clamped);
    float gub[3], gvb[3];
    adjoint_tail(a, bn, inv, up, vp, k, inv_den, ab, bb, clamped, gub, gvb);
    A = pk2(gub[0], gub[1]);
    B = pk2(gvb[0], gvb[1]);
    C = pk2(gub[2], gvb[2]);
}

template <bool L1>
__global__ void __launch_bounds__(kLossThreads, 3) normals_loss_bwd_pairs_kernel(const LossParams p) {
    extern __shared__ __align__(128) unsigned char bwd_smem[];
    float2 (*T)[kBPitch] = reinterpret_cast<float2 (*)[kBPitch]>(bwd_smem);                                     // (gt, pred) tile, halo 2
    float2 (*G)[kGH][kBPitch] = reinterpret_cast<float2 (*)[kGH][kBPitch]>(bwd_smem + sizeof(float2) * kBTileRows * kBPitch);   // A, B, C
    const int b = blockIdx.z, x0 = blockIdx.x * kLW, y0 = blockIdx.y * kBwdH;
    const size_t hw = (size_t)p.H * p.W;
    const float* Gt = p.gt + b * hw;
    const float* Pr = p.pred + b * hw;
    const CamRaw cam_raw = cam_fetch(p.K, b);                      // requested before the tile, used after its barrier
    const float go = __ldg(p.grad_out), gl1 = (L1 && p.grad_l1) ? __ldg(p.grad_l1) : 0.0f;
    const double msum = __ldg(p.sums2 + 1);
    // ---- stage rows y0 - 2 .. y0 + kBwdH + 1, columns x0 - 2 .. x0 + 129 (replicate padding by clamping) ----
    stage_pairs<kBwdH, 2, 3, kBPitch>(T, Gt, Pr, p.H, p.W, x0, y0);
    __syncthreads();
    const Cam cam = cam_finish(cam_raw);
    const float scale = -go / (float)msum;    // -grad_out / sum(mask)
    const float l1_scale = (L1 && p.grad_l1) ? gl1 / (float)msum : 0.0f;

    // ---- phase 1: adjoints of the two gradients at every pixel of the tile and its 1-pixel ring ----
    {
        const int tx0 = 4 * (threadIdx.x & 31), x = x0 + tx0;
        f32x2 fx6[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) fx6[c] = dup2(((float)min(max(x + c - 1, 0), p.W - 1) - cam.cx) * cam.inv_fx);
#pragma unroll 1
        for (int gy = threadIdx.x >> 5; gy < kGH; gy += kLossThreads / 32) {
            const int ty = gy - 1, y = y0 + ty;
            f32x2 oa[4] = {0ull, 0ull, 0ull, 0ull}, ob[4] = {0ull, 0ull, 0ull, 0ull}, oc[4] = {0ull, 0ull, 0ull, 0ull};
            if (y >= 0 && y < p.H && x < p.W) {
                f32x2 fy3[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) fy3[r] = dup2(((float)min(max(y + r - 1, 0), p.H - 1) - cam.cy) * cam.inv_fy);
                f32x2 gu[3][4], gv[3][4], centre[4];
                gradients4_pairs(&T[ty + 1][0], tx0 + 2, kBPitch, fx6, fy3, gu, gv, centre);
                float mk4[4];
                if constexpr (!L1) {
                    const float4 mv = __ldg(reinterpret_cast<const float4*>(p.mask + b * hw + (size_t)y * p.W + x));
                    mk4[0] = mv.x; mk4[1] = mv.y; mk4[2] = mv.z; mk4[3] = mv.w;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float mk;
                    if constexpr (L1) {
                        float zg, zp;
                        unpk2(centre[j], zg, zp);
                        mk = (zg >= p.min_d && zg <= p.max_d) ? 1.0f : 0.0f;
                    } else {
                        mk = mk4[j];
                    }
                    const f32x2 u[3] = {gu[0][j], gu[1][j], gu[2][j]}, v[3] = {gv[0][j], gv[1][j], gv[2][j]};
                    adjoint_pairs(u, v, scale * mk, oa[j], ob[j], oc[j]);
                }
            }
            // pairs 3 + tx0 .. 6 + tx0 of the row: the upper half of a unit, a whole unit, the lower half of the next
            const int pa = swz_pair(3 + tx0), pb = swz_pair(4 + tx0), pc = swz_pair(6 + tx0);
#pragma unroll
            for (int f = 0; f < 3; ++f) {
                const f32x2* o = f == 0 ? oa : (f == 1 ? ob : oc);
                float2* row = &G[f][gy][0];
                *reinterpret_cast<f32x2*>(row + pa) = o[0];
                *reinterpret_cast<ulonglong2*>(row + pb) = make_ulonglong2(o[1], o[2]);
                *reinterpret_cast<f32x2*>(row + pc) = o[3];
            }
        }
    }
    for (int i = threadIdx.x; i < kGH * 2; i += kLossThreads) {     // the two ring columns, one pixel at a time
        const int gy = i >> 1, tx = (i & 1) ? kLW : -1;
        const int ty = gy - 1;
        const int x = x0 + tx, y = y0 + ty;
        f32x2 A = 0ull, B = 0ull, C = 0ull;
        if (x >= 0 && x < p.W && y >= 0 && y < p.H) {
            f32x2 gu[3], gv[3], centre;
            gradients1_pairs(&T[ty + 1][0], tx + 2, kBPitch, x, y, p.H, p.W, cam, gu, gv, centre);
            float zg, zp;
            unpk2(centre, zg, zp);
            const float mk = mask_value<L1>(p, b * hw + (size_t)y * p.W + x, zg);
            adjoint_pairs(gu, gv, scale * mk, A, B, C);
        }
        *reinterpret_cast<f32x2*>(&G[0][gy][swz_pair(3 + tx)]) = A;
        *reinterpret_cast<f32x2*>(&G[1][gy][swz_pair(3 + tx)]) = B;
        *reinterpret_cast<f32x2*>(&G[2][gy][swz_pair(3 + tx)]) = C;
    }
    __syncthreads();

    // ---- phase 2: gather, four pixels at a time ----
    {
        const int tx0 = 4 * (threadIdx.x & 31), x = x0 + tx0;
        if (x >= p.W) return;
#pragma unroll 1
        for (int ty = threadIdx.x >> 5; ty < kBwdH; ty += kLossThreads / 32) {
            const int y = y0 + ty;
            if (y >= p.H) break;
            const float Sv[3] = {1.0f, 2.0f + (y == 0) + (y == p.H - 1), 1.0f};
            const float Dv[3] = {1.0f, (float)(y == p.H - 1) - (float)(y == 0), -1.0f};
            f32x2 VA[6], VB[6], VC[6];      // vertical combinations: (VU_x, VU_y), (VV_x, VV_y), (VU_z, VV_z) of the six window columns
#pragma unroll
            for (int k6 = 0; k6 < 6; ++k6) VA[k6] = VB[k6] = VC[k6] = 0ull;
#pragma unroll
            for (int di = 0; di < 3; ++di) {
                const f32x2 ws = dup2(Sv[di]), wd = dup2(Dv[di]), wsd = pk2(Sv[di], Dv[di]);
                const float2* ra = &G[0][ty + di][0];      // columns x - 1 .. x + 4 = pairs tx0 + 2 .. tx0 + 7 = three swizzled units
                const float2* rb = &G[1][ty + di][0];
                const float2* rc = &G[2][ty + di][0];
#pragma unroll
                for (int h = 0; h < 3; ++h) {
                    const int u = swz_pair(tx0 + 2 + 2 * h);
                    const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(ra + u), bq = *reinterpret_cast<const ulonglong2*>(rb + u),
                                     c = *reinterpret_cast<const ulonglong2*>(rc + u);
                    VA[2 * h] = fma2(ws, a.x, VA[2 * h]);
                    VA[2 * h + 1] = fma2(ws, a.y, VA[2 * h + 1]);
                    VB[2 * h] = fma2(wd, bq.x, VB[2 * h]);
                    VB[2 * h + 1] = fma2(wd, bq.y, VB[2 * h + 1]);
                    VC[2 * h] = fma2(wsd, c.x, VC[2 * h]);
                    VC[2 * h + 1] = fma2(wsd, c.y, VC[2 * h + 1]);
                }
            }
            const float fyq = ((float)y - cam.cy) * cam.inv_fy;
            float res[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int xj = x + j;
                const float Sh[3] = {1.0f, 2.0f + (xj == 0) + (xj == p.W - 1), 1.0f};
                const float Dh[3] = {1.0f, (float)(xj == p.W - 1) - (float)(xj == 0), -1.0f};
                f32x2 acc01 = 0ull;
                float acc2 = 0.0f;
#pragma unroll
                for (int dj = 0; dj < 3; ++dj) {
                    acc01 = fma2(dup2(Dh[dj]), VA[j + dj], fma2(dup2(Sh[dj]), VB[j + dj], acc01));
                    float vu2, vv2;
                    unpk2(VC[j + dj], vu2, vv2);
                    acc2 = fmaf(Dh[dj], vu2, fmaf(Sh[dj], vv2, acc2));
                }
                float a0, a1;
                unpk2(acc01, a0, a1);
                const float fxq = ((float)xj - cam.cx) * cam.inv_fx;
                res[j] = fmaf(fxq, a0, fmaf(fyq, a1, acc2));
                if constexpr (L1) {
                    float zg, zp;
                    unpk2(*reinterpret_cast<const f32x2*>(&T[ty + 2][swz_pair(3 + tx0 + j)]), zg, zp);
                    const float mk = (zg >= p.min_d && zg <= p.max_d) ? 1.0f : 0.0f;
                    const float sgn = (zp > zg) ? 1.0f : ((zp < zg) ? -1.0f : 0.0f);
                    res[j] = fmaf(l1_scale * mk, sgn, res[j]);
                }
            }
            float* o = p.grad_pred + b * hw + (size_t)y * p.W + x;
            if ((reinterpret_cast<uintptr_t>(o) & 15) == 0) {
                *reinterpret_cast<float4*>(o) = make_float4(res[0], res[1], res[2], res[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = res[j];
            }
        }
    }
}

// 72 KB of shared memory per CTA: three CTAs per SM, so up to 85 registers per thread (no spills)
template <bool L1>
__global__ void __launch_bounds__(kLossThreads, 3) normals_loss_bwd_kernel(const LossParams p) {
    // dynamic shared memory: two halo'd depth tiles and the six adjoint fields of the tile + 1-pixel ring (46 KB)
    extern __shared__ __align__(128) unsigned char bwd_smem[];
    float (*tg)[kLBoxW] = reinterpret_cast<float (*)[kLBoxW]>(bwd_smem);
    float (*tp)[kLBoxW] = reinterpret_cast<float (*)[kLBoxW]>(bwd_smem + kLTilePad);
    float (*G)[kGH][kGPitch] = reinterpret_cast<float (*)[kGH][kGPitch]>(bwd_smem + 2 * kLTilePad);   // gu_bar xyz, gv_bar xyz
    const int b = blockIdx.z, x0 = blockIdx.x * kLW, y0 = blockIdx.y * kBwdH;
    const size_t hw = (size_t)p.H * p.W;
    stage_tile<kBwdH>(tg, p.gt + b * hw, p.H, p.W, x0, y0);
    stage_tile<kBwdH>(tp, p.pred + b * hw, p.H, p.W, x0, y0);
    __syncthreads();
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

int check_common(const float* gt, const float* pred, const float* K, const float* mask, int B, int H, int W, int tile_h, bool cap,
                 dim3& grid) {
    if (!gt || !pred || !K || !mask || B < 0 || H <= 0 || W <= 0 || B > 65535) return POLCUE_EINVAL;
    if ((reinterpret_cast<uintptr_t>(gt) | reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(K) |
         reinterpret_cast<uintptr_t>(mask)) & 3)
        return POLCUE_EINVAL;
    grid = dim3((W + kLW - 1) / kLW, (H + tile_h - 1) / tile_h, B);
    if (grid.y > 65535 || (unsigned long long)H * W >= (1ull << 31)) return POLCUE_E2BIG;   // 32-bit pixel offsets inside an image
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
    cudaStream_t s = (cudaStream_t)stream;
    // rows that are 16-byte multiples: the packed (GT, prediction) kernel; anything else: the scalar kernel with a hand-staged tile
    const bool pairs = W % 4 == 0 && ((reinterpret_cast<uintptr_t>(depth_gt) | reinterpret_cast<uintptr_t>(depth_pred) |
                                       (range_mask ? 0 : reinterpret_cast<uintptr_t>(mask))) & 15) == 0;
    static_assert(kPW == kLW && kPH == kFwdH, "both forward kernels must tile alike: one partial per CTA, same fold order");
    if (pairs) {
        auto kern = with_l1 ? normals_loss_fwd_pairs_kernel<true> : normals_loss_fwd_pairs_kernel<false>;
        const size_t smem = kFwdPairsTileBytes + (with_l1 ? 0 : kFwdPairsMaskBytes);
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<grid, kLossThreads, smem, s>>>(p);
    } else if (with_l1) {
        normals_loss_fwd_kernel<true><<<grid, kLossThreads, 0, s>>>(p);
    } else {
        normals_loss_fwd_kernel<false><<<grid, kLossThreads, 0, s>>>(p);
    }
    const int rc1 = launch_status();
    if (rc1 != POLCUE_OK) return rc1;
    loss_fold_kernel<<<1, kFoldThreads, 0, s>>>(p.partials, grid.x * grid.y * grid.z, with_l1 ? 1 : 0, sums, loss);
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
    cudaStream_t s = (cudaStream_t)stream;
    const bool pairs = W % 4 == 0 && ((reinterpret_cast<uintptr_t>(depth_gt) | reinterpret_cast<uintptr_t>(depth_pred) |
                                       (range_mask ? 0 : reinterpret_cast<uintptr_t>(mask))) & 15) == 0;
    if (pairs) {
        auto kern = with_l1 ? normals_loss_bwd_pairs_kernel<true> : normals_loss_bwd_pairs_kernel<false>;
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdPairsSmem);
        if (e != cudaSuccess) return (int)e;
        kern<<<grid, kLossThreads, kBwdPairsSmem, s>>>(p);
        return launch_status();
    }
    auto kern = with_l1 ? normals_loss_bwd_kernel<true> : normals_loss_bwd_kernel<false>;
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, kLossThreads, kBwdSmem, s>>>(p);
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
