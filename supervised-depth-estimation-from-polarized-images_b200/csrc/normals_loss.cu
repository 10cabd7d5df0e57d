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
// Both kernels evaluate the stencil in a CANCELLATION-FREE form (the derivation is at scaled_normals4 / adjoint_terms below):
// the gradients of P = (fx(u) Z, fy(v) Z, Z) are written through six small window functionals G, V, A, B, Cu, Cv of the
// depth, f_x f_y (gu x gv) is a handful of products of them, and the backward propagates to those six functionals
// (separable 3 x 3 tap patterns, gathered in a second phase through shared memory) instead of to the six gradient
// components -- half the FP operations of the reference's operation order (which stencil.cu, the public
// depth_to_normals, keeps) and none of its differences of nearly equal products.
//     k_p   = -grad_out * mask_p / sum(mask)
//     m_bar = k_p (g - <g,b> b) / |m|,   g = dcos/db = a/D - <a,b> b / (D |b|^2),  D = max(|a||b|, 1e-8),  b = m/|m| (clamps as the reference)
// Phase 1 of the backward kernel writes the six adjoints of the tile + 1-pixel ring to shared memory, phase 2 gathers.
// Two kernel families: packed (rows that are 16-byte multiples: GT and prediction in the two lanes of packed FP32
// instructions, interleaved shared tile) and scalar (any width / alignment); both stage their tiles by hand and run the
// same operation sequence per lane, so they agree bit for bit.
#include "polcue_device.cuh"
#include "polcue_host.h"

namespace polcue {
namespace {

constexpr int kLW = 128, kLossThreads = 256;
constexpr int kFwdH = 32, kBwdH = 14;                        // tile heights: the forward keeps the depth tile(s), the backward also six adjoint fields;
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

// The four intrinsics are REQUESTED at the top of a kernel (cam_fetch) and turned into reciprocals after the tile's barrier
// (fwd_cam): fetched after the barrier, every warp of the CTA sat out an L2 round trip there (7 % of the backward's stall samples).
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

// Column combinations of a 3 x N window (Z[r][c] = depth at row y - 1 + r, column first + c, replicate padding applied).
// BORDER_ROW (first / last image row, a warp-uniform case the callers branch on): T and D take the row weights; elsewhere
// they are the plain column sums (same bits as the weighted form with both weights 1), which keeps the common path short.
template <typename V, int N>
struct StencilCols {
    V S2[N], D2[N], T[N], D[N];
};
template <typename V, int N, bool BORDER_ROW>
__device__ __forceinline__ void stencil_columns(const V (&Z)[3][N], V wm, V wp, StencilCols<V, N>& s) {
    const V two = vdup<V>(2.0f);
#pragma unroll
    for (int c = 0; c < N; ++c) {
        const V t = vadd(Z[0][c], Z[2][c]);
        s.S2[c] = vfma(two, Z[1][c], t);
        s.D2[c] = vsub(Z[2][c], Z[0][c]);
        if constexpr (BORDER_ROW) {
            const V lo = vmul(wm, Z[0][c]);
            s.T[c] = vfma(wp, Z[2][c], lo);
            s.D[c] = vfma(wp, Z[2][c], vneg(lo));
        } else {
            s.T[c] = t;
            s.D[c] = s.D2[c];
        }
    }
}

// The six functionals and m of the pixel at window column j + 1.  `mode`: which of the pixel's horizontal neighbours may
// be a replicated image border (a compile-time constant at every call site): kWNone: neither; kWLeft / kWRight: wl / wr
// is its weight; kWBoth: both (the general form).  With weights 1 every form rounds identically.
// (Differences are formed in the sign each product needs: a packed negation is two LOP3 on the half-rate ALU pipe.)
enum { kWNone = 0, kWLeft = 1, kWRight = 2, kWBoth = 3 };
template <typename V>
struct PixelTerms {
    V nG, Vz, A, B, Cu, Cv;    // -G, V, A, B, Cu, Cv of the file comment above
    V m[3];
};
template <typename V, bool BORDER_ROW, int N>
__device__ __forceinline__ PixelTerms<V> pixel_terms(const StencilCols<V, N>& s, int j, int mode, V wl, V nwl, V wr, V pfX, V nfY,
                                                     V nfx0, V nfy0) {
    const V two = vdup<V>(2.0f);
    PixelTerms<V> t;
    t.nG = vsub(s.S2[j], s.S2[j + 2]);
    t.Vz = vfma(two, s.D2[j + 1], vadd(s.D2[j], s.D2[j + 2]));
    t.B = vfma(two, s.T[j + 1], vadd(s.T[j], s.T[j + 2]));
    t.Cu = vsub(s.D[j + 2], s.D[j]);
    const V nCu = vsub(s.D[j], s.D[j + 2]);
    if (mode == kWBoth) {
        t.A = vfma(wr, s.S2[j + 2], vmul(wl, s.S2[j]));
        t.Cv = vfma(wr, s.D2[j + 2], vmul(nwl, s.D2[j]));
    } else if (mode == kWLeft) {
        t.A = vfma(wl, s.S2[j], s.S2[j + 2]);
        t.Cv = vfma(nwl, s.D2[j], s.D2[j + 2]);
    } else if (mode == kWRight) {
        t.A = vfma(wr, s.S2[j + 2], s.S2[j]);
        t.Cv = vfma(wr, s.D2[j + 2], vmul(vdup<V>(-1.0f), s.D2[j]));
    } else {
        t.A = vadd(s.S2[j + 2], s.S2[j]);
        t.Cv = BORDER_ROW ? vsub(s.D2[j + 2], s.D2[j]) : t.Cu;
    }
    t.m[0] = vmul(pfX, vfma(t.Vz, t.Cu, vmul(t.nG, t.B)));              //  f_x (V Cu - G B)
    t.m[1] = vmul(nfY, vfma(t.nG, t.Cv, vmul(t.Vz, t.A)));              // -f_y (V A - G Cv)
    t.m[2] = vfma(nfx0, t.m[0], vfma(nfy0, t.m[1], vfma(nCu, t.Cv, vmul(t.A, t.B))));
    return t;
}

// m of the four pixels at window columns 1..4 of a 3 x 6 window.  GENERAL = false: x is a multiple of 4 and so is W, so
// only pixel 0 can sit on the left image border and only pixel 3 on the right one (wls[0], wrs[3]); GENERAL = true: any
// pixel may (wls[j], wrs[j]).
template <typename V, bool GENERAL, bool BORDER_ROW>
__device__ __forceinline__ void scaled_normals4(const V (&Z)[3][6], V wm, V wp, const V (&wls)[4], const V (&wrs)[4],
                                                V nfX, V nfY, const V (&nfx0)[4], V nfy0, V (&m)[3][4]) {
    const V pfX = vneg(nfX), nwl0 = vneg(wls[0]);      // loop-invariant for the callers (hoisted)
    StencilCols<V, 6> s;
    stencil_columns<V, 6, BORDER_ROW>(Z, wm, wp, s);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int mode = GENERAL ? kWBoth : (j == 0 ? kWLeft : (j == 3 ? kWRight : kWNone));
        const PixelTerms<V> t = pixel_terms<V, BORDER_ROW, 6>(s, j, mode, wls[j], GENERAL ? vneg(wls[j]) : nwl0, wrs[j], pfX, nfY, nfx0[j], nfy0);
        m[0][j] = t.m[0];
        m[1][j] = t.m[1];
        m[2][j] = t.m[2];
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

// BACKWARD in the same form.  With m = (f_x (V Cu - G B), f_y (G Cv - V A), -fx m_x - fy m_y + A B - Cu Cv) and m_bar = dL/dm
// (from the cosine and the two clamped normalisations, exactly as the reference's autograd forms it), the adjoints of the
// six window functionals of the PREDICTED depth at one pixel are, with px = f_x (m_bar_x - fx m_bar_z), py = f_y (m_bar_y - fy m_bar_z):
//     G_bar = py Cv - px B    V_bar = px Cu - py A    A_bar = m_bar_z B - py V
//     B_bar = m_bar_z A - px G    Cu_bar = px V - m_bar_z Cv    Cv_bar = py G - m_bar_z Cu
// and each functional is a separable 3 x 3 tap pattern of Z: rows x columns = G: s x d, V: d x s, A: s x e, B: e x s,
// Cu: d' x d, Cv: d x d', with s = (1, 2, 1), d = (-1, 0, 1) (replicate padding folds their outer taps onto border
// pixels) and the weighted e = (w-, 0, w+), d' = (-w-, 0, w+), whose weights cancel the folded taps exactly: in gather
// form, for the output pixel q and F = 0 outside the image,
//     s: F(q-1) + F(q+1) + (2 + [q first] + [q last]) F(q)      d: F(q-1) - F(q+1) + ([q last] - [q first]) F(q)
//     e: F(q-1) + F(q+1)                                        d': F(q-1) - F(q+1).
// Phase 1 stores the six adjoints of every pixel of the tile + 1-pixel ring, phase 2 gathers: rows first, then
//     Zbar(q) = U(q-1) + W(q+1) + Dh0 Xd(q) + Sh0 Xs(q),   Xd = Rs(G_bar) + Rd'(Cu_bar),  Xs = Rd(V_bar) + Re(B_bar),
//     P1 = Xd + Rd(Cv_bar),  P2 = Xs + Rs(A_bar),  U = P1 + P2,  W = P2 - P1         (R = the row combinations above).
// (Checked against float64 autograd of the reference formulation: tests, and tools/probes/bwd_terms_proto.py.)
struct Adj6 {
    float G, A, V, Cv, B, Cu;
};
__device__ __forceinline__ Adj6 adjoint_terms(const float (&mg)[3], const float (&mp)[3], float qg, float qp, float nG, float Vz, float A,
                                              float B, float Cu, float Cv, float k, float fx0, float fy0, const FwdCam& fc) {
    Adj6 o{0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    if (k == 0.0f) return o;
    // a = n_gt / max(|n_gt|, 1e-12), b = n_pred / max(|n_pred|, 1e-12) in terms of m = 64 f_x f_y n
    const float inv_a = fminf(rsqrt_approx(qg), fc.kap), inv_b = fminf(rsqrt_approx(qp), fc.kap);
    float a[3], bn[3], mb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        a[c] = __fmul_rn(mg[c], inv_a);
        bn[c] = __fmul_rn(mp[c], inv_b);
    }
    const float ki = __fmul_rn(k, inv_b);
    if (fminf(qg, qp) >= fc.qthr) {
        // neither normalisation is capped: a and b are unit vectors, the eps clamp of the cosine is far away, and
        // m_bar = k (a - <a,b> b) / |m|
        const float c = fmaf(a[0], bn[0], fmaf(a[1], bn[1], __fmul_rn(a[2], bn[2])));
#pragma unroll
        for (int i = 0; i < 3; ++i) mb[i] = __fmul_rn(ki, fmaf(-c, bn[i], a[i]));
    } else {
        // a zero-depth hole under one of the two windows: the reference's clamps decide.  g = dcos/db: above the eps clamp
        // a/D - <a,b> b / (D |b|^2), inside it a / eps;  b = m inv_b: the capped normalisation is linear, otherwise project
        // out b and divide by |m|
        float inv_den, ab, bb;
        bool clamped;
        cosine(a, bn, inv_den, ab, bb, clamped);
        const float w_b = clamped ? 0.0f : __fmul_rn(__fmul_rn(ab, inv_den), rcp_approx(fmaxf(bb, 1e-30f)));
        float g[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) g[c] = fmaf(-w_b, bn[c], __fmul_rn(a[c], inv_den));
        const float gb = (inv_b >= fc.kap) ? 0.0f : fmaf(g[0], bn[0], fmaf(g[1], bn[1], __fmul_rn(g[2], bn[2])));
#pragma unroll
        for (int c = 0; c < 3; ++c) mb[c] = __fmul_rn(ki, fmaf(-gb, bn[c], g[c]));
    }
    const float mz = mb[2];
    const float px = __fmul_rn(-fc.nfX, fmaf(-fx0, mz, mb[0]));
    const float py = __fmul_rn(-fc.nfY, fmaf(-fy0, mz, mb[1]));
    const float G = -nG;
    o.G = fmaf(py, Cv, -__fmul_rn(px, B));
    o.V = fmaf(px, Cu, -__fmul_rn(py, A));
    o.A = fmaf(mz, B, -__fmul_rn(py, Vz));
    o.B = fmaf(mz, A, -__fmul_rn(px, G));
    o.Cu = fmaf(px, Vz, -__fmul_rn(mz, Cv));
    o.Cv = fmaf(py, G, -__fmul_rn(mz, Cu));
    return o;
}

// Row combinations of one window column -> the four per-column values of the gather (see above).
__device__ __forceinline__ void gather_column(float RsG, float RsA, float RdV, float RdCv, float ReB, float RdpCu, float& U, float& Wq,
                                              float& Xd, float& Xs) {
    Xd = __fadd_rn(RsG, RdpCu);
    Xs = __fadd_rn(RdV, ReB);
    const float p1 = __fadd_rn(Xd, RdCv), p2 = __fadd_rn(Xs, RsA);
    U = __fadd_rn(p1, p2);
    Wq = __fsub_rn(p2, p1);
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
// in half the issue slots, tools/probes/int_pipe_probe.cu): every lane runs exactly the operation sequence of the
// scalar kernel (the same templated core), so losses are bit-identical to the scalar kernel's.  The tile is staged
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
// in the two lanes of packed FP32 instructions (as the packed forward does), forms the six adjoints of the prediction's
// functionals per pixel (adjoint_terms, scalar: only one lane needs it) and stores them as three PAIR fields paired by
// their row taps, PA = (G_bar, A_bar), PB = (V_bar, Cv_bar), PC = (B_bar, Cu_bar); phase 2 forms the row combinations of
// both members of a pair in packed instructions, then four values per window column (gather_column) and the output.
// Every lane performs exactly the operation sequence of the scalar kernel below: gradients are bit-identical to it.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kBPitch = 136;                          // pairs per shared row; tile column c lives at index c + 3 (c = -2 .. 129)
constexpr int kBTileRows = kBwdH + 4;                 // rows y0 - 2 .. y0 + kBwdH + 1
constexpr size_t kBwdPairsSmem = (size_t)(kBTileRows + 3 * kGH) * kBPitch * sizeof(float2);

// Phase-1 work of one 4-pixel group: the (gt, pred) window `top` (shared row of image row y - 1; `p0` = even pair index of
// column x - 1) -> the adjoints of the six functionals at the four pixels as three pair fields
// PA = (G_bar, A_bar), PB = (V_bar, Cv_bar), PC = (B_bar, Cu_bar)   (paired by their ROW taps: s, d, and e / d').
struct BwdConsts {            // per-thread constants of phase 1, kept as scalars (duplicated into lane pairs at their use:
    float wl0, wr3, pfX, nfY; // registers are what limits this kernel to three CTAs per SM)
    float nfx0[4];
};
template <bool L1, bool BORDER_ROW>
__device__ __forceinline__ void bwd_group_pairs(const LossParams& p, const FwdCam& fc, const float2* top, int p0, int pitch, f32x2 wm,
                                                f32x2 wp, const BwdConsts& k, f32x2 nfy0, float scale, const float (&mk4)[4],
                                                float2* ga, float2* gb, float2* gc, int g0) {    // G rows of the three pair fields; g0: pair index of pixel 0
    f32x2 Z[3][6];
    const int u0 = swz_pair(p0), u1 = swz_pair(p0 + 2), u2 = swz_pair(p0 + 4);    // the window's three 16-byte units
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float2* row = top + r * pitch;
        const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(row + u0), bq = *reinterpret_cast<const ulonglong2*>(row + u1),
                         c = *reinterpret_cast<const ulonglong2*>(row + u2);
        Z[r][0] = a.x; Z[r][1] = a.y; Z[r][2] = bq.x; Z[r][3] = bq.y; Z[r][4] = c.x; Z[r][5] = c.y;
    }
    StencilCols<f32x2, 6> s;
    stencil_columns<f32x2, 6, BORDER_ROW>(Z, wm, wp, s);
    float fy0, dummy;
    unpk2(nfy0, dummy, fy0);
    fy0 = -fy0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int mode = j == 0 ? kWLeft : (j == 3 ? kWRight : kWNone);
        const PixelTerms<f32x2> t = pixel_terms<f32x2, BORDER_ROW, 6>(s, j, mode, dup2(k.wl0), dup2(-k.wl0), dup2(k.wr3), dup2(k.pfX), dup2(k.nfY),
                                                                      dup2(k.nfx0[j]), nfy0);
        float qg, qp, mg[3], mp[3], nG, Vz, A, B, Cu, Cv, zg, zp, nfx;
        unpk2(fma2(t.m[0], t.m[0], fma2(t.m[1], t.m[1], mul2(t.m[2], t.m[2]))), qg, qp);
#pragma unroll
        for (int c = 0; c < 3; ++c) unpk2(t.m[c], mg[c], mp[c]);
        unpk2(t.nG, dummy, nG);
        unpk2(t.Vz, dummy, Vz);
        unpk2(t.A, dummy, A);
        unpk2(t.B, dummy, B);
        unpk2(t.Cu, dummy, Cu);
        unpk2(t.Cv, dummy, Cv);
        unpk2(Z[1][1 + j], zg, zp);
        nfx = k.nfx0[j];
        float mk;
        if constexpr (L1) mk = (zg >= p.min_d && zg <= p.max_d) ? 1.0f : 0.0f;
        else mk = mk4[j];
        const Adj6 o = adjoint_terms(mg, mp, qg, qp, nG, Vz, A, B, Cu, Cv, scale * mk, -nfx, fy0, fc);
        const int gi = swz_pair(g0 + j);      // stored pixel by pixel: three 4-pixel output arrays would not fit the register budget
        *reinterpret_cast<f32x2*>(ga + gi) = pk2(o.G, o.A);
        *reinterpret_cast<f32x2*>(gb + gi) = pk2(o.V, o.Cv);
        *reinterpret_cast<f32x2*>(gc + gi) = pk2(o.B, o.Cu);
    }
}

template <bool L1>
__global__ void __launch_bounds__(kLossThreads, 3) normals_loss_bwd_pairs_kernel(const LossParams p) {
    extern __shared__ __align__(128) unsigned char bwd_smem[];
    float2 (*T)[kBPitch] = reinterpret_cast<float2 (*)[kBPitch]>(bwd_smem);                                     // (gt, pred) tile, halo 2
    float2 (*G)[kGH][kBPitch] = reinterpret_cast<float2 (*)[kGH][kBPitch]>(bwd_smem + sizeof(float2) * kBTileRows * kBPitch);   // PA, PB, PC
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
    const FwdCam fc = fwd_cam(cam_raw);
    const float scale = -go / (float)msum;    // -grad_out / sum(mask)
    const float l1_scale = (L1 && p.grad_l1) ? gl1 / (float)msum : 0.0f;

    // ---- phase 1: adjoints of the six functionals at every pixel of the tile and its 1-pixel ring ----
    {
        const int tx0 = 4 * (threadIdx.x & 31), x = x0 + tx0;
        BwdConsts kc;
        kc.wl0 = x == 0 ? 0.0f : 1.0f;
        kc.wr3 = x + 3 == p.W - 1 ? 0.0f : 1.0f;
        kc.pfX = -fc.nfX;
        kc.nfY = fc.nfY;
#pragma unroll
        for (int j = 0; j < 4; ++j) kc.nfx0[j] = -(((float)(x + j) - fc.cx) * fc.inv_fx);
#pragma unroll 1
        for (int gy = threadIdx.x >> 5; gy < kGH; gy += kLossThreads / 32) {
            const int ty = gy - 1, y = y0 + ty;
            float2 *ga = &G[0][gy][0], *gb = &G[1][gy][0], *gc = &G[2][gy][0];
            if (y >= 0 && y < p.H && x < p.W) {
                float mk4[4] = {0.f, 0.f, 0.f, 0.f};
                if constexpr (!L1) {
                    const float4 mv = __ldg(reinterpret_cast<const float4*>(p.mask + b * hw + (size_t)y * p.W + x));
                    mk4[0] = mv.x; mk4[1] = mv.y; mk4[2] = mv.z; mk4[3] = mv.w;
                }
                const f32x2 wm = dup2(y > 0 ? 1.0f : 0.0f), wp = dup2(y < p.H - 1 ? 1.0f : 0.0f);
                const f32x2 nfy0 = dup2(-(((float)y - fc.cy) * fc.inv_fy));
                if ((y == 0) | (y == p.H - 1)) bwd_group_pairs<L1, true>(p, fc, &T[ty + 1][0], tx0 + 2, kBPitch, wm, wp, kc, nfy0, scale, mk4, ga, gb, gc, 3 + tx0);
                else bwd_group_pairs<L1, false>(p, fc, &T[ty + 1][0], tx0 + 2, kBPitch, wm, wp, kc, nfy0, scale, mk4, ga, gb, gc, 3 + tx0);
            } else {      // outside the image: no contribution (pairs 3 + tx0 .. 6 + tx0: half a unit, a whole unit, half of the next)
                const int pa = swz_pair(3 + tx0), pb = swz_pair(4 + tx0), pc = swz_pair(6 + tx0);
#pragma unroll
                for (int f = 0; f < 3; ++f) {
                    float2* row = f == 0 ? ga : (f == 1 ? gb : gc);
                    *reinterpret_cast<f32x2*>(row + pa) = 0ull;
                    *reinterpret_cast<ulonglong2*>(row + pb) = make_ulonglong2(0ull, 0ull);
                    *reinterpret_cast<f32x2*>(row + pc) = 0ull;
                }
            }
        }
    }
    for (int i = threadIdx.x; i < kGH * 2; i += kLossThreads) {     // the two ring columns, one pixel at a time (general weights)
        const int gy = i >> 1, tx = (i & 1) ? kLW : -1;
        const int ty = gy - 1;
        const int x = x0 + tx, y = y0 + ty;
        f32x2 A2 = 0ull, B2 = 0ull, C2 = 0ull;
        if (x >= 0 && x < p.W && y >= 0 && y < p.H) {
            f32x2 Z[3][3];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) Z[r][c] = *reinterpret_cast<const f32x2*>(&T[ty + 1 + r][swz_pair(tx + 2 + c)]);
            StencilCols<f32x2, 3> s;
            stencil_columns<f32x2, 3, true>(Z, dup2(y > 0 ? 1.0f : 0.0f), dup2(y < p.H - 1 ? 1.0f : 0.0f), s);
            const float fx0 = ((float)x - fc.cx) * fc.inv_fx, fy0 = ((float)y - fc.cy) * fc.inv_fy;
            const float wl = x == 0 ? 0.0f : 1.0f, wr = x == p.W - 1 ? 0.0f : 1.0f;
            const PixelTerms<f32x2> t = pixel_terms<f32x2, true, 3>(s, 0, kWBoth, dup2(wl), dup2(-wl), dup2(wr), dup2(-fc.nfX), dup2(fc.nfY),
                                                                    dup2(-fx0), dup2(-fy0));
            float qg, qp, mg[3], mp[3], nG, Vz, A, B, Cu, Cv, zg, zp, dummy;
            unpk2(fma2(t.m[0], t.m[0], fma2(t.m[1], t.m[1], mul2(t.m[2], t.m[2]))), qg, qp);
#pragma unroll
            for (int c = 0; c < 3; ++c) unpk2(t.m[c], mg[c], mp[c]);
            unpk2(t.nG, dummy, nG);
            unpk2(t.Vz, dummy, Vz);
            unpk2(t.A, dummy, A);
            unpk2(t.B, dummy, B);
            unpk2(t.Cu, dummy, Cu);
            unpk2(t.Cv, dummy, Cv);
            unpk2(Z[1][1], zg, zp);
            const float mk = mask_value<L1>(p, b * hw + (size_t)y * p.W + x, zg);
            const Adj6 o = adjoint_terms(mg, mp, qg, qp, nG, Vz, A, B, Cu, Cv, scale * mk, fx0, fy0, fc);
            A2 = pk2(o.G, o.A);
            B2 = pk2(o.V, o.Cv);
            C2 = pk2(o.B, o.Cu);
        }
        *reinterpret_cast<f32x2*>(&G[0][gy][swz_pair(3 + tx)]) = A2;
        *reinterpret_cast<f32x2*>(&G[1][gy][swz_pair(3 + tx)]) = B2;
        *reinterpret_cast<f32x2*>(&G[2][gy][swz_pair(3 + tx)]) = C2;
    }
    __syncthreads();

    // ---- phase 2: gather, four pixels at a time: rows first (packed: both members of a pair field share their row taps) ----
    {
        const int tx0 = 4 * (threadIdx.x & 31), x = x0 + tx0;
        if (x >= p.W) return;
        const f32x2 pm = pk2(1.0f, -1.0f);
#pragma unroll 1
        for (int ty = threadIdx.x >> 5; ty < kBwdH; ty += kLossThreads / 32) {
            const int y = y0 + ty;
            if (y >= p.H) break;
            const f32x2 sv = dup2(2.0f + (y == 0) + (y == p.H - 1)), dv = dup2((float)(y == p.H - 1) - (float)(y == 0));
            float U[6], Wq[6], Xd[6], Xs[6];
#pragma unroll
            for (int h = 0; h < 3; ++h) {      // columns x - 1 .. x + 4 = pairs tx0 + 2 .. tx0 + 7 = three swizzled units
                const int u = swz_pair(tx0 + 2 + 2 * h);
                ulonglong2 w[3][3];            // [field][row above / at / below]
#pragma unroll
                for (int f = 0; f < 3; ++f)
#pragma unroll
                    for (int di = 0; di < 3; ++di) w[f][di] = *reinterpret_cast<const ulonglong2*>(&G[f][ty + di][u]);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const f32x2 au = e ? w[0][0].y : w[0][0].x, am = e ? w[0][1].y : w[0][1].x, ad = e ? w[0][2].y : w[0][2].x;
                    const f32x2 bu = e ? w[1][0].y : w[1][0].x, bm = e ? w[1][1].y : w[1][1].x, bd = e ? w[1][2].y : w[1][2].x;
                    const f32x2 cu = e ? w[2][0].y : w[2][0].x, cd = e ? w[2][2].y : w[2][2].x;
                    float RsG, RsA, RdV, RdCv, ReB, RdpCu;
                    unpk2(fma2(sv, am, add2(au, ad)), RsG, RsA);
                    unpk2(fma2(dv, bm, sub2(bu, bd)), RdV, RdCv);
                    unpk2(fma2(cd, pm, cu), ReB, RdpCu);
                    gather_column(RsG, RsA, RdV, RdCv, ReB, RdpCu, U[2 * h + e], Wq[2 * h + e], Xd[2 * h + e], Xs[2 * h + e]);
                }
            }
            float res[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int xj = x + j;
                const float sh0 = 2.0f + (xj == 0) + (xj == p.W - 1), dh0 = (float)(xj == p.W - 1) - (float)(xj == 0);
                res[j] = fmaf(sh0, Xs[j + 1], fmaf(dh0, Xd[j + 1], __fadd_rn(U[j], Wq[j + 2])));
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

// Scalar twin (any width / alignment): the same operation sequence per pixel as the packed kernel (general border weights,
// which round identically), planar depth tiles and six planar adjoint fields 0: G_bar, 1: A_bar, 2: V_bar, 3: Cv_bar,
// 4: B_bar, 5: Cu_bar.  72 KB of shared memory per CTA: three CTAs per SM.
__device__ __forceinline__ Adj6 bwd_pixel_scalar(const LossParams& p, const FwdCam& fc, const float (*tg)[kLBoxW], const float (*tp)[kLBoxW],
                                                 int ty, int tx, int x, int y, float k) {
    float Zg[3][3], Zp[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            Zg[r][c] = tg[kLRow + ty + r - 1][kLCol + tx + c - 1];
            Zp[r][c] = tp[kLRow + ty + r - 1][kLCol + tx + c - 1];
        }
    const float wm = y > 0 ? 1.0f : 0.0f, wp = y < p.H - 1 ? 1.0f : 0.0f;
    const float wl = x == 0 ? 0.0f : 1.0f, wr = x == p.W - 1 ? 0.0f : 1.0f;
    const float fx0 = ((float)x - fc.cx) * fc.inv_fx, fy0 = ((float)y - fc.cy) * fc.inv_fy;
    StencilCols<float, 3> sg, sp;
    stencil_columns<float, 3, true>(Zg, wm, wp, sg);
    stencil_columns<float, 3, true>(Zp, wm, wp, sp);
    const PixelTerms<float> g = pixel_terms<float, true, 3>(sg, 0, kWBoth, wl, -wl, wr, -fc.nfX, fc.nfY, -fx0, -fy0);
    const PixelTerms<float> t = pixel_terms<float, true, 3>(sp, 0, kWBoth, wl, -wl, wr, -fc.nfX, fc.nfY, -fx0, -fy0);
    const float qg = fmaf(g.m[0], g.m[0], fmaf(g.m[1], g.m[1], __fmul_rn(g.m[2], g.m[2])));
    const float qp = fmaf(t.m[0], t.m[0], fmaf(t.m[1], t.m[1], __fmul_rn(t.m[2], t.m[2])));
    const float mg[3] = {g.m[0], g.m[1], g.m[2]}, mp[3] = {t.m[0], t.m[1], t.m[2]};
    return adjoint_terms(mg, mp, qg, qp, t.nG, t.Vz, t.A, t.B, t.Cu, t.Cv, k, fx0, fy0, fc);
}

template <bool L1>
__global__ void __launch_bounds__(kLossThreads, 3) normals_loss_bwd_kernel(const LossParams p) {
    extern __shared__ __align__(128) unsigned char bwd_smem[];
    float (*tg)[kLBoxW] = reinterpret_cast<float (*)[kLBoxW]>(bwd_smem);
    float (*tp)[kLBoxW] = reinterpret_cast<float (*)[kLBoxW]>(bwd_smem + kLTilePad);
    float (*G)[kGH][kGPitch] = reinterpret_cast<float (*)[kGH][kGPitch]>(bwd_smem + 2 * kLTilePad);   // the six adjoint fields
    const int b = blockIdx.z, x0 = blockIdx.x * kLW, y0 = blockIdx.y * kBwdH;
    const size_t hw = (size_t)p.H * p.W;
    stage_tile<kBwdH>(tg, p.gt + b * hw, p.H, p.W, x0, y0);
    stage_tile<kBwdH>(tp, p.pred + b * hw, p.H, p.W, x0, y0);
    __syncthreads();
    const FwdCam fc = fwd_cam(cam_fetch(p.K, b));
    const float scale = -__ldg(p.grad_out) / (float)p.sums2[1];    // -grad_out / sum(mask)
    const float l1_scale = (L1 && p.grad_l1) ? __ldg(p.grad_l1) / (float)p.sums2[1] : 0.0f;

    // phase 1: the six adjoints at every pixel of the tile and its 1-pixel ring (G column kGCol + tx, tx = -1 .. kLW)
    for (int i = threadIdx.x; i < kGH * (kLW + 2); i += kLossThreads) {
        const int gy = i / (kLW + 2), tx = i - gy * (kLW + 2) - 1;
        const int ty = gy - 1, x = x0 + tx, y = y0 + ty;
        Adj6 o{0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        if (x >= 0 && x < p.W && y >= 0 && y < p.H) {
            const float k = scale * mask_value<L1>(p, b * hw + (size_t)y * p.W + x, tg[kLRow + ty][kLCol + tx]);
            o = bwd_pixel_scalar(p, fc, tg, tp, ty, tx, x, y, k);
        }
        const int gx = kGCol + tx;
        G[0][gy][gx] = o.G;
        G[1][gy][gx] = o.A;
        G[2][gy][gx] = o.V;
        G[3][gy][gx] = o.Cv;
        G[4][gy][gx] = o.B;
        G[5][gy][gx] = o.Cu;
    }
    __syncthreads();

    // phase 2: gather, one pixel per thread and step: rows first, then the four per-column values of the three window columns
    for (int i = threadIdx.x; i < kLW * kBwdH; i += kLossThreads) {
        const int ty = i / kLW, tx = i - ty * kLW;
        const int x = x0 + tx, y = y0 + ty;
        if (x >= p.W || y >= p.H) continue;
        const float sv0 = 2.0f + (y == 0) + (y == p.H - 1), dv0 = (float)(y == p.H - 1) - (float)(y == 0);
        float U[3], Wq[3], Xd[3], Xs[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int gx = kGCol + tx + c - 1;
            float u[6], m[6], d[6];
#pragma unroll
            for (int f = 0; f < 6; ++f) {
                u[f] = G[f][ty][gx];
                m[f] = G[f][ty + 1][gx];
                d[f] = G[f][ty + 2][gx];
            }
            const float RsG = fmaf(sv0, m[0], __fadd_rn(u[0], d[0])), RsA = fmaf(sv0, m[1], __fadd_rn(u[1], d[1]));
            const float RdV = fmaf(dv0, m[2], __fsub_rn(u[2], d[2])), RdCv = fmaf(dv0, m[3], __fsub_rn(u[3], d[3]));
            const float ReB = fmaf(d[4], 1.0f, u[4]), RdpCu = fmaf(d[5], -1.0f, u[5]);
            gather_column(RsG, RsA, RdV, RdCv, ReB, RdpCu, U[c], Wq[c], Xd[c], Xs[c]);
        }
        const float sh0 = 2.0f + (x == 0) + (x == p.W - 1), dh0 = (float)(x == p.W - 1) - (float)(x == 0);
        float res = fmaf(sh0, Xs[1], fmaf(dh0, Xd[1], __fadd_rn(U[0], Wq[2])));
        if constexpr (L1) {     // d/dpred of sum |gt - pred| m / sum m: sign(pred - gt) m / sum m (sign(0) = 0, as torch.abs)
            const float zg = tg[kLRow + ty][kLCol + tx], zp = tp[kLRow + ty][kLCol + tx];
            const float mk = mask_value<L1>(p, b * hw + (size_t)y * p.W + x, zg);
            const float sgn = (zp > zg) ? 1.0f : ((zp < zg) ? -1.0f : 0.0f);
            res = fmaf(l1_scale * mk, sgn, res);
        }
        p.grad_pred[b * hw + (size_t)y * p.W + x] = res;
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
