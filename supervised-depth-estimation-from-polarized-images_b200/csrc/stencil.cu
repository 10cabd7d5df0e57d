// Depth -> surface-normal 3x3 stencil (forward).
//
// Replaces kornia.geometry.depth.depth_to_normals (kornia 0.5.11, environment.yml:41) as called at
// manydepth/trainer.py:1305-1306,1477,1484:
//     xyz  = ((u - cx)/fx * Z, (v - cy)/fy * Z, Z)        u in [0, W-1], v in [0, H-1]   (depth_to_3d)
//     grad = cross-correlation of the REPLICATE-padded xyz with sobel_x / 8 and sobel_y / 8
//     n    = cross(d xyz/du, d xyz/dv);  n / max(||n||, 1e-12)
// Replicate padding acts on xyz, so a border neighbour contributes the xyz of the CLAMPED pixel
// (clamped u and v as well as clamped depth).
//
// Roofline: HBM, 4 B read + 12 B written per pixel.  A CTA stages a (TH+2) x (TW+2) depth tile with
// its halo in shared memory (each depth value is fetched from DRAM once and re-used 9 times from
// shared memory); a thread produces four consecutive pixels and writes one 16-byte vector per plane.
#include "polcue_device.cuh"
#include "polcue_host.h"

namespace polcue {
namespace {

constexpr int kTW = 128, kTH = 32, kRowGroups = 4, kStencilThreads = (kTW / 4) * (kTH / kRowGroups);  // 256 threads, 4 rows each
constexpr int kPitch = kTW + 4;  // [halo | TW | halo | pad], keeps rows 16-byte multiples

struct StencilParams {
    const float* depth;
    const float* K;
    float* normals;
    int H, W;
    bool vec4;  // W % 4 == 0 and 16-byte aligned output
};

__global__ void __launch_bounds__(kStencilThreads) depth_to_normals_kernel(const StencilParams p) {
    __shared__ float tile[(kTH + 2) * kPitch];
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;
    const size_t hw = (size_t)p.H * p.W;
    const float* z = p.depth + (size_t)b * hw;

    // stage tile + halo with clamped (replicate) coordinates
    for (int i = threadIdx.x; i < (kTH + 2) * (kTW + 2); i += kStencilThreads) {
        const int r = i / (kTW + 2), c = i - r * (kTW + 2);
        const int yy = min(max(y0 + r - 1, 0), p.H - 1);
        const int xx = min(max(x0 + c - 1, 0), p.W - 1);
        tile[r * kPitch + c] = __ldg(z + (size_t)yy * p.W + xx);
    }
    const float* k = p.K + (size_t)b * 9;
    const float inv_fx = 1.0f / __ldg(k + 0), cx = __ldg(k + 2);
    const float inv_fy = 1.0f / __ldg(k + 4), cy = __ldg(k + 5);
    __syncthreads();

    const int ty0 = threadIdx.x / (kTW / 4), tx = threadIdx.x - ty0 * (kTW / 4);
    const int xb = x0 + 4 * tx;
    if (xb >= p.W) return;
    float fx6[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) fx6[c] = ((float)min(max(xb + c - 1, 0), p.W - 1) - cx) * inv_fx;

#pragma unroll 1
    for (int rg = 0; rg < kRowGroups; ++rg) {
    const int ty = ty0 + rg * (kTH / kRowGroups);
    const int y = y0 + ty;
    if (y >= p.H) break;

    // ray factors of the three rows (clamped rows repeat the border row's factor)
    float fy3[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) fy3[r] = ((float)min(max(y + r - 1, 0), p.H - 1) - cy) * inv_fy;
    // xyz of the 3 x 6 window, shared by the thread's four pixels
    float X[3][6], Y[3][6], Z[3][6];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const float d = tile[(ty + r) * kPitch + 4 * tx + c];
            Z[r][c] = d;
            X[r][c] = fx6[c] * d;
            Y[r][c] = fy3[r] * d;
        }
    // separable Sobel: per column the vertical smoothing S = top + 2 mid + bottom and the vertical difference
    // D = bottom - top; then d/du = S[c+2] - S[c], d/dv = D[c] + 2 D[c+1] + D[c+2].  The common factor 1/8 of
    // both gradients only scales the cross product by 1/64, so it is folded into the normalisation guard below.
    float Su[3][6], Dv[3][6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        Su[0][c] = fmaf(2.0f, X[1][c], X[0][c] + X[2][c]);
        Su[1][c] = fmaf(2.0f, Y[1][c], Y[0][c] + Y[2][c]);
        Su[2][c] = fmaf(2.0f, Z[1][c], Z[0][c] + Z[2][c]);
        Dv[0][c] = X[2][c] - X[0][c];
        Dv[1][c] = Y[2][c] - Y[0][c];
        Dv[2][c] = Z[2][c] - Z[0][c];
    }

    float out[3][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float gu[3], gv[3];  // 8 * d/du, 8 * d/dv of (X, Y, Z)
#pragma unroll
        for (int comp = 0; comp < 3; ++comp) {
            gu[comp] = Su[comp][j + 2] - Su[comp][j];
            gv[comp] = fmaf(2.0f, Dv[comp][j + 1], Dv[comp][j] + Dv[comp][j + 2]);
        }
        // products rounded separately (no FMA contraction): parallel gradients next to zero-depth holes then cancel
        // to an exact zero vector, as they do in the reference's torch.cross, instead of leaving round-off
        const float nx = __fsub_rn(__fmul_rn(gu[1], gv[2]), __fmul_rn(gu[2], gv[1]));
        const float ny = __fsub_rn(__fmul_rn(gu[2], gv[0]), __fmul_rn(gu[0], gv[2]));
        const float nz = __fsub_rn(__fmul_rn(gu[0], gv[1]), __fmul_rn(gu[1], gv[0]));
        // n / max(|n|, eps) with n = 64 x the reference's cross product: 1 / max(|n|, 64 eps) = min(rsqrt(|n|^2), 1 / (64 eps))
        const float inv = fminf(rsqrt_approx(fmaf(nx, nx, fmaf(ny, ny, nz * nz))), 1.0f / (64.0f * 1e-12f));
        out[0][j] = nx * inv;
        out[1][j] = ny * inv;
        out[2][j] = nz * inv;
    }

    float* o = p.normals + (size_t)b * 3 * hw + (size_t)y * p.W + xb;
    if (p.vec4 && xb + 3 < p.W) {
#pragma unroll
        for (int comp = 0; comp < 3; ++comp) st_stream_vec<4>(o + comp * hw, out[comp]);
    } else {
#pragma unroll
        for (int comp = 0; comp < 3; ++comp)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (xb + j < p.W) st_stream_f32(o + comp * hw + j, out[comp][j]);
    }
    }
}

}  // namespace
}  // namespace polcue

using namespace polcue;

extern "C" int polcue_depth_to_normals_f32(const float* depth, const float* K, int B, int H, int W, float* normals,
                                           polcue_stream_t stream) {
    if (!depth || !K || !normals || B < 0 || H <= 0 || W <= 0 || B > 65535) return POLCUE_EINVAL;
    if ((reinterpret_cast<uintptr_t>(depth) | reinterpret_cast<uintptr_t>(normals) | reinterpret_cast<uintptr_t>(K)) & 3)
        return POLCUE_EINVAL;
    if (B == 0) return POLCUE_OK;
    StencilParams p;
    p.depth = depth;
    p.K = K;
    p.normals = normals;
    p.H = H;
    p.W = W;
    p.vec4 = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(normals) & 15) == 0);
    const dim3 grid((W + kTW - 1) / kTW, (H + kTH - 1) / kTH, B);
    if (grid.y > 65535) return POLCUE_E2BIG;
    depth_to_normals_kernel<<<grid, kStencilThreads, 0, (cudaStream_t)stream>>>(p);
    return launch_status();
}
