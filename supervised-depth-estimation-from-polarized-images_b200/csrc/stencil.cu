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
// Roofline: HBM, 4 B read + 12 B written per pixel.  A CTA owns a 128 x 32 pixel tile.  Its depth tile with a
// one-pixel halo is staged in shared memory by ONE TMA tensor copy (cp.async.bulk.tensor.3d, box 136 x 34 x 1
// over the (W, H, B) depth tensor, out-of-bounds elements zero-filled by the hardware), completion on an
// mbarrier; CTAs on an image border then overwrite the zero halo with the replicated edge.  Each depth value is
// fetched from DRAM once and re-used 9 times from shared memory.  A thread produces four consecutive pixels of
// four rows, reads its 3 x 6 window as one LDS.128 + two LDS.32 per row and writes one 16-byte vector per plane.
// Images whose width is not a multiple of 4 (TMA needs 16-byte row strides) take the same arithmetic with a
// manually staged tile.
#include <cuda.h>

#include <mutex>

#include "polcue_device.cuh"
#include "polcue_host.h"

namespace polcue {
namespace {

constexpr int kTW = 128, kTH = 32, kRowGroups = 4;
constexpr int kStencilThreads = (kTW / 4) * (kTH / kRowGroups);  // 256 threads, each 4 px of 4 rows
constexpr int kBoxW = kTW + 8, kBoxH = kTH + 2;                  // halo'd box; interior starts at column 4 (16-byte aligned)
constexpr int kColOff = 4;                                       // shared column of image column x0
constexpr uint32_t kTileBytes = kBoxW * kBoxH * sizeof(float);

struct StencilParams {
    const float* depth;
    const float* K;
    float* normals;
    int H, W;
    bool vec4;  // W % 4 == 0 and 16-byte aligned output
};

// Normals of rows ty0, ty0 + 8, ... of the tile for the four pixels starting at image column xb.
// tile[r][kColOff + c] holds depth(y0 - 1 + r, x0 + c) with replicate padding already applied.
// kraw: (fx, cx, fy, cy) of image b, requested by the caller BEFORE it waits for the tile (fetched afterwards, every warp
// of the CTA sat out an L2 round trip between the tile's arrival and its first arithmetic).
__device__ __forceinline__ float4 fetch_intrinsics(const StencilParams& p, int b) {
    const float* k = p.K + (size_t)b * 9;
    return make_float4(__ldg(k + 0), __ldg(k + 2), __ldg(k + 4), __ldg(k + 5));
}
__device__ __forceinline__ void stencil_rows(const StencilParams& p, const float (*tile)[kBoxW], int b, int x0, int y0, const float4 kraw) {
    const size_t hw = (size_t)p.H * p.W;
    const float inv_fx = 1.0f / kraw.x, cx = kraw.y;
    const float inv_fy = 1.0f / kraw.z, cy = kraw.w;
    const int ty0 = threadIdx.x / (kTW / 4), tx = threadIdx.x - ty0 * (kTW / 4);
    const int xb = x0 + 4 * tx;
    if (xb >= p.W) return;
    float fx6[6];   // ray factors of the six window columns (clamped columns repeat the border factor)
#pragma unroll
    for (int c = 0; c < 6; ++c) fx6[c] = ((float)min(max(xb + c - 1, 0), p.W - 1) - cx) * inv_fx;

#pragma unroll 1
    for (int rg = 0; rg < kRowGroups; ++rg) {
        const int ty = ty0 + rg * (kTH / kRowGroups);
        const int y = y0 + ty;
        if (y >= p.H) break;
        float fy3[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) fy3[r] = ((float)min(max(y + r - 1, 0), p.H - 1) - cy) * inv_fy;

        // depth of the 3 x 6 window, shared by the thread's four pixels
        float Z[3][6];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float* row = &tile[ty + r][kColOff + 4 * tx];
            const float4 mid = *reinterpret_cast<const float4*>(row);
            Z[r][0] = row[-1];
            Z[r][1] = mid.x; Z[r][2] = mid.y; Z[r][3] = mid.z; Z[r][4] = mid.w;
            Z[r][5] = row[4];
        }
        // separable Sobel on xyz = (fx[c] Z, fy[r] Z, Z): per column the vertical smoothing S = top + 2 mid + bottom
        // and the vertical difference D = bottom - top; then d/du = S[c+2] - S[c], d/dv = D[c] + 2 D[c+1] + D[c+2].
        // The column factor fx[c] is common to a column's three rows, so X needs one multiply per column.  The common
        // factor 1/8 of both gradients only scales the cross product by 1/64; it is folded into the guard below.
        // The arithmetic below is the library's DEFINITION of the stencil, rounding by rounding (normals_loss.cu repeats it in
        // its scalar and packed kernels, so the normals inside the loss are these, bit for bit):
        //   S2 = fma(2, Z1, Z0 + Z2), D2 = Z2 - Z0                     vertical smoothing / difference of Z per column
        //   X: P = fl(fx S2), Q = fl(fx D2);  gu_x = P[+1] - P[-1];  gv_x = fma(2, Q[0], Q[-1] + Q[+1])
        //   Y: lo = fl(fy[-1] Z0), hi = fl(fy[+1] Z2), S1 = fma(fy[+1], Z2, fma(2 fy[0], Z1, lo)), D1 = hi - lo;
        //      gu_y = S1[+1] - S1[-1];  gv_y = fma(2, D1[0], D1[-1] + D1[+1])
        //   Z: gu_z = S2[+1] - S2[-1];  gv_z = fma(2, D2[0], D2[-1] + D2[+1])
        //   n = gu x gv with n_0 = fl(gu_1 gv_2) - fl(gu_2 gv_1), ...;  unit = fl(n * min(rsqrt(|n|^2), 1 / (64 eps)))
        // Products that feed a difference are rounded separately: replicated rows / columns and parallel gradients (next to
        // zero-depth holes) cancel to exact zeros, as they do in the float64 oracle.
        // The common factor 1/8 of both gradients only scales the cross product by 1/64; it is folded into the guard.
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
        float out[3][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float gu[3], gv[3];  // 8 * d/du, 8 * d/dv of (X, Y, Z)
            gu[0] = __fsub_rn(P[j + 2], P[j]);
            gv[0] = fmaf(2.0f, Q[j + 1], __fadd_rn(Q[j], Q[j + 2]));
            gu[1] = __fsub_rn(S1[j + 2], S1[j]);
            gv[1] = fmaf(2.0f, D1[j + 1], __fadd_rn(D1[j], D1[j + 2]));
            gu[2] = __fsub_rn(S2[j + 2], S2[j]);
            gv[2] = fmaf(2.0f, D2[j + 1], __fadd_rn(D2[j], D2[j + 2]));
            const float nx = __fsub_rn(__fmul_rn(gu[1], gv[2]), __fmul_rn(gu[2], gv[1]));
            const float ny = __fsub_rn(__fmul_rn(gu[2], gv[0]), __fmul_rn(gu[0], gv[2]));
            const float nz = __fsub_rn(__fmul_rn(gu[0], gv[1]), __fmul_rn(gu[1], gv[0]));
            // n / max(|n|, eps) with n = 64 x the reference's cross product: 1 / max(|n|, 64 eps) = min(rsqrt(|n|^2), 1 / (64 eps))
            const float inv = fminf(rsqrt_approx(fmaf(nx, nx, fmaf(ny, ny, __fmul_rn(nz, nz)))), 1.0f / (64.0f * 1e-12f));
            out[0][j] = __fmul_rn(nx, inv);
            out[1][j] = __fmul_rn(ny, inv);
            out[2][j] = __fmul_rn(nz, inv);
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

// ---- TMA-staged variant ----------------------------------------------------------------------
__global__ void __launch_bounds__(kStencilThreads) depth_to_normals_tma_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                               const StencilParams p) {
    __shared__ __align__(128) float tile[kBoxH][kBoxW];
    __shared__ uint64_t bar;
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;
    const float4 kraw = fetch_intrinsics(p, b);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(kTileBytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                smem_u32(&tile[0][0])),
            "l"(&tmap), "r"(x0 - kColOff), "r"(y0 - 1), "r"(b), "r"(smem_u32(&bar))
            : "memory");
    }
    lut_stage_wait(&bar);   // __syncthreads + mbarrier phase-0 wait (same helper the table staging uses)

    // replicate padding on image borders: the hardware zero-filled everything outside the image
    const int last_x = p.W - 1 - x0, last_y = p.H - 1 - y0;     // tile-local index of the last image column / row
    const bool left = x0 == 0, right = last_x < kTW, top = y0 == 0, bottom = last_y < kTH;
    if (left | right | top | bottom) {                           // block-uniform
        if (left | right) {
            for (int r = threadIdx.x; r < kBoxH; r += kStencilThreads) {
                if (left) tile[r][kColOff - 1] = tile[r][kColOff];
                if (right) tile[r][kColOff + last_x + 1] = tile[r][kColOff + last_x];
            }
            __syncthreads();
        }
        if (top | bottom) {
            for (int c = threadIdx.x; c < kBoxW; c += kStencilThreads) {
                if (top) tile[0][c] = tile[1][c];
                if (bottom) tile[last_y + 2][c] = tile[last_y + 1][c];
            }
        }
        __syncthreads();
    }
    stencil_rows(p, tile, b, x0, y0, kraw);
}

// ---- manually staged variant (any width / alignment) ----------------------------------------------
__global__ void __launch_bounds__(kStencilThreads) depth_to_normals_kernel(const StencilParams p) {
    __shared__ __align__(16) float tile[kBoxH][kBoxW];
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;
    const float* z = p.depth + (size_t)b * p.H * p.W;
    const float4 kraw = fetch_intrinsics(p, b);
    for (int i = threadIdx.x; i < kBoxH * (kTW + 2); i += kStencilThreads) {
        const int r = i / (kTW + 2), c = i - r * (kTW + 2);      // c = 0 is image column x0 - 1
        const int yy = min(max(y0 + r - 1, 0), p.H - 1);
        const int xx = min(max(x0 + c - 1, 0), p.W - 1);
        tile[r][kColOff - 1 + c] = __ldg(z + (size_t)yy * p.W + xx);
    }
    __syncthreads();
    stencil_rows(p, tile, b, x0, y0, kraw);
}

std::atomic<unsigned long long> g_tma_launches{0};

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
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
    cudaStream_t s = (cudaStream_t)stream;

    EncodeTiledFn encode = encode_tiled();
    const bool tma_ok = encode && (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(depth) & 15) == 0);
    if (tma_ok) {
        CUtensorMap tmap;
        const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)W * sizeof(float), (cuuint64_t)W * H * sizeof(float)};
        const cuuint32_t box[3] = {kBoxW, kBoxH, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(depth), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS) {
            depth_to_normals_tma_kernel<<<grid, kStencilThreads, 0, s>>>(tmap, p);
            g_tma_launches.fetch_add(1, std::memory_order_relaxed);
            return launch_status();
        }
    }
    depth_to_normals_kernel<<<grid, kStencilThreads, 0, s>>>(p);
    return launch_status();
}

extern "C" unsigned long long polcue_debug_stencil_tma_launches(void) { return g_tma_launches.load(); }
