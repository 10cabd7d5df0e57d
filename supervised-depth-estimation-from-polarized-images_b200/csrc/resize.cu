// Loader front end: the Lanczos ("ANTIALIAS") resize of the 8-bit polarizer images, bit-exact with Pillow.
//
// Reference being replaced (file:line relative to the reference root):
//   manydepth/datasets/indoor_dataset.py:77       self.interp = Image.ANTIALIAS
//   manydepth/datasets/indoor_dataset.py:115      self.resize_pol = transforms.Resize((height, width), interpolation=self.interp)
//   manydepth/datasets/indoor_dataset.py:335-349  resize_pol(get_gray(...)) for pol00 / pol10 / pol01 / pol11
//   manydepth/datasets/hammer_dataset.py:68-75    get_gray: 'L' image, optional FLIP_LEFT_RIGHT
// The arithmetic itself lives in Pillow (pinned 6.2.1, environment.yml:14), src/libImaging/Resample.c, restated here
// from its published algorithm: precompute_coeffs (float64 Lanczos-3 weights, support scaled by the shrink factor,
// normalised per output sample), normalize_coeffs_8bpc (22-bit fixed point, round half away from zero), a horizontal
// pass into an 8-bit intermediate and a vertical pass, each  clip8((2^21 + sum pixel * k) >> 22).
//
// Roofline: this is integer multiply-add work, ~56 taps per output sample per plane (15 horizontal on 2.6 input rows +
// 17 vertical for 832x1088 -> 320x480); the input is read once from HBM (0.9 MB per plane) and the intermediate
// (Hin x Wout bytes per plane) stays in L2.  Bound: shared-memory loads / IMAD issue, not HBM; no tensor cores (the
// banded 22-bit fixed-point weights are neither dense nor 8-bit).
#include <cmath>
#include <cstring>

#include "polcue_device.cuh"
#include "polcue_host.h"

struct polcue_resize_plan {
    int in_h = 0, in_w = 0, out_h = 0, out_w = 0;
    int ksize[2] = {0, 0};               // 0: horizontal, 1: vertical
    std::vector<int> bounds[2];          // (first input index, tap count) per output index
    std::vector<int> kk[2];              // out x ksize fixed-point weights, row-major as Pillow holds them
    int* d_blob = nullptr;               // bounds_h | kk_h TRANSPOSED [ksize][out_w] | bounds_v | kk_v
    size_t off_kk_h = 0, off_bounds_v = 0, off_kk_v = 0, blob_ints = 0;
    int device = -1;
};

namespace polcue {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;   // Resample.c PRECISION_BITS
constexpr int kHalf = 1 << (kPrecisionBits - 1);

double sinc_filter(double x) {
    if (x == 0.0) return 1.0;
    x = x * M_PI;
    return std::sin(x) / x;
}
double lanczos_filter(double x) {   // truncated sinc, support 3
    if (-3.0 <= x && x < 3.0) return sinc_filter(x) * sinc_filter(x / 3);
    return 0.0;
}

// precompute_coeffs + normalize_coeffs_8bpc for the whole-image box.  Equal sizes: Pillow skips the pass; the
// single-tap identity below is the same function and lets one kernel handle flips and copies.
int make_axis(int in_size, int out_size, std::vector<int>& bounds, std::vector<int>& kk) {
    bounds.assign((size_t)out_size * 2, 0);
    if (in_size == out_size) {
        kk.assign((size_t)out_size, 1 << kPrecisionBits);
        for (int i = 0; i < out_size; ++i) {
            bounds[2 * i] = i;
            bounds[2 * i + 1] = 1;
        }
        return 1;
    }
    const float in0 = 0.0f, in1 = (float)in_size;   // the box is held in C floats
    const double scale = (double)(in1 - in0) / out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 3.0 * filterscale;
    const int ksize = (int)std::ceil(support) * 2 + 1;
    kk.assign((size_t)out_size * ksize, 0);
    std::vector<double> w(ksize);
    const double ss = 1.0 / filterscale;
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = in0 + (xx + 0.5) * scale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            w[x] = lanczos_filter((x + xmin - center + 0.5) * ss);
            ww += w[x];
        }
        for (int x = 0; x < xmax; ++x) {
            if (ww != 0.0) w[x] /= ww;
            kk[(size_t)xx * ksize + x] = w[x] < 0 ? (int)(-0.5 + w[x] * (1 << kPrecisionBits)) : (int)(0.5 + w[x] * (1 << kPrecisionBits));
        }
        bounds[2 * xx] = xmin;
        bounds[2 * xx + 1] = xmax;
    }
    return ksize;
}

int build_plan(int in_h, int in_w, int out_h, int out_w, polcue_resize_plan** out) {
    if (!out) return POLCUE_EINVAL;
    *out = nullptr;
    if (in_h <= 0 || in_w <= 0 || out_h <= 0 || out_w <= 0) return POLCUE_EINVAL;
    if (in_h >= (1 << 24) || in_w >= (1 << 24) || out_h >= (1 << 24) || out_w >= (1 << 24)) return POLCUE_E2BIG;
    auto* plan = new polcue_resize_plan();
    plan->in_h = in_h; plan->in_w = in_w; plan->out_h = out_h; plan->out_w = out_w;
    plan->ksize[0] = make_axis(in_w, out_w, plan->bounds[0], plan->kk[0]);
    plan->ksize[1] = make_axis(in_h, out_h, plan->bounds[1], plan->kk[1]);
    *out = plan;
    return POLCUE_OK;
}

__device__ __forceinline__ uint32_t clip8(int acc) {
    return (uint32_t)min(max(acc >> kPrecisionBits, 0), 255);   // arithmetic shift, then the clip8 lookup's clamp
}

// ---------------------------------------------------------------------------------------------
// Horizontal pass.  One CTA = kRowsH input rows of one image, staged in shared memory; one thread = one output
// column with its <= KMAX weights in registers (KMAX = 0: any tap count, weights re-read through L1).
// ---------------------------------------------------------------------------------------------
constexpr int kRowsH = 16;
constexpr int kPadH = 64;   // bytes before and after the staged rows: taps with zero weight may index past a row end

struct HParams {
    const uint8_t* src[4];      // image i lives at src[i % nsrc] + (i / nsrc) * image_stride
    int nsrc;
    long long image_stride;
    const uint8_t* flip;        // per (i / nsrc): mirror the image left-right before resizing; may be null
    uint8_t* dst;               // [images, in_h, out_w]
    const int2* bounds;         // [out_w]
    const int* kkT;             // [ksize][out_w]
    int ksize, in_h, in_w, out_w;
    int vec16;                  // rows may be staged with 16-byte loads
};

template <int KMAX, bool FLIP>
__device__ __forceinline__ void hpass_columns(const HParams& p, const uint8_t* tile, int rows, uint8_t* dst_rows) {
    for (int xx = threadIdx.x; xx < p.out_w; xx += blockDim.x) {
        const int2 bd = p.bounds[xx];
        const uint8_t* base = tile + (FLIP ? p.in_w - 1 - bd.x : bd.x);
        if constexpr (KMAX > 0) {
            int k[KMAX];
#pragma unroll
            for (int j = 0; j < KMAX; ++j) k[j] = j < p.ksize ? __ldg(p.kkT + (size_t)j * p.out_w + xx) : 0;
            for (int r = 0; r < rows; ++r) {
                const uint8_t* px = base + r * p.in_w;
                int acc = kHalf;
#pragma unroll
                for (int j = 0; j < KMAX; ++j) acc += (int)px[FLIP ? -j : j] * k[j];
                dst_rows[(size_t)r * p.out_w + xx] = (uint8_t)clip8(acc);
            }
        } else {
            for (int r = 0; r < rows; ++r) {
                const uint8_t* px = base + r * p.in_w;
                int acc = kHalf;
                for (int j = 0; j < bd.y; ++j) acc += (int)px[FLIP ? -j : j] * __ldg(p.kkT + (size_t)j * p.out_w + xx);
                dst_rows[(size_t)r * p.out_w + xx] = (uint8_t)clip8(acc);
            }
        }
    }
}

template <int KMAX>
__global__ void __launch_bounds__(512) resize_h_kernel(const __grid_constant__ HParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint8_t* tile = smem_raw + kPadH;
    const int img = blockIdx.y;
    const int row0 = blockIdx.x * kRowsH;
    const int rows = min(kRowsH, p.in_h - row0);
    const uint8_t* src = p.src[img % p.nsrc] + (size_t)(img / p.nsrc) * p.image_stride + (size_t)row0 * p.in_w;
    const int bytes = rows * p.in_w;   // the rows are contiguous in memory
    if (p.vec16) {
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        uint4* t4 = reinterpret_cast<uint4*>(tile);
        for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x) {
            uint4 v;
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(s4 + i));
            t4[i] = v;
        }
    } else {
        for (int i = threadIdx.x; i < bytes; i += blockDim.x) tile[i] = ld_stream_u8(src + i);
    }
    if (threadIdx.x < kPadH) {      // zero-weight taps multiply these
        smem_raw[threadIdx.x] = 0;
        tile[bytes + threadIdx.x] = 0;
    }
    __syncthreads();
    uint8_t* dst_rows = p.dst + ((size_t)img * p.in_h + row0) * p.out_w;
    const bool flip = p.flip && p.flip[img / p.nsrc];
    if (flip) hpass_columns<KMAX, true>(p, tile, rows, dst_rows);
    else hpass_columns<KMAX, false>(p, tile, rows, dst_rows);
}

// ---------------------------------------------------------------------------------------------
// Vertical pass.  One thread = V adjacent columns of one output row; the 8-bit intermediate is read through L1/L2
// (adjacent output rows of a CTA share most of their input rows), the row's weights are warp-uniform loads.
// ---------------------------------------------------------------------------------------------
struct VParams {
    const uint8_t* src;        // [images, in_h, w]
    uint8_t* dst;              // [images, out_h, w]
    const int2* bounds;        // [out_h]
    const int* kk;             // [out_h][ksize]
    int ksize, in_h, out_h, w;
};

constexpr int kRowsV = 4;

template <int V>
__global__ void __launch_bounds__(128 * kRowsV) resize_v_kernel(const __grid_constant__ VParams p) {
    const int yy = blockIdx.y * kRowsV + threadIdx.y;
    const int xg = blockIdx.x * 128 + threadIdx.x;
    if (yy >= p.out_h || xg * V >= p.w) return;
    const int img = blockIdx.z;
    const int2 bd = p.bounds[yy];
    const int* k = p.kk + (size_t)yy * p.ksize;
    const uint8_t* col = p.src + ((size_t)img * p.in_h + bd.x) * p.w + (size_t)xg * V;
    int acc[V];
#pragma unroll
    for (int c = 0; c < V; ++c) acc[c] = kHalf;
#pragma unroll 4
    for (int j = 0; j < bd.y; ++j) {
        const int kj = __ldg(k + j);
        if constexpr (V == 4) {
            const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(col + (size_t)j * p.w));
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[c] += (int)__byte_perm(w, 0, 0x4440 + c) * kj;
        } else {
            acc[0] += (int)__ldg(col + (size_t)j * p.w) * kj;
        }
    }
    uint8_t* out = p.dst + ((size_t)img * p.out_h + yy) * p.w + (size_t)xg * V;
    if constexpr (V == 4) {
        *reinterpret_cast<uint32_t*>(out) = clip8(acc[0]) | (clip8(acc[1]) << 8) | (clip8(acc[2]) << 16) | (clip8(acc[3]) << 24);
    } else {
        *out = (uint8_t)clip8(acc[0]);
    }
}

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

int launch_resize(const polcue_resize_plan* plan, const uint8_t* const* src, int nsrc, long long image_stride, int images,
                  const uint8_t* flip, uint8_t* workspace, uint8_t* dst, cudaStream_t stream) {
    if (!plan || !plan->d_blob || !src || nsrc < 1 || nsrc > 4 || images < 0 || images % nsrc || !workspace || !dst)
        return POLCUE_EINVAL;
    for (int s = 0; s < nsrc; ++s)
        if (!src[s]) return POLCUE_EINVAL;
    if (images == 0) return POLCUE_OK;
    if (images > 65535) return POLCUE_E2BIG;
    if ((unsigned long long)plan->in_h * plan->in_w >= (1ull << 31) || (unsigned long long)kRowsH * plan->in_w > 200 * 1024)
        return POLCUE_E2BIG;
    HParams h;
    bool v16 = plan->in_w % 16 == 0 && image_stride % 16 == 0;
    for (int s = 0; s < 4; ++s) {
        h.src[s] = src[s < nsrc ? s : 0];
        v16 = v16 && aligned(h.src[s], 16);
    }
    h.nsrc = nsrc;
    h.image_stride = image_stride;
    h.flip = flip;
    h.dst = workspace;
    h.bounds = reinterpret_cast<const int2*>(plan->d_blob);
    h.kkT = plan->d_blob + plan->off_kk_h;
    h.ksize = plan->ksize[0];
    h.in_h = plan->in_h;
    h.in_w = plan->in_w;
    h.out_w = plan->out_w;
    h.vec16 = v16 ? 1 : 0;
    const int threads = std::min(512, (plan->out_w + 31) / 32 * 32);
    const size_t smem = (size_t)kRowsH * plan->in_w + 2 * kPadH;
    const dim3 grid_h((plan->in_h + kRowsH - 1) / kRowsH, images);
    auto launch_h = [&](auto kern) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<grid_h, threads, smem, stream>>>(h);
        return launch_status();
    };
    int rc;
    if (h.ksize <= 16) rc = launch_h(resize_h_kernel<16>);
    else if (h.ksize <= 32) rc = launch_h(resize_h_kernel<32>);
    else rc = launch_h(resize_h_kernel<0>);
    if (rc != POLCUE_OK) return rc;

    VParams v;
    v.src = workspace;
    v.dst = dst;
    v.bounds = reinterpret_cast<const int2*>(plan->d_blob + plan->off_bounds_v);
    v.kk = plan->d_blob + plan->off_kk_v;
    v.ksize = plan->ksize[1];
    v.in_h = plan->in_h;
    v.out_h = plan->out_h;
    v.w = plan->out_w;
    const int vec = (plan->out_w % 4 == 0 && aligned(workspace, 4) && aligned(dst, 4)) ? 4 : 1;
    const int groups = plan->out_w / vec;
    const dim3 grid_v((groups + 127) / 128, (plan->out_h + kRowsV - 1) / kRowsV, images);
    if (grid_v.y > 65535) return POLCUE_E2BIG;
    if (vec == 4) resize_v_kernel<4><<<grid_v, dim3(128, kRowsV), 0, stream>>>(v);
    else resize_v_kernel<1><<<grid_v, dim3(128, kRowsV), 0, stream>>>(v);
    return launch_status();
}

}  // namespace
}  // namespace polcue

using namespace polcue;

extern "C" {

int polcue_resize_plan_host_build(int in_h, int in_w, int out_h, int out_w, polcue_resize_plan** out) {
    return build_plan(in_h, in_w, out_h, out_w, out);
}

int polcue_resize_plan_create(int in_h, int in_w, int out_h, int out_w, polcue_resize_plan** out) {
    polcue_resize_plan* plan = nullptr;
    int rc = build_plan(in_h, in_w, out_h, out_w, &plan);
    if (rc != POLCUE_OK) return rc;
    // device blob: bounds_h | kk_h transposed | bounds_v | kk_v
    const int kh = plan->ksize[0], kv = plan->ksize[1];
    plan->off_kk_h = (size_t)2 * out_w;
    plan->off_bounds_v = plan->off_kk_h + (size_t)kh * out_w;
    plan->off_bounds_v = (plan->off_bounds_v + 1) & ~(size_t)1;   // int2 alignment
    plan->off_kk_v = plan->off_bounds_v + (size_t)2 * out_h;
    plan->blob_ints = plan->off_kk_v + (size_t)kv * out_h;
    std::vector<int> blob(plan->blob_ints, 0);
    std::memcpy(blob.data(), plan->bounds[0].data(), sizeof(int) * 2 * out_w);
    for (int xx = 0; xx < out_w; ++xx)
        for (int j = 0; j < kh; ++j) blob[plan->off_kk_h + (size_t)j * out_w + xx] = plan->kk[0][(size_t)xx * kh + j];
    std::memcpy(blob.data() + plan->off_bounds_v, plan->bounds[1].data(), sizeof(int) * 2 * out_h);
    std::memcpy(blob.data() + plan->off_kk_v, plan->kk[1].data(), sizeof(int) * (size_t)kv * out_h);
    cudaError_t e = cudaGetDevice(&plan->device);
    if (e == cudaSuccess) e = cudaMalloc(&plan->d_blob, sizeof(int) * plan->blob_ints);
    if (e == cudaSuccess) e = cudaMemcpy(plan->d_blob, blob.data(), sizeof(int) * plan->blob_ints, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        if (plan->d_blob) cudaFree(plan->d_blob);
        delete plan;
        *out = nullptr;
        return (int)e;
    }
    *out = plan;
    return POLCUE_OK;
}

void polcue_resize_plan_destroy(polcue_resize_plan* plan) {
    if (!plan) return;
    if (plan->d_blob) cudaFree(plan->d_blob);
    delete plan;
}

int polcue_resize_plan_coeffs(const polcue_resize_plan* plan, int axis, int* bounds, int* kk, size_t kk_capacity) {
    if (!plan || axis < 0 || axis > 1) return POLCUE_EINVAL;
    if (bounds) std::memcpy(bounds, plan->bounds[axis].data(), sizeof(int) * plan->bounds[axis].size());
    if (kk) {
        if (kk_capacity < plan->kk[axis].size()) return POLCUE_EINVAL;
        std::memcpy(kk, plan->kk[axis].data(), sizeof(int) * plan->kk[axis].size());
    }
    return plan->ksize[axis];
}

size_t polcue_resize_workspace_bytes(const polcue_resize_plan* plan, int images) {
    if (!plan || images < 0) return 0;
    return (size_t)images * plan->in_h * plan->out_w;
}

int polcue_resize_lanczos_u8(const polcue_resize_plan* plan, const uint8_t* src, int images, const uint8_t* flip,
                             uint8_t* workspace, uint8_t* dst, polcue_stream_t stream) {
    if (!plan) return POLCUE_EINVAL;
    const uint8_t* one[1] = {src};
    return launch_resize(plan, one, 1, (long long)plan->in_h * plan->in_w, images, flip, workspace, dst, (cudaStream_t)stream);
}

int polcue_loader_front_end_u8(const polcue_resize_plan* plan, const uint8_t* i0, const uint8_t* i45, const uint8_t* i90,
                               const uint8_t* i135, int B, const uint8_t* flip, const polcue_lut* lut, uint8_t* workspace,
                               uint8_t* planes, float* iun, float* xolp, float* normals, polcue_stream_t stream) {
    if (!plan || !planes || !xolp || B < 0) return POLCUE_EINVAL;
    if (B == 0) return POLCUE_OK;
    const uint8_t* four[4] = {i0, i45, i90, i135};
    int rc = launch_resize(plan, four, 4, (long long)plan->in_h * plan->in_w, 4 * B, flip, workspace, planes, (cudaStream_t)stream);
    if (rc != POLCUE_OK) return rc;
    return fused_planes_strided(planes, B, plan->out_h, plan->out_w, lut, iun, xolp, normals, (cudaStream_t)stream);
}

}  // extern "C"
