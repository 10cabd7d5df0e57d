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
    int* d_blob = nullptr;               // bounds_h | kk_h TRANSPOSED [ksize][out_w] | bounds_v | kk_v | packed kk_h
    size_t off_kk_h = 0, off_bounds_v = 0, off_kk_v = 0, off_pack_h = 0, off_pack_v = 0, blob_ints = 0;
    int v_tile_rows = 0;                 // most input rows any tile of kRowsV output rows touches in the dp4a form
    int pack_words_v = 0;                // vertical weights as byte planes: [out_h][plane 3][word]; 0 = not packable
    int pack_words = 0;                  // horizontal weights as byte planes for dp4a: [dir 2][plane 3][word][out_w]; 0 = not packable
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
    const uint32_t* pack;       // [dir 2][plane 3][KW][out_w] byte-plane weights for the dp4a form, or null
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

// dp4a form of the same sum.  A 22-bit weight is split into byte planes k = lo + 256 mid + 65536 hi (lo, mid unsigned, hi
// signed), four taps per 32-bit word, so  sum px k = dp4a(px4, lo4) + 256 dp4a(px4, mid4) + 65536 dp4a(px4, hi4)  exactly.
// Per row a thread loads the KW + 1 aligned words that cover its 4 KW-byte window, funnels them to the window's
// alignment with PRMT, and issues 3 KW dp4a: a third of the shared-memory instructions of the byte-load form.
// Mirrored images use the weights in reverse order over the window that ends at in_w - 1 - xmin.
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int dp4a_uu(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {   // a unsigned bytes, b signed bytes
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// One output column's byte-plane weights (normal or mirrored order) and the first byte of its window.
template <int KW>
struct ColumnWeights {
    uint32_t lo[KW], mid[KW], hi[KW];
    int start;
    __device__ __forceinline__ void load(const HParams& p, int xx, bool flip) {
        const uint32_t* pk = p.pack + (size_t)(flip ? 3 * KW : 0) * p.out_w + xx;
#pragma unroll
        for (int i = 0; i < KW; ++i) {
            lo[i] = __ldg(pk + (size_t)i * p.out_w);
            mid[i] = __ldg(pk + (size_t)(KW + i) * p.out_w);
            hi[i] = __ldg(pk + (size_t)(2 * KW + i) * p.out_w);
        }
        const int xmin = p.bounds[xx].x;
        start = flip ? p.in_w - 1 - xmin - (4 * KW - 1) : xmin;
    }
};

// `rows` rows of one output column: per row KW + 1 aligned LDS.32 cover the 4 KW-byte window at shared address `addr`,
// PRMT funnels them to the window's alignment, 3 KW dp4a reduce them.  ROW4: in_w % 4 == 0, so the alignment is the same
// in every row and the selector is computed once.
template <int KW, bool ROW4>
__device__ __forceinline__ void dp4a_column_rows(const ColumnWeights<KW>& k, uint32_t addr, int rows, int in_w, uint8_t* out, int out_w) {
    uint32_t sel = 0x3210u + 0x1111u * (addr & 3u);
    if constexpr (ROW4) addr &= ~3u;
#pragma unroll 2
    for (int r = 0; r < rows; ++r) {
        uint32_t a = addr;
        if constexpr (!ROW4) {
            a = addr & ~3u;
            sel = 0x3210u + 0x1111u * (addr & 3u);
        }
        uint32_t w[KW + 1];
#pragma unroll
        for (int i = 0; i <= KW; ++i) w[i] = lds_u32(a + 4 * i);
        int lo = kHalf, mid = 0, hi = 0;
#pragma unroll
        for (int i = 0; i < KW; ++i) {
            const uint32_t v = __byte_perm(w[i], w[i + 1], sel);
            lo = dp4a_uu(v, k.lo[i], lo);
            mid = dp4a_uu(v, k.mid[i], mid);
            hi = dp4a_us(v, k.hi[i], hi);
        }
        *out = (uint8_t)clip8(lo + (mid << 8) + (hi << 16));
        out += out_w;
        addr += in_w;
    }
}

template <int KW>
__device__ __forceinline__ void hpass_columns_dp4a(const HParams& p, uint32_t tile_addr, int rows, uint8_t* dst_rows, bool flip) {
    for (int xx = threadIdx.x; xx < p.out_w; xx += blockDim.x) {
        ColumnWeights<KW> k;
        k.load(p, xx, flip);
        dp4a_column_rows<KW, false>(k, tile_addr + (uint32_t)k.start, rows, p.in_w, dst_rows + xx, p.out_w);
    }
}

template <int KW>   // KW > 0: window of 4 KW taps (dp4a form when the weights pack, else bytes with weights in registers); 0: any tap count
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
    if constexpr (KW > 0) {
        if (p.pack) {
            hpass_columns_dp4a<KW>(p, smem_u32(tile), rows, dst_rows, flip);
            return;
        }
    }
    if (flip) hpass_columns<4 * KW, true>(p, tile, rows, dst_rows);
    else hpass_columns<4 * KW, false>(p, tile, rows, dst_rows);
}

// The same pass as a pipeline: a CTA owns one output column per thread (weights stay in registers) and walks a run of
// row tiles of one image; while tile t is being reduced, tile t + 1 is already landing in the other shared-memory
// buffer through a bulk-async (TMA) copy signalled on an mbarrier.  Rows are contiguous in memory, so a tile is ONE
// copy.  Widths that are not a multiple of 16 bytes (or unaligned images) stage with ordinary loads instead.
struct HPipeParams {
    HParams h;
    int tiles_per_image, images;        // row tiles per image; images
    int step_img, step_t;               // grid size as (whole images, remaining tiles): one step of a CTA's tile walk
    int nsrc_shift;                     // log2(nsrc)
    int tile_rows;                      // rows per tile: 32, or 16 for wide images (three stages must fit twice per SM)
    uint32_t buf_stride;        // bytes between the two staging buffers (multiple of 16)
};

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}

constexpr int kStagesH = 3;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Warp-specialised: the last warp of the CTA is the producer (one lane issues the bulk copies), every other warp owns
// 32 output columns.  Stages are handed over through full/empty mbarriers, so no CTA-wide barrier sits in the loop and
// warps drift up to kStagesH - 1 tiles apart.
template <int KW, bool ROW4>   // ROW4: in_w % 4 == 0, so a column's window has the same word alignment in every row
__global__ void __launch_bounds__(1024, 1) resize_h_pipe_kernel(const __grid_constant__ HPipeParams pp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t full[kStagesH], empty[kStagesH];
    const HParams& p = pp.h;
    const int consumer_warps = blockDim.x / 32 - 1;
    const int warp = threadIdx.x / 32;
    const uint32_t buf0 = smem_u32(smem_raw), full0 = smem_u32(&full[0]), empty0 = smem_u32(&empty[0]);
    const int tile_rows = pp.tile_rows;
    const int tile_bytes = tile_rows * p.in_w;
    for (int i = threadIdx.x; i < kPadH * kStagesH; i += blockDim.x) {     // zero-weight taps multiply these
        const int st = i / kPadH, o = i - st * kPadH;
        smem_raw[st * pp.buf_stride + o] = 0;
        smem_raw[st * pp.buf_stride + kPadH + tile_bytes + o] = 0;
    }
    if (threadIdx.x == 0) {
        for (int st = 0; st < kStagesH; ++st) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full0 + 8 * st));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty0 + 8 * st), "r"(consumer_warps));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // The launch is persistent: CTA c reduces tiles c, c + grid, c + 2 grid, ... of the flat (image, row tile) list; the
    // last round is the only partially filled one.
    // (image, row tile) of this CTA's k-th tile, advanced without divisions: one step is grid = step_img images + step_t tiles
    int img = blockIdx.x / pp.tiles_per_image, t = blockIdx.x - img * pp.tiles_per_image;
    auto advance = [&]() {
        img += pp.step_img;
        t += pp.step_t;
        if (t >= pp.tiles_per_image) {
            t -= pp.tiles_per_image;
            ++img;
        }
    };
    const int nsrc_mask = p.nsrc - 1;      // nsrc is 1 or 4
    if (warp == consumer_warps) {
        if (threadIdx.x % 32 == 0) {
            for (int k = 0; img < pp.images; ++k, advance()) {
                const int st = k % kStagesH;
                mbar_wait(empty0 + 8 * st, ((k / kStagesH) & 1) ^ 1);              // passes at once on a fresh barrier
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // consumers' reads before the async write
                const int rows = min(tile_rows, p.in_h - t * tile_rows);
                const uint8_t* src = p.src[img & nsrc_mask] + (size_t)(img >> pp.nsrc_shift) * p.image_stride + (size_t)t * tile_bytes;
                bulk_load(buf0 + st * pp.buf_stride + kPadH, src, (uint32_t)(rows * p.in_w), full0 + 8 * st);
            }
        }
        return;
    }
    const int xx = threadIdx.x;
    const bool active = xx < p.out_w;
    ColumnWeights<KW> k;
    k.start = 0;
    int have = -1;                  // which weight order is loaded (0 normal, 1 mirrored)
    int st = 0, parity = 0;
    for (; img < pp.images; advance()) {
        const int rows = min(tile_rows, p.in_h - t * tile_rows);
        const int flip = (p.flip && p.flip[img >> pp.nsrc_shift]) ? 1 : 0;
        if (active && flip != have) {
            have = flip;
            k.load(p, xx, flip != 0);
        }
        mbar_wait(full0 + 8 * st, parity);
        if (active)
            dp4a_column_rows<KW, ROW4>(k, buf0 + st * pp.buf_stride + kPadH + (uint32_t)k.start, rows, p.in_w,
                                       p.dst + ((size_t)img * p.in_h + (size_t)t * tile_rows) * p.out_w + xx, p.out_w);
        __syncwarp();
        if (threadIdx.x % 32 == 0) mbar_arrive(empty0 + 8 * st);    // this warp is done with the stage
        if (++st == kStagesH) {
            st = 0;
            parity ^= 1;
        }
    }
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
    const uint32_t* pack;      // [out_h][plane 3][KW] byte-plane weights for the dp4a form, or null
    int vec16;                 // rows may be staged with 16-byte loads
    int ksize, in_h, out_h, w;
    long long tiles_total;     // dp4a form: images x row tiles x column tiles
    uint32_t buf_bytes;        // dp4a form: bytes of one staging buffer
};

constexpr int kRowsV = 4;
constexpr int kWorkspacePadRows = 32;   // zero-weight taps of the dp4a form may read this far past the last image

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

// dp4a form of the vertical pass: four input rows x four columns are transposed in registers (8 PRMT) so that each
// column's four taps share a word, then 3 dp4a per column as in the horizontal pass.  The row's byte-plane weights are
// warp-uniform and live in registers.  Taps past the row's count have zero weight; their loads may run into the next
// image of the workspace (or its tail padding), never out of it.
template <int BYTES>
__device__ __forceinline__ void cp_async(uint32_t dst, const void* src) {
    if constexpr (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}

// Persistent: CTA c reduces tiles c, c + grid, ... of the flat (image, row tile, column tile) list; a tile is kTileV
// output rows x 512 columns (each thread: 4 columns of kTileV / kRowsV rows).  The input rows of tile k + 1 are copied
// into the other shared-memory buffer with cp.async while tile k is reduced.
constexpr int kTileV = 16;

template <int KW>
__global__ void __launch_bounds__(128 * kRowsV) resize_v_dp4a_kernel(const __grid_constant__ VParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];     // 2 x [tile_rows][512]: the input rows of a tile
    const uint32_t smem0 = smem_u32(smem_raw);
    const int tid = threadIdx.y * 128 + threadIdx.x;
    const int tiles_x = (p.w + 511) / 512, tiles_y = (p.out_h + kTileV - 1) / kTileV;
    struct Tile {
        int img, yy0, x0, y_first, nrows, cols;
        bool valid;
    };
    auto locate = [&](int k) {
        Tile t;
        const long long id = blockIdx.x + (long long)k * gridDim.x;
        t.valid = id < p.tiles_total;
        if (!t.valid) return t;
        const int per_img = tiles_x * tiles_y;
        t.img = (int)(id / per_img);
        const int r = (int)(id - (long long)t.img * per_img);
        const int ty = r / tiles_x;
        t.yy0 = ty * kTileV;
        t.x0 = (r - ty * tiles_x) * 512;
        t.y_first = p.bounds[t.yy0].x;
        t.nrows = p.bounds[min(t.yy0 + kTileV, p.out_h) - 1].x + 4 * KW - t.y_first;     // bounds are non-decreasing
        t.cols = min(512, p.w - t.x0);                                                    // multiple of 4
        return t;
    };
    auto stage = [&](const Tile& t, int buf) {
        if (t.valid) {
            const uint8_t* src = p.src + ((size_t)t.img * p.in_h + t.y_first) * p.w + t.x0;
            const uint32_t dst = smem0 + buf * p.buf_bytes;
            if (p.vec16) {
                for (int i = tid; i < t.nrows * 32; i += 128 * kRowsV) {
                    const int r = i >> 5, c = (i & 31) * 16;
                    if (c < t.cols) cp_async<16>(dst + r * 512 + c, src + (size_t)r * p.w + c);
                }
            } else {
                for (int i = tid; i < t.nrows * 128; i += 128 * kRowsV) {
                    const int r = i >> 7, c = (i & 127) * 4;
                    if (c < t.cols) cp_async<4>(dst + r * 512 + c, src + (size_t)r * p.w + c);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    Tile cur = locate(0);
    stage(cur, 0);
    for (int k = 0; cur.valid; ++k) {
        const Tile nxt = locate(k + 1);
        stage(nxt, (k + 1) & 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        if (threadIdx.x * 4 < cur.cols) {
            const uint32_t tile0 = smem0 + (k & 1) * p.buf_bytes + threadIdx.x * 4u;
            uint8_t* dst = p.dst + ((size_t)cur.img * p.out_h) * p.w + cur.x0 + threadIdx.x * 4;
#pragma unroll 1
            for (int yy = cur.yy0 + threadIdx.y; yy < min(cur.yy0 + kTileV, p.out_h); yy += kRowsV) {
                const uint32_t* pk = p.pack + (size_t)yy * 3 * KW;
                uint32_t klo[KW], kmid[KW], khi[KW];
#pragma unroll
                for (int i = 0; i < KW; ++i) {
                    klo[i] = __ldg(pk + i);
                    kmid[i] = __ldg(pk + KW + i);
                    khi[i] = __ldg(pk + 2 * KW + i);
                }
                const uint32_t col = tile0 + (uint32_t)(p.bounds[yy].x - cur.y_first) * 512u;
                int lo[4] = {kHalf, kHalf, kHalf, kHalf}, mid[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
#pragma unroll
                for (int g = 0; g < KW; ++g) {
                    uint32_t r[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) r[i] = lds_u32(col + (4 * g + i) * 512);
                    const uint32_t t0 = __byte_perm(r[0], r[1], 0x5140), t1 = __byte_perm(r[2], r[3], 0x5140);   // columns 0, 1
                    const uint32_t t2 = __byte_perm(r[0], r[1], 0x7362), t3 = __byte_perm(r[2], r[3], 0x7362);   // columns 2, 3
                    const uint32_t c[4] = {__byte_perm(t0, t1, 0x5410), __byte_perm(t0, t1, 0x7632), __byte_perm(t2, t3, 0x5410),
                                           __byte_perm(t2, t3, 0x7632)};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        lo[q] = dp4a_uu(c[q], klo[g], lo[q]);
                        mid[q] = dp4a_uu(c[q], kmid[g], mid[q]);
                        hi[q] = dp4a_us(c[q], khi[g], hi[q]);
                    }
                }
                uint32_t word = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) word |= clip8(lo[q] + (mid[q] << 8) + (hi[q] << 16)) << (8 * q);
                *reinterpret_cast<uint32_t*>(dst + (size_t)yy * p.w) = word;
            }
        }
        __syncthreads();    // the buffer of tile k may be refilled (by the stage() of iteration k + 1)
        cur = nxt;
    }
}

cudaEvent_t g_resize_ev[3] = {nullptr, nullptr, nullptr};   // profiling aid: start / between the passes / end of the last call
int g_resize_timing = 0;
int g_resize_force_bytes = 0;   // tests: take the byte-load form of the horizontal pass even when the weights pack

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

int launch_resize(const polcue_resize_plan* plan, const uint8_t* const* src, int nsrc, long long image_stride, int images,
                  const uint8_t* flip, uint8_t* workspace, uint8_t* dst, cudaStream_t stream) {
    if (!plan || !plan->d_blob || !src || nsrc < 1 || nsrc > 4 || images < 0 || images % nsrc || !workspace || !dst)
        return POLCUE_EINVAL;
    for (int s = 0; s < nsrc; ++s)
        if (!src[s]) return POLCUE_EINVAL;
    if (images == 0) return POLCUE_OK;
    if (images > 65535) return POLCUE_E2BIG;
    if ((unsigned long long)plan->in_h * plan->in_w >= (1ull << 31) || (unsigned long long)kRowsH * plan->in_w > 100 * 1024)
        return POLCUE_E2BIG;
    HParams h;
    bool base16 = image_stride % 16 == 0;
    for (int s = 0; s < 4; ++s) {
        h.src[s] = src[s < nsrc ? s : 0];
        base16 = base16 && aligned(h.src[s], 16);
    }
    const bool v16 = base16 && plan->in_w % 16 == 0;
    // A 32-row tile is one contiguous run of 32 in_w bytes: always a 16-byte multiple, and 16-byte aligned when the
    // image is; only a shorter last tile needs (in_h % 32) in_w to be one as well.  (1224-wide quadrants qualify.)
    const size_t stage_budget = 110 * 1024;      // per CTA, so that two CTAs share an SM
    const int tile_rows = kStagesH * (32 * (size_t)plan->in_w + 2 * kPadH + 16) <= stage_budget ? 32 : 16;
    const bool tma_tiles = base16 && ((long long)(plan->in_h % tile_rows) * plan->in_w) % 16 == 0 &&
                           kStagesH * (tile_rows * (size_t)plan->in_w + 2 * kPadH + 16) <= stage_budget;
    h.nsrc = nsrc;
    h.image_stride = image_stride;
    h.flip = flip;
    h.dst = workspace;
    h.bounds = reinterpret_cast<const int2*>(plan->d_blob);
    h.kkT = plan->d_blob + plan->off_kk_h;
    h.pack = (plan->pack_words && !g_resize_force_bytes) ? reinterpret_cast<const uint32_t*>(plan->d_blob + plan->off_pack_h) : nullptr;
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
    if (g_resize_timing) cudaEventRecord(g_resize_ev[0], stream);
    if (h.pack && plan->out_w <= 992 && tma_tiles && (nsrc == 1 || nsrc == 4)) {      // one output column per consumer thread
        HPipeParams pp;
        pp.h = h;
        pp.tile_rows = tile_rows;
        pp.tiles_per_image = (plan->in_h + tile_rows - 1) / tile_rows;
        if ((long long)pp.tiles_per_image * images >= (1ll << 30)) return POLCUE_E2BIG;
        const int tiles_total = pp.tiles_per_image * images;
        pp.images = images;
        pp.nsrc_shift = nsrc == 4 ? 2 : 0;
        pp.buf_stride = (uint32_t)(((size_t)tile_rows * plan->in_w + 2 * kPadH + 15) / 16 * 16);
        const size_t smem2 = kStagesH * (size_t)pp.buf_stride;
        const int threads_p = (plan->out_w + 31) / 32 * 32 + 32;                  // + the producer warp
        auto launch_p = [&](auto kern) -> int {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
            int per_sm = 0;
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads_p, smem2);
            if (e != cudaSuccess) return (int)e;
            if (per_sm < 1) return POLCUE_ERANGE;
            const int grid_p = std::min(tiles_total, per_sm * device_info().sms);     // persistent: every CTA resident
            pp.step_img = grid_p / pp.tiles_per_image;
            pp.step_t = grid_p % pp.tiles_per_image;
            kern<<<grid_p, threads_p, smem2, stream>>>(pp);
            return launch_status();
        };
        const bool row4 = plan->in_w % 4 == 0;
        switch (plan->pack_words) {
#define POLCUE_H_CASE(K) case K: rc = row4 ? launch_p(resize_h_pipe_kernel<K, true>) : launch_p(resize_h_pipe_kernel<K, false>); break
            POLCUE_H_CASE(4); POLCUE_H_CASE(5); POLCUE_H_CASE(6); POLCUE_H_CASE(7); POLCUE_H_CASE(8);
#undef POLCUE_H_CASE
            default: return POLCUE_EINVAL;
        }
    } else {
        switch (h.ksize <= 32 ? std::max(4, (h.ksize + 3) / 4) : 0) {      // the same word count the weights were packed with
            case 4: rc = launch_h(resize_h_kernel<4>); break;
            case 5: rc = launch_h(resize_h_kernel<5>); break;
            case 6: rc = launch_h(resize_h_kernel<6>); break;
            case 7: rc = launch_h(resize_h_kernel<7>); break;
            case 8: rc = launch_h(resize_h_kernel<8>); break;
            default: rc = launch_h(resize_h_kernel<0>); break;
        }
    }
    if (rc != POLCUE_OK) return rc;
    if (g_resize_timing) cudaEventRecord(g_resize_ev[1], stream);

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
    v.pack = (plan->pack_words_v && vec == 4 && !g_resize_force_bytes) ? reinterpret_cast<const uint32_t*>(plan->d_blob + plan->off_pack_v)
                                                                         : nullptr;
    v.vec16 = (plan->out_w % 16 == 0 && aligned(workspace, 16)) ? 1 : 0;
    if (v.pack) {
        // shared-memory tile: the input rows kRowsV consecutive output rows need (first row's start .. last row's start + 4 KW)
        v.buf_bytes = (uint32_t)plan->v_tile_rows * 512u;
        const size_t smem_v = 2 * (size_t)v.buf_bytes;
        v.tiles_total = (long long)images * grid_v.x * ((plan->out_h + kTileV - 1) / kTileV);

        switch (plan->pack_words_v) {
#define POLCUE_V_CASE(K)                                                                                                  \
    case K: {                                                                                                             \
        cudaError_t e = cudaFuncSetAttribute(resize_v_dp4a_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v); \
        int per_sm = 0;                                                                                                   \
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, resize_v_dp4a_kernel<K>, 128 * kRowsV, smem_v); \
        if (e != cudaSuccess) return (int)e;                                                                              \
        if (per_sm < 1) return POLCUE_ERANGE;                                                                             \
        const int grid_pv = (int)std::min<long long>(v.tiles_total, (long long)per_sm * device_info().sms);               \
        resize_v_dp4a_kernel<K><<<grid_pv, dim3(128, kRowsV), smem_v, stream>>>(v);                                       \
        break;                                                                                                            \
    }
            POLCUE_V_CASE(1); POLCUE_V_CASE(2); POLCUE_V_CASE(3); POLCUE_V_CASE(4); POLCUE_V_CASE(5); POLCUE_V_CASE(6); POLCUE_V_CASE(7);
            POLCUE_V_CASE(8);
#undef POLCUE_V_CASE
            default: return POLCUE_EINVAL;
        }
    } else if (vec == 4) resize_v_kernel<4><<<grid_v, dim3(128, kRowsV), 0, stream>>>(v);
    else resize_v_kernel<1><<<grid_v, dim3(128, kRowsV), 0, stream>>>(v);
    if (g_resize_timing) cudaEventRecord(g_resize_ev[2], stream);
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
    // byte planes of the horizontal weights for the dp4a form: window words KW = 4 (<= 16 taps) or 8 (<= 32 taps)
    const int kw = kh <= 32 ? std::max(4, (kh + 3) / 4) : 0;      // 4-tap words of the window: 4..8
    bool packable = kw > 0;
    for (int v : plan->kk[0]) packable = packable && (v >> 16) >= -128 && (v >> 16) <= 127;
    plan->pack_words = packable ? kw : 0;
    plan->off_pack_h = plan->blob_ints;
    if (packable) plan->blob_ints += (size_t)2 * 3 * kw * out_w;
    const int kwv = (kv + 3) / 4;
    bool packable_v = kwv <= 8;
    for (int v : plan->kk[1]) packable_v = packable_v && (v >> 16) >= -128 && (v >> 16) <= 127;
    for (int y0 = 0; packable_v && y0 < out_h; y0 += kTileV) {
        const int y1 = std::min(y0 + kTileV, out_h) - 1;
        plan->v_tile_rows = std::max(plan->v_tile_rows, plan->bounds[1][2 * y1] + 4 * kwv - plan->bounds[1][2 * y0]);
    }
    if (plan->v_tile_rows * 512 > 100 * 1024) packable_v = false;     // extreme shrink factors: byte-load form
    plan->pack_words_v = packable_v ? kwv : 0;
    plan->off_pack_v = plan->blob_ints;
    if (packable_v) plan->blob_ints += (size_t)3 * kwv * out_h;
    std::vector<int> blob(plan->blob_ints, 0);
    if (packable_v) {
        uint32_t* pack = reinterpret_cast<uint32_t*>(blob.data() + plan->off_pack_v);
        for (int yy = 0; yy < out_h; ++yy)
            for (int m = 0; m < 4 * kwv; ++m) {
                const int k = m < kv ? plan->kk[1][(size_t)yy * kv + m] : 0;
                const uint32_t bytes[3] = {(uint32_t)k & 0xffu, ((uint32_t)k >> 8) & 0xffu, (uint32_t)(k >> 16) & 0xffu};
                for (int pl = 0; pl < 3; ++pl) pack[((size_t)yy * 3 + pl) * kwv + (m >> 2)] |= bytes[pl] << (8 * (m & 3));
            }
    }
    if (packable) {
        uint32_t* pack = reinterpret_cast<uint32_t*>(blob.data() + plan->off_pack_h);
        for (int dir = 0; dir < 2; ++dir)
            for (int xx = 0; xx < out_w; ++xx)
                for (int m = 0; m < 4 * kw; ++m) {
                    const int j = dir ? 4 * kw - 1 - m : m;      // mirrored images walk the taps backwards
                    const int k = j < kh ? plan->kk[0][(size_t)xx * kh + j] : 0;
                    const uint32_t bytes[3] = {(uint32_t)k & 0xffu, ((uint32_t)k >> 8) & 0xffu, (uint32_t)(k >> 16) & 0xffu};
                    for (int pl = 0; pl < 3; ++pl)
                        pack[((size_t)(dir * 3 + pl) * kw + (m >> 2)) * out_w + xx] |= bytes[pl] << (8 * (m & 3));
                }
    }
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
    return ((size_t)images * plan->in_h + kWorkspacePadRows) * plan->out_w;
}

int polcue_debug_resize_pass_times(int enable, float* ms_h, float* ms_v) {
    if (enable && !g_resize_ev[0])
        for (auto& e : g_resize_ev) cudaEventCreate(&e);
    if (g_resize_timing && ms_h && ms_v) {
        if (cudaEventSynchronize(g_resize_ev[2]) != cudaSuccess) return POLCUE_EINVAL;
        cudaEventElapsedTime(ms_h, g_resize_ev[0], g_resize_ev[1]);
        cudaEventElapsedTime(ms_v, g_resize_ev[1], g_resize_ev[2]);
    }
    g_resize_timing = enable ? 1 : 0;
    return POLCUE_OK;
}

int polcue_debug_resize_force_bytes(int on) {
    g_resize_force_bytes = on ? 1 : 0;
    return POLCUE_OK;
}

int polcue_resize_lanczos_u8(const polcue_resize_plan* plan, const uint8_t* src, int images, const uint8_t* flip,
                             uint8_t* workspace, uint8_t* dst, polcue_stream_t stream) {
    if (!plan) return POLCUE_EINVAL;
    const uint8_t* one[1] = {src};
    return launch_resize(plan, one, 1, (long long)plan->in_h * plan->in_w, images, flip, workspace, dst, (cudaStream_t)stream);
}

int polcue_loader_front_end_u8(const polcue_resize_plan* plan, const uint8_t* i0, const uint8_t* i45, const uint8_t* i90,
                               const uint8_t* i135, int B, const uint8_t* flip, const polcue_lut* lut, uint8_t* workspace,
                               uint8_t* planes, float* iun, float* xolp, float* normals, const float* xolp_mean_std, float* xolp_norm,
                               polcue_stream_t stream) {
    if (!plan || !planes || !xolp || B < 0 || (xolp_norm && !xolp_mean_std)) return POLCUE_EINVAL;
    if (B == 0) return POLCUE_OK;
    const uint8_t* four[4] = {i0, i45, i90, i135};
    int rc = launch_resize(plan, four, 4, (long long)plan->in_h * plan->in_w, 4 * B, flip, workspace, planes, (cudaStream_t)stream);
    if (rc != POLCUE_OK) return rc;
    return fused_planes_strided(planes, B, plan->out_h, plan->out_w, lut, iun, xolp, normals, (cudaStream_t)stream, xolp_norm,
                                xolp_norm ? xolp_mean_std[0] : 0.0f, xolp_norm ? xolp_mean_std[1] : 1.0f);
}

}  // extern "C"
