// Quadrant split of the stored 2x2-tiled polarizer image (the reference's "demosaic"), bit-exact.
//
// Replaces polarisation/pol_split_and_save.py:10-27:  split_pol(img) -> (im00, im10, im01, im11)
//   im00 = img[:H/2, :W/2]  im10 = img[H/2:, :W/2]  im01 = img[:H/2, W/2:]  im11 = img[H/2:, W/2:]
// The reference returns numpy views; a device library has to materialise them, which is a strided 2-D
// copy: every output row is a contiguous run of (W/2)*px_bytes input bytes.  Moved in the widest unit
// (16, 8, 4, 2 or 1 bytes) that divides the run and the addresses.  Roofline: HBM, 1 B read + 1 B written
// per input byte.
#include "polcue_device.cuh"
#include "polcue_host.h"

namespace polcue {
namespace {

struct SplitParams {
    const unsigned char* img;
    unsigned char* out[4];   // im00, im10, im01, im11
    uint32_t rows_total;     // B * 4 * Hs output rows
    uint32_t units_per_row;  // (Ws * px_bytes) / sizeof(T)
    FastDiv rows_per_quad;   // Hs
    uint32_t Hs;
    size_t row_bytes;        // W * px_bytes
    size_t half_row_bytes;   // Ws * px_bytes
    size_t frame_bytes;      // H * W * px_bytes
    size_t quad_bytes;       // Hs * Ws * px_bytes
};

constexpr int kSplitThreads = 256, kSplitRowsPerWarp = 4, kSplitUnroll = 4;
constexpr int kSplitRowsPerCta = (kSplitThreads / 32) * kSplitRowsPerWarp;

// A warp copies whole output rows: the (frame, quadrant, row) decode is done once per row (warp-uniform), the
// lanes then stream the row's units with kSplitUnroll independent copies in flight.  One CTA per 32 rows,
// launched plainly so the hardware CTA queue balances the SMs.
template <typename T>
__global__ void __launch_bounds__(kSplitThreads) split_pol_kernel(const SplitParams p) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t row0 = (blockIdx.x * (kSplitThreads / 32) + warp) * kSplitRowsPerWarp;
#pragma unroll 1
    for (uint32_t k = 0; k < kSplitRowsPerWarp; ++k) {
        const uint32_t r = row0 + k;
        if (r >= p.rows_total) return;
        const uint32_t r2 = fastdiv(r, p.rows_per_quad);
        const uint32_t y = r - r2 * p.Hs;
        const uint32_t q = r2 & 3;        // 0: im00, 1: im10, 2: im01, 3: im11
        const uint32_t b = r2 >> 2;
        const size_t src_row = (size_t)y + ((q & 1) ? p.Hs : 0);           // im10 / im11: bottom half
        const size_t src_col = (q & 2) ? p.half_row_bytes : 0;             // im01 / im11: right half
        const T* src = reinterpret_cast<const T*>(p.img + (size_t)b * p.frame_bytes + src_row * p.row_bytes + src_col);
        T* dst = reinterpret_cast<T*>(p.out[q] + (size_t)b * p.quad_bytes + (size_t)y * p.half_row_bytes);
        for (uint32_t u0 = lane; u0 < p.units_per_row; u0 += 32 * kSplitUnroll) {
            T v[kSplitUnroll];
#pragma unroll
            for (int j = 0; j < kSplitUnroll; ++j)
                if (u0 + 32 * j < p.units_per_row) v[j] = src[u0 + 32 * j];
#pragma unroll
            for (int j = 0; j < kSplitUnroll; ++j)
                if (u0 + 32 * j < p.units_per_row) dst[u0 + 32 * j] = v[j];
        }
    }
}

template <typename T>
int launch_split(SplitParams p, unsigned long long rows, cudaStream_t s) {
    const unsigned long long upr = p.half_row_bytes / sizeof(T);
    if (rows >= (1ull << 31) || upr >= (1ull << 31)) return POLCUE_E2BIG;
    p.rows_total = (uint32_t)rows;
    p.units_per_row = (uint32_t)upr;
    p.rows_per_quad.div = p.Hs;
    make_fastdiv(p.rows_per_quad.div, p.rows_per_quad.mul, p.rows_per_quad.shift);
    split_pol_kernel<T><<<(unsigned)((rows + kSplitRowsPerCta - 1) / kSplitRowsPerCta), kSplitThreads, 0, s>>>(p);
    return launch_status();
}

}  // namespace
}  // namespace polcue

using namespace polcue;

extern "C" int polcue_split_pol(const void* img, int B, int H, int W, int px_bytes, void* im00, void* im10, void* im01,
                                void* im11, polcue_stream_t stream) {
    if (!img || !im00 || !im10 || !im01 || !im11 || B < 0 || H <= 0 || W <= 0 || px_bytes <= 0) return POLCUE_EINVAL;
    if ((H & 1) || (W & 1)) return POLCUE_EINVAL;   // np.split raises on an unequal division
    if (B == 0) return POLCUE_OK;
    SplitParams p;
    p.img = static_cast<const unsigned char*>(img);
    p.out[0] = static_cast<unsigned char*>(im00);
    p.out[1] = static_cast<unsigned char*>(im10);
    p.out[2] = static_cast<unsigned char*>(im01);
    p.out[3] = static_cast<unsigned char*>(im11);
    p.Hs = (unsigned)(H / 2);
    p.row_bytes = (size_t)W * px_bytes;
    p.half_row_bytes = p.row_bytes / 2;
    p.frame_bytes = (size_t)H * p.row_bytes;
    p.quad_bytes = (size_t)p.Hs * p.half_row_bytes;
    const unsigned long long rows = (unsigned long long)B * 4 * p.Hs;   // output rows
    uintptr_t bits = reinterpret_cast<uintptr_t>(img) | p.half_row_bytes;
    for (int q = 0; q < 4; ++q) bits |= reinterpret_cast<uintptr_t>(p.out[q]);
    cudaStream_t s = (cudaStream_t)stream;
    if ((bits & 15) == 0) return launch_split<uint4>(p, rows, s);
    if ((bits & 7) == 0) return launch_split<uint2>(p, rows, s);
    if ((bits & 3) == 0) return launch_split<uint32_t>(p, rows, s);
    if ((bits & 1) == 0) return launch_split<uint16_t>(p, rows, s);
    return launch_split<uint8_t>(p, rows, s);
}
