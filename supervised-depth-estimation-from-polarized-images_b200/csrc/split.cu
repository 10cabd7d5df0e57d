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
    uint32_t units_total;    // B * H * 2 * units_per_half_row  (< 2^31)
    FastDiv units_per_row;   // 2 * units_per_half_row : units in one full input row
    FastDiv rows_per_frame;  // H
    uint32_t upr_half;       // (Ws * px_bytes) / sizeof(T)
    uint32_t Hs;
    size_t half_row_bytes;   // Ws * px_bytes
    size_t quad_bytes;       // Hs * Ws * px_bytes
};

constexpr int kSplitThreads = 256;

// The input is read strictly linearly (unit i is bytes [i, i+1) * sizeof(T) of the batch), so DRAM sees one
// sequential read stream; each unit is scattered to its quadrant, whose rows are contiguous runs of half a row.
// A thread keeps ITEMS independent copies (64 bytes) in flight; one CTA per tile, launched plainly.
template <typename T>
__global__ void __launch_bounds__(kSplitThreads) split_pol_kernel(const SplitParams p) {
    constexpr int kItems = sizeof(T) >= 16 ? 4 : (sizeof(T) >= 8 ? 8 : 16);
    const uint32_t base = blockIdx.x * (kSplitThreads * kItems) + threadIdx.x;
    const T* src = reinterpret_cast<const T*>(p.img);
    T v[kItems];
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const uint32_t i = base + k * kSplitThreads;
        if (i < p.units_total) v[k] = src[i];
    }
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const uint32_t i = base + k * kSplitThreads;
        if (i < p.units_total) {
            const uint32_t row = fastdiv(i, p.units_per_row);
            uint32_t u = i - row * p.units_per_row.div;
            const uint32_t b = fastdiv(row, p.rows_per_frame);
            uint32_t y = row - b * p.rows_per_frame.div;
            const bool right = u >= p.upr_half, bottom = y >= p.Hs;
            u -= right ? p.upr_half : 0;
            y -= bottom ? p.Hs : 0;
            // return order of split_pol: im00 (top-left), im10 (bottom-left), im01 (top-right), im11 (bottom-right)
            unsigned char* q = right ? (bottom ? p.out[3] : p.out[2]) : (bottom ? p.out[1] : p.out[0]);
            reinterpret_cast<T*>(q + (size_t)b * p.quad_bytes + (size_t)y * p.half_row_bytes)[u] = v[k];
        }
    }
}

template <typename T>
int launch_split(SplitParams p, unsigned long long rows, cudaStream_t s) {
    const unsigned long long upr_half = p.half_row_bytes / sizeof(T);
    const unsigned long long units = rows * 2 * upr_half;      // rows = B * H input rows
    if (units >= (1ull << 31)) return POLCUE_E2BIG;
    p.units_total = (uint32_t)units;
    p.upr_half = (uint32_t)upr_half;
    p.units_per_row.div = (uint32_t)(2 * upr_half);
    make_fastdiv(p.units_per_row.div, p.units_per_row.mul, p.units_per_row.shift);
    p.rows_per_frame.div = 2 * p.Hs;
    make_fastdiv(p.rows_per_frame.div, p.rows_per_frame.mul, p.rows_per_frame.shift);
    const unsigned per_cta = kSplitThreads * (sizeof(T) >= 16 ? 4 : (sizeof(T) >= 8 ? 8 : 16));
    split_pol_kernel<T><<<(unsigned)((units + per_cta - 1) / per_cta), kSplitThreads, 0, s>>>(p);
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
    p.half_row_bytes = (size_t)W * px_bytes / 2;
    p.quad_bytes = (size_t)p.Hs * p.half_row_bytes;
    const unsigned long long rows = (unsigned long long)B * H;          // input rows
    uintptr_t bits = reinterpret_cast<uintptr_t>(img) | p.half_row_bytes;
    for (int q = 0; q < 4; ++q) bits |= reinterpret_cast<uintptr_t>(p.out[q]);
    cudaStream_t s = (cudaStream_t)stream;
    if ((bits & 15) == 0) return launch_split<uint4>(p, rows, s);
    if ((bits & 7) == 0) return launch_split<uint2>(p, rows, s);
    if ((bits & 3) == 0) return launch_split<uint32_t>(p, rows, s);
    if ((bits & 1) == 0) return launch_split<uint16_t>(p, rows, s);
    return launch_split<uint8_t>(p, rows, s);
}
