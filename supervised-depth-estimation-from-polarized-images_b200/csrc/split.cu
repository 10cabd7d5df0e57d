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
    unsigned long long units_total;  // B * 4 * Hs * units_per_row
    unsigned units_per_row;  // (Ws * px_bytes) / sizeof(T)
    unsigned Hs;
    size_t row_bytes;        // W * px_bytes
    size_t half_row_bytes;   // Ws * px_bytes
    size_t frame_bytes;      // H * W * px_bytes
    size_t quad_bytes;       // Hs * Ws * px_bytes
};

template <typename T>
__global__ void __launch_bounds__(256) split_pol_kernel(const SplitParams p) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.units_total; i += stride) {
        const unsigned u = (unsigned)(i % p.units_per_row);
        unsigned long long r = i / p.units_per_row;
        const unsigned y = (unsigned)(r % p.Hs);
        r /= p.Hs;
        const unsigned q = (unsigned)(r & 3);        // 0: im00, 1: im10, 2: im01, 3: im11
        const unsigned long long b = r >> 2;
        const size_t src_row = (size_t)y + ((q & 1) ? p.Hs : 0);           // im10 / im11: bottom half
        const size_t src_col = (q & 2) ? p.half_row_bytes : 0;             // im01 / im11: right half
        const T* src = reinterpret_cast<const T*>(p.img + b * p.frame_bytes + src_row * p.row_bytes + src_col) + u;
        T* dst = reinterpret_cast<T*>(p.out[q] + b * p.quad_bytes + (size_t)y * p.half_row_bytes) + u;
        *dst = *src;
    }
}

template <typename T>
int launch_split(SplitParams p, cudaStream_t s) {
    p.units_per_row = (unsigned)(p.half_row_bytes / sizeof(T));
    p.units_total = p.units_total * p.units_per_row;
    const unsigned long long want = (p.units_total + 255) / 256;
    const unsigned long long cap = (unsigned long long)device_info().sms * 32;
    split_pol_kernel<T><<<(unsigned)(want < cap ? want : cap), 256, 0, s>>>(p);
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
    p.units_total = (unsigned long long)B * 4 * p.Hs;   // rows; launch_split multiplies by units per row
    uintptr_t bits = reinterpret_cast<uintptr_t>(img) | p.half_row_bytes;
    for (int q = 0; q < 4; ++q) bits |= reinterpret_cast<uintptr_t>(p.out[q]);
    cudaStream_t s = (cudaStream_t)stream;
    if ((bits & 15) == 0) return launch_split<uint4>(p, s);
    if ((bits & 7) == 0) return launch_split<uint2>(p, s);
    if ((bits & 3) == 0) return launch_split<uint32_t>(p, s);
    if ((bits & 1) == 0) return launch_split<uint16_t>(p, s);
    return launch_split<uint8_t>(p, s);
}
