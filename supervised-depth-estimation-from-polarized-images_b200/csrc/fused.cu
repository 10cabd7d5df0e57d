// Fused polarization-cue kernels: quadrant split -> Stokes -> Iun/DoLP/AoLP -> normal candidates,
// plus the XOLP-only and normals-only members of the same family.
//
// Roofline: HBM.  Algorithmic bytes per output pixel: fused 4 R + 44 W = 48 B
// (+4 Iun, +4 planes); XOLP-only 4 R + 8 W; normals-from-XOLP 8 R + 36 W.  Nothing here is a
// contraction, so no tensor cores: the work is MUFU/FMA per pixel and coalesced 128-bit stores.
//
// Layout: every output is planar (B x C x Hs x Ws) so each of the 11 float planes is written as
// one contiguous 512-byte run per warp instruction.  A thread owns VEC consecutive pixels of one
// row: it reads VEC bytes from each of the four quadrants (one 32-bit load each for VEC = 4) and
// writes one 16-byte vector per plane.  One CTA per tile of 256 pixel groups; resident CTAs take over
// pending tiles through cluster launch control (polcue_device.cuh), so the zenith tables (49 KB for
// n = 1.5) are staged once per RESIDENT CTA into shared memory with a single bulk-async (TMA) copy.
#include <cfloat>

#include "polcue_device.cuh"
#include "polcue_host.h"

namespace polcue {

std::atomic<unsigned long long> g_launches{0};

const DeviceInfo& device_info() {
    static DeviceInfo cache[64];
    static std::atomic<int> ready[64];
    int dev = 0;
    cudaGetDevice(&dev);
    dev = (dev < 0 || dev >= 64) ? 0 : dev;
    if (!ready[dev].load(std::memory_order_acquire)) {
        DeviceInfo d;
        cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cache[dev] = d;
        ready[dev].store(1, std::memory_order_release);
    }
    return cache[dev];
}

namespace {

#ifndef POLCUE_FUSED_THREADS
#define POLCUE_FUSED_THREADS 256
#endif
constexpr int kFusedThreads = POLCUE_FUSED_THREADS;            // pixel groups per tile = threads per CTA
constexpr int kFusedMinBlocks = 1024 / kFusedThreads;           // 1024 resident threads per SM at <= 64 registers

struct LutArgs {
    const float4* blob;
    uint32_t bytes;
    int offset[3];
    float scale[3];
    float last[3];
    LutSteep steep;
};

LutArgs lut_args(const polcue_lut* lut) {
    LutArgs a;
    a.blob = lut->d_blob;
    a.bytes = (uint32_t)lut->bytes();
    for (int t = 0; t < 3; ++t) {
        a.offset[t] = lut->offset[t];
        a.scale[t] = lut->scale[t];
        a.last[t] = (float)lut->cells[t] - 0.5f;
    }
    a.steep.mask = lut->steep_mask;
    a.steep.from = INFINITY;
    for (int t = 0; t < 3; ++t) {
        a.steep.x[t] = lut->steep_x[t];
        a.steep.y[t] = lut->steep_y[t];
        a.steep.slope[t] = lut->steep_slope[t];
        if ((lut->steep_mask >> t) & 1) {
            float f = (float)lut->steep_x[t];
            if ((double)f > lut->steep_x[t]) f = nextafterf(f, -INFINITY);   // round DOWN: never miss a query
            a.steep.from = fminf(a.steep.from, f);
        }
    }
    return a;
}

__device__ __forceinline__ LutShared lut_shared(const void* smem_base, const LutArgs& a) {
    LutShared v;
    for (int t = 0; t < 3; ++t) {
        uint32_t addr = smem_u32(smem_base) + 16u * (uint32_t)a.offset[t];
        asm volatile("" : "+r"(addr));   // keep the three table bases as values: lookup = one LEA, not add + LEA
        v.addr[t] = addr;
        v.scale[t] = a.scale[t];
        v.last[t] = a.last[t];
    }
    return v;
}

__device__ __forceinline__ LutView lut_view(const float4* base, const LutArgs& a) {
    LutView v;
    for (int t = 0; t < 3; ++t) {
        v.cells[t] = base + a.offset[t];
        v.scale[t] = a.scale[t];
        v.last[t] = a.last[t];
    }
    return v;
}

// ---------------------------------------------------------------------------------------------
// Fused kernel
// ---------------------------------------------------------------------------------------------
struct FusedParams {
    const uint8_t* mosaic;   // address of the 0-degree sample of pixel (0,0) of frame 0
    long long off45, off90, off135;   // byte offsets from a pixel's 0-degree sample to its 45 / 90 / 135-degree samples:
                                      // quadrant mosaic: Ws, Hs*W, Hs*W + Ws;  four separate planes: pointer differences
    uint8_t* planes;   // may be null
    float* iun;        // may be null
    float* xolp;
    float* xolp_norm;  // may be null: (xolp - norm_mean) / norm_std, ShallowEncoder.normalizeInput (pre_encoders.py:75-83)
    float norm_mean, norm_inv_std;
    float* normals;    // may be null
    LutArgs lut;
    uint32_t groups_total;   // B * Hs * (Ws / VEC)
    FastDiv groups_per_frame, groups_per_row;
    int superpixel;          // 0: samples are `offXX` apart (quadrants / planes); 1: interleaved 2x2 super-pixels
    int angle_at[4];         // superpixel only: angle index (0..3 = 0, 45, 90, 135 deg) at (0,0), (0,1), (1,0), (1,1)
    uint32_t W;              // input row stride in bytes (mosaic width, or Ws for separate planes)
    uint32_t row_px;         // output pixels per row (Ws)
    uint32_t frame_bytes;    // input frame stride in bytes                 (all strides < 2^32, checked on the host)
    uint32_t plane;          // Hs * Ws
    uint32_t plane_bytes;    // 4 * Hs * Ws
    float* tile_records;     // STATS kernels only: [tiles][kSumRecord] per-tile sums (rho, phi, 9 normal channels, rho^2, phi^2)
};

template <int VEC>
struct Packed;
template <>
struct Packed<4> {
    using type = uint32_t;
    static __device__ __forceinline__ type load(const uint8_t* p) { return ld_stream_u32(p); }
};
template <>
struct Packed<2> {
    using type = uint16_t;
    static __device__ __forceinline__ type load(const uint8_t* p) { return ld_stream_u16(p); }
};
template <>
struct Packed<1> {
    using type = uint8_t;
    static __device__ __forceinline__ type load(const uint8_t* p) { return ld_stream_u8(p); }
};

// The four quadrant words of one pixel group, plus where the group lives.
template <int VEC>
struct GroupIn {
    typename Packed<VEC>::type w0, w45, w90, w135;
    uint32_t b, rem;   // frame, group index inside the frame
    bool valid;
};

template <int VEC>
__device__ __forceinline__ GroupIn<VEC> load_group(const FusedParams& p, uint32_t tile) {
    using PK = Packed<VEC>;
    GroupIn<VEC> g;
    const uint32_t gid = tile * kFusedThreads + threadIdx.x;
    g.valid = gid < p.groups_total;
    g.b = 0; g.rem = 0;
    g.w0 = g.w45 = g.w90 = g.w135 = 0;
    if (g.valid) {
        g.b = fastdiv(gid, p.groups_per_frame);
        g.rem = gid - g.b * p.groups_per_frame.div;
        const uint32_t y = fastdiv(g.rem, p.groups_per_row);
        const uint32_t xg = g.rem - y * p.groups_per_row.div;
        if (!p.superpixel) {
            const uint8_t* src = p.mosaic + ((size_t)g.b * p.frame_bytes + (y * p.W + xg * VEC));
            g.w0 = PK::load(src);                  // TL:   0 deg
            g.w45 = PK::load(src + p.off45);       // TR:  45 deg
            g.w90 = PK::load(src + p.off90);       // BL:  90 deg
            g.w135 = PK::load(src + p.off135);     // BR: 135 deg
        } else {
            // raw sensor layout: output pixel (y, x) owns mosaic bytes (2y, 2x), (2y, 2x+1), (2y+1, 2x), (2y+1, 2x+1)
            const uint8_t* src = p.mosaic + ((size_t)g.b * p.frame_bytes + ((2 * y) * p.W + 2 * xg * VEC));
            uint32_t pos[4] = {0, 0, 0, 0};        // VEC samples of each super-pixel position, packed like Packed<VEC>
            if constexpr (VEC == 4) {
                const uint2 r0 = *reinterpret_cast<const uint2*>(src), r1 = *reinterpret_cast<const uint2*>(src + p.W);
                pos[0] = __byte_perm(r0.x, r0.y, 0x6420);   // even bytes of row 2y
                pos[1] = __byte_perm(r0.x, r0.y, 0x7531);   // odd bytes
                pos[2] = __byte_perm(r1.x, r1.y, 0x6420);
                pos[3] = __byte_perm(r1.x, r1.y, 0x7531);
            } else {
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    pos[0] |= (uint32_t)ld_stream_u8(src + 2 * j) << (8 * j);
                    pos[1] |= (uint32_t)ld_stream_u8(src + 2 * j + 1) << (8 * j);
                    pos[2] |= (uint32_t)ld_stream_u8(src + p.W + 2 * j) << (8 * j);
                    pos[3] |= (uint32_t)ld_stream_u8(src + p.W + 2 * j + 1) << (8 * j);
                }
            }
            uint32_t ang[4] = {0, 0, 0, 0};
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int a = 0; a < 4; ++a)
                    if (p.angle_at[k] == a) ang[a] = pos[k];    // warp-uniform selects
            g.w0 = (typename PK::type)ang[0];
            g.w45 = (typename PK::type)ang[1];
            g.w90 = (typename PK::type)ang[2];
            g.w135 = (typename PK::type)ang[3];
        }
    }
    return g;
}

// Pixels whose rho may lie on a steep end segment (LutSteep; bit j of `bits` = pixel pix + j of frame b): the samples are
// read again, rho is formed in float64 and the candidates of the steep tables are stored over the float32 ones (same
// thread, same addresses, program order).  Out of line so the hot loop carries one flag register for it.
__device__ __noinline__ void steep_redo(const FusedParams& p, uint32_t b, uint32_t pix, uint32_t bits) {
    for (int j = 0; j < 4; ++j) {
        if (!((bits >> j) & 1)) continue;
        const uint32_t px = pix + j;
        const uint32_t ws = p.superpixel ? p.W / 2 : p.row_px;
        const uint32_t y = px / ws, x = px - y * ws;
        float i[4];
        if (!p.superpixel) {
            const uint8_t* src = p.mosaic + ((size_t)b * p.frame_bytes + (y * p.W + x));
            i[0] = (float)src[0]; i[1] = (float)src[p.off45]; i[2] = (float)src[p.off90]; i[3] = (float)src[p.off135];
        } else {
            const uint8_t* src = p.mosaic + ((size_t)b * p.frame_bytes + ((2 * y) * p.W + 2 * x));
            const float pos[4] = {(float)src[0], (float)src[1], (float)src[p.W], (float)src[p.W + 1]};
            for (int k = 0; k < 4; ++k) i[p.angle_at[k]] = pos[k];
        }
        const Cues q = cues_from_u8<true>(i[0], i[1], i[2], i[3]);
        const double rho = rho_exact_u8(i[0], i[1], i[2], i[3]);
        float* no = p.normals + ((size_t)(9 * b) * p.plane + px);
        for (int t = 0; t < 3; ++t) {
            if (!((p.lut.steep.mask >> t) & 1) || !(rho > p.lut.steep.x[t])) continue;
            const float2 sc = steep_sincos(steep_theta(p.lut.steep, t, rho));
            no[(size_t)(3 * t + 0) * p.plane] = (t == 0 ? q.cos_phi : -q.sin_phi) * sc.x;
            no[(size_t)(3 * t + 1) * p.plane] = (t == 0 ? q.sin_phi : q.cos_phi) * sc.x;
            no[(size_t)(3 * t + 2) * p.plane] = sc.y;
        }
    }
}

constexpr int kStatValues = 13;   // sums of rho, phi and the nine normal channels, then the squares of rho and phi

template <int VEC, bool MUFU, bool NORMALS, bool STATS = false>
__device__ __forceinline__ void process_group(const FusedParams& p, const LutShared& lut, const GroupIn<VEC>& g, float* part = nullptr) {
    using PK = Packed<VEC>;
    const uint32_t b = g.b;
    const uint32_t pix = g.rem * VEC;   // first pixel inside the Hs x Ws plane (Ws = groups_per_row * VEC)

    if (p.planes) {
        typename PK::type* dst = reinterpret_cast<typename PK::type*>(p.planes + (size_t)b * 4 * p.plane + pix);
        const size_t ps = p.plane / VEC;    // plane stride in packed words
        dst[0] = g.w0;
        dst[ps] = g.w45;
        dst[2 * ps] = g.w90;
        dst[3 * ps] = g.w135;
    }

    float rho[VEC], phi[VEC], iun[VEC], sp[VEC], cp[VEC];
    if constexpr (VEC >= 2) {       // two pixels per packed-FP32 instruction (bit-identical to the scalar form).  Measured: the
                                    // software-pipelined kernels gain (XOLP-only 0.75 -> 0.80 of the HBM peak, the full kernel 1 % under the
                                    // power cap); the plain XOLP kernels below lose ILP with it (0.92 -> 0.85) and stay scalar.
#pragma unroll
        for (int j = 0; j < VEC; j += 2) {
            Cues qa, qb;
            cues2_from_u8<NORMALS>(bytes_to_float2(g.w0, j), bytes_to_float2(g.w45, j), bytes_to_float2(g.w90, j),
                                   bytes_to_float2(g.w135, j), qa, qb);
            rho[j] = qa.rho; phi[j] = qa.phi; iun[j] = qa.iun; sp[j] = qa.sin_phi; cp[j] = qa.cos_phi;
            rho[j + 1] = qb.rho; phi[j + 1] = qb.phi; iun[j + 1] = qb.iun; sp[j + 1] = qb.sin_phi; cp[j + 1] = qb.cos_phi;
        }
    } else {
        const Cues q = cues_from_u8<NORMALS>(byte_to_float(g.w0, 0), byte_to_float(g.w45, 0), byte_to_float(g.w90, 0),
                                             byte_to_float(g.w135, 0));
        rho[0] = q.rho;
        phi[0] = q.phi;
        iun[0] = q.iun;
        sp[0] = q.sin_phi;
        cp[0] = q.cos_phi;
    }
    // plane pointers advance by a 32-bit stride
    float* xo = p.xolp + ((size_t)(2 * b) * p.plane + pix);
    asm volatile("" : "+l"(xo));   // keep a pointer VALUE (not base + 64-bit index) across the plane steps
    st_stream_vec<VEC>(xo, rho);
    xo += p.plane;
    st_stream_vec<VEC>(xo, phi);
    if (p.iun) st_stream_vec<VEC>(p.iun + ((size_t)b * p.plane + pix), iun);
    if (p.xolp_norm) {      // the encoder's input normalisation folded into the store, rounded as torch's CUDA kernels round it:
                            // x - float(mean), then times float(1 / std) (BinaryDivTrueKernel multiplies by the reciprocal of a scalar)
        float nr[VEC], nf[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            nr[j] = __fmul_rn(__fsub_rn(rho[j], p.norm_mean), p.norm_inv_std);
            nf[j] = __fmul_rn(__fsub_rn(phi[j], p.norm_mean), p.norm_inv_std);
        }
        float* xn = p.xolp_norm + ((size_t)(2 * b) * p.plane + pix);
        st_stream_vec<VEC>(xn, nr);
        st_stream_vec<VEC>(xn + p.plane, nf);
    }

    if constexpr (NORMALS) {
        float nrm[9][VEC];
        uint32_t steep_bits = 0;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float n9[9];
            normals_from_trig<MUFU>(lut, rho[j], sp[j], cp[j], n9);
            steep_bits |= (rho[j] >= p.lut.steep.from ? 1u : 0u) << j;
#pragma unroll
            for (int c = 0; c < 9; ++c) nrm[c][j] = n9[c];
        }
        float* no = p.normals + ((size_t)(9 * b) * p.plane + pix);
        asm volatile("" : "+l"(no));
#pragma unroll
        for (int c = 0; c < 9; ++c) {
            st_stream_vec<VEC>(no, nrm[c]);
            if constexpr (STATS && VEC == 4) part[2 + c] = group_sum4(nrm[c][0], nrm[c][1], nrm[c][2], nrm[c][3]);
            no += p.plane;
        }
        if (steep_bits) steep_redo(p, b, pix, steep_bits);   // rare: float64 evaluation of a steep end segment (never with STATS)
    }
    if constexpr (STATS && VEC == 4) {      // group partials in the canonical order (polcue_device.cuh)
        part[0] = group_sum4(rho[0], rho[1], rho[2], rho[3]);
        part[1] = group_sum4(phi[0], phi[1], phi[2], phi[3]);
        part[11] = group_squares4(rho[0], rho[1], rho[2], rho[3]);
        part[12] = group_squares4(phi[0], phi[1], phi[2], phi[3]);
    }
}

// Statistics by-product: the 13 group partials of this thread -> the eight warp sums of the tile, canonical order.  Per
// warp one butterfly level in registers (xor 16), the 16 surviving lane sums of every value go through a 832-byte shared
// slab, lane v finishes the tree of value v and leaves the warp's sum in warp_sums[warp][v]; after a barrier the first 13
// threads add the eight warp sums in order (tile_record_flush).  (Deferring that fold past the loop's own barrier, to save
// the extra barrier, was measured 1.7x SLOWER: 1.18 vs 0.69 ms per 64 frames.)
constexpr int kSlabValues = 7;      // the slab holds 7 of the 13 values at a time (two passes): 3.5 KB instead of 6.5 KB per CTA,
                                    // which keeps four CTAs per SM next to the 49 KB table
__device__ __forceinline__ void tile_warp_sums(float (&part)[kStatValues], float (*slab)[kSlabValues][16], float (*warp_sums)[16]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int v = 0; v < kStatValues; ++v) part[v] = __fadd_rn(part[v], __shfl_xor_sync(0xffffffffu, part[v], 16));
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int v0 = pass * kSlabValues, nv = pass ? kStatValues - kSlabValues : kSlabValues;
        if (lane < 16) {
#pragma unroll
            for (int v = 0; v < kSlabValues; ++v)
                if (v < nv) slab[warp][v][lane] = part[v0 + v];
        }
        __syncwarp();
        if (lane < nv) {
            const float4* row = reinterpret_cast<const float4*>(&slab[warp][lane][0]);
            const float4 a = row[0], b = row[1], c = row[2], d = row[3];
            // xor 8: l + (l ^ 8); xor 4; xor 2; xor 1
            const float e0 = __fadd_rn(a.x, c.x), e1 = __fadd_rn(a.y, c.y), e2 = __fadd_rn(a.z, c.z), e3 = __fadd_rn(a.w, c.w);
            const float e4 = __fadd_rn(b.x, d.x), e5 = __fadd_rn(b.y, d.y), e6 = __fadd_rn(b.z, d.z), e7 = __fadd_rn(b.w, d.w);
            const float f0 = __fadd_rn(e0, e4), f1 = __fadd_rn(e1, e5), f2 = __fadd_rn(e2, e6), f3 = __fadd_rn(e3, e7);
            const float g0 = __fadd_rn(f0, f2), g1 = __fadd_rn(f1, f3);
            warp_sums[warp][v0 + lane] = __fadd_rn(g0, g1);
        }
        __syncwarp();        // the slab is rewritten by the next pass / the next tile
    }
}

__device__ __forceinline__ void tile_record_flush(const float (*warp_sums)[16], float* record) {
    if (threadIdx.x < kStatValues) {
        float t = warp_sums[0][threadIdx.x];
#pragma unroll
        for (int w = 1; w < 8; ++w) t = __fadd_rn(t, warp_sums[w][threadIdx.x]);
        record[threadIdx.x] = t;
    }
}

// One tile = one group of VEC pixels per thread (512 groups per CTA).  Tiles arrive in order through cluster launch
// control; the loop is software-pipelined two deep: while tile k is computed and stored, the quadrant words of tile
// k+1 are already in flight and the query for tile k+2 is outstanding, so no warp waits on DRAM latency.
template <int VEC, bool MUFU, bool NORMALS, bool STATS = false>
__global__ void __launch_bounds__(kFusedThreads, kFusedMinBlocks) fused_mosaic_kernel(const __grid_constant__ FusedParams p) {
    static_assert(!STATS || (VEC == 4 && NORMALS && kFusedThreads == kSumTileGroups), "the statistics by-product needs 4-pixel groups, 256 per tile");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ __align__(16) uint4 clc_resp;
    __shared__ uint64_t clc_bar;
    if constexpr (NORMALS) lut_stage_begin(smem_raw, p.lut.blob, p.lut.bytes, &bar);
    ClcTiles clc;
    clc.init(&clc_resp, &clc_bar);
    clc.prefetch();                                      // query for the tile after blockIdx.x
    uint32_t cur_tile = blockIdx.x;
    GroupIn<VEC> cur = load_group<VEC>(p, cur_tile);
    if constexpr (NORMALS) lut_stage_wait(&bar);
    const LutShared lut = lut_shared(smem_raw, p.lut);

    for (;;) {
        uint32_t next_tile;
        const bool more = clc.next(next_tile);           // response to the query issued one tile ago
        __syncthreads();                                 // every thread has read it: the slot may be re-armed
        GroupIn<VEC> nxt;
        nxt.valid = false;
        if (more) {
            clc.prefetch();
            nxt = load_group<VEC>(p, next_tile);         // in flight while `cur` is processed
        }
        if constexpr (STATS) {
            __shared__ __align__(16) float slab[kFusedThreads / 32][kSlabValues][16];
            __shared__ float warp_sums[kFusedThreads / 32][16];
            float part[kStatValues];
#pragma unroll
            for (int v = 0; v < kStatValues; ++v) part[v] = 0.0f;       // threads past the end of the batch add +0
            if (cur.valid) process_group<VEC, MUFU, NORMALS, true>(p, lut, cur, part);
            tile_warp_sums(part, slab, warp_sums);
            __syncthreads();                                             // (the loop's own barrier protects warp_sums' reuse)
            tile_record_flush(warp_sums, p.tile_records + (size_t)cur_tile * kSumRecord);
        } else {
            if (cur.valid) process_group<VEC, MUFU, NORMALS>(p, lut, cur);
        }
        if (!more) break;
        cur = nxt;
        cur_tile = next_tile;
    }
}

// XOLP only (no normals wanted): nothing to stage per CTA and a third of the arithmetic, so the tile machinery above costs
// more than it hides (0.73 of the HBM peak).  A plain launch instead: four groups per thread, all sixteen quadrant words
// requested before the first is used, the hardware's own CTA queue as the scheduler (as the XOLP kernels below).
constexpr int kStreamItems = 4;
template <int VEC>
__global__ void __launch_bounds__(kFusedThreads) fused_xolp_stream_kernel(const __grid_constant__ FusedParams p) {
    GroupIn<VEC> g[kStreamItems];
#pragma unroll
    for (int item = 0; item < kStreamItems; ++item) g[item] = load_group<VEC>(p, blockIdx.x * kStreamItems + item);
    const LutShared none{};
#pragma unroll
    for (int item = 0; item < kStreamItems; ++item)
        if (g[item].valid) process_group<VEC, true, false>(p, none, g[item]);
}

// `mufu`: zenith-angle sincos of this launch, a property of the caller's table handle (polcue_lut_set_trig): MUFU sin/cos
// (3.6e-7 abs, default) or the polynomial (1.4e-7).  DESIGN.md 5.
template <int VEC>
int launch_fused(const FusedParams& p, size_t smem, cudaStream_t stream, bool mufu) {
    if (!p.normals) {
        const uint32_t tiles = (p.groups_total + kFusedThreads - 1) / kFusedThreads;
        fused_xolp_stream_kernel<VEC><<<(tiles + kStreamItems - 1) / kStreamItems, kFusedThreads, 0, stream>>>(p);
        return launch_status();
    }
    auto kern = !p.normals ? fused_mosaic_kernel<VEC, true, false>
                           : (mufu ? fused_mosaic_kernel<VEC, true, true> : fused_mosaic_kernel<VEC, false, true>);
    if constexpr (VEC == 4) {
        if (p.tile_records) kern = mufu ? fused_mosaic_kernel<4, true, true, true> : fused_mosaic_kernel<4, false, true, true>;
    }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kFusedThreads, smem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) return POLCUE_ERANGE;
    const uint32_t tiles = (p.groups_total + kFusedThreads - 1) / kFusedThreads;   // one CTA per tile (see ClcTiles)
    kern<<<tiles, kFusedThreads, smem, stream>>>(p);
    return launch_status();
}

// ---------------------------------------------------------------------------------------------
// XOLP-only kernels (loader path): interleaved H x W x 4 stacks or four planes.
// ---------------------------------------------------------------------------------------------
struct Pinv {
    float m[12];
};

// General angles: x = pinv(A) I, then xolp.py:22-30 with its inf/nan scrub.
__device__ __forceinline__ Cues cues_general(const float (&i)[4], const Pinv& pv) {
    const float x0 = fmaf(pv.m[3], i[3], fmaf(pv.m[2], i[2], fmaf(pv.m[1], i[1], pv.m[0] * i[0])));
    const float x1 = fmaf(pv.m[7], i[3], fmaf(pv.m[6], i[2], fmaf(pv.m[5], i[1], pv.m[4] * i[0])));
    const float x2 = fmaf(pv.m[11], i[3], fmaf(pv.m[10], i[2], fmaf(pv.m[9], i[1], pv.m[8] * i[0])));
    const float amp = sqrtf(fmaf(x1, x1, x2 * x2));
    float rho = amp / x0;                             // IEEE: 0/0 = nan, a/0 = +-inf like numpy
    if (isnan(rho) || rho == INFINITY) rho = 0.0f;    // xolp.py:28-29
    else if (rho == -INFINITY) rho = -FLT_MAX;        // np.nan_to_num(-inf)
    Cues q;
    q.iun = x0;
    q.rho = rho;
    q.phi = 0.5f * atan2_poly(x2, x1);
    return q;
}

// Canonical angles, float samples.
__device__ __forceinline__ Cues cues_canonical_f32(const float (&i)[4]) {
    const float s1 = i[0] - i[2], s2 = i[1] - i[3];
    const float x0 = 0.25f * ((i[0] + i[2]) + (i[1] + i[3]));
    const float amp = 0.5f * sqrtf(fmaf(s1, s1, s2 * s2));
    float rho = amp / x0;
    if (isnan(rho) || rho == INFINITY) rho = 0.0f;
    else if (rho == -INFINITY) rho = -FLT_MAX;
    Cues q;
    q.iun = x0;
    q.rho = rho;
    q.phi = 0.5f * atan2_poly(s2, s1);
    return q;
}

template <typename T>
__device__ __forceinline__ void load_stack_px(const T* stack, size_t px, float (&v)[4], int (&iv)[4]);
template <>
__device__ __forceinline__ void load_stack_px<uint8_t>(const uint8_t* stack, size_t px, float (&v)[4], int (&iv)[4]) {
    const uint32_t w = ld_stream_u32(stack + 4 * px);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        iv[k] = (w >> (8 * k)) & 0xff;
        v[k] = (float)iv[k];
    }
}
template <>
__device__ __forceinline__ void load_stack_px<float>(const float* stack, size_t px, float (&v)[4], int (&iv)[4]) {
    const float4 w = ld_stream_f32x4(stack + 4 * px);
    v[0] = w.x; v[1] = w.y; v[2] = w.z; v[3] = w.w;
    iv[0] = iv[1] = iv[2] = iv[3] = 0;
}

constexpr int kXolpItems = 4;   // pixel groups per thread: a CTA moves ~48 KB, long enough to amortise its launch

// XOLP-only kernels.  A thread owns V consecutive pixels (V = 4 when the image size allows: one 16-byte load of
// four interleaved u8 pixels, one 16-byte store per output plane); one CTA per tile of 256 * V pixels, launched
// plainly (no per-CTA setup to amortise, so the hardware's own CTA queue is the dynamic scheduler).
template <typename T, bool GENERAL, int V>
__global__ void __launch_bounds__(256) xolp_stack_kernel(const T* __restrict__ stack, size_t hw, size_t total, Pinv pv,
                                                          float* __restrict__ iun, float* __restrict__ xolp, FastDiv hw_div) {
#pragma unroll 2
    for (int item = 0; item < kXolpItems; ++item) {
    const size_t i = (((size_t)blockIdx.x * kXolpItems + item) * 256 + threadIdx.x) * V;
    if (i >= total) return;
    float rho[V], phi[V], un[V];
    if constexpr (sizeof(T) == 1 && V == 4) {
        const uint4 w = *reinterpret_cast<const uint4*>(stack + 4 * i);   // 4 pixels x 4 samples
        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v[4] = {byte_to_float(ws[j], 0), byte_to_float(ws[j], 1), byte_to_float(ws[j], 2), byte_to_float(ws[j], 3)};
            const Cues q = GENERAL ? cues_general(v, pv) : cues_from_u8<false>(v[0], v[1], v[2], v[3]);
            rho[j] = q.rho; phi[j] = q.phi; un[j] = q.iun;
        }
    } else {
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float v[4];
            int iv[4];
            load_stack_px<T>(stack, i + j, v, iv);
            Cues q;
            if constexpr (GENERAL) q = cues_general(v, pv);
            else if constexpr (sizeof(T) == 1) q = cues_from_u8<false>(v[0], v[1], v[2], v[3]);
            else q = cues_canonical_f32(v);
            rho[j] = q.rho; phi[j] = q.phi; un[j] = q.iun;
        }
    }
    // image index: a 32-bit multiply-shift when the batch has < 2^31 pixels (hw_div.div != 0), else the 64-bit division
    const size_t b = hw_div.div ? (size_t)fastdiv((uint32_t)i, hw_div) : i / hw, r = i - b * hw;      // V divides hw when V > 1
    float* xo = xolp + b * 2 * hw + r;
    st_stream_vec<V>(xo, rho);
    st_stream_vec<V>(xo + hw, phi);
    if (iun) st_stream_vec<V>(iun + i, un);
    }
}

template <int V>
__global__ void __launch_bounds__(256) xolp_planes_kernel(const uint8_t* __restrict__ i0, const uint8_t* __restrict__ i45,
                                                           const uint8_t* __restrict__ i90, const uint8_t* __restrict__ i135,
                                                           size_t hw, size_t total, float* __restrict__ iun,
                                                           float* __restrict__ xolp, FastDiv hw_div) {
    // all sixteen plane words of the thread's four groups are requested before the first is used (four narrow input streams:
    // with two groups in flight the kernel sat at 0.88 of the HBM peak)
    uint32_t w[kXolpItems][4];
    size_t idx[kXolpItems];
#pragma unroll
    for (int item = 0; item < kXolpItems; ++item) {
        const size_t i = (((size_t)blockIdx.x * kXolpItems + item) * 256 + threadIdx.x) * V;
        idx[item] = i;
        if (i < total) {
            if constexpr (V == 4) {
                w[item][0] = ld_stream_u32(i0 + i);
                w[item][1] = ld_stream_u32(i45 + i);
                w[item][2] = ld_stream_u32(i90 + i);
                w[item][3] = ld_stream_u32(i135 + i);
            } else {
                w[item][0] = ld_stream_u8(i0 + i);
                w[item][1] = ld_stream_u8(i45 + i);
                w[item][2] = ld_stream_u8(i90 + i);
                w[item][3] = ld_stream_u8(i135 + i);
            }
        }
    }
#pragma unroll
    for (int item = 0; item < kXolpItems; ++item) {
        const size_t i = idx[item];
        if (i >= total) return;
        float rho[V], phi[V], un[V];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const Cues q = cues_from_u8<false>(byte_to_float(w[item][0], j), byte_to_float(w[item][1], j), byte_to_float(w[item][2], j),
                                               byte_to_float(w[item][3], j));
            rho[j] = q.rho; phi[j] = q.phi; un[j] = q.iun;
        }
        const size_t b = hw_div.div ? (size_t)fastdiv((uint32_t)i, hw_div) : i / hw, r = i - b * hw;
        float* xo = xolp + b * 2 * hw + r;
        st_stream_vec<V>(xo, rho);
        st_stream_vec<V>(xo + hw, phi);
        if (iun) st_stream_vec<V>(iun + i, un);
    }
}

// ---------------------------------------------------------------------------------------------
// Normals from XOLP (get_normals) and the fine-grained normals_vec functions.
// ---------------------------------------------------------------------------------------------
struct NormalsParams {
    const float* xolp;
    float* normals;
    LutArgs lut;
    uint32_t groups_total;    // B * HW / VEC
    FastDiv groups_per_image;
    size_t hw;
};

template <int VEC>
struct XolpIn {
    float rho[VEC], phi[VEC];
    uint32_t b, rem;
    bool valid;
};

template <int VEC>
__device__ __forceinline__ XolpIn<VEC> load_xolp(const NormalsParams& p, uint32_t tile) {
    XolpIn<VEC> g;
    const uint32_t gid = tile * kFusedThreads + threadIdx.x;
    g.valid = gid < p.groups_total;
    g.b = 0; g.rem = 0;
#pragma unroll
    for (int j = 0; j < VEC; ++j) g.rho[j] = g.phi[j] = 0.0f;
    if (g.valid) {
        g.b = fastdiv(gid, p.groups_per_image);
        g.rem = gid - g.b * p.groups_per_image.div;
        const float* xi = p.xolp + ((size_t)g.b * 2 * p.hw + (size_t)g.rem * VEC);
        if constexpr (VEC == 4) {
            const float4 r = ld_stream_f32x4(xi), f = ld_stream_f32x4(xi + p.hw);
            g.rho[0] = r.x; g.rho[1] = r.y; g.rho[2] = r.z; g.rho[3] = r.w;
            g.phi[0] = f.x; g.phi[1] = f.y; g.phi[2] = f.z; g.phi[3] = f.w;
        } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                g.rho[j] = ld_stream_f32(xi + j);
                g.phi[j] = ld_stream_f32(xi + p.hw + j);
            }
        }
    }
    return g;
}

// As steep_redo, for the XOLP-fed kernel: rho is the float32 input widened exactly, as scipy widens it.
__device__ __noinline__ void steep_redo_xolp(const NormalsParams& p, uint32_t b, uint32_t pix, uint32_t bits) {
    for (int j = 0; j < 4; ++j) {
        if (!((bits >> j) & 1)) continue;
        const float* xi = p.xolp + ((size_t)b * 2 * p.hw + pix + j);
        const double rho = (double)xi[0];
        float sp, cp;
        sincos_poly(xi[p.hw], sp, cp);
        float* no = p.normals + ((size_t)b * 9 * p.hw + pix + j);
        for (int t = 0; t < 3; ++t) {
            if (!((p.lut.steep.mask >> t) & 1) || !(rho > p.lut.steep.x[t])) continue;
            const float2 sc = steep_sincos(steep_theta(p.lut.steep, t, rho));
            no[(size_t)(3 * t + 0) * p.hw] = (t == 0 ? cp : -sp) * sc.x;
            no[(size_t)(3 * t + 1) * p.hw] = (t == 0 ? sp : cp) * sc.x;
            no[(size_t)(3 * t + 2) * p.hw] = sc.y;
        }
    }
}

// get_normals: same tile scheduling and two-deep software pipeline as the fused kernel.
template <int VEC, bool MUFU>
__global__ void __launch_bounds__(kFusedThreads, kFusedMinBlocks) normals_from_xolp_kernel(const __grid_constant__ NormalsParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ __align__(16) uint4 clc_resp;
    __shared__ uint64_t clc_bar;
    lut_stage_begin(smem_raw, p.lut.blob, p.lut.bytes, &bar);
    ClcTiles clc;
    clc.init(&clc_resp, &clc_bar);
    clc.prefetch();
    XolpIn<VEC> cur = load_xolp<VEC>(p, blockIdx.x);
    lut_stage_wait(&bar);
    const LutShared lut = lut_shared(smem_raw, p.lut);
    for (;;) {
        uint32_t next_tile;
        const bool more = clc.next(next_tile);
        __syncthreads();
        XolpIn<VEC> nxt;
        nxt.valid = false;
        if (more) {
            clc.prefetch();
            nxt = load_xolp<VEC>(p, next_tile);
        }
        if (cur.valid) {
            float nrm[9][VEC];
            uint32_t steep_bits = 0;
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                float n9[9];
                normals_from_cues<MUFU>(lut, cur.rho[j], cur.phi[j], n9);
                steep_bits |= (cur.rho[j] >= p.lut.steep.from ? 1u : 0u) << j;
#pragma unroll
                for (int c = 0; c < 9; ++c) nrm[c][j] = n9[c];
            }
            float* no = p.normals + ((size_t)cur.b * 9 * p.hw + (size_t)cur.rem * VEC);
            asm volatile("" : "+l"(no));
#pragma unroll
            for (int c = 0; c < 9; ++c) {
                st_stream_vec<VEC>(no, nrm[c]);
                no += p.hw;
            }
            if (steep_bits) steep_redo_xolp(p, cur.b, cur.rem * VEC, steep_bits);
        }
        if (!more) break;
        cur = nxt;
    }
}

// theta lookups straight from the global-memory blob (L1-resident after first touch).
template <int WHICH>  // 0: diffuse, 1: both specular branches
__global__ void __launch_bounds__(256) theta_kernel(const float* __restrict__ rho, size_t count, LutArgs a,
                                                     float* __restrict__ out0, float* __restrict__ out1) {
    const LutView lut = lut_view(a.blob, a);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const float r = ld_stream_f32(rho + i);
        const float g = lut_coord(r);
        float th[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int t = (WHICH == 0 ? 0 : 1); t < (WHICH == 0 ? 1 : 3); ++t) {
            th[t] = lut_eval(lut.cells[t], lut.scale[t], lut.last[t], r, g);
            if (r >= a.steep.from && ((a.steep.mask >> t) & 1) && (double)r > a.steep.x[t])
                th[t] = (float)steep_theta(a.steep, t, (double)r);    // float64 line, rounded once
        }
        if constexpr (WHICH == 0) {
            st_stream_f32(out0 + i, th[0]);
        } else {
            st_stream_f32(out0 + i, th[1]);
            st_stream_f32(out1 + i, th[2]);
        }
    }
}

// calc_normals: (phi, theta)[B, HW] -> [B, 3, HW]   (normals_vec.py:53-60)
__global__ void __launch_bounds__(256) calc_normals_kernel(const float* __restrict__ phi, const float* __restrict__ theta,
                                                            size_t hw, size_t total, float* __restrict__ out, FastDiv hw_div) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        float sp, cp, st, ct;
        sincos_poly(ld_stream_f32(phi + i), sp, cp);
        sincos_poly(ld_stream_f32(theta + i), st, ct);
        const size_t b = hw_div.div ? (size_t)fastdiv((uint32_t)i, hw_div) : i / hw, r = i - b * hw;
        float* o = out + b * 3 * hw + r;
        st_stream_f32(o, cp * st);
        st_stream_f32(o + hw, sp * st);
        st_stream_f32(o + 2 * hw, ct);
    }
}

// ppp "channel" variant (physical_normals_channels.py:15-36): s0 = I0 + I90, no nan scrub, masked.
__global__ void __launch_bounds__(256) stokes_channel_kernel(const float* __restrict__ stack, const uint8_t* __restrict__ mask,
                                                              size_t total, float* __restrict__ rho, float* __restrict__ phi,
                                                              float* __restrict__ iun) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const float4 w = ld_stream_f32x4(stack + 4 * i);
        const bool in = mask[i] != 0;
        const float s0 = w.x + w.z, s1 = w.x - w.z, s2 = w.y - w.w;
        const float r = sqrtf(fmaf(s1, s1, s2 * s2)) / s0;   // 0/0 stays NaN, as np.divide
        rho[i] = in ? r : 0.0f;
        phi[i] = in ? 0.5f * atan2_poly(s2, s1) : 0.0f;
        iun[i] = in ? 0.5f * s0 : 0.0f;
    }
}

// calc_normals_channel (physical_normals_channels.py:75-83): H x W x 3 interleaved, zero outside mask.
__global__ void __launch_bounds__(256) calc_normals_channel_kernel(const float* __restrict__ phi, const float* __restrict__ theta,
                                                                    const uint8_t* __restrict__ mask, size_t total,
                                                                    float phi_offset, float* __restrict__ out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        float sp, cp, st, ct;
        sincos_poly(phi[i] + phi_offset, sp, cp);
        sincos_poly(theta[i], st, ct);
        const bool in = mask[i] != 0;
        out[3 * i + 0] = in ? cp * st : 0.0f;
        out[3 * i + 1] = in ? sp * st : 0.0f;
        out[3 * i + 2] = in ? ct : 0.0f;
    }
}

inline unsigned grid_for(size_t total, int threads) {
    const size_t want = (total + threads - 1) / threads;
    const size_t cap = (size_t)device_info().sms * 16;
    return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

}  // namespace
}  // namespace polcue

using namespace polcue;

extern "C" {

static int fused_mosaic_common(const uint8_t* mosaic, int B, int H, int W, bool layout_superpixel, const int* angle_at,
                               const polcue_lut* lut, uint8_t* planes, float* iun, float* xolp, float* normals,
                               polcue_stream_t stream, void* stats_workspace = nullptr, bool* stats_fused = nullptr) {
    if (!mosaic || !xolp || B < 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1)) return POLCUE_EINVAL;
    if (layout_superpixel) {
        if (!angle_at) return POLCUE_EINVAL;
        int seen = 0;
        for (int k = 0; k < 4; ++k) {
            if (angle_at[k] < 0 || angle_at[k] > 3) return POLCUE_EINVAL;
            seen |= 1 << angle_at[k];
        }
        if (seen != 15) return POLCUE_EINVAL;          // every angle exactly once
    }
    if (normals && (!lut || !lut->d_blob)) return POLCUE_EINVAL;
    if (B == 0) return POLCUE_OK;
    const int Hs = H / 2, Ws = W / 2;
    int vec = 1;
    if (Ws % 4 == 0 && aligned(mosaic, layout_superpixel ? 8 : 4) && aligned(xolp, 16) && aligned(normals, 16) && aligned(iun, 16) &&
        aligned(planes, 4))
        vec = 4;
    else if (Ws % 2 == 0 && aligned(mosaic, 2) && aligned(xolp, 8) && aligned(normals, 8) && aligned(iun, 8) && aligned(planes, 2))
        vec = 2;
    else if (!aligned(xolp, 4) || !aligned(normals, 4) || !aligned(iun, 4))
        return POLCUE_EINVAL;
    const unsigned long long groups = (unsigned long long)B * Hs * (Ws / vec);
    if (groups >= (1ull << 31)) return POLCUE_E2BIG;
    FusedParams p;
    p.mosaic = mosaic;
    p.planes = planes;
    p.iun = iun;
    p.xolp = xolp;
    p.xolp_norm = nullptr;
    p.norm_mean = 0.0f;
    p.norm_inv_std = 1.0f;
    p.normals = normals;
    if (normals) p.lut = lut_args(lut);
    else p.lut = LutArgs{};
    p.groups_total = (uint32_t)groups;
    p.groups_per_frame.div = (uint32_t)Hs * (Ws / vec);
    make_fastdiv(p.groups_per_frame.div, p.groups_per_frame.mul, p.groups_per_frame.shift);
    p.groups_per_row.div = (uint32_t)(Ws / vec);
    make_fastdiv(p.groups_per_row.div, p.groups_per_row.mul, p.groups_per_row.shift);
    p.W = (uint32_t)W;
    p.row_px = (uint32_t)Ws;
    if ((unsigned long long)H * W >= (1ull << 30)) return POLCUE_E2BIG;   // 32-bit strides inside the kernel
    p.frame_bytes = (uint32_t)H * (uint32_t)W;
    p.off45 = Ws;
    p.off90 = (long long)Hs * W;
    p.off135 = (long long)Hs * W + Ws;
    p.superpixel = layout_superpixel ? 1 : 0;
    for (int k = 0; k < 4; ++k) p.angle_at[k] = layout_superpixel ? angle_at[k] : k;
    p.plane = (uint32_t)Hs * (uint32_t)Ws;
    p.plane_bytes = 4u * p.plane;
    p.tile_records = nullptr;
    // statistics by-product: 4-pixel groups, normals present, no steep table (its float64 redo overwrites stored values)
    const uint32_t tiles = (p.groups_total + kFusedThreads - 1) / kFusedThreads;
    if (stats_workspace && vec == 4 && normals && !lut->steep_mask) {
        p.tile_records = fold_records(stats_workspace, tiles);
        if (stats_fused) *stats_fused = true;
    }
    const size_t smem = normals ? lut->bytes() : 0;
    cudaStream_t s = (cudaStream_t)stream;
    const bool mufu = !lut || lut->trig_mufu != 0;
    switch (vec) {
        case 4: return launch_fused<4>(p, smem, s, mufu);
        case 2: return launch_fused<2>(p, smem, s, mufu);
        default: return launch_fused<1>(p, smem, s, mufu);
    }
}

static size_t stats_scratch_offset(int B, int H, int W) {
    const size_t hw = (size_t)(H / 2) * (W / 2);
    // the by-product's tile records, or (fallback) the workspace of the separate statistics pass over the 9 normal channels
    const unsigned long long tiles = ((unsigned long long)B * (hw / 4) + kFusedThreads - 1) / kFusedThreads + 1;
    const size_t fused = tiles < (1ull << 31) ? fold_workspace_bytes((uint32_t)tiles, kSumRecord) : 64;
    const size_t separate = polcue_channel_stats_workspace_bytes(B, 9, hw);
    return ((fused > separate ? fused : separate) + 63) / 64 * 64;
}

size_t polcue_fused_stats_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 64;
    return stats_scratch_offset(B, H, W) + 22 * sizeof(double) + 64;
}

__global__ void pack_stats13_kernel(const double* __restrict__ xolp4, const double* __restrict__ normals18, double* __restrict__ out) {
    const int i = threadIdx.x;
    if (i < 2) out[i] = xolp4[2 * i];                   // sum rho, sum phi
    else if (i < 11) out[i] = normals18[2 * (i - 2)];   // sums of the nine normal channels
    else if (i < 13) out[i] = xolp4[2 * (i - 11) + 1];  // squares of rho, phi
}

int polcue_fused_mosaic_stats_u8(const uint8_t* mosaic, int B, int H, int W, const polcue_lut* lut, uint8_t* planes, float* iun,
                                 float* xolp, float* normals, void* workspace, double* stats13, polcue_stream_t stream) {
    if (!workspace || !stats13 || !normals || B <= 0) return POLCUE_EINVAL;
    if ((reinterpret_cast<uintptr_t>(workspace) & 63) || (reinterpret_cast<uintptr_t>(stats13) & 7)) return POLCUE_EINVAL;
    bool fused = false;
    int rc = fused_mosaic_common(mosaic, B, H, W, false, nullptr, lut, planes, iun, xolp, normals, stream, workspace, &fused);
    if (rc != POLCUE_OK) return rc;
    const size_t hw = (size_t)(H / 2) * (W / 2);
    if (fused) {
        const uint32_t tiles = (uint32_t)(((unsigned long long)B * (hw / 4) + kFusedThreads - 1) / kFusedThreads);
        return launch_fold_tiles(workspace, tiles, kSumRecord, kStatValues, stats13, (cudaStream_t)stream);
    }
    // Shapes / tables the by-product does not cover (widths that are not multiples of 8, steep end segments): the separate
    // statistics pass over the stored outputs, which adds in the same canonical order.
    double* tmp = reinterpret_cast<double*>(static_cast<char*>(workspace) + stats_scratch_offset(B, H, W));
    rc = polcue_channel_stats_f32(xolp, B, 2, hw, workspace, tmp, stream);
    if (rc != POLCUE_OK) return rc;
    rc = polcue_channel_stats_f32(normals, B, 9, hw, workspace, tmp + 4, stream);
    if (rc != POLCUE_OK) return rc;
    pack_stats13_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(tmp, tmp + 4, stats13);
    return launch_status();
}

int polcue_fused_mosaic_u8(const uint8_t* mosaic, int B, int H, int W, const polcue_lut* lut, uint8_t* planes, float* iun,
                           float* xolp, float* normals, polcue_stream_t stream) {
    return fused_mosaic_common(mosaic, B, H, W, false, nullptr, lut, planes, iun, xolp, normals, stream);
}

int polcue_fused_superpixel_u8(const uint8_t* mosaic, int B, int H, int W, const int* angle_at, const polcue_lut* lut,
                               uint8_t* planes, float* iun, float* xolp, float* normals, polcue_stream_t stream) {
    return fused_mosaic_common(mosaic, B, H, W, true, angle_at, lut, planes, iun, xolp, normals, stream);
}

// Four sample planes at `i0 + {0, off45, off90, off135}`, consecutive frames `frame_stride` bytes apart.
static int fused_planes_common(const uint8_t* i0, long long off45, long long off90, long long off135, unsigned long long frame_stride,
                               bool aligned4, bool aligned2, int B, int H, int W, const polcue_lut* lut, float* iun, float* xolp,
                               float* normals, polcue_stream_t stream, float* xolp_norm = nullptr, float norm_mean = 0.0f,
                               float norm_std = 1.0f) {
    if (normals && (!lut || !lut->d_blob)) return POLCUE_EINVAL;
    if (xolp_norm && !(norm_std != 0.0f)) return POLCUE_EINVAL;
    if (B == 0) return POLCUE_OK;
    int vec = 1;
    if (W % 4 == 0 && aligned4 && aligned(xolp, 16) && aligned(normals, 16) && aligned(iun, 16) && aligned(xolp_norm, 16)) vec = 4;
    else if (W % 2 == 0 && aligned2 && aligned(xolp, 8) && aligned(normals, 8) && aligned(iun, 8) && aligned(xolp_norm, 8)) vec = 2;
    else if (!aligned(xolp, 4) || !aligned(normals, 4) || !aligned(iun, 4) || !aligned(xolp_norm, 4)) return POLCUE_EINVAL;
    const unsigned long long groups = (unsigned long long)B * H * (W / vec);
    if (groups >= (1ull << 31) || frame_stride >= (1ull << 30)) return POLCUE_E2BIG;
    FusedParams p;
    p.mosaic = i0;
    p.off45 = off45;
    p.off90 = off90;
    p.off135 = off135;
    p.superpixel = 0;
    for (int k = 0; k < 4; ++k) p.angle_at[k] = k;
    p.planes = nullptr;
    p.iun = iun;
    p.xolp = xolp;
    p.xolp_norm = xolp_norm;
    p.norm_mean = norm_mean;
    p.norm_inv_std = 1.0f / norm_std;
    p.normals = normals;
    if (normals) p.lut = lut_args(lut);
    else p.lut = LutArgs{};
    p.groups_total = (uint32_t)groups;
    p.groups_per_frame.div = (uint32_t)H * (W / vec);
    make_fastdiv(p.groups_per_frame.div, p.groups_per_frame.mul, p.groups_per_frame.shift);
    p.groups_per_row.div = (uint32_t)(W / vec);
    make_fastdiv(p.groups_per_row.div, p.groups_per_row.mul, p.groups_per_row.shift);
    p.W = (uint32_t)W;
    p.row_px = (uint32_t)W;
    p.frame_bytes = (uint32_t)frame_stride;
    p.plane = (uint32_t)H * (uint32_t)W;
    p.plane_bytes = 4u * p.plane;
    p.tile_records = nullptr;
    const size_t smem = normals ? lut->bytes() : 0;
    cudaStream_t s = (cudaStream_t)stream;
    const bool mufu = !lut || lut->trig_mufu != 0;
    switch (vec) {
        case 4: return launch_fused<4>(p, smem, s, mufu);
        case 2: return launch_fused<2>(p, smem, s, mufu);
        default: return launch_fused<1>(p, smem, s, mufu);
    }
}

int polcue_fused_planes_u8(const uint8_t* i0, const uint8_t* i45, const uint8_t* i90, const uint8_t* i135, int B, int H, int W,
                           const polcue_lut* lut, float* iun, float* xolp, float* normals, polcue_stream_t stream) {
    if (!i0 || !i45 || !i90 || !i135 || !xolp || B < 0 || H <= 0 || W <= 0) return POLCUE_EINVAL;
    auto all4 = [&](size_t a) { return aligned(i0, a) && aligned(i45, a) && aligned(i90, a) && aligned(i135, a); };
    return fused_planes_common(i0, i45 - i0, i90 - i0, i135 - i0, (unsigned long long)H * W, all4(4), all4(2), B, H, W, lut, iun,
                               xolp, normals, stream);
}

static int xolp_stack_common(const void* stack, bool is_u8, int B, int H, int W, const float* pinv, float* iun, float* xolp,
                             polcue_stream_t stream) {
    if (!stack || !xolp || B < 0 || H <= 0 || W <= 0) return POLCUE_EINVAL;
    if (!aligned(stack, is_u8 ? 4 : 16) || !aligned(xolp, 4) || !aligned(iun, 4)) return POLCUE_EINVAL;
    if (B == 0) return POLCUE_OK;
    const size_t hw = (size_t)H * W, total = hw * B;
    Pinv pv{};
    if (pinv)
        for (int k = 0; k < 12; ++k) pv.m[k] = pinv[k];
    const bool v4 = hw % 4 == 0 && aligned(stack, 16) && aligned(xolp, 16) && aligned(iun, 16);
    const size_t per_cta = 256 * kXolpItems * (v4 ? 4 : 1);
    const size_t ctas = (total + per_cta - 1) / per_cta;
    if (ctas >= (1ull << 31)) return POLCUE_E2BIG;
    const unsigned grid = (unsigned)ctas;
    cudaStream_t s = (cudaStream_t)stream;
    FastDiv hw_div{0, 0, 0};
    if (total < (1ull << 31) && hw < (1ull << 31)) {
        hw_div.div = (uint32_t)hw;
        make_fastdiv(hw_div.div, hw_div.mul, hw_div.shift);
    }
#define POLCUE_LAUNCH_STACK(T, G)                                                                                      \
    do {                                                                                                               \
        if (v4) xolp_stack_kernel<T, G, 4><<<grid, 256, 0, s>>>((const T*)stack, hw, total, pv, iun, xolp, hw_div);    \
        else xolp_stack_kernel<T, G, 1><<<grid, 256, 0, s>>>((const T*)stack, hw, total, pv, iun, xolp, hw_div);       \
    } while (0)
    if (is_u8) {
        if (pinv) POLCUE_LAUNCH_STACK(uint8_t, true);
        else POLCUE_LAUNCH_STACK(uint8_t, false);
    } else {
        if (pinv) POLCUE_LAUNCH_STACK(float, true);
        else POLCUE_LAUNCH_STACK(float, false);
    }
#undef POLCUE_LAUNCH_STACK
    return launch_status();
}

int polcue_xolp_stack_u8(const uint8_t* stack, int B, int H, int W, const float* pinv, float* iun, float* xolp,
                         polcue_stream_t stream) {
    return xolp_stack_common(stack, true, B, H, W, pinv, iun, xolp, stream);
}

int polcue_xolp_stack_f32(const float* stack, int B, int H, int W, const float* pinv, float* iun, float* xolp,
                          polcue_stream_t stream) {
    return xolp_stack_common(stack, false, B, H, W, pinv, iun, xolp, stream);
}

int polcue_xolp_planes_u8(const uint8_t* i0, const uint8_t* i45, const uint8_t* i90, const uint8_t* i135, int B, int H, int W,
                          float* iun, float* xolp, polcue_stream_t stream) {
    if (!i0 || !i45 || !i90 || !i135 || !xolp || B < 0 || H <= 0 || W <= 0) return POLCUE_EINVAL;
    if (!aligned(xolp, 4) || !aligned(iun, 4)) return POLCUE_EINVAL;
    if (B == 0) return POLCUE_OK;
    const size_t hw = (size_t)H * W, total = hw * B;
    const bool v4 = hw % 4 == 0 && aligned(i0, 4) && aligned(i45, 4) && aligned(i90, 4) && aligned(i135, 4) && aligned(xolp, 16) &&
                    aligned(iun, 16);
    if (v4 && W % 4 == 0 && (unsigned long long)B * H * (W / 4) < (1ull << 31) && hw < (1ull << 30)) {
        // rows of whole 4-pixel groups: the streaming member of the fused family (all loads of four groups first), measured faster
        return fused_planes_common(i0, i45 - i0, i90 - i0, i135 - i0, (unsigned long long)hw, true, true, B, H, W, nullptr, iun, xolp,
                                   nullptr, stream);
    }
    const size_t per_cta = 256 * kXolpItems * (v4 ? 4 : 1);
    const size_t ctas = (total + per_cta - 1) / per_cta;
    if (ctas >= (1ull << 31)) return POLCUE_E2BIG;
    FastDiv hw_div{0, 0, 0};
    if (total < (1ull << 31) && hw < (1ull << 31)) {
        hw_div.div = (uint32_t)hw;
        make_fastdiv(hw_div.div, hw_div.mul, hw_div.shift);
    }
    if (v4) xolp_planes_kernel<4><<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(i0, i45, i90, i135, hw, total, iun, xolp, hw_div);
    else xolp_planes_kernel<1><<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(i0, i45, i90, i135, hw, total, iun, xolp, hw_div);
    return launch_status();
}

int polcue_normals_from_xolp_f32(const float* xolp, int B, int H, int W, const polcue_lut* lut, float* normals,
                                 polcue_stream_t stream) {
    if (!xolp || !normals || !lut || !lut->d_blob || B < 0 || H <= 0 || W <= 0) return POLCUE_EINVAL;
    if (!aligned(xolp, 4) || !aligned(normals, 4)) return POLCUE_EINVAL;
    if (B == 0) return POLCUE_OK;
    const size_t hw = (size_t)H * W;
    const int vec = (hw % 4 == 0 && aligned(xolp, 16) && aligned(normals, 16)) ? 4 : 1;
    const unsigned long long groups = (unsigned long long)B * (hw / vec);
    if (groups >= (1ull << 31)) return POLCUE_E2BIG;
    NormalsParams p;
    p.xolp = xolp;
    p.normals = normals;
    p.lut = lut_args(lut);
    p.groups_total = (uint32_t)groups;
    p.groups_per_image.div = (uint32_t)(hw / vec);
    make_fastdiv(p.groups_per_image.div, p.groups_per_image.mul, p.groups_per_image.shift);
    p.hw = hw;
    const size_t smem = lut->bytes();
    auto launch = [&](auto kern) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kFusedThreads, smem);
        if (e != cudaSuccess) return (int)e;
        if (per_sm < 1) return POLCUE_ERANGE;
        const uint32_t tiles = (p.groups_total + kFusedThreads - 1) / kFusedThreads;
        kern<<<tiles, kFusedThreads, smem, (cudaStream_t)stream>>>(p);
        return launch_status();
    };
    if (vec == 4) return lut->trig_mufu ? launch(normals_from_xolp_kernel<4, true>) : launch(normals_from_xolp_kernel<4, false>);
    return lut->trig_mufu ? launch(normals_from_xolp_kernel<1, true>) : launch(normals_from_xolp_kernel<1, false>);
}

int polcue_rho_diffuse_f32(const float* rho, size_t count, const polcue_lut* lut, float* theta, polcue_stream_t stream) {
    if ((!rho || !theta) && count) return POLCUE_EINVAL;
    if (!lut || !lut->d_blob) return POLCUE_EINVAL;
    if (!count) return POLCUE_OK;
    theta_kernel<0><<<grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>(rho, count, lut_args(lut), theta, nullptr);
    return launch_status();
}

int polcue_rho_spec_f32(const float* rho, size_t count, const polcue_lut* lut, float* theta1, float* theta2,
                        polcue_stream_t stream) {
    if ((!rho || !theta1 || !theta2) && count) return POLCUE_EINVAL;
    if (!lut || !lut->d_blob) return POLCUE_EINVAL;
    if (!count) return POLCUE_OK;
    theta_kernel<1><<<grid_for(count, 256), 256, 0, (cudaStream_t)stream>>>(rho, count, lut_args(lut), theta1, theta2);
    return launch_status();
}

int polcue_calc_normals_f32(const float* phi, const float* theta, int B, size_t hw, float* normals, polcue_stream_t stream) {
    if (!phi || !theta || !normals || B < 0) return POLCUE_EINVAL;
    const size_t total = hw * (size_t)B;
    if (!total) return POLCUE_OK;
    FastDiv hw_div{0, 0, 0};
    if (total < (1ull << 31) && hw && hw < (1ull << 31)) {
        hw_div.div = (uint32_t)hw;
        make_fastdiv(hw_div.div, hw_div.mul, hw_div.shift);
    }
    calc_normals_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(phi, theta, hw, total, normals, hw_div);
    return launch_status();
}

int polcue_stokes_channel_f32(const float* stack, const uint8_t* mask, int H, int W, float* rho, float* phi, float* iun,
                              polcue_stream_t stream) {
    if (!stack || !mask || !rho || !phi || !iun || H <= 0 || W <= 0 || !aligned(stack, 16)) return POLCUE_EINVAL;
    const size_t total = (size_t)H * W;
    stokes_channel_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(stack, mask, total, rho, phi, iun);
    return launch_status();
}

int polcue_calc_normals_channel_f32(const float* phi, const float* theta, const uint8_t* mask, size_t hw, float phi_offset,
                                    float* normals_hw3, polcue_stream_t stream) {
    if (!phi || !theta || !mask || !normals_hw3) return POLCUE_EINVAL;
    if (!hw) return POLCUE_OK;
    calc_normals_channel_kernel<<<grid_for(hw, 256), 256, 0, (cudaStream_t)stream>>>(phi, theta, mask, hw, phi_offset, normals_hw3);
    return launch_status();
}

}  // extern "C"

// planes: B x 4 x H x W (angle order) -- the layout the loader front end produces (resize.cu)
int polcue::fused_planes_strided(const uint8_t* planes, int B, int H, int W, const polcue_lut* lut, float* iun, float* xolp,
                                 float* normals, cudaStream_t stream, float* xolp_norm, float norm_mean, float norm_std) {
    if (!planes || !xolp || B < 0 || H <= 0 || W <= 0) return POLCUE_EINVAL;
    const long long hw = (long long)H * W;
    return fused_planes_common(planes, hw, 2 * hw, 3 * hw, 4ull * hw, aligned(planes, 4) && hw % 4 == 0,
                               aligned(planes, 2) && hw % 2 == 0, B, H, W, lut, iun, xolp, normals, (polcue_stream_t)stream, xolp_norm,
                               norm_mean, norm_std);
}
