// Device-side building blocks shared by the polcue kernels (sm_100a).
//   - approximate-unit math wrappers (MUFU) with known error bounds
//   - atan2 / sincos polynomials accurate to ~1.5e-7 abs (fitted for this project)
//   - the search-free zenith-angle table lookup
//   - the per-pixel Stokes -> XOLP -> normal-candidate pipeline
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace polcue {

constexpr float kPi = 3.14159265358979323846f;
constexpr float kHalfPi = 1.57079632679489661923f;
constexpr float kSqrt2 = 1.41421356237309504880f;

// ------------------------------------------------------------------------------------------
// MUFU wrappers.  rel. error: rcp 1 ulp, sqrt 1 ulp, rsqrt 2 ulp (PTX ISA approx variants).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float lg2_approx(float x) {   // abs error 2^-22 on [0.5, 2]; denormals flush (never depths)
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ------------------------------------------------------------------------------------------
// Streaming global access: inputs are read once, outputs written once (no reuse in L1/L2).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_stream_u32(const void* p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint16_t ld_stream_u16(const void* p) {
    uint16_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint8_t ld_stream_u8(const void* p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return (uint8_t)v;
}
__device__ __forceinline__ float ld_stream_f32(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_stream_f32x4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream_f32(float* p, float v) {
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void st_stream_f32x2(float* p, float a, float b) {
    asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void st_stream_f32x4(float* p, float a, float b, float c, float d) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ float4 lds_f32x4(uint32_t shared_addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(shared_addr));
    return v;
}

// Vector store of V consecutive floats (V in {1,2,4}); p is V*4-byte aligned.
template <int V>
__device__ __forceinline__ void st_stream_vec(float* p, const float (&v)[V]) {
    if constexpr (V == 4) st_stream_f32x4(p, v[0], v[1], v[2], v[3]);
    else if constexpr (V == 2) st_stream_f32x2(p, v[0], v[1]);
    else st_stream_f32(p, v[0]);
}

// ------------------------------------------------------------------------------------------
// atan2(y, x) in [-pi, pi].  a = min/max in [0,1]; atan(a) = a + a s P7(s), s = a^2.
// Max abs error 1.2e-7 rad (float32 evaluation, fitted on [0,1]).  atan2(0,0) = 0, atan2(+0, x<0) = +pi
// -- the closed-form tie behaviour DESIGN.md documents (the reference's lstsq leaves it to round-off).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float atan2_poly(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float a = mn * rcp_approx(mx);
    a = (mx == 0.0f) ? 0.0f : a;  // 0 * inf
    const float s = a * a;
    float p = 0.004059889819473028f;
    p = fmaf(p, s, -0.020706569775938988f);
    p = fmaf(p, s, 0.049855392426252365f);
    p = fmaf(p, s, -0.08074377477169037f);
    p = fmaf(p, s, 0.10888639092445374f);
    p = fmaf(p, s, -0.142609104514122f);
    p = fmaf(p, s, 0.19998927414417267f);
    p = fmaf(p, s, -0.33333325386047363f);
    float r = fmaf(a * s, p, a);
    r = (ay > ax) ? kHalfPi - r : r;
    r = (x < 0.0f) ? kPi - r : r;
    return copysignf(r, y);
}

// ------------------------------------------------------------------------------------------
// sincos.  Polynomials fitted on |x| <= 1.60 (covers every in-table zenith angle and every AoLP
// without reduction): abs error 1.4e-7.  Larger arguments (extrapolated zenith angles reach
// +-140 rad) take a 3-term Cody-Waite reduction by pi/2 first; valid to |x| ~ 1e5.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void sincos_core(float r, float& s, float& c) {
    const float t = r * r;
    float ps = 2.6301038360543316e-06f;
    ps = fmaf(ps, t, -0.00019821283058263361f);
    ps = fmaf(ps, t, 0.008333231322467327f);
    ps = fmaf(ps, t, -0.1666666567325592f);
    s = fmaf(r * t, ps, r);
    float pc = -2.624870205636398e-07f;
    pc = fmaf(pc, t, 2.4772434699116275e-05f);
    pc = fmaf(pc, t, -0.0013888622634112835f);
    pc = fmaf(pc, t, 0.041666656732559204f);
    pc = fmaf(pc, t, -0.5f);
    c = fmaf(t, pc, 1.0f);
}

__device__ __forceinline__ void sincos_poly(float x, float& s, float& c) {
    if (fabsf(x) <= 1.6f) {  // every in-table zenith angle and every AoLP: no reduction, no quadrant fix-up
        sincos_core(x, s, c);
        return;
    }
    const float k = rintf(x * 0.636619772367581343f);  // rare: extrapolated zenith angles only
    float r = fmaf(k, -1.57079601e+00f, x);
    r = fmaf(k, -3.13916473e-07f, r);
    r = fmaf(k, -5.39030253e-15f, r);
    float ss, cc;
    sincos_core(r, ss, cc);
    const int q = (int)k;
    const float s1 = (q & 1) ? cc : ss;
    const float c1 = (q & 1) ? ss : cc;
    s = (q & 2) ? -s1 : s1;
    c = ((q + 1) & 2) ? -c1 : c1;
}

// MUFU variant: sin.approx/cos.approx, abs error 2^-21.4 = 3.6e-7 on [-pi, pi].  Arguments beyond
// that are first reduced by 2 pi (2-term Cody-Waite) so the error stays flat out to |x| ~ 1e4.
__device__ __forceinline__ void sincos_mufu(float x, float& s, float& c) {
    float r = x;
    if (fabsf(x) > kPi) {
        const float k = rintf(x * 0.159154943091895336f);
        r = fmaf(k, -6.28318548202514648f, x);       // 2 pi rounded to float32
        r = fmaf(k, 1.74845553146951715e-07f, r);    // (float32(2 pi) - 2 pi)
    }
    s = __sinf(r);
    c = __cosf(r);
}

template <bool kMufu>
__device__ __forceinline__ void sincos_sel(float x, float& s, float& c) {
    if constexpr (kMufu) sincos_mufu(x, s, c);
    else sincos_poly(x, s, c);
}

// ------------------------------------------------------------------------------------------
// Zenith-angle tables.  Each table is a uniform grid of cells over g(rho) in [0, sqrt 2]:
//     g = sqrt(rho)                for rho <  0.5   (knots of rho_d, rho_s crowd quadratically at 0)
//     g = sqrt 2 - sqrt(1 - rho)   for rho >= 0.5   (... and at the specular peak rho -> 1)
// A cell holds (x_k, y_k, slope_left, slope_right) of the single knot it contains (or of the next
// knot to the right); theta = y_k + (rho - x_k) * (rho <= x_k ? slope_left : slope_right), which is the
// same line scipy's interp1d(linear, extrapolate) evaluates on that segment.  Built in lut.cu.
// ------------------------------------------------------------------------------------------
struct LutView {
    const float4* cells[3];  // diffuse, spec1, spec2 (global memory, L1-resident)
    float scale[3];          // cells per unit g
    float last[3];           // cells - 0.5
};

__device__ __forceinline__ float lut_coord(float rho) {
    const bool low = rho < 0.5f;
    const float t = fmaxf(low ? rho : 1.0f - rho, 0.0f);
    const float g = sqrt_approx(t);
    return low ? g : kSqrt2 - g;
}

// cell = (x_k, y_k, slope_left, slope_right - slope_left); the cell coordinate is clamped to the last cell.
__device__ __forceinline__ float lut_line(const float4 e, float rho) {
    const float d = rho - e.x;
    return fmaf(fmaxf(d, 0.0f), e.w, fmaf(d, e.z, e.y));
}
__device__ __forceinline__ float lut_eval(const float4* __restrict__ cells, float scale, float last, float rho, float g) {
    return lut_line(cells[(int)fminf(g * scale, last)], rho);
}

// Shared-memory tables addressed by 32-bit shared-window addresses (one IMAD per lookup).
struct LutShared {
    uint32_t addr[3];
    float scale[3];
    float last[3];    // cells - 0.5: clamp of the cell coordinate (beyond the last knot = extrapolation cell)
};
__device__ __forceinline__ float lut_eval_shared(const LutShared& lut, int t, float rho, float g) {
    return lut_line(lds_f32x4(lut.addr[t] + 16u * (uint32_t)(int)fminf(g * lut.scale[t], lut.last[t])), rho);
}

// ------------------------------------------------------------------------------------------
// Steep end segments.  Where a table's last segment is steeper than 512 rad per unit rho (second specular branch
// at n = 1.8: -10797, the knot next to the peak lies 1.5e-7 below it) neither rho nor theta fit float32 to the
// 1e-3 rad parity bound: theta reaches -1e4 rad for rho -> 2.  Queries beyond the second-to-last knot of such a
// table (rare: rho > ~0.99999) are evaluated in float64 exactly as scipy's _call_linear does on that segment,
// theta = slope (rho - x_lo) + y_lo, reduced by 2 pi in float64 before the float32 sin/cos.
// ------------------------------------------------------------------------------------------
struct LutSteep {
    int mask;                      // bit t: table t has a steep end segment
    float from;                    // rho >= from may lie on one (float32 rounded down); +inf when mask == 0
    double x[3], y[3], slope[3];   // (x_lo, y_lo) = second-to-last knot
};

__device__ __forceinline__ double steep_theta(const LutSteep& st, int t, double rho) {
    return fma(st.slope[t], rho - st.x[t], st.y[t]);
}

static __device__ __noinline__ float2 steep_sincos(double theta) {
    const double k = rint(theta * 0.15915494309189535);
    double r = fma(k, -6.283185307179586232, theta);      // 2 pi = hi + lo
    r = fma(k, -2.4492935982947064e-16, r);
    float s, c;
    sincos_poly((float)r, s, c);
    return make_float2(s, c);
}

// Overwrites the candidates of steep tables in n[9] (layout of normals_from_trig) for a float64 rho.
__device__ __forceinline__ void steep_fix(const LutSteep& st, double rho, float sp, float cp, float (&n)[9]) {
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        if (((st.mask >> t) & 1) && rho > st.x[t]) {
            const float2 sc = steep_sincos(steep_theta(st, t, rho));
            n[3 * t + 0] = (t == 0 ? cp : -sp) * sc.x;
            n[3 * t + 1] = (t == 0 ? sp : cp) * sc.x;
            n[3 * t + 2] = sc.y;
        }
    }
}

// rho of cues_from_u8 in float64 (every intermediate is an exact integer)
__device__ __forceinline__ double rho_exact_u8(float i0, float i45, float i90, float i135) {
    const float s1 = i0 - i90, s2 = i45 - i135, sum = (i0 + i90) + (i45 + i135);
    return 2.0 * sqrt((double)fmaf(s1, s1, s2 * s2)) / (double)sum;
}

// ------------------------------------------------------------------------------------------
// Per-pixel pipeline pieces.
// ------------------------------------------------------------------------------------------
struct Cues {
    float iun, rho, phi;
    float sin_phi, cos_phi;  // filled by cues_from_u8 only
};

// atan on [0,1]: a + a s P7(s), s = a^2; max abs error 1.2e-7.
__device__ __forceinline__ float atan_unit(float a, float s) {
    float p = 0.004059889819473028f;
    p = fmaf(p, s, -0.020706569775938988f);
    p = fmaf(p, s, 0.049855392426252365f);
    p = fmaf(p, s, -0.08074377477169037f);
    p = fmaf(p, s, 0.10888639092445374f);
    p = fmaf(p, s, -0.142609104514122f);
    p = fmaf(p, s, 0.19998927414417267f);
    p = fmaf(p, s, -0.33333325386047363f);
    return fmaf(a * s, p, a);
}

// Canonical angles, 8-bit samples already widened to float (exact): polarisation/xolp.py:8-34 in closed form.
//   x0 = (I0+I45+I90+I135)/4, x1 = (I0-I90)/2, x2 = (I45-I135)/2
//   rho = sqrt(x1^2+x2^2)/x0 (x0 == 0 -> 0, the inf/nan scrub of xolp.py:26-29), phi = atan2(x2,x1)/2
// The half angle is taken algebraically: with r = |(s1,s2)|, u = |s2| / (r + |s1|) = tan(phi'), phi' in
// [0, pi/4], so atan needs no octant swap and cos/sin(phi) come from one rsqrt:
//   s1 >= 0: |phi| = phi'           (cos, sin) = (c', s')        c' = rsqrt(1+u^2), s' = u c'
//   s1 <  0: |phi| = pi/2 - phi'    (cos, sin) = (s', c')        sign(phi) = sign(s2), s2 = +0 -> +
// Ties resolve as exact arithmetic does: s1 = s2 = 0 -> phi = 0; s2 = 0, s1 < 0 -> phi = +pi/2.
template <bool kTrig = true>   // kTrig = false: XOLP only, skip cos/sin(phi)
__device__ __forceinline__ Cues cues_from_u8(float i0, float i45, float i90, float i135) {
    const float s1 = i0 - i90, s2 = i45 - i135, sum = (i0 + i90) + (i45 + i135);
    const float amp = sqrt_approx(fmaf(s1, s1, s2 * s2));   // integer < 2^24 under the root: exact
    Cues q;
    q.iun = 0.25f * sum;
    q.rho = (2.0f * amp) * rcp_approx(sum + 1e-30f);        // sum >= 1 unless every sample is 0 (then amp = 0)
    const float u = fabsf(s2) * rcp_approx((amp + fabsf(s1)) + 1e-30f);
    const float t = u * u;
    const float half = atan_unit(u, t);
    const bool flip = s1 < 0.0f;
    q.phi = copysignf(flip ? kHalfPi - half : half, s2);
    if constexpr (kTrig) {
        const float cq = rsqrt_approx(1.0f + t), sq = u * cq;
        q.cos_phi = flip ? sq : cq;
        q.sin_phi = copysignf(flip ? cq : sq, s2);
    } else {
        q.cos_phi = q.sin_phi = 0.0f;
    }
    return q;
}

// ------------------------------------------------------------------------------------------
// Packed FP32 (Blackwell FFMA2 / FADD2 / FMUL2): two pixels per instruction.  tools/probes/int_pipe_probe.cu measures
// FFMA2 at the scalar FFMA FLOP rate in half the issue slots (60.7 packed instructions/clk/SM; immediates broadcast to
// both lanes at no register cost).  Each lane runs exactly the operation sequence of cues_from_u8, so the results are
// bit-identical to the scalar form.  Used by the fused kernel family (fused.cu: process_group).
// ------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) {
    f32x2 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 dup2(float c) { return pk2(c, c); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 abs2(f32x2 a) { return a & 0x7FFFFFFF7FFFFFFFull; }
__device__ __forceinline__ f32x2 neg2(f32x2 a) { return a ^ 0x8000000080000000ull; }   // folds into the consumer's operand modifier
// NOTE (measured, ptxas 12.9): a mul.rn.f32x2 whose result feeds an add.rn.f32x2 / sub.rn.f32x2 IS contracted into FFMA2
// by ptxas, .rn qualifier and -fmad=false notwithstanding (the scalar mul.rn.f32 / add.rn.f32 pair is not).  Code that
// must agree bit for bit with a scalar twin, or that relies on separately rounded products cancelling exactly, therefore
// never feeds a packed product into a packed add / sub: it subtracts the two lanes with scalar instructions
// (sub2_unfused below: two FADD instead of one FADD2).
__device__ __forceinline__ f32x2 sub2_unfused(f32x2 a, f32x2 b) {
    float a0, a1, b0, b1;
    unpk2(a, a0, a1);
    unpk2(b, b0, b1);
    return pk2(__fsub_rn(a0, b0), __fsub_rn(a1, b1));
}
__device__ __forceinline__ f32x2 add2_unfused(f32x2 a, f32x2 b) {
    float a0, a1, b0, b1;
    unpk2(a, a0, a1);
    unpk2(b, b0, b1);
    return pk2(__fadd_rn(a0, b0), __fadd_rn(a1, b1));
}

// byte ka of word wa and byte kb of word wb as an exact float pair (see byte_to_float): two PRMT, one FADD2
__device__ __forceinline__ f32x2 bytes_to_float2(uint32_t wa, int ka, uint32_t wb, int kb) {
    const f32x2 raw = pk2(__uint_as_float(__byte_perm(wa, 0x4B000000u, 0x7650 + ka)), __uint_as_float(__byte_perm(wb, 0x4B000000u, 0x7650 + kb)));
    return add2(raw, dup2(-8388608.0f));
}
// bytes k and k + 1 of one packed word
__device__ __forceinline__ f32x2 bytes_to_float2(uint32_t w, int k) { return bytes_to_float2(w, k, w, k + 1); }

// cues_from_u8 for two pixels at once (lane 0 -> a, lane 1 -> b).
template <bool kTrig = true>
__device__ __forceinline__ void cues2_from_u8(f32x2 i0, f32x2 i45, f32x2 i90, f32x2 i135, Cues& a, Cues& b) {
    const f32x2 s1 = sub2(i0, i90), s2 = sub2(i45, i135), sum = add2(add2(i0, i90), add2(i45, i135));
    float q0, q1;
    unpk2(fma2(s1, s1, mul2(s2, s2)), q0, q1);
    const f32x2 amp = pk2(sqrt_approx(q0), sqrt_approx(q1));
    float d0, d1;
    unpk2(add2(sum, dup2(1e-30f)), d0, d1);
    const f32x2 rho = mul2(mul2(dup2(2.0f), amp), pk2(rcp_approx(d0), rcp_approx(d1)));
    const f32x2 as1 = abs2(s1), as2 = abs2(s2);
    float e0, e1;
    unpk2(add2(add2(amp, as1), dup2(1e-30f)), e0, e1);
    const f32x2 u = mul2(as2, pk2(rcp_approx(e0), rcp_approx(e1)));
    const f32x2 t = mul2(u, u);
    f32x2 pl = dup2(0.004059889819473028f);
    pl = fma2(pl, t, dup2(-0.020706569775938988f));
    pl = fma2(pl, t, dup2(0.049855392426252365f));
    pl = fma2(pl, t, dup2(-0.08074377477169037f));
    pl = fma2(pl, t, dup2(0.10888639092445374f));
    pl = fma2(pl, t, dup2(-0.142609104514122f));
    pl = fma2(pl, t, dup2(0.19998927414417267f));
    pl = fma2(pl, t, dup2(-0.33333325386047363f));
    const f32x2 half = fma2(mul2(u, t), pl, u);
    const f32x2 other = sub2(dup2(kHalfPi), half);
    float s1a, s1b, s2a, s2b, ha, hb, oa, ob;
    unpk2(s1, s1a, s1b);
    unpk2(s2, s2a, s2b);
    unpk2(half, ha, hb);
    unpk2(other, oa, ob);
    const bool fa = s1a < 0.0f, fb = s1b < 0.0f;
    unpk2(mul2(dup2(0.25f), sum), a.iun, b.iun);
    unpk2(rho, a.rho, b.rho);
    a.phi = copysignf(fa ? oa : ha, s2a);
    b.phi = copysignf(fb ? ob : hb, s2b);
    if constexpr (kTrig) {
        float t0, t1, u0, u1;
        unpk2(add2(dup2(1.0f), t), t0, t1);
        unpk2(u, u0, u1);
        const f32x2 cq = pk2(rsqrt_approx(t0), rsqrt_approx(t1));
        float ca, cb, sa, sb;
        unpk2(cq, ca, cb);
        unpk2(mul2(u, cq), sa, sb);
        a.cos_phi = fa ? sa : ca;
        a.sin_phi = copysignf(fa ? ca : sa, s2a);
        b.cos_phi = fb ? sb : cb;
        b.sin_phi = copysignf(fb ? cb : sb, s2b);
        (void)u0; (void)u1;
    } else {
        a.cos_phi = a.sin_phi = b.cos_phi = b.sin_phi = 0.0f;
    }
}

// Three normal candidates in the channel order of get_normals (pre_encoders.py:99-113):
//   N_diff = (cos phi sin td, sin phi sin td, cos td)
//   N_spec = (cos(phi+pi/2) sin ts, sin(phi+pi/2) sin ts, cos ts) = (-sin phi sin ts, cos phi sin ts, cos ts)
template <bool kMufu>
__device__ __forceinline__ void normals_from_trig(const LutShared& lut, float rho, float sp, float cp, float (&n)[9]) {
    const float g = lut_coord(rho);
    const float td = lut_eval_shared(lut, 0, rho, g);
    const float t1 = lut_eval_shared(lut, 1, rho, g);
    const float t2 = lut_eval_shared(lut, 2, rho, g);
    float s, c;
    sincos_sel<kMufu>(td, s, c);
    n[0] = cp * s; n[1] = sp * s; n[2] = c;
    sincos_sel<kMufu>(t1, s, c);
    n[3] = -sp * s; n[4] = cp * s; n[5] = c;
    sincos_sel<kMufu>(t2, s, c);
    n[6] = -sp * s; n[7] = cp * s; n[8] = c;
}

template <bool kMufu>
__device__ __forceinline__ void normals_from_cues(const LutShared& lut, float rho, float phi, float (&n)[9]) {
    float sp, cp;
    sincos_sel<kMufu>(phi, sp, cp);
    normals_from_trig<kMufu>(lut, rho, sp, cp, n);
}

// byte k of a packed word as float, exactly: PRMT builds 0x4B0000bb = 2^23 + bb, one FADD removes 2^23.
__device__ __forceinline__ float byte_to_float(uint32_t w, int k) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650 + k)) - 8388608.0f;
}

// ------------------------------------------------------------------------------------------
// Bulk (TMA) copy of the table blob into shared memory, completion on an mbarrier.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void lut_stage_begin(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
            "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
            : "memory");
    }
}

__device__ __forceinline__ void lut_stage_wait(uint64_t* bar) {
    __syncthreads();  // orders the init by thread 0 before anyone polls
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar))
            : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// Dynamic tile scheduling with Cluster Launch Control (sm_100).  The grid has one CTA per tile; only as many
// CTAs as fit are resident.  A resident CTA that finishes a tile asks the hardware work distributor to CANCEL a
// still-pending CTA of the same grid and takes over its blockIdx, so tiles are handed out in order to
// whichever SM is free: no global counter, no host-side reset, table staging paid once per resident CTA.
// On the two-die B200 this matters: under HBM saturation SMs do not all get the same bandwidth, and a static
// grid-stride split finishes with its slowest SM (measured 0.73 ms static vs 0.60 ms dynamic for this kernel's
// access pattern, tools/hbm_probe.cu).
// Usage (all threads):  clc.init(..); clc.prefetch(); tile = blockIdx.x;
//                        for (;;) { more = clc.next(nxt); __syncthreads(); if (more) clc.prefetch();
//                                   work(tile); if (!more) break; tile = nxt; }
// ------------------------------------------------------------------------------------------
struct ClcTiles {
    uint4* resp;       // 16-byte response slot (shared)
    uint64_t* bar;     // mbarrier (shared)
    uint32_t phase;

    __device__ __forceinline__ void init(uint4* resp_slot, uint64_t* barrier) {
        resp = resp_slot;
        bar = barrier;
        phase = 0;
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    // ask for the next tile while the current one is being processed
    __device__ __forceinline__ void prefetch() {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 16;" ::"r"(smem_u32(bar)) : "memory");
            asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];" ::"r"(
                             smem_u32(resp)),
                         "r"(smem_u32(bar))
                         : "memory");
        }
    }
    // false when no pending CTA was left to cancel (this CTA is done)
    __device__ __forceinline__ bool next(uint32_t& tile) {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(smem_u32(bar)), "r"(phase)
                : "memory");
        }
        phase ^= 1;
        uint32_t valid, first;
        asm volatile(
            "{\n\t.reg .pred p1;\n\t.reg .b128 r;\n\t"
            "ld.shared.b128 r, [%2];\n\t"
            "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, r;\n\t"
            "selp.u32 %1, 1, 0, p1;\n\t"
            "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, _, _, _}, r;\n\t}"
            : "=r"(first), "=r"(valid)
            : "r"(smem_u32(resp))
            : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tile = first;
        return valid != 0;   // caller: __syncthreads() before the next prefetch() re-arms the slot
    }
};

// ------------------------------------------------------------------------------------------
// Deterministic sums ("canonical order"), shared by the statistics kernel (stats.cu) and the statistics by-product of
// the fused kernel (fused.cu) so that both give the SAME BITS for the same data under any CTA scheduling:
//   group   4 consecutive elements of a channel's batch-flattened sequence (1 element when hw % 4 != 0):
//           sum  = (x + y) + (z + w),   squares = fma(x, x, fma(y, y, fma(z, z, w * w)))            float32
//   tile    256 consecutive groups = 8 "warps" of 32: per warp the xor butterfly 16, 8, 4, 2, 1 (float add commutes, so
//           the lane that evaluates a node does not matter), then the 8 warp sums added in order 0..7          float32
//   segment 1024 consecutive tiles: thread j adds tiles j, j + 256, j + 512, j + 768 in float64, block_sum_f64
//   total   thread j adds segments j, j + 256, ... in float64, block_sum_f64
// Missing groups / tiles (ragged ends) contribute +0.
// ------------------------------------------------------------------------------------------
constexpr int kSumTileGroups = 256;
constexpr int kSumSegmentTiles = 1024;
constexpr int kSumRecord = 16;          // floats per tile record of the fused by-product (13 used)

__device__ __forceinline__ float group_sum4(float x, float y, float z, float w) { return __fadd_rn(__fadd_rn(x, y), __fadd_rn(z, w)); }
__device__ __forceinline__ float group_squares4(float x, float y, float z, float w) {
    return fmaf(x, x, fmaf(y, y, fmaf(z, z, __fmul_rn(w, w))));
}
__device__ __forceinline__ float warp_tree_sum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}
// float64 sum over the 256 threads of a CTA in a fixed order (shuffle-down tree, then the 8 warp sums in order); the
// result is valid in thread 0.  scratch: 8 doubles of shared memory.
__device__ __forceinline__ double block_sum_f64(double v, double* scratch) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();                      // scratch may still be read from a previous call
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < 8; ++w) t += scratch[w];
    return t;
}

// Exact unsigned division by a runtime constant (d >= 1, n < 2^31): q = (n * mul) >> 32 >> shift.
struct FastDiv {
    uint32_t mul, shift, div;
};
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, const FastDiv& d) {
    return (d.div == 1) ? n : (__umulhi(n, d.mul) >> d.shift);
}

}  // namespace polcue
