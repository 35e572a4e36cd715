// Shared device helpers for the ocean kernels (sm_100a).
//
// Arithmetic contract: reference shaders under /root/reference/shader/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace ocean {

// `const float pi = 3.1415926;` (propagate.comp:6, fft_row.comp:5): rounds to 0x40490FDA.
constexpr float kPi32 = 3.1415926f;

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// complex_mul of the shaders: (c0.x*c1.x - c0.y*c1.y, c0.y*c1.x + c0.x*c1.y)
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.y * b.x + a.x * b.y);
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// propagate.comp:45-46,50-53. The shader forms `uint x = 2*gid - resolution - 1` in u32
// (wraps for gid <= N/2) and converts it with an UNSIGNED int->float conversion
// (OpConvertUToF in shader/spv/propagate.comp.spv), then k = pi * float(x) / domain_size.
__device__ __forceinline__ float wave_number(uint32_t g, uint32_t n, float domain_size)
{
    const uint32_t u = 2u * g - n - 1u;
    return __fdiv_rn(__fmul_rn(kPi32, __uint2float_rn(u)), domain_size);
}

// propagate.comp:64-67: k / length(k) if length(k) > 1e-10 else 0.
__device__ __forceinline__ float2 unit_wave_vector(float kx, float ky)
{
    const float len = sqrtf(kx * kx + ky * ky);
    float2 r = make_float2(0.f, 0.f);
    if (len > 1.0e-10f) r = make_float2(kx / len, ky / len);
    return r;
}

// Same quantity with one MUFU.RSQ + one Newton step instead of sqrt and two IEEE divisions
// (<= 1 ulp-level difference on k/|k|; length(k) > 1e-10 <=> |k|^2 > 1e-20).
__device__ __forceinline__ float2 unit_wave_vector_fast(float kx, float ky)
{
    const float l2 = fmaf(kx, kx, ky * ky);
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l2));    // |k|^2 is never denormal on these grids
    r = r * fmaf(-0.5f * l2, r * r, 1.5f);
    r = l2 > 1.0e-20f ? r : 0.f;
    return make_float2(kx * r, ky * r);
}

// Full-range sincos for the propagation phase omega*t (which reaches 1e3..1e4 rad).
// |x| <= 1e5: Cody-Waite reduction by pi/2 in three FMA steps (constants sum to pi/2 within
// 3.3e-22) + degree-7/8 minimax polynomials on [-pi/4, pi/4]; max abs error 7e-8 (measured
// against f64 over 2M samples by tests/test_host_fft.py::test_sincos_reduced_accuracy, which compiles this
// header for the host). Larger arguments take libdevice's
// Payne-Hanek path, kept out of line so the unrolled callers stay small.
static __device__ __noinline__ float2 sincos_huge(float x)
{
    float s, c;
    sincosf(x, &s, &c);
    return make_float2(s, c);
}

__host__ __device__ __forceinline__ void sincos_reduced(float x, float& sn, float& cs)
{
    float j = fmaf(x, 0.636619747f, 12582912.f);     // 1.5 * 2^23: rint(x * 2/pi) in the low mantissa bits
#ifdef __CUDA_ARCH__
    const int q = __float_as_int(j);
#else
    int q;
    memcpy(&q, &j, sizeof q);
#endif
    j -= 12582912.f;
    float r = fmaf(j, -1.57079601e+00f, x);
    r = fmaf(j, -3.13916473e-07f, r);
    r = fmaf(j, -5.39030253e-15f, r);
    const float s2 = r * r;
    float ps = fmaf(-1.9495291e-4f, s2, 8.3319759e-3f);
    ps = fmaf(ps, s2, -1.66666508e-1f);
    const float sr = fmaf(r * s2, ps, r);
    float pc = fmaf(2.4438088e-5f, s2, -1.38873642e-3f);
    pc = fmaf(pc, s2, 4.16666456e-2f);
    pc = fmaf(pc, s2, -0.5f);
    const float cr = fmaf(pc, s2, 1.0f);
    const bool swap = q & 1;
    float so = swap ? cr : sr, co = swap ? sr : cr;
    sn = (q & 2) ? -so : so;
    cs = ((q + 1) & 2) ? -co : co;
}

// The default: Cody-Waite reduction by 2 pi (the same three constants, times four) to [-pi, pi], then the
// special-function unit (MUFU.SIN / MUFU.COS, absolute error ~4e-7 on that range): 9 instructions instead of ~30.
// Measured on B200 against the f64 oracle the end-to-end error is unchanged (<= 1.9e-6 of the field maximum at
// N = 64..2048, t = 0..25000; profiles/r02_sincos_mufu_accuracy.txt) and k_rows gets 3-5 % faster.
// -DOCEAN_SINCOS_POLY selects the polynomial version (7e-8) instead.
__device__ __forceinline__ void sincos_mufu(float x, float& sn, float& cs)
{
    const float n = fmaf(x, 0.159154943f, 12582912.f) - 12582912.f;       // rint(x / 2 pi)
    float r = fmaf(n, -6.28318405e+00f, x);
    r = fmaf(n, -1.25566589e-06f, r);
    r = fmaf(n, -2.15612101e-14f, r);
    sn = __sinf(r);
    cs = __cosf(r);
}

// the fast path of sincos_full, valid for |x| <= 1e5
__device__ __forceinline__ void sincos_in_range(float x, float& sn, float& cs)
{
#ifdef OCEAN_SINCOS_POLY
    sincos_reduced(x, sn, cs);
#else
    sincos_mufu(x, sn, cs);
#endif
}

__device__ __forceinline__ void sincos_full(float x, float& sn, float& cs)
{
    if (fabsf(x) <= 1.0e5f) sincos_in_range(x, sn, cs);
    else {
        const float2 sc = sincos_huge(x);
        sn = sc.x;
        cs = sc.y;
    }
}

// propagate.comp:55-62 with the sine and cosine of the phase given.
__device__ __forceinline__ float2 propagate_point_sc(float2 a, float2 b, float s, float c)
{
    return make_float2((a.x + b.x) * c - (a.y - b.y) * s, (a.y + b.y) * c + (a.x - b.x) * s);
}

// Same as propagate_point with the sincos above.
__device__ __forceinline__ float2 propagate_point_fast(float2 a, float2 b, float omega, float time)
{
    float s, c;
    sincos_full(__fmul_rn(omega, time), s, c);
    return make_float2((a.x + b.x) * c - (a.y - b.y) * s, (a.y + b.y) * c + (a.x - b.x) * s);
}

// propagate.comp:55-62: h = h0[idx]*(cos,sin)(w t) + h0[N*N-1-idx]*(cos,-sin)(w t).
// The phase product is fp32 (it reaches thousands of radians), the sincos is the
// full-range accurate one: never compile this file with --use_fast_math.
__device__ __forceinline__ float2 propagate_point(float2 a, float2 b, float omega, float time)
{
    float s, c;
    sincosf(__fmul_rn(omega, time), &s, &c);
    // (a.x c - a.y s) + (b.x c + b.y s),  (a.y c + a.x s) + (b.y c - b.x s)
    return make_float2((a.x + b.x) * c - (a.y - b.y) * s, (a.y + b.y) * c + (a.x - b.x) * s);
}

}  // namespace ocean
