// In-register radix-R inverse DFT building block (R = 2..64, power of two) with
// compile-time twiddles.
//
//   X[n] = sum_k x[k] * exp(+2 pi i k n / R),  natural order in, natural order out,
//
// i.e. the unnormalised inverse transform the reference's Stockham passes compute
// (shader/fft_row.comp:25-40: dst[2t-k] = a + w b, dst[2t-k+bs] = a - w b, w = e^{+i pi k/bs}).
// The reference runs log2(N) radix-2 stages through shared memory with one barrier triple
// per stage; here a thread owns R points and does log2(R) stages in registers, so a length-N
// line needs only ceil(log_R N) trips through shared memory.
//
// Twiddles use the true pi (rounded once from double), not the shader's 3.1415926 literal:
// see DESIGN.md "numerics" -- the difference is ~1e-6 of the field maximum, a 10x margin
// under the 1e-5 parity bar, and keeping w^(R/4) = i exact is what makes the butterflies cheap.
#pragma once
#include <cuda_runtime.h>

#include <utility>

namespace ocean {

constexpr double kPiD = 3.14159265358979323846264338327950288;

__host__ __device__ constexpr double taylor_sin(double x)  // |x| <= pi/4
{
    double x2 = x * x, term = x, sum = x;
    for (int i = 1; i < 12; ++i) {
        term *= -x2 / double((2 * i) * (2 * i + 1));
        sum += term;
    }
    return sum;
}

__host__ __device__ constexpr double taylor_cos(double x)  // |x| <= pi/4
{
    double x2 = x * x, term = 1.0, sum = 1.0;
    for (int i = 1; i < 12; ++i) {
        term *= -x2 / double((2 * i - 1) * (2 * i));
        sum += term;
    }
    return sum;
}

// cos(2 pi k / n), sin(2 pi k / n) for n a power of two, with exact values on the axes and
// diagonals (so that w = 1, i, -1, -i fold away at compile time).
__host__ __device__ constexpr double cos2pi(long k, long n)
{
    k = ((k % n) + n) % n;
    if (2 * k > n) k = n - k;                 // cos(2pi - t) = cos t       -> t in [0, pi]
    double sgn = 1.0;
    if (4 * k > n) { k = n / 2 - k; sgn = -1.0; }   // cos(pi - t) = -cos t  -> t in [0, pi/2]
    if (k == 0) return sgn;
    if (4 * k == n) return 0.0;
    if (8 * k == n) return sgn * 0.70710678118654752440084436210485;
    if (8 * k > n) return sgn * taylor_sin(2.0 * kPiD * double(n - 4 * k) / double(4 * n));
    return sgn * taylor_cos(2.0 * kPiD * double(k) / double(n));
}

__host__ __device__ constexpr double sin2pi(long k, long n) { return cos2pi(4 * k - n, 4 * n); }  // sin t = cos(t - pi/2)

// OCEAN_FFT_PACKED (device code, sm_100+): 1 = the twiddle-free butterflies (w = 1: 31 of the 80 butterflies of a
// 32-point transform) use the packed add.rn.f32x2 / fma.rn.f32x2 on the (re, im) register pair -- two instructions
// instead of four; 2 = the general butterflies use fma.rn.f32x2 as well (three packed FMAs + the swapped operand).
// 0 = scalar. Measured on B200 (same box, 0 / 1 / 2): 1024^2 x 8 tiles 83.1 k / 82.4 k / 82.8 k frames/s (single-tile
// frames 13.2 / 13.1 / 12.9 us), 512^2 312.9 k / 312.5 k / 313.2 k, 2048^2 16.4 k / 17.0 k / 17.0 k -- the static
// instruction count of k_rows_t drops 2880 -> 2800 -> 2640 but the two-pass kernels are latency bound, not issue bound,
// and the scalar form keeps its twiddles as FFMA immediates. Level 1 is bit-identical to scalar, level 2 rounds differently.
// The macro is the default of the PK template parameter; the line configurations pick level 1 for the three-pass
// lines (N = 2048), see LineCfg::PK.
#ifndef OCEAN_FFT_PACKED
#define OCEAN_FFT_PACKED 0
#endif

template <int R, int K, int PK>
__host__ __device__ __forceinline__ void dit_butterfly(const float2 e, const float2 o, float2& lo, float2& hi)
{
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
    if constexpr (K == 0 && PK >= 1) {
        lo = __fadd2_rn(e, o);
        hi = __ffma2_rn(o, make_float2(-1.0f, -1.0f), e);      // e - o, exact like a subtraction
        return;
    }
    if constexpr (4 * K != R && K != 0 && PK >= 2) {
        constexpr float wr = float(cos2pi(K, R));
        constexpr float wi = float(sin2pi(K, R));
        const float2 t = __ffma2_rn(make_float2(wr, wr), o, e);                        // e + wr o
        lo = __ffma2_rn(make_float2(-wi, wi), make_float2(o.y, o.x), t);               // + (-wi o.y, wi o.x)
        hi = __ffma2_rn(make_float2(2.0f, 2.0f), e, make_float2(-lo.x, -lo.y));        // 2e - lo
        return;
    }
#endif
    if constexpr (K == 0) {
        lo = make_float2(e.x + o.x, e.y + o.y);
        hi = make_float2(e.x - o.x, e.y - o.y);
    } else if constexpr (4 * K == R) {        // w = +i
        lo = make_float2(e.x - o.y, e.y + o.x);
        hi = make_float2(e.x + o.y, e.y - o.x);
    } else {
        constexpr float wr = float(cos2pi(K, R));
        constexpr float wi = float(sin2pi(K, R));
        lo.x = fmaf(wr, o.x, fmaf(-wi, o.y, e.x));
        lo.y = fmaf(wr, o.y, fmaf(wi, o.x, e.y));
        hi.x = fmaf(2.0f, e.x, -lo.x);        // e - w o = 2e - (e + w o)
        hi.y = fmaf(2.0f, e.y, -lo.y);
    }
}

// PK: packed level of the butterflies (see OCEAN_FFT_PACKED); the line configurations choose it per N (LineCfg::PK)
template <int R, int PK = OCEAN_FFT_PACKED>
struct RegFft {
    template <int... K>
    __host__ __device__ __forceinline__ static void combine(const float2 (&e)[R / 2], const float2 (&o)[R / 2],
                                                   float2 (&v)[R], std::integer_sequence<int, K...>)
    {
        (dit_butterfly<R, K, PK>(e[K], o[K], v[K], v[K + R / 2]), ...);
    }

    __host__ __device__ __forceinline__ static void run(float2 (&v)[R])
    {
        float2 e[R / 2], o[R / 2];
#pragma unroll
        for (int k = 0; k < R / 2; ++k) {
            e[k] = v[2 * k];
            o[k] = v[2 * k + 1];
        }
        RegFft<R / 2, PK>::run(e);
        RegFft<R / 2, PK>::run(o);
        combine(e, o, v, std::make_integer_sequence<int, R / 2>{});
    }
};

template <int PK>
struct RegFft<1, PK> {
    __host__ __device__ __forceinline__ static void run(float2 (&)[1]) {}
};

// v *= w
__host__ __device__ __forceinline__ float2 cmul_tw(float2 v, float2 w)
{
    return make_float2(fmaf(v.x, w.x, -v.y * w.y), fmaf(v.x, w.y, v.y * w.x));
}

}  // namespace ocean
