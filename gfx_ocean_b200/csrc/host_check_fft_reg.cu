// Host-side check of the in-register transform (fft_reg.cuh) against a naive f64 DFT.
// Built and run by tests/test_host_fft.py (no GPU needed): nvcc compiles the same
// __host__ __device__ templates for the CPU.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "fft_reg.cuh"

template <int R>
static double check()
{
    float2 v[R];
    double xr[R], xi[R];
    for (int k = 0; k < R; ++k) {
        xr[k] = std::sin(1.0 + 3.7 * k) + 0.25 * k;
        xi[k] = std::cos(2.0 + 1.3 * k * k);
        v[k] = make_float2(float(xr[k]), float(xi[k]));
        xr[k] = v[k].x;
        xi[k] = v[k].y;
    }
    ocean::RegFft<R>::run(v);
    double worst = 0, scale = 0;
    for (int n = 0; n < R; ++n) {
        double sr = 0, si = 0;
        for (int k = 0; k < R; ++k) {
            const double th = 2.0 * ocean::kPiD * double((k * n) % R) / R;
            sr += xr[k] * std::cos(th) - xi[k] * std::sin(th);
            si += xr[k] * std::sin(th) + xi[k] * std::cos(th);
        }
        worst = std::fmax(worst, std::fmax(std::fabs(sr - v[n].x), std::fabs(si - v[n].y)));
        scale = std::fmax(scale, std::fmax(std::fabs(sr), std::fabs(si)));
    }
    return worst / scale;
}

int main()
{
    int bad = 0;
    const double e[] = {check<2>(), check<4>(), check<8>(), check<16>(), check<32>(), check<64>()};
    const int r[] = {2, 4, 8, 16, 32, 64};
    for (int i = 0; i < 6; ++i) {
        std::printf("R=%d max_rel_err=%.3e\n", r[i], e[i]);
        if (!(e[i] < 1e-6)) bad = 1;
    }
    // compile-time twiddles: exact on the axes, accurate elsewhere
    static_assert(ocean::cos2pi(0, 32) == 1.0 && ocean::cos2pi(8, 32) == 0.0 && ocean::cos2pi(16, 32) == -1.0, "axes");
    static_assert(ocean::sin2pi(8, 32) == 1.0 && ocean::sin2pi(16, 32) == 0.0 && ocean::sin2pi(24, 32) == -1.0, "axes");
    double tw = 0;
    for (int k = 0; k < 1024; ++k) {
        tw = std::fmax(tw, std::fabs(ocean::cos2pi(k, 1024) - std::cos(2 * ocean::kPiD * k / 1024)));
        tw = std::fmax(tw, std::fabs(ocean::sin2pi(k, 1024) - std::sin(2 * ocean::kPiD * k / 1024)));
    }
    std::printf("twiddle max_abs_err=%.3e\n", tw);
    if (!(tw < 1e-15)) bad = 1;
    return bad;
}
