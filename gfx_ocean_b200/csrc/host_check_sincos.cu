// Host-side accuracy check of sincos_reduced (ocean_common.cuh) against f64 over the range it is used on
// (|x| <= 1e5 rad). Built and run by tests/test_host_fft.py (no GPU needed): the function is
// __host__ __device__ and made of fmaf / multiplies only, so the host evaluation is bit-identical to the device's.
#include <cmath>
#include <cstdio>

#include "ocean_common.cuh"

int main()
{
    double worst = 0.0;
    float worst_x = 0.f;
    unsigned long long n = 0;
    auto probe = [&](float x) {
        float s, c;
        ocean::sincos_reduced(x, s, c);
        const double es = std::fabs(double(s) - std::sin(double(x))), ec = std::fabs(double(c) - std::cos(double(x)));
        const double e = es > ec ? es : ec;
        if (e > worst) { worst = e; worst_x = x; }
        ++n;
    };
    for (int i = -1000000; i <= 1000000; ++i) probe(0.1f * float(i) + 0.0137f);             // uniform over [-1e5, 1e5]
    for (int i = 0; i < 400000; ++i) probe(std::ldexp(1.0f + float(i % 1000) * 1e-3f, -20 + i / 11000));   // 2^-20 .. 2^16
    for (int k = -60000; k <= 60000; ++k) {                                                      // next to the quadrant boundaries
        const float b = float(double(k) * 0.78539816339744830962);
        probe(std::nextafterf(b, 1e9f));
        probe(std::nextafterf(b, -1e9f));
        probe(b);
    }
    std::printf("samples=%llu max_abs_err=%.3e at x=%.9g\n", n, worst, double(worst_x));
    return worst < 1.0e-7 ? 0 : 1;
}
