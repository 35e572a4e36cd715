/*
 * ocean_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU oracle for the gfx-ocean per-frame compute path
 *   propagate -> fft_row x3 -> fft_col x3 -> correction
 * (reference: shader/propagate.comp, shader/fft_row.comp, shader/fft_col.comp,
 * shader/correction.comp, recorded by src/render.rs:1122-1287).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library. The product path
 * (gfx_ocean_b200/, include/ocean_b200.h) never links or calls it.
 *
 * Two instantiations of ocean_oracle_impl.h:
 *   oracle_*_f32 : literal fp32 restatement (cosf/sinf, fp32 Stockham), the
 *                  arithmetic a GPU driver would run; also the CPU baseline.
 *   oracle_*_f64 : parity-defining oracle. f64 everywhere except
 *                  (1) phi = fl32(omega * t) and
 *                  (2) k   = fl32(fl32(pi32 * fl32(u32 x)) / L),
 *                  where fp32 rounding changes results above 1e-5.
 *
 * PARITY UNPINNED: the reference holds no golden vectors for this path.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define REAL float
#define REAL_IS_DOUBLE 0
#define NAME(x) CAT(CAT(oracle_, x), _f32)
#include "ocean_oracle_impl.h"
#undef REAL
#undef REAL_IS_DOUBLE
#undef NAME
#undef PI32

#define REAL double
#define REAL_IS_DOUBLE 1
#define NAME(x) CAT(CAT(oracle_, x), _f64)
#include "ocean_oracle_impl.h"
#undef REAL
#undef REAL_IS_DOUBLE
#undef NAME

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* Decode of the reference's input files (src/render.rs:769-771, 808-810):
 * bincode 1.3.1 Vec<f32> / Vec<[f32;2]> = u64-LE element count + raw LE f32s.
 * Returns the element count, or <0 on error. elem_floats = 1 (omega) or 2. */
int64_t oracle_read_bincode(const char *path, uint32_t elem_floats, float *dst,
                            uint64_t dst_capacity_elems)
{
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    uint64_t count = 0;
    if (fread(&count, 8, 1, f) != 1) { fclose(f); return -2; }
    if (count > dst_capacity_elems) { fclose(f); return -3; }
    const size_t want = (size_t)count * elem_floats;
    if (fread(dst, sizeof(float), want, f) != want) { fclose(f); return -4; }
    fclose(f);
    return (int64_t)count;
}
