/*
 * ocean_oracle_impl.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Body of the CPU oracle, included twice by ocean_oracle.c: once with
 * REAL=float (a literal fp32 restatement, also the timed CPU baseline) and once
 * with REAL=double (the parity-defining oracle: f64 everywhere except the two
 * places where fp32 rounding is behaviour-defining, see propagate below).
 *
 * It restates the arithmetic of the reference's four compute shaders
 * (paths relative to /root/reference):
 *   shader/propagate.comp:42-72    -> NAME(propagate)
 *   shader/fft_row.comp:25-63      -> NAME(fft_row)   (butterfly :25-40, main :44-63)
 *   shader/fft_col.comp:44-63      -> NAME(fft_col)
 *   shader/correction.comp:24-35   -> NAME(correction)
 *   src/render.rs:1122-1287        -> NAME(frame)     (dispatch order: propagate,
 *                                     3x row, 3x col, correction)
 * generalised from the hard-coded 512 to any power-of-two N (stage count log2 N,
 * partner offset N/2, row stride N), which is what the reference's constants
 * mean (src/render.rs:42-46: RESOLUTION = 16*32).
 *
 * PARITY UNPINNED: the reference ships no tests, golden outputs or known-answer
 * vectors for this path and cannot be compiled here
 * (no Rust toolchain, no Vulkan). The oracle is pinned only by the shader
 * sources + the shipped SPIR-V (OpConvertUToF checked), by the reference-owned
 * inputs data/omega.bin + data/spectrum.bin and by an independent numpy twin.
 */

#ifndef REAL
#error "include from ocean_oracle.c"
#endif

/* const float pi = 3.1415926;  (propagate.comp:6, fft_row.comp:5) -- the literal
 * rounds to the fp32 value 0x40490FDA = 3.14159250259..., NOT fp32(pi). */
#define PI32 3.1415926f

/* ------------------------------------------------------------------------- */
/* shader/propagate.comp:42-72                                                */
/* ------------------------------------------------------------------------- */
void NAME(propagate)(const float *h0 /* N*N*2 */, const float *omega /* N*N */,
                     float time, int32_t resolution, float domain_size,
                     REAL *height_spec, REAL *disp_x_spec, REAL *disp_z_spec)
{
    const uint32_t n = (uint32_t)resolution;
#pragma omp parallel for schedule(static)
    for (int64_t gy_ = 0; gy_ < (int64_t)n; ++gy_) {
        const uint32_t gy = (uint32_t)gy_;
        for (uint32_t gx = 0; gx < n; ++gx) {
            /* :43  uint index = gid.x + resolution * gid.y */
            const uint32_t index = gx + n * gy;
            /* :45-46  uint x = 2*gid.x - resolution - 1  (u32, wraps for gx <= N/2) */
            const uint32_t xu = 2u * gx - n - 1u;
            const uint32_t yu = 2u * gy - n - 1u;
            /* :48 */
            const uint32_t index_neg = (n - gy - 1u) * n + n - gx - 1u;
            /* :50-53  k = pi * float(x) / domain_size; float(uint) is OpConvertUToF.
             * fp32 rounding is behaviour-defining here: both products/quotients
             * are rounded to fp32 in either instantiation. */
            const float kxf = (float)(PI32 * (float)xu) / domain_size;
            const float kyf = (float)(PI32 * (float)yu) / domain_size;
            /* :55  float disp = omega[index] * time -- fp32 product (behaviour-defining:
             * the phase reaches thousands of radians). */
            const float dispf = omega[index] * time;
#if REAL_IS_DOUBLE
            const REAL c = cos((double)dispf), s = sin((double)dispf);
#else
            const REAL c = cosf(dispf), s = sinf(dispf);
#endif
            /* :56-62  h = h0[idx]*(c,s) + h0[idx_neg]*(c,-s) */
            const REAL ar = h0[2 * index], ai = h0[2 * index + 1];
            const REAL br = h0[2 * index_neg], bi = h0[2 * index_neg + 1];
            const REAL hr = (ar * c - ai * s) + (br * c - bi * (-s));
            const REAL hi = (ai * c + ar * s) + (bi * c + br * (-s));
            /* :64-67  k_norm = k / length(k) if length(k) > 1e-10 */
            const REAL kx = kxf, ky = kyf;
#if REAL_IS_DOUBLE
            const REAL len = sqrt(kx * kx + ky * ky);
#else
            const REAL len = sqrtf(kx * kx + ky * ky);
#endif
            REAL nx = 0, nz = 0;
            if (len > (REAL)1.0e-10) { nx = kx / len; nz = ky / len; }
            /* :69-71  (0,-n)*h = (n*h.y, -n*h.x) */
            height_spec[2 * index] = hr;
            height_spec[2 * index + 1] = hi;
            disp_x_spec[2 * index] = (REAL)0 * hr - (-nx) * hi;
            disp_x_spec[2 * index + 1] = (-nx) * hr + (REAL)0 * hi;
            disp_z_spec[2 * index] = (REAL)0 * hr - (-nz) * hi;
            disp_z_spec[2 * index + 1] = (-nz) * hr + (REAL)0 * hi;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* shader/fft_row.comp:25-40 butterfly + :51-59 stage loop, one line of N      */
/* points held in a ping-pong pair (the shader's shared_row[2][512]).          */
/* tw holds, per stage i, the 2^i twiddles (cos,sin)(pi32*k/2^i): the same     */
/* values the shader recomputes per butterfly, hoisted (bit-identical).        */
/* Returns the index (0/1) of the buffer holding the result.                   */
/* ------------------------------------------------------------------------- */
static int NAME(stockham_line)(REAL *buf0, REAL *buf1, uint32_t n, uint32_t stages,
                               const REAL *tw)
{
    REAL *pp[2] = {buf0, buf1};
    const uint32_t half = n >> 1;
    const REAL *twi = tw;
    for (uint32_t i = 0; i < stages; ++i) {
        const uint32_t bs = 1u << i;
        const REAL *src = pp[i % 2];
        REAL *dst = pp[(i + 1) % 2];
        for (uint32_t index = 0; index < half; ++index) {
            const uint32_t k = index & (bs - 1u);
            const REAL in0r = src[2 * index], in0i = src[2 * index + 1];
            const REAL in1r = src[2 * (index + half)], in1i = src[2 * (index + half) + 1];
            const REAL cr = twi[2 * k], ci = twi[2 * k + 1];
            const REAL tr = in1r * cr - in1i * ci;
            const REAL ti = in1i * cr + in1r * ci;
            const uint32_t dest = (index << 1) - k;
            dst[2 * dest] = in0r + tr;
            dst[2 * dest + 1] = in0i + ti;
            dst[2 * (dest + bs)] = in0r - tr;
            dst[2 * (dest + bs) + 1] = in0i - ti;
        }
        twi += 2 * bs;
    }
    return (int)(stages % 2);
}

static uint32_t NAME(log2u)(uint32_t n)
{
    uint32_t s = 0;
    while ((1u << s) < n) ++s;
    return s;
}

/* fft_row.comp:32-33: theta = pi * float(k) / float(block_size); c = (cos, sin) */
static REAL *NAME(make_twiddles)(uint32_t n)
{
    const uint32_t stages = NAME(log2u)(n);
    REAL *tw = (REAL *)malloc(sizeof(REAL) * 2 * (size_t)n);
    REAL *p = tw;
    for (uint32_t i = 0; i < stages; ++i) {
        const uint32_t bs = 1u << i;
        for (uint32_t k = 0; k < bs; ++k) {
#if REAL_IS_DOUBLE
            const double theta = (double)PI32 * (double)k / (double)bs;
            p[2 * k] = cos(theta);
            p[2 * k + 1] = sin(theta);
#else
            const float theta = PI32 * (float)k / (float)bs;
            p[2 * k] = cosf(theta);
            p[2 * k + 1] = sinf(theta);
#endif
        }
        p += 2 * bs;
    }
    return tw;
}

/* shader/fft_row.comp:44-63: one workgroup per row y, in place. */
void NAME(fft_row)(REAL *data /* N*N*2 */, uint32_t n)
{
    const uint32_t stages = NAME(log2u)(n);
    REAL *tw = NAME(make_twiddles)(n);
#pragma omp parallel
    {
        REAL *b0 = (REAL *)malloc(sizeof(REAL) * 2 * n);
        REAL *b1 = (REAL *)malloc(sizeof(REAL) * 2 * n);
        REAL *pp[2] = {b0, b1};
#pragma omp for schedule(static)
        for (int64_t y = 0; y < (int64_t)n; ++y) {
            REAL *row = data + 2 * (size_t)n * (size_t)y;
            memcpy(b0, row, sizeof(REAL) * 2 * n);          /* :45-47 */
            const int r = NAME(stockham_line)(b0, b1, n, stages, tw);
            memcpy(row, pp[r], sizeof(REAL) * 2 * n);       /* :61-62 */
        }
        free(b0);
        free(b1);
    }
    free(tw);
}

/* shader/fft_col.comp:44-63: one workgroup per column c, element j at c + N*j. */
void NAME(fft_col)(REAL *data /* N*N*2 */, uint32_t n)
{
    const uint32_t stages = NAME(log2u)(n);
    REAL *tw = NAME(make_twiddles)(n);
#pragma omp parallel
    {
        REAL *b0 = (REAL *)malloc(sizeof(REAL) * 2 * n);
        REAL *b1 = (REAL *)malloc(sizeof(REAL) * 2 * n);
        REAL *pp[2] = {b0, b1};
#pragma omp for schedule(static)
        for (int64_t c = 0; c < (int64_t)n; ++c) {
            for (uint32_t j = 0; j < n; ++j) {               /* :45-47 */
                b0[2 * j] = data[2 * ((size_t)c + (size_t)n * j)];
                b0[2 * j + 1] = data[2 * ((size_t)c + (size_t)n * j) + 1];
            }
            const int r = NAME(stockham_line)(b0, b1, n, stages, tw);
            const REAL *res = pp[r];
            for (uint32_t j = 0; j < n; ++j) {               /* :61-62 */
                data[2 * ((size_t)c + (size_t)n * j)] = res[2 * j];
                data[2 * ((size_t)c + (size_t)n * j) + 1] = res[2 * j + 1];
            }
        }
        free(b0);
        free(b1);
    }
    free(tw);
}

/* shader/correction.comp:24-35: out(x,y) = (dx.re, h.re, dz.re)*sign, w = 0.0.
 * Output is the linear row-major float4[y][x] stand-in for the RGBA32F image. */
void NAME(correction)(const REAL *height, const REAL *disp_x, const REAL *disp_z,
                      uint32_t n, REAL *out_rgba /* N*N*4 */)
{
#pragma omp parallel for schedule(static)
    for (int64_t gy_ = 0; gy_ < (int64_t)n; ++gy_) {
        const uint32_t gy = (uint32_t)gy_;
        for (uint32_t gx = 0; gx < n; ++gx) {
            const size_t index = (size_t)gx + (size_t)n * gy;
            const REAL sign_mul = ((gx + gy) % 2u == 0u) ? (REAL)-1.0 : (REAL)1.0;  /* :29 */
            out_rgba[4 * index + 0] = disp_x[2 * index] * sign_mul;                  /* :31 */
            out_rgba[4 * index + 1] = height[2 * index] * sign_mul;
            out_rgba[4 * index + 2] = disp_z[2 * index] * sign_mul;
            out_rgba[4 * index + 3] = (REAL)0.0;                                     /* :34 */
        }
    }
}

/* Consumer step: the "lazy" normal map of shader/ocean.frag:50-66 evaluated at the
 * texel centres of the displacement map (sampler: Linear filter, Tile wrap, src/render.rs:398, so a one-texel
 * textureOffset at a texel centre is the wrapped neighbour). It differentiates channel .x (the dx displacement),
 * as the shader does; `diff = 2.0 / dim` with dim hard-coded to 512 there (":50 TODO textureSize") is 2/N here.
 *   na = normalize(-diff, (x1-x0)/height_scale, 0), nb = normalize(0, (z1-z0)/height_scale, diff),
 *   N = normalize(cross(na, nb)); height_scale = 180 (:19). out = (N.x, N.y, N.z, 0). */
void NAME(normal_map)(const REAL *disp_rgba /* N*N*4 */, uint32_t n, REAL *out_nrm /* N*N*4 */)
{
    const REAL diff = (REAL)2.0 / (REAL)n;
    const REAL height_scale = (REAL)180.0;
#pragma omp parallel for schedule(static)
    for (int64_t y_ = 0; y_ < (int64_t)n; ++y_) {
        const uint32_t y = (uint32_t)y_;
        for (uint32_t x = 0; x < n; ++x) {
            const REAL x0 = disp_rgba[4 * ((size_t)((x + n - 1) % n) + (size_t)n * y)];
            const REAL x1 = disp_rgba[4 * ((size_t)((x + 1) % n) + (size_t)n * y)];
            const REAL z0 = disp_rgba[4 * ((size_t)x + (size_t)n * ((y + n - 1) % n))];
            const REAL z1 = disp_rgba[4 * ((size_t)x + (size_t)n * ((y + 1) % n))];
            REAL nax = -diff, nay = (x1 - x0) / height_scale;
            REAL nby = (z1 - z0) / height_scale, nbz = diff;
#if REAL_IS_DOUBLE
            const REAL la = sqrt(nax * nax + nay * nay), lb = sqrt(nby * nby + nbz * nbz);
#else
            const REAL la = sqrtf(nax * nax + nay * nay), lb = sqrtf(nby * nby + nbz * nbz);
#endif
            nax /= la; nay /= la; nby /= lb; nbz /= lb;
            /* cross((nax, nay, 0), (0, nby, nbz)) */
            REAL cx = nay * nbz, cy = -nax * nbz, cz = nax * nby;
#if REAL_IS_DOUBLE
            const REAL lc = sqrt(cx * cx + cy * cy + cz * cz);
#else
            const REAL lc = sqrtf(cx * cx + cy * cy + cz * cz);
#endif
            REAL *o = out_nrm + 4 * ((size_t)x + (size_t)n * y);
            o[0] = cx / lc; o[1] = cy / lc; o[2] = cz / lc; o[3] = (REAL)0.0;
        }
    }
}

/* src/render.rs:1122-1287: propagate; barrier; row pass on dx,dy,dz; barrier;
 * col pass on dx,dy,dz; barrier; correction. work = 3 * N*N*2 REALs of scratch. */
int NAME(frame)(const float *h0, const float *omega, float time, uint32_t n,
                float domain_size, REAL *out_rgba)
{
    if (n < 2 || (n & (n - 1)) != 0) return -1;
    const size_t field = 2 * (size_t)n * n;
    REAL *work = (REAL *)malloc(sizeof(REAL) * 3 * field);
    if (!work) return -2;
    REAL *hs = work, *dx = work + field, *dz = work + 2 * field;
    NAME(propagate)(h0, omega, time, (int32_t)n, domain_size, hs, dx, dz);
    NAME(fft_row)(dx, n);
    NAME(fft_row)(hs, n);
    NAME(fft_row)(dz, n);
    NAME(fft_col)(dx, n);
    NAME(fft_col)(hs, n);
    NAME(fft_col)(dz, n);
    NAME(correction)(hs, dx, dz, n, out_rgba);
    free(work);
    return 0;
}
