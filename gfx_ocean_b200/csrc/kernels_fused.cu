// The product path: two fused sm_100a kernels per frame.
//
//   k_rows  = propagate (shader/propagate.comp:42-72) + Hermitian fold + row transforms
//             (shader/fft_row.comp:44-63 for the three fields)
//   k_cols  = column transforms (shader/fft_col.comp:44-63) + sign correction + RGBA pack
//             (shader/correction.comp:24-35)
//
// Only the REAL parts of the three inverse transforms are consumed (correction.comp:31), so
// instead of three complex 2-D transforms the kernels run 1.5:
//   * F_S(k) = F(k) + conj(F(-k))  (indices mod N) has the real 2-D transform 2 Re f;
//   * P = dx_S + i dz_S packs two real fields into one complex transform: f_P = 2 Re dx + 2i Re dz;
//   * h_S is Hermitian, so only rows 0..N/2 of its row pass exist (rows 0 and N/2 are real and
//     share one transform), and its column pass packs two columns per complex transform.
// A block of k_rows owns the row pair (y, N-y): both rows' spectra are needed for either fold,
// so nothing is read twice. Intermediate traffic is 12 B per grid point (GP: N x N complex,
// GH: N/2 x N complex) instead of the reference's 24, and the butterflies are halved.
//
// Line transforms are 2-pass: N = R1 * R2, each thread owns P = R1 points in registers
// (fft_reg.cuh), with one trip through padded shared memory between the passes.
#include <cuda_runtime.h>

#include <cmath>
#include <new>
#include <vector>

#include "fft_reg.cuh"
#include "kernels.h"
#include "ocean_common.cuh"

namespace ocean {

template <int N_, int P_>
struct LineCfg {
    static constexpr int N = N_;
    static constexpr int P = P_;        // points per thread
    static constexpr int T = N / P;     // threads per line transform
    static constexpr int R1 = P;        // pass-1 radix (stride T)
    static constexpr int R2 = T;        // pass-2 radix (contiguous)
    static constexpr int SUB2 = P / R2; // pass-2 sub-transforms per thread
    static_assert(R1 * R2 == N && R2 <= P && T <= 32 && 32 % T == 0, "unsupported 2-pass factorisation");
    static constexpr int PADQ = R2 < 32 ? R2 : 32;
    __host__ __device__ static constexpr int pad(int p) { return p + p / PADQ; }
    // line stride (in float2): >= pad(N-1)+1 and == 2 (mod 16) so that 8 lines x 2 rows of
    // 64-bit accesses fall into 16 distinct bank pairs
    static constexpr int LINE = ((pad(N - 1) + 1 - 2 + 15) / 16) * 16 + 2;
};

// ------------------------------------------------------------------------------------------
// k_rows
// ------------------------------------------------------------------------------------------
// (b - i a) * h            with kh = (a, b) = unit wave vector, h complex
__device__ __forceinline__ float2 rot_mul(float4 s)
{
    return make_float2(fmaf(s.w, s.x, s.z * s.y), fmaf(s.w, s.y, -s.z * s.x));
}
// (b - i a) * conj(h)
__device__ __forceinline__ float2 rot_mul_conj(float4 s)
{
    return make_float2(fmaf(s.w, s.x, -s.z * s.y), -fmaf(s.w, s.y, s.z * s.x));
}

template <int N, int P, int PAIRS, int MINB>
__global__ void __launch_bounds__(3 * PAIRS * (N / P), MINB)
k_rows(const float2* __restrict__ h0_all, const float* __restrict__ omega_all, const float2* __restrict__ tw_g,
       const float* __restrict__ kx_g, float2* __restrict__ gp_all, float2* __restrict__ gh_all, float time,
       uint32_t first_tile)
{
    using Cfg = LineCfg<N, P>;
    constexpr int T = Cfg::T, R1 = Cfg::R1, R2 = Cfg::R2;
    constexpr int NT = 3 * PAIRS * T;
    constexpr int NROWS = 2 * PAIRS;
    constexpr int SP = N + 1;             // row pitch of S: element N duplicates element 0, so S[N - x] is -x mod N

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* S = reinterpret_cast<float4*>(smem_raw);                     // [NROWS][N+1]  (h.re, h.im, khat.x, khat.z)
    float2* X = reinterpret_cast<float2*>(smem_raw);                     // [3 PAIRS][LINE], reuses S after phase B's loads
    static_assert(sizeof(float2) * 3 * PAIRS * Cfg::LINE <= sizeof(float4) * NROWS * SP, "exchange lines must fit in S");

    const uint32_t tile = first_tile + blockIdx.y;
    const float2* __restrict__ h0 = h0_all + size_t(tile) * N * N;
    const float* __restrict__ omega = omega_all + size_t(tile) * N * N;
    float2* __restrict__ gp = gp_all + size_t(blockIdx.y) * N * N;
    float2* __restrict__ gh = gh_all + size_t(blockIdx.y) * (N / 2) * N;

    const int tid = threadIdx.x;

    // ---- phase A: propagate.comp for the block's 2*PAIRS rows -> shared memory.
    // Two adjacent points per step (128-bit loads), UB steps batched so their loads are in flight together.
    {
        constexpr int NPAIR = NROWS * N / 2;
        constexpr int ITER = (NPAIR + NT - 1) / NT;
        constexpr int UB = 4;
#pragma unroll 1
        for (int it0 = 0; it0 < ITER; it0 += UB) {
            float4 a[UB], b[UB];
            float2 w[UB], kx[UB];
            float ky[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int i = tid + (it0 + u) * NT;
                const bool valid = (it0 + u < ITER) && i < NPAIR;
                const int slot = valid ? i / (N / 2) : 0;
                const uint32_t x = valid ? 2 * (i % (N / 2)) : 0;
                const uint32_t jp = blockIdx.x * PAIRS + (slot >> 1);    // pair index, 0 = the self-paired rows 0, N/2
                const uint32_t r = jp == 0 ? ((slot & 1) ? N / 2 : 0) : ((slot & 1) ? N - jp : jp);
                const uint32_t index = x + N * r;                        // propagate.comp:43
                const uint32_t index_neg = (N - r - 1u) * N + N - x - 2u;   // :48 for x+1; the partner of x is one further
                a[u] = __ldg(reinterpret_cast<const float4*>(h0 + index));
                b[u] = __ldg(reinterpret_cast<const float4*>(h0 + index_neg));
                w[u] = __ldg(reinterpret_cast<const float2*>(omega + index));
                // kx_g[g] = pi32 * float(uint(2g - N - 1)) / L, tabulated on the host with the shader's fp32 ops
                kx[u] = __ldg(reinterpret_cast<const float2*>(kx_g + x));
                ky[u] = __ldg(kx_g + r);
            }
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int i = tid + (it0 + u) * NT;
                if ((it0 + u < ITER) && i < NPAIR) {
                    const int slot = i / (N / 2);
                    const uint32_t x = 2 * (i % (N / 2));
                    const float2 h_0 = propagate_point_fast(make_float2(a[u].x, a[u].y), make_float2(b[u].z, b[u].w), w[u].x, time);
                    const float2 h_1 = propagate_point_fast(make_float2(a[u].z, a[u].w), make_float2(b[u].x, b[u].y), w[u].y, time);
                    const float2 k_0 = unit_wave_vector_fast(kx[u].x, ky[u]);
                    const float2 k_1 = unit_wave_vector_fast(kx[u].y, ky[u]);
                    float4* row = S + slot * SP;
                    row[x] = make_float4(h_0.x, h_0.y, k_0.x, k_0.y);
                    row[x + 1] = make_float4(h_1.x, h_1.y, k_1.x, k_1.y);
                    if (x == 0) row[N] = make_float4(h_0.x, h_0.y, k_0.x, k_0.y);
                }
            }
        }
    }
    __syncthreads();

    // ---- phase B: three line transforms per row pair
    const int f = tid / T;                // line index in the block
    const int pair = f / 3, seq = f % 3;
    const int k2 = tid % T;
    const uint32_t j = blockIdx.x * PAIRS + pair;
    const bool self_paired = (j == 0);
    const float4* S0 = S + (2 * pair) * SP;
    const float4* S1 = S0 + SP;
    float2* line = X + f * Cfg::LINE;

    float2 v[R1];
    if (seq < 2) {
        // P_S(A; B) = (b_A - i a_A) h_A - (b_B - i a_B) conj(h_B), kh = (a, b):
        //   seq 0: A = (x, y),    B = (-x, N-y)   -> row y of GP
        //   seq 1: A = (-x, N-y), B = (x, y)      -> row N-y of GP, transformed and stored mirrored
        //   rows 0 and N/2 are their own partners: A = (x, r), B = (-x, r), natural order
        const bool mirror_a = !self_paired && seq == 1;
        const float4* PA = (seq == 0) ? S0 : S1;
        const float4* PB = self_paired ? PA : (seq == 0 ? S1 : S0);
        // element k1: x = k1 R2 + k2; A at (mirror_a ? N - x : x), B at the other one
        const int step = mirror_a ? -R2 : R2;
        const float4* pa = PA + (mirror_a ? N - k2 : k2);
        const float4* pb = PB + (mirror_a ? k2 : N - k2);
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) {
            const float2 qa = rot_mul(pa[k1 * step]), qb = rot_mul_conj(pb[-k1 * step]);
            v[k1] = make_float2(qa.x - qb.x, qa.y - qb.y);
        }
    } else if (!self_paired) {
        // h_S(x, y) = h(x, y) + conj h(-x, N-y)
        const float2* pa = reinterpret_cast<const float2*>(S0 + k2);
        const float2* pb = reinterpret_cast<const float2*>(S1 + N - k2);
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) {
            const float2 A = pa[2 * k1 * R2], B = pb[-2 * k1 * R2];
            v[k1] = make_float2(A.x + B.x, A.y - B.y);
        }
    } else {
        // h_S(x, 0) + i h_S(x, N/2): both row transforms are real, so they share one complex transform
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) {
            const int x = k1 * R2 + k2;
            const float4 a0 = S0[x], b0 = S0[N - x], a1 = S1[x], b1 = S1[N - x];
            v[k1] = make_float2((a0.x + b0.x) - (a1.y - b1.y), (a0.y - b0.y) + (a1.x + b1.x));
        }
    }
    __syncthreads();                      // every warp has its inputs: S may be overwritten by the lines
    RegFft<R1>::run(v);
#pragma unroll
    for (int n1 = 0; n1 < R1; ++n1) {
        const float2 y = n1 == 0 ? v[0] : cmul_tw(v[n1], __ldg(tw_g + n1 * R2 + k2));
        line[Cfg::pad(n1 * R2 + k2)] = y;
    }
    __syncwarp();

    float2* dst;
    bool mirrored = false;
    if (seq == 0) dst = gp + size_t(j) * N;
    else if (seq == 1) { dst = gp + size_t(self_paired ? N / 2 : N - j) * N; mirrored = !self_paired; }
    else dst = gh + size_t(j) * N;
#pragma unroll
    for (int i = 0; i < Cfg::SUB2; ++i) {
        const int n1 = k2 + R2 * i;
        float2 u[R2];
#pragma unroll
        for (int k = 0; k < R2; ++k) u[k] = line[Cfg::pad(n1 * R2 + k)];
        RegFft<R2>::run(u);
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) {
            const uint32_t n = n1 + R1 * n2;
            dst[mirrored ? ((N - n) & (N - 1)) : n] = u[n2];
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_cols
// ------------------------------------------------------------------------------------------
template <int N, int P, int C>
__global__ void __launch_bounds__(3 * C * (N / P) / 2, 2)
k_cols(const float2* __restrict__ gp_all, const float2* __restrict__ gh_all, const float2* __restrict__ tw_g,
       float4* __restrict__ out_all, uint32_t first_tile)
{
    using Cfg = LineCfg<N, P>;
    constexpr int T = Cfg::T, R1 = Cfg::R1, R2 = Cfg::R2, LINE = Cfg::LINE;
    constexpr int NTP = C * T;            // threads on the packed (dx, dz) columns
    constexpr int HC = C / 2;             // packed height columns

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* XP = reinterpret_cast<float2*>(smem_raw);      // [C][LINE]
    float2* XH = XP + C * LINE;                            // [C/2][LINE]
    float* HR = reinterpret_cast<float*>(XH);              // [N][C] height results, reuses XH once it is consumed
    static_assert(sizeof(float) * N * C <= sizeof(float2) * HC * LINE, "height results must fit in XH");

    const float2* __restrict__ gp = gp_all + size_t(blockIdx.y) * N * N;
    const float2* __restrict__ gh = gh_all + size_t(blockIdx.y) * (N / 2) * N;
    float4* __restrict__ out = out_all + size_t(first_tile + blockIdx.y) * N * N;

    const int tid = threadIdx.x;
    const uint32_t n0 = blockIdx.x * C;

    const bool is_p = tid < NTP;
    // lanes run over columns first so that a warp's loads cover whole 32/64-byte row segments
    const int c = is_p ? tid % C : (tid - NTP) % HC;
    const int k2 = is_p ? tid / C : (tid - NTP) / HC;
    float2* line = is_p ? XP + c * LINE : XH + c * LINE;

    float2 v[R1];
    if (is_p) {
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) v[k1] = __ldg(gp + size_t(k1 * R2 + k2) * N + n0 + c);
    } else {
        // Z(y) = G(nA, y) + i G(nB, y) over the full column, G(n, N-y) = conj G(n, y),
        // G(n, 0) = Re GH[0][n], G(n, N/2) = Im GH[0][n]
        const uint32_t nA = n0 + c, nB = nA + HC;
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) {
            const uint32_t y = k1 * R2 + k2;
            const uint32_t yy = (y == N / 2) ? 0u : (y > N / 2 ? N - y : y);
            const float2 a = __ldg(gh + size_t(yy) * N + nA), b = __ldg(gh + size_t(yy) * N + nB);
            float2 z;
            if (y == 0) z = make_float2(a.x, b.x);
            else if (y == N / 2) z = make_float2(a.y, b.y);
            else if (y < N / 2) z = make_float2(a.x - b.y, a.y + b.x);
            else z = make_float2(a.x + b.y, b.x - a.y);
            v[k1] = z;
        }
    }
    RegFft<R1>::run(v);
#pragma unroll
    for (int n1 = 0; n1 < R1; ++n1) {
        const float2 y = n1 == 0 ? v[0] : cmul_tw(v[n1], __ldg(tw_g + n1 * R2 + k2));
        line[Cfg::pad(n1 * R2 + k2)] = y;
    }
    __syncthreads();

    // pass 2 (+ output). Thread (c, g) owns outputs m = n1 + R1 n2, n1 = g + R2 i.
    float2 u[Cfg::SUB2][R2];
#pragma unroll
    for (int i = 0; i < Cfg::SUB2; ++i) {
        const int n1 = k2 + R2 * i;
#pragma unroll
        for (int k = 0; k < R2; ++k) u[i][k] = line[Cfg::pad(n1 * R2 + k)];
    }
    // the height threads have drained XH; only they need to agree before HR overwrites it
    if (!is_p) asm volatile("bar.sync 1, %0;" ::"n"(HC * T) : "memory");
#pragma unroll
    for (int i = 0; i < Cfg::SUB2; ++i) {
        const int n1 = k2 + R2 * i;
        RegFft<R2>::run(u[i]);
        if (!is_p) {                      // height columns: park the two real results for the packers
#pragma unroll
            for (int n2 = 0; n2 < R2; ++n2) {
                const int m = n1 + R1 * n2;
                HR[m * C + c] = u[i][n2].x;
                HR[m * C + c + HC] = u[i][n2].y;
            }
        }
    }
    __syncthreads();                      // HR complete
    if (is_p) {
#pragma unroll
        for (int i = 0; i < Cfg::SUB2; ++i) {
            const int n1 = k2 + R2 * i;
#pragma unroll
            for (int n2 = 0; n2 < R2; ++n2) {
                const uint32_t m = n1 + R1 * n2;
                // correction.comp:29 sign, times the 1/2 of the Hermitian fold
                const float s = ((n0 + c + m) & 1u) ? 0.5f : -0.5f;
                out[size_t(m) * N + n0 + c] = make_float4(u[i][n2].x * s, HR[m * C + c] * s, u[i][n2].y * s, 0.0f);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct FusedPlan {
    uint32_t n = 0, n_tiles = 0;
    float domain_size = 0.f;
    float2* d_tw = nullptr;      // [R1][R2] inter-pass twiddles
    float* d_kx = nullptr;       // [N] wave numbers, propagate.comp:45-46,50-53
    float2* d_gp = nullptr;      // [tiles][N][N]
    float2* d_gh = nullptr;      // [tiles][N/2][N]
};

template <int N, int P, int PAIRS, int C, int MINB>
struct Launch {
    using Cfg = LineCfg<N, P>;
    static constexpr size_t smem_rows = sizeof(float4) * 2 * PAIRS * (N + 1);
    static constexpr size_t smem_cols = sizeof(float2) * (C + C / 2) * Cfg::LINE;

    static cudaError_t prepare()
    {
        cudaError_t e = cudaFuncSetAttribute(k_rows<N, P, PAIRS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_rows));
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(k_cols<N, P, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_cols));
    }

    static cudaError_t run(FusedPlan* p, const float2* h0, const float* omega, float4* out, float time,
                           uint32_t first_tile, uint32_t count, cudaStream_t s, cudaEvent_t* ev)
    {
        if (ev) cudaEventRecord(ev[0], s);
        const dim3 grid_rows(N / 2 / PAIRS, count), grid_cols(N / C, count);
        k_rows<N, P, PAIRS, MINB><<<grid_rows, 3 * PAIRS * Cfg::T, smem_rows, s>>>(h0, omega, p->d_tw, p->d_kx, p->d_gp, p->d_gh,
                                                                             time, first_tile);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (ev) cudaEventRecord(ev[1], s);
        k_cols<N, P, C><<<grid_cols, 3 * C * Cfg::T / 2, smem_cols, s>>>(p->d_gp, p->d_gh, p->d_tw, out, first_tile);
        e = cudaGetLastError();
        if (ev) cudaEventRecord(ev[2], s);
        return e;
    }
};

using L256 = Launch<256, 16, 2, 8, 4>;
using L512 = Launch<512, 32, 2, 8, 4>;
using L1024 = Launch<1024, 32, 1, 8, 6>;

bool fused_supports(uint32_t n) { return n == 256 || n == 512 || n == 1024; }

static void line_factors(uint32_t n, uint32_t& r1, uint32_t& r2)
{
    switch (n) {
        case 256: r1 = 16; r2 = 16; break;
        case 512: r1 = 32; r2 = 16; break;
        default: r1 = 32; r2 = n / 32; break;
    }
}

cudaError_t fused_plan_create(FusedPlan** out, uint32_t n, uint32_t n_tiles, float domain_size, int /*device*/)
{
    *out = nullptr;
    if (!fused_supports(n)) return cudaErrorInvalidValue;
    FusedPlan* p = new (std::nothrow) FusedPlan;
    if (!p) return cudaErrorMemoryAllocation;
    p->n = n;
    p->n_tiles = n_tiles;
    p->domain_size = domain_size;
    cudaError_t e;
    auto bail = [&](cudaError_t err) { fused_plan_destroy(p); return err; };

    uint32_t r1, r2;
    line_factors(n, r1, r2);
    std::vector<float2> tw(n);
    for (uint32_t n1 = 0; n1 < r1; ++n1)
        for (uint32_t k2 = 0; k2 < r2; ++k2) {
            const double th = 2.0 * kPiD * double((n1 * k2) % n) / double(n);
            tw[n1 * r2 + k2] = make_float2(float(std::cos(th)), float(std::sin(th)));
        }
    if ((e = cudaMalloc(&p->d_tw, n * sizeof(float2))) != cudaSuccess) return bail(e);
    if ((e = cudaMemcpy(p->d_tw, tw.data(), n * sizeof(float2), cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e);
    // the shader's fp32 arithmetic, op for op: k = pi * float(uint(2g - N - 1)) / domain_size
    std::vector<float> kx(n);
    for (uint32_t g = 0; g < n; ++g) {
        const uint32_t u = 2u * g - n - 1u;
        volatile float prod = kPi32 * float(u);
        kx[g] = prod / domain_size;
    }
    if ((e = cudaMalloc(&p->d_kx, n * sizeof(float))) != cudaSuccess) return bail(e);
    if ((e = cudaMemcpy(p->d_kx, kx.data(), n * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e);
    const size_t np = size_t(n) * n;
    if ((e = cudaMalloc(&p->d_gp, size_t(n_tiles) * np * sizeof(float2))) != cudaSuccess) return bail(e);
    if ((e = cudaMalloc(&p->d_gh, size_t(n_tiles) * (np / 2) * sizeof(float2))) != cudaSuccess) return bail(e);
    switch (n) {
        case 256: e = L256::prepare(); break;
        case 512: e = L512::prepare(); break;
        default: e = L1024::prepare(); break;
    }
    if (e != cudaSuccess) return bail(e);
    *out = p;
    return cudaSuccess;
}

void fused_plan_destroy(FusedPlan* p)
{
    if (!p) return;
    cudaFree(p->d_tw);
    cudaFree(p->d_kx);
    cudaFree(p->d_gp);
    cudaFree(p->d_gh);
    delete p;
}

cudaError_t fused_enqueue(FusedPlan* p, const float2* h0, const float* omega, float4* out, float time,
                          uint32_t first_tile, uint32_t count, cudaStream_t s, uint32_t* launches, cudaEvent_t* ev)
{
    cudaError_t e;
    switch (p->n) {
        case 256: e = L256::run(p, h0, omega, out, time, first_tile, count, s, ev); break;
        case 512: e = L512::run(p, h0, omega, out, time, first_tile, count, s, ev); break;
        case 1024: e = L1024::run(p, h0, omega, out, time, first_tile, count, s, ev); break;
        default: return cudaErrorInvalidValue;
    }
    if (launches) *launches = 2;
    return e;
}

}  // namespace ocean
