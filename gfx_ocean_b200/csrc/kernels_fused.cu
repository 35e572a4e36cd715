// The product path: two fused sm_100a kernels per frame.
//
//   k_rows  = propagate (shader/propagate.comp:42-72) + Hermitian fold + row transforms
//             (shader/fft_row.comp:44-63 for the three fields). Three implementations of the same arithmetic:
//             k_rows_t "tma" (default): one block per row pair; one thread bulk-copies (cp.async.bulk / TMA) the pair's
//                      raw rows into shared memory, the propagated (h, khat) records overwrite them in place, each line
//                      warp folds its own inputs. Always launched with programmatic dependent launch.
//             k_rows   "staged": round 1's kernel, the same with register-staged 128-bit global loads.
//             k_rows_p "persistent" / "fold": fold at the source -- the thread that owns x evaluates both points of a
//                      fold and writes the three folded sequences directly -- as a persistent kernel with a bulk-copy fed
//                      raw-row ring (or an L2-prefetch warp), or with one unit per block. A third of the shared-memory
//                      traffic, but slower on B200: kept as measured alternatives.
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
// Line transforms take two or three passes, N = R1 * R2 (* R3): each thread owns P = R1 points in
// registers (fft_reg.cuh), with one trip through padded shared memory between passes (LineCfg).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <new>
#include <vector>

#include "fft_reg.cuh"
#include "kernels.h"
#include "ocean_common.cuh"

namespace ocean {

// Programmatic dependent launch (both kernels are launched with programmaticStreamSerialization): a
// kernel lets its successor start as SMs drain, and waits for its predecessor's memory only where it
// first touches data the predecessor owns.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait_prior_grid() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// A length-N line transform is done in two or three passes, N = R1 * R2 (* R3); every thread owns P = R1
// points in registers and T = N / P threads share a line:
//   pass 1  radix R1 over the most significant input digit (stride T), twiddle w_N^(n1 t)
//   pass 2  radix R2 (stride R3), twiddle w_T^(n2 k3)            [P / R2 sub-transforms per thread]
//   pass 3  radix R3 (contiguous), only when T > P               [P / R3 sub-transforms per thread]
// with one trip through shared memory between passes; outputs appear at n = n1 + R1 n2 + R1 R2 n3.
#ifndef OCEAN_FFT_PACKED_3PASS
#define OCEAN_FFT_PACKED_3PASS 1
#endif
template <int N_, int P_>
struct LineCfg {
    static constexpr int N = N_;
    static constexpr int P = P_;                  // points per thread
    static constexpr int T = N / P;               // threads per line transform
    static constexpr int R1 = P;
    static constexpr int R2 = T < P ? T : P;
    static constexpr int R3 = T / R2;
    static constexpr int SUB2 = P / R2;           // pass-2 sub-transforms per thread
    static constexpr int SUB3 = P / R3;           // pass-3 sub-transforms per thread (R3 > 1)
    static_assert(R1 * R2 * R3 == N && R3 <= P && (T <= 32 ? 32 % T == 0 : T % 32 == 0), "unsupported factorisation");
    static constexpr int GS = R3 == 1 ? R2 : 32;  // row-group size of the intermediate layout (see Inter)
    // packed f32x2 butterflies (fft_reg.cuh): the twiddle-free ones for the three-pass lines (N = 2048: +3.5 %
    // frames/s, bit-identical), scalar elsewhere (1024: -0.5 %, 512: +-0) unless the build overrides it
    static constexpr int PK = (OCEAN_FFT_PACKED == 0 && R3 > 1) ? OCEAN_FFT_PACKED_3PASS : OCEAN_FFT_PACKED;
    static constexpr int PADQ = T < 32 ? T : 32;
    __host__ __device__ static constexpr int pad(int p) { return p + p / PADQ; }
    // line stride (in float2): >= pad(N-1)+1 and == 2 (mod 16) so that 8 lines x 2 rows of
    // 64-bit accesses fall into 16 distinct bank pairs
    static constexpr int LINE = ((pad(N - 1) + 1 - 2 + 15) / 16) * 16 + 2;
    // same for 4 lines x 4 rows per half-warp (the packed height columns of k_cols): == 4 (mod 16)
    static constexpr int LINE_H = ((pad(N - 1) + 1 - 4 + 15) / 16) * 16 + 4;
    // and for 2 lines x 8 rows per half-warp (4-column strips: two packed height columns): == 8 (mod 16)
    static constexpr int LINE_H2 = ((pad(N - 1) + 1 - 8 + 15) / 16) * 16 + 8;
};

// Passes 2 (and 3) of a line held in shared memory at line[pad(p)], for the three-pass factorisation.
// `t` is the thread's index in the line; `sync` separates the passes; out(n, value) receives natural-order
// results. Generic (not the tuned two-pass path): used for N = 2048.
template <class Cfg, class Sync, class Out>
__device__ __forceinline__ void line_passes_23(float2* line, int t, const float2* __restrict__ tw2, Sync sync, Out out)
{
    constexpr int T = Cfg::T, R1 = Cfg::R1, R2 = Cfg::R2, R3 = Cfg::R3;
    const int g = t / R3, k3 = t % R3;
#pragma unroll
    for (int i = 0; i < Cfg::SUB2; ++i) {
        const int n1 = g + R2 * i;
        float2 u[R2];
#pragma unroll
        for (int k = 0; k < R2; ++k) u[k] = line[Cfg::pad(n1 * T + k * R3 + k3)];
        RegFft<R2, Cfg::PK>::run(u);
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2)
            line[Cfg::pad(n1 * T + n2 * R3 + k3)] = n2 == 0 ? u[0] : cmul_tw(u[n2], tw2[n2 * R3 + k3]);
    }
    sync();
#pragma unroll
    for (int i = 0; i < Cfg::SUB3; ++i) {
        const int q = t + T * i, n1 = q % R1, n2 = q / R1;
        float2 w[R3];
#pragma unroll
        for (int k = 0; k < R3; ++k) w[k] = line[Cfg::pad(n1 * T + n2 * R3 + k)];
        RegFft<R3, Cfg::PK>::run(w);
#pragma unroll
        for (int n3 = 0; n3 < R3; ++n3) out(n1 + R1 * n2 + R1 * R2 * n3, w[n3]);
    }
}

// The three-pass factorisation with R3 = 2 (N = 2048), closing radix-2 as a WARP-SHUFFLE butterfly. Threads t = 2g and
// 2g + 1 of a line (LANE_XOR lanes apart in their warp) hold, after pass 2, the two k3 inputs of every (n1 = g, n2)
// butterfly. Instead of a second trip through shared memory (32 stores + barrier + 32 loads per thread) each lane sends
// its partner the half it does not finish itself: lane k3 = 0 completes the even n2, lane k3 = 1 the odd n2 -- 16
// complex shuffles per thread and no barrier. Same arithmetic as line_passes_23 (a + b, a - b), bit for bit.
// ld(p) reads position p of the pass-1 output; loaded() runs once the thread holds its inputs (the line may then be
// released); out(n, value) receives natural-order results. Every lane of the warp must take part.
#ifndef OCEAN_SHFL_RADIX2
#define OCEAN_SHFL_RADIX2 1
#endif
template <class Cfg, int LANE_XOR, class Ld, class Loaded, class Out>
__device__ __forceinline__ void line_pass2_shfl(int t, const float2* __restrict__ tw2, Ld ld, Loaded loaded, Out out)
{
    constexpr int T = Cfg::T, R1 = Cfg::R1, R2 = Cfg::R2;
    static_assert(Cfg::R3 == 2 && Cfg::SUB2 == 1 && R2 % 2 == 0, "one pass-2 sub-transform per thread, lane pairs close the line");
    const int g = t >> 1, k3 = t & 1;
    float2 u[R2];
#pragma unroll
    for (int k = 0; k < R2; ++k) u[k] = ld(g * T + k * 2 + k3);
    loaded();
    RegFft<R2, Cfg::PK>::run(u);
#pragma unroll
    for (int n2 = 1; n2 < R2; ++n2) u[n2] = cmul_tw(u[n2], tw2[n2 * 2 + k3]);
#pragma unroll
    for (int i = 0; i < R2 / 2; ++i) {
        const float2 send = k3 ? u[2 * i] : u[2 * i + 1];
        const float2 recv = make_float2(__shfl_xor_sync(0xffffffffu, send.x, LANE_XOR), __shfl_xor_sync(0xffffffffu, send.y, LANE_XOR));
        const float2 a = k3 ? recv : u[2 * i], b = k3 ? u[2 * i + 1] : recv;        // the k3 = 0 / k3 = 1 input at n2 = 2 i + k3
        const int n = g + R1 * (2 * i + k3);
        out(n, make_float2(a.x + b.x, a.y + b.y));
        out(n + R1 * R2, make_float2(a.x - b.x, a.y - b.y));
    }
}

// Layout of the row-pass output in HBM/L2 (ours to choose: nothing outside the two kernels sees it).
// Strip-major: all rows of the C columns [s*C, (s+1)*C) are contiguous, so the column pass pulls its
// whole working set with ONE bulk copy; inside a strip rows are grouped by GS (= R2 for two-pass lines,
// 32 for three-pass lines) with one pad row per group, which is exactly the shared-memory image the
// in-place column transform wants (group pitch = 64 (mod 128) bytes for C = 8, 32 (mod 128) for C = 4:
// pass-2 reads of neighbouring groups hit disjoint banks).
//   GP: N rows (packed dx/dz field)      GH: N/2 rows (height field, row 0 = rows 0 and N/2 packed)
template <int N, int C, int GS>
struct Inter {
    static constexpr int GROUP_PITCH = (GS + 1) * C;                     // float2 per row group
    static constexpr int P_STRIP = (N / GS) * GROUP_PITCH;               // float2 per strip of GP
    static constexpr int H_STRIP = (N / 2 / GS) * GROUP_PITCH;           // float2 per strip of GH
    static constexpr size_t P_TILE = size_t(P_STRIP) * (N / C);
    static constexpr size_t H_TILE = size_t(H_STRIP) * (N / C);
    __host__ __device__ static constexpr uint32_t row_off(uint32_t y) { return (y / GS) * GROUP_PITCH + (y % GS) * C; }
    __host__ __device__ static constexpr size_t p_off(uint32_t y, uint32_t n) { return size_t(n / C) * P_STRIP + row_off(y) + n % C; }
    __host__ __device__ static constexpr size_t h_off(uint32_t y, uint32_t n) { return size_t(n / C) * H_STRIP + row_off(y) + n % C; }
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

// Second half of a row line, shared by the row kernels that stage propagated rows (k_rows, k_rows_t): pass 1 in
// registers on the folded inputs v, twiddle, exchange through the line, pass 2 (and 3), strip-major stores.
template <int N, int P, int C>
__device__ __forceinline__ void rows_line_finish(float2 (&v)[LineCfg<N, P>::R1], float2* line, int k2, int seq, uint32_t j,
                                                 const float2* __restrict__ tw_g, float2* __restrict__ gp, float2* __restrict__ gh)
{
    using Cfg = LineCfg<N, P>;
    using IL = Inter<N, C, Cfg::GS>;
    constexpr int T = Cfg::T, R1 = Cfg::R1, R2 = Cfg::R2;
    const bool self_paired = (j == 0);
    RegFft<R1, Cfg::PK>::run(v);
#pragma unroll
    for (int n1 = 0; n1 < R1; ++n1) {
        const float2 y = n1 == 0 ? v[0] : cmul_tw(v[n1], __ldg(tw_g + n1 * T + k2));
        line[Cfg::pad(n1 * T + k2)] = y;
    }
    if constexpr (T <= 32) __syncwarp(); else __syncthreads();
    // The intermediate is still being read by the previous frame's k_cols until that grid has completed: every
    // storing block waits itself (no assumption on block dispatch order; ~0.3 us per block, +2 % on the kernel).
    pdl_wait_prior_grid();

    // Destination row in the strip-major intermediate. Thread n1 owns columns n = n1 + R1 n2: strip n / C and
    // in-strip column n % C advance by a constant per n2 (R1 is a multiple of C), also for the mirrored
    // sequence, whose column (N - n) mod N runs backwards; only n = 0 maps to itself there.
    static_assert(R1 % C == 0, "pass-1 radix must cover whole strips");
    float2* dst;
    bool mirrored = false;
    if (seq == 0) dst = gp + IL::row_off(j);
    else if (seq == 1) { dst = gp + IL::row_off(self_paired ? N / 2 : N - j); mirrored = !self_paired; }
    else dst = gh + IL::row_off(j);
    const long strip_stride = seq == 2 ? IL::H_STRIP : IL::P_STRIP;
    if constexpr (Cfg::R3 > 1) {
        auto store = [&](int n, float2 val) {
            const uint32_t col = mirrored ? ((N - n) & (N - 1)) : n;
            dst[long(col / C) * strip_stride + col % C] = val;
        };
        if constexpr (Cfg::R3 == 2 && OCEAN_SHFL_RADIX2)      // k2 = tid % T with T a multiple of 32: the pair sits in lanes 2g, 2g + 1
            line_pass2_shfl<Cfg, 1>(k2, tw_g + N, [&](int p) { return line[Cfg::pad(p)]; }, [] {}, store);
        else
            line_passes_23<Cfg>(line, k2, tw_g + N, [] { __syncthreads(); }, store);
        return;
    }
    const long step = (mirrored ? -long(R1 / C) : long(R1 / C)) * strip_stride;
#pragma unroll
    for (int i = 0; i < Cfg::SUB2; ++i) {
        const int n1 = k2 + R2 * i;
        float2 u[R2];
#pragma unroll
        for (int k = 0; k < R2; ++k) u[k] = line[Cfg::pad(n1 * R2 + k)];
        RegFft<R2, Cfg::PK>::run(u);
        const uint32_t col0 = mirrored ? ((N - n1) & (N - 1)) : n1;          // column of n2 = 0
        float2* q = dst + long(col0 / C) * strip_stride + col0 % C;
        if (mirrored && n1 == 0) {
            q[0] = u[0];                                                       // n = 0 -> column 0
            q += long(N / C) * strip_stride;                                   // n = R1 n2 -> column N - R1 n2
#pragma unroll
            for (int n2 = 1; n2 < R2; ++n2) q[n2 * step] = u[n2];
        } else {
#pragma unroll
            for (int n2 = 0; n2 < R2; ++n2) q[n2 * step] = u[n2];
        }
    }
}

template <int N, int P, int PAIRS, int C, int MINB>
__global__ void __launch_bounds__(3 * PAIRS * (N / P), MINB)
k_rows(const float2* __restrict__ h0_all, const float* __restrict__ omega_all, const float2* __restrict__ tw_g,
       const float* __restrict__ kx_g, float2* __restrict__ gp_all, float2* __restrict__ gh_all, float time,
       uint32_t first_tile, uint32_t /*unused*/)
{
    using Cfg = LineCfg<N, P>;
    constexpr int T = Cfg::T, R1 = Cfg::R1;
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
    using IL = Inter<N, C, Cfg::GS>;
    float2* __restrict__ gp = gp_all + size_t(blockIdx.y) * IL::P_TILE;
    float2* __restrict__ gh = gh_all + size_t(blockIdx.y) * IL::H_TILE;

    const int tid = threadIdx.x;
    pdl_launch_dependents();

    // ---- phase A: propagate.comp for the block's 2*PAIRS rows -> shared memory.
    // One row at a time; a thread takes point pairs x = 2 (tid + k NT) (128-bit loads), all of a row's
    // loads are issued before any of them is consumed.
#pragma unroll 1
    for (int slot = 0; slot < NROWS; ++slot) {
        const uint32_t jp = blockIdx.x * PAIRS + (slot >> 1);            // pair index, 0 = the self-paired rows 0, N/2
        const uint32_t r = jp == 0 ? ((slot & 1) ? N / 2 : 0) : ((slot & 1) ? N - jp : jp);
        constexpr int ITER = (N / 2 + NT - 1) / NT;
        const float4* __restrict__ pf = reinterpret_cast<const float4*>(h0 + size_t(r) * N) + tid;             // propagate.comp:43
        const float4* __restrict__ pr = reinterpret_cast<const float4*>(h0 + size_t(N - 1 - r) * N + N) - 1 - tid;   // :48
        const float2* __restrict__ pw = reinterpret_cast<const float2*>(omega + size_t(r) * N) + tid;
        const float2* __restrict__ pk = reinterpret_cast<const float2*>(kx_g) + tid;
        // kx_g[g] = pi32 * float(uint(2g - N - 1)) / L, tabulated on the host with the shader's fp32 ops
        const float ky = __ldg(kx_g + r);
        float4 a[ITER], b[ITER];
        float2 w[ITER], kx[ITER];
#pragma unroll
        for (int u = 0; u < ITER; ++u) {
            if (tid + u * NT < N / 2) {
                a[u] = __ldg(pf + u * NT);
                b[u] = __ldg(pr - u * NT);          // .zw is the partner of x, .xy the partner of x + 1
                w[u] = __ldg(pw + u * NT);
                kx[u] = __ldg(pk + u * NT);
            }
        }
        float4* row = S + slot * SP;
#pragma unroll
        for (int u = 0; u < ITER; ++u) {
            if (tid + u * NT < N / 2) {
                const int x = 2 * (tid + u * NT);
                const float2 h_0 = propagate_point_fast(make_float2(a[u].x, a[u].y), make_float2(b[u].z, b[u].w), w[u].x, time);
                const float2 h_1 = propagate_point_fast(make_float2(a[u].z, a[u].w), make_float2(b[u].x, b[u].y), w[u].y, time);
                const float2 k_0 = unit_wave_vector_fast(kx[u].x, ky);
                const float2 k_1 = unit_wave_vector_fast(kx[u].y, ky);
                row[x] = make_float4(h_0.x, h_0.y, k_0.x, k_0.y);
                row[x + 1] = make_float4(h_1.x, h_1.y, k_1.x, k_1.y);
                if (x == 0) row[N] = make_float4(h_0.x, h_0.y, k_0.x, k_0.y);
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
        const int step = mirror_a ? -T : T;
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
            const float2 A = pa[2 * k1 * T], B = pb[-2 * k1 * T];
            v[k1] = make_float2(A.x + B.x, A.y - B.y);
        }
    } else {
        // h_S(x, 0) + i h_S(x, N/2): both row transforms are real, so they share one complex transform
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) {
            const int x = k1 * T + k2;
            const float4 a0 = S0[x], b0 = S0[N - x], a1 = S1[x], b1 = S1[N - x];
            v[k1] = make_float2((a0.x + b0.x) - (a1.y - b1.y), (a0.y - b0.y) + (a1.x + b1.x));
        }
    }
    __syncthreads();                      // every warp has its inputs: S may be overwritten by the lines
    rows_line_finish<N, P, C>(v, line, k2, seq, j, tw_g, gp, gh);
}


// mbarrier / bulk-copy (TMA) / named-barrier wrappers shared by k_rows_p and k_cols
namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        "WAIT_LOOP:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra WAIT_DONE;\n"
        " bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// Same, for the producer thread: back off between polls so the spin does not eat issue slots of the
// compute warps sharing its scheduler (a buffer frees up once per item, microseconds apart).
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity)
{
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(64);
    }
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
template <int ID, int COUNT>
__device__ __forceinline__ void named_bar_sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }
}  // namespace ptx

// ------------------------------------------------------------------------------------------
// k_rows_t: the staged row kernel fed by bulk copies (TMA), records written in place
// ------------------------------------------------------------------------------------------
// Same block shape and phase B as k_rows. Phase A's global loads are replaced by cp.async.bulk: at block start one
// thread copies, per row, the forward row of h0, its reversed partner row and the row of omega into shared memory
// (one mbarrier, one memory round trip for everything, no register staging). The propagated records then overwrite
// the raw rows they were computed from: H[x] = h(x) goes where h0(x) was, K[x] = khat(x) where the thread's partner
// h0(N-1-x) was (so K is stored reversed). Per row: [FWD: N+2 float2][REV: N+2 float2][OM: N float]; FWD[N] duplicates
// H[0] and the slot before REV duplicates K[0], so that index N means 0 for both (the -x mod N of the folds).
#ifndef OCEAN_ROWS_T_PREFETCH
#define OCEAN_ROWS_T_PREFETCH 0
#endif
#ifndef OCEAN_ROWS_T_UNROLL
#define OCEAN_ROWS_T_UNROLL 3           // point pairs of phase A in flight per thread (2 -> 3: +1 % at 1024, same box)
#endif
#ifndef OCEAN_ROWS_T_SPLIT_BAR
#define OCEAN_ROWS_T_SPLIT_BAR 1        // +1.2 % at 1024, +2.8 % at 2048 (same box)
#endif
// OCEAN_ROWS_T_OMEGA_LDG = 1 (A/B builds): omega is not staged in shared memory; every thread fetches the values of its
// own points with read-only global loads issued before it waits for the bulk copies (they stay in registers through
// phase A). The slots shrink from 20 to 16 KB per row at N = 1024, which makes room for a sixth block per SM.
#ifndef OCEAN_ROWS_T_OMEGA_LDG
#define OCEAN_ROWS_T_OMEGA_LDG 0
#endif
template <int N>
struct RowSlot {
    static constexpr uint32_t REV_OFF = 8u * (N + 2), OM_OFF = 16u * (N + 2), BYTES = OM_OFF + (OCEAN_ROWS_T_OMEGA_LDG ? 0u : 4u * N);
    static_assert(REV_OFF % 16 == 0 && BYTES % 16 == 0, "bulk copies need 16-byte granules");
    unsigned char* base;
    __device__ __forceinline__ float2* fwd() const { return reinterpret_cast<float2*>(base); }
    __device__ __forceinline__ float2* rev() const { return reinterpret_cast<float2*>(base + REV_OFF); }
    __device__ __forceinline__ float* om() const { return reinterpret_cast<float*>(base + OM_OFF); }
    // after phase A: record i in [0, N]
    __device__ __forceinline__ float2 h(int i) const { return fwd()[i]; }
    __device__ __forceinline__ float4 rec(int i) const
    {
        const float2 a = fwd()[i], k = rev()[N - 1 - i];
        return make_float4(a.x, a.y, k.x, k.y);
    }
};

template <int N, int P, int PAIRS, int C, int MINB>
__global__ void __launch_bounds__(3 * PAIRS * (N / P), MINB)
k_rows_t(const float2* __restrict__ h0_all, const float* __restrict__ omega_all, const float2* __restrict__ tw_g,
         const float* __restrict__ kx_g, float2* __restrict__ gp_all, float2* __restrict__ gh_all, float time, uint32_t first_tile,
         uint32_t /*keeps the parameter list of all row kernels alike: ocean_update_graph patches `time` by position*/)
{
    using Cfg = LineCfg<N, P>;
    using Slot = RowSlot<N>;
    constexpr int T = Cfg::T, R1 = Cfg::R1;
    constexpr int NT = 3 * PAIRS * T;
    constexpr int NROWS = 2 * PAIRS;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* X = reinterpret_cast<float2*>(smem_raw);                     // [3 PAIRS][LINE], reuses the slots after phase B's loads
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + NROWS * Slot::BYTES);
    static_assert(sizeof(float2) * 3 * PAIRS * Cfg::LINE <= NROWS * Slot::BYTES, "exchange lines must fit in the slots");

    const uint32_t tile = first_tile + blockIdx.y;
    const float2* __restrict__ h0 = h0_all + size_t(tile) * N * N;
    const float* __restrict__ omega = omega_all + size_t(tile) * N * N;
    using IL = Inter<N, C, Cfg::GS>;
    float2* __restrict__ gp = gp_all + size_t(blockIdx.y) * IL::P_TILE;
    float2* __restrict__ gh = gh_all + size_t(blockIdx.y) * IL::H_TILE;

    const int tid = threadIdx.x;
    pdl_launch_dependents();
    auto row_of = [&](int slot) -> uint32_t {
        const uint32_t jp = blockIdx.x * PAIRS + (slot >> 1);            // pair index, 0 = the self-paired rows 0, N/2
        return jp == 0 ? ((slot & 1) ? N / 2 : 0) : ((slot & 1) ? N - jp : jp);
    };
    // OCEAN_ROWS_T_SPLIT_BAR = 1: one mbarrier per row, so that phase A of the first row starts while the second is in flight
    constexpr int NBAR = OCEAN_ROWS_T_SPLIT_BAR ? NROWS : 1;
    if (tid == 0) {
        for (int i = 0; i < NBAR; ++i) ptx::mbar_init(bar + i, 1);
        ptx::fence_mbar_init();
        // h0 / omega are never written by a frame kernel: no griddepcontrol.wait before reading them
        constexpr uint32_t ROW_TX = 2 * N * sizeof(float2) + (OCEAN_ROWS_T_OMEGA_LDG ? 0 : N * sizeof(float));
        for (int i = 0; i < NBAR; ++i) ptx::mbar_arrive_expect_tx(bar + i, (NROWS / NBAR) * ROW_TX);
        for (int slot = 0; slot < NROWS; ++slot) {
            const uint32_t r = row_of(slot);
            const Slot sl{smem_raw + slot * Slot::BYTES};
            uint64_t* b = bar + (NBAR > 1 ? slot : 0);
            ptx::bulk_g2s(sl.fwd(), h0 + size_t(r) * N, N * sizeof(float2), b);                 // propagate.comp:43
            ptx::bulk_g2s(sl.rev(), h0 + size_t(N - 1 - r) * N, N * sizeof(float2), b);         // :48, read reversed below
#if !OCEAN_ROWS_T_OMEGA_LDG
            ptx::bulk_g2s(sl.om(), omega + size_t(r) * N, N * sizeof(float), b);
#endif
        }
    }
#if OCEAN_ROWS_T_OMEGA_LDG
    constexpr int ITER_A = (N / 2 + NT - 1) / NT;
    float2 wreg[NROWS][ITER_A];
#pragma unroll
    for (int slot = 0; slot < NROWS; ++slot) {
        const float2* __restrict__ wrow = reinterpret_cast<const float2*>(omega + size_t(row_of(slot)) * N);
#pragma unroll
        for (int it = 0; it < ITER_A; ++it) {
            const int i = tid + it * NT;
            wreg[slot][it] = i < N / 2 ? __ldg(wrow + i) : make_float2(0.f, 0.f);
        }
    }
#endif
#if OCEAN_ROWS_T_PREFETCH > 0
    {
        // While this block waits for its own rows, pull the rows of the block that will run about one wave later from
        // DRAM into L2, so that its bulk copies are L2 hits (blocks are dispatched roughly in linear order).
        const uint32_t lin = blockIdx.x + gridDim.x * blockIdx.y + OCEAN_ROWS_T_PREFETCH;
        if (lin < gridDim.x * gridDim.y) {
            const uint32_t pbx = lin % gridDim.x, pby = lin / gridDim.x;
            const char* ph0 = reinterpret_cast<const char*>(h0_all + size_t(first_tile + pby) * N * N);
            const char* pom = reinterpret_cast<const char*>(omega_all + size_t(first_tile + pby) * N * N);
            for (int slot = 0; slot < NROWS; ++slot) {
                const uint32_t jp = pbx * PAIRS + (slot >> 1);
                const uint32_t r = jp == 0 ? ((slot & 1) ? N / 2 : 0) : ((slot & 1) ? N - jp : jp);
                for (uint32_t o = tid * 128; o < N * sizeof(float2); o += NT * 128) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(ph0 + size_t(r) * N * sizeof(float2) + o));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(ph0 + size_t(N - 1 - r) * N * sizeof(float2) + o));
                }
                for (uint32_t o = tid * 128; o < N * sizeof(float); o += NT * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pom + size_t(r) * N * sizeof(float) + o));
            }
        }
    }
#endif
    __syncthreads();                      // the barrier is initialised before anyone polls it
    if constexpr (NBAR == 1) ptx::mbar_wait(bar, 0);

    // ---- phase A: propagate.comp, in place. A thread takes point pairs x = 2 (tid + k NT), U pairs (2 U points) at a
    // time: all loads first, then every point's phase and MUFU sincos in straight-line code -- the 2 U dependent
    // chains (multiply, reduce, MUFU, rotate) interleave -- and ONE test per batch for the arguments beyond the fast
    // path's range (|omega t| > 1e5, i.e. t of the order of hours), which then takes libdevice's Payne-Hanek sincos.
    // A test per point, as propagate_point_fast has it, fences every chain between two reconvergence points and
    // leaves each warp with one ~70-cycle dependent chain at a time. Same arithmetic per point, bit for bit.
    auto phase_a_batch = [&](const Slot& sl, float ky, const int (&idx)[OCEAN_ROWS_T_UNROLL], const float2 (&w)[OCEAN_ROWS_T_UNROLL]) {
        constexpr int U = OCEAN_ROWS_T_UNROLL;
        float4* F4 = reinterpret_cast<float4*>(sl.fwd());
        float4* R4 = reinterpret_cast<float4*>(sl.rev());
        const float2* __restrict__ pk = reinterpret_cast<const float2*>(kx_g);
        float4 a[U], b[U];
        float2 kx[U];
        float ph[2 * U], sn[2 * U], cs[2 * U];
        float worst = 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool on = idx[u] < N / 2;
            const int i = on ? idx[u] : 0;
            const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
            a[u] = on ? F4[i] : zero;                    // (an idle slot of the batch reads nothing: predicated loads)
            b[u] = on ? R4[N / 2 - 1 - i] : zero;        // b.zw is the partner of x = 2 i, b.xy the partner of x + 1
            kx[u] = __ldg(pk + i);
            ph[2 * u] = on ? __fmul_rn(w[u].x, time) : 0.f;
            ph[2 * u + 1] = on ? __fmul_rn(w[u].y, time) : 0.f;
            worst = fmaxf(worst, fmaxf(fabsf(ph[2 * u]), fabsf(ph[2 * u + 1])));
        }
#pragma unroll
        for (int q = 0; q < 2 * U; ++q) sincos_in_range(ph[q], sn[q], cs[q]);
        if (!(worst <= 1.0e5f)) {                        // also taken for NaN phases, like sincos_full's test
#pragma unroll
            for (int q = 0; q < 2 * U; ++q)
                if (!(fabsf(ph[q]) <= 1.0e5f)) {
                    const float2 sc = sincos_huge(ph[q]);
                    sn[q] = sc.x;
                    cs[q] = sc.y;
                }
        }
        float2 h_0[U], h_1[U], k_0[U], k_1[U];           // computed for every slot of the batch (an idle slot works on zeros), stored by the live ones
#pragma unroll
        for (int u = 0; u < U; ++u) {
            h_0[u] = propagate_point_sc(make_float2(a[u].x, a[u].y), make_float2(b[u].z, b[u].w), sn[2 * u], cs[2 * u]);
            h_1[u] = propagate_point_sc(make_float2(a[u].z, a[u].w), make_float2(b[u].x, b[u].y), sn[2 * u + 1], cs[2 * u + 1]);
            k_0[u] = unit_wave_vector_fast(kx[u].x, ky);
            k_1[u] = unit_wave_vector_fast(kx[u].y, ky);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = idx[u];
            if (i < N / 2) {
                F4[i] = make_float4(h_0[u].x, h_0[u].y, h_1[u].x, h_1[u].y);             // H[x], H[x + 1]
                R4[N / 2 - 1 - i] = make_float4(k_1[u].x, k_1[u].y, k_0[u].x, k_0[u].y); // K[x + 1] at REV[N-2-x], K[x] at REV[N-1-x]
                if (i == 0) {
                    sl.fwd()[N] = h_0[u];                                                // H[N] = H[0]
                    sl.rev()[-1] = k_0[u];                                               // K[N] = K[0] (the spare slot before REV)
                }
            }
        }
    };
    {
        constexpr int U = OCEAN_ROWS_T_UNROLL;
#if OCEAN_ROWS_T_OMEGA_LDG
#pragma unroll
        for (int slot = 0; slot < NROWS; ++slot) {
            const Slot sl{smem_raw + slot * Slot::BYTES};
            const float ky = __ldg(kx_g + row_of(slot));
            if constexpr (NBAR > 1) ptx::mbar_wait(bar + slot, 0);
#pragma unroll
            for (int it = 0; it < ITER_A; it += U) {
                int idx[U];
                float2 w[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    idx[u] = it + u < ITER_A ? tid + (it + u) * NT : N;
                    w[u] = wreg[slot][it + u < ITER_A ? it + u : 0];
                }
                phase_a_batch(sl, ky, idx, w);
            }
        }
#else
#pragma unroll 1
        for (int slot = 0; slot < NROWS; ++slot) {
            const Slot sl{smem_raw + slot * Slot::BYTES};
            const float ky = __ldg(kx_g + row_of(slot));
            const float2* W2 = reinterpret_cast<const float2*>(sl.om());
            if constexpr (NBAR > 1) ptx::mbar_wait(bar + slot, 0);
#pragma unroll 1
            for (int i0 = tid; i0 < N / 2; i0 += U * NT) {
                int idx[U];
                float2 w[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    idx[u] = i0 + u * NT;
                    w[u] = W2[idx[u] < N / 2 ? idx[u] : 0];
                }
                phase_a_batch(sl, ky, idx, w);
            }
        }
#endif
    }
    __syncthreads();

    // ---- phase B: three line transforms per row pair
    const int f = tid / T;                // line index in the block
    const int pair = f / 3, seq = f % 3;
    const int k2 = tid % T;
    const uint32_t j = blockIdx.x * PAIRS + pair;
    const bool self_paired = (j == 0);
    const Slot S0{smem_raw + (2 * pair) * Slot::BYTES}, S1{smem_raw + (2 * pair + 1) * Slot::BYTES};
    float2* line = X + f * Cfg::LINE;

    float2 v[R1];
    if (seq < 2) {
        // as in k_rows: seq 0: A = (x, y), B = (-x, N-y); seq 1: A = (-x, N-y), B = (x, y); rows 0 and N/2: A = (x, r), B = (-x, r)
        const bool mirror_a = !self_paired && seq == 1;
        const Slot PA = (seq == 0) ? S0 : S1;
        const Slot PB = self_paired ? PA : (seq == 0 ? S1 : S0);
        const int step = mirror_a ? -T : T;
        const int ia = mirror_a ? N - k2 : k2, ib = mirror_a ? k2 : N - k2;
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) {
            const float2 qa = rot_mul(PA.rec(ia + k1 * step)), qb = rot_mul_conj(PB.rec(ib - k1 * step));
            v[k1] = make_float2(qa.x - qb.x, qa.y - qb.y);
        }
    } else if (!self_paired) {
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) {
            const float2 A = S0.h(k2 + k1 * T), B = S1.h(N - k2 - k1 * T);
            v[k1] = make_float2(A.x + B.x, A.y - B.y);
        }
    } else {
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) {
            const int x = k1 * T + k2;
            const float2 a0 = S0.h(x), b0 = S0.h(N - x), a1 = S1.h(x), b1 = S1.h(N - x);
            v[k1] = make_float2((a0.x + b0.x) - (a1.y - b1.y), (a0.y - b0.y) + (a1.x + b1.x));
        }
    }
    __syncthreads();                      // every warp has its inputs: the slots may be overwritten by the lines
    rows_line_finish<N, P, C>(v, line, k2, seq, j, tw_g, gp, gh);
}

// ------------------------------------------------------------------------------------------
// k_rows_p: persistent, bulk-copy fed k_rows with the Hermitian fold done at the source
// ------------------------------------------------------------------------------------------
// One CTA per SM: GROUPS independent groups of 3*PAIRS*T threads plus (SLOTS > 0) one producer warp. A group takes
// "units" (PAIRS consecutive row pairs of one tile) from the CTA's contiguous share of the launch through a
// shared-memory ticket. Per unit:
//   producer one elected thread bulk-copies (cp.async.bulk + mbarrier complete_tx) the unit's raw rows -- per row pair
//            four rows of h0 and two of omega, 40 KB at N = 1024 -- into a ring of SLOTS unit slots, running ahead of
//            the groups: phase A never waits on DRAM and no register stages a global load.
//   phase A  all threads of the group walk x: the thread that owns x evaluates propagate.comp for BOTH points a
//            fold needs -- (x, y) and (-x, N-y) -- and writes the three folded sequences (row y of P, row N-y of
//            P, row y of h_S) straight into the lines the transforms run on. Every spectrum value and unit wave
//            vector is computed once and never parked in shared memory: per row pair the shared-memory traffic
//            of the fold is 24 KB written + 24 KB read, against 32 KB written + 96 KB read when the propagated
//            rows were staged as (h, khat) records and folded by each of the three line warps (round 1; its ncu
//            profile had the L1/shared pipe, not DRAM, as the busiest unit).
//   phase B  one line per T threads: pass 1 in registers, twiddle, IN-PLACE exchange through the same line
//            (a thread writes back exactly the padded slots it read), pass 2 (and 3), strip-major stores.
// A thread executes griddepcontrol.wait once, before its first store to the intermediate, so the write-after-read
// hazard against the previous frame's k_cols is closed per storing thread (no assumption on block dispatch order).
// The inputs h0 / omega are never written by a frame kernel, so the producer's copies need no such wait.
#ifndef OCEAN_ROWS_UNROLL
#define OCEAN_ROWS_UNROLL 4
#endif
#ifndef OCEAN_ROWS_PF_OP
#define OCEAN_ROWS_PF_OP "prefetch.global.L2"
#endif
template <int N, int P, int PAIRS, int C, int GROUPS, int SLOTS>
struct RowsCfg {
    using Line = LineCfg<N, P>;
    static constexpr int T = Line::T;
    static constexpr int LINES = 3 * PAIRS;               // lines per group
    static constexpr int GT = LINES * T;                  // threads per group
    static constexpr int RING = SLOTS > 0 ? SLOTS : 0;    // SLOTS > 0: bulk-copy ring; SLOTS < 0: L2 prefetch warp, -SLOTS units ahead
    static constexpr int AHEAD = SLOTS < 0 ? -SLOTS : 0;
    static constexpr int NTHREADS = GROUPS * GT + (SLOTS != 0 ? 32 : 0);
    static constexpr int UPT = N / 2 / PAIRS;             // units per tile
    static constexpr size_t GROUP_SMEM = sizeof(float2) * LINES * Line::LINE;
    static constexpr size_t LINES_SMEM = GROUPS * GROUP_SMEM;
    // raw rows of one row pair: h0 rows [F0][P0][F1][P1], omega rows [W0][W1]
    static constexpr uint32_t PAIR_RAW = 4 * N * sizeof(float2) + 2 * N * sizeof(float);
    static constexpr uint32_t SLOT_BYTES = PAIRS * PAIR_RAW;
    static constexpr size_t SMEM = LINES_SMEM + size_t(RING) * SLOT_BYTES + (RING > 0 ? 2 * RING * sizeof(uint64_t) : 0);
    static_assert(GT % 32 == 0 && GROUPS >= 1 && GROUPS <= 15, "groups synchronise on named barriers 1..GROUPS");
    static_assert(LINES_SMEM % 16 == 0 && PAIR_RAW % 16 == 0, "bulk copies need 16-byte granules");
};

// propagate.comp:42-72 for one grid point, as the record the folds consume: (h.re, h.im, khat.x, khat.z)
__device__ __forceinline__ float4 propagate_record(float2 h0, float2 h0_rev, float omega, float kx, float ky, float time)
{
    const float2 h = propagate_point_fast(h0, h0_rev, omega, time);
    const float2 k = unit_wave_vector_fast(kx, ky);
    return make_float4(h.x, h.y, k.x, k.y);
}

// rows of h0 / omega a row pair needs (in slot order F0 P0 F1 P1 | W0 W1). h0's partner is the REVERSED array
// (propagate.comp:48): the partner of (x, r) is (N-1-x, N-1-r).
//   pair j > 0: A = (x, j) and B = (-x, N-j):  F0 = j, P0 = N-1-j, F1 = N-j, P1 = j-1;      W0 = j, W1 = N-j
//   pair 0    : rows 0 and N/2 are their own partners: F0 = 0, P0 = N-1, F1 = N/2, P1 = N/2-1;  W0 = 0, W1 = N/2
template <int N>
__device__ __forceinline__ void pair_rows(uint32_t j, uint32_t (&h)[4], uint32_t (&w)[2])
{
    if (j) { h[0] = j; h[1] = N - 1 - j; h[2] = N - j; h[3] = j - 1; w[0] = j; w[1] = N - j; }
    else { h[0] = 0; h[1] = N - 1; h[2] = N / 2; h[3] = N / 2 - 1; w[0] = 0; w[1] = N / 2; }
}

// Phase A for one row pair. F0 P0 F1 P1 W0 W1: the six rows (global memory with read-only loads, or the raw slot in
// shared memory). Writes the three folded sequences into L0 L1 L2 at pad(x).
template <class Cfg, bool GLOBAL_SRC, class Ld2, class Ld1>
__device__ __forceinline__ void fold_pair(uint32_t jp, const float2* F0, const float2* P0, const float2* F1, const float2* P1,
                                          const float* W0, const float* W1, const float* __restrict__ kx_g, float time,
                                          float2* L0, float2* L1, float2* L2, int gt, int GT, Ld2 ld2, Ld1 ld1)
{
    constexpr int N = Cfg::N;
    const int ITER = (N + GT - 1) / GT;
    if (jp != 0) {
        // A = (x, y) from F0 / P0 / W0, B = ((N-x)%N, N-y) from F1 / P1 / W1 (its partner sits at column (x-1)%N)
        const float kya = __ldg(kx_g + jp), kyb = __ldg(kx_g + (N - jp));
        constexpr int U = GLOBAL_SRC ? OCEAN_ROWS_UNROLL : 2;          // x values in flight per thread
#pragma unroll 1
        for (int u0 = 0; u0 < ITER; u0 += U) {
            float2 a0[U], a1[U], b0[U], b1[U];
            float oa[U], ob[U], ka[U], kb[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t x = gt + (u0 + u) * GT;
                if (x < N) {
                    const uint32_t xm = (N - x) & (N - 1), xp = (x - 1) & (N - 1);
                    a0[u] = ld2(F0 + x);
                    a1[u] = ld2(P0 + (N - 1 - x));
                    oa[u] = ld1(W0 + x);
                    b0[u] = ld2(F1 + xm);
                    b1[u] = ld2(P1 + xp);
                    ob[u] = ld1(W1 + xm);
                    ka[u] = __ldg(kx_g + x);
                    kb[u] = __ldg(kx_g + xm);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t x = gt + (u0 + u) * GT;
                if (x < N) {
                    const float4 A = propagate_record(a0[u], a1[u], oa[u], ka[u], kya, time);
                    const float4 B = propagate_record(b0[u], b1[u], ob[u], kb[u], kyb, time);
                    const float2 qa = rot_mul(A), qac = rot_mul_conj(A), qb = rot_mul(B), qbc = rot_mul_conj(B);
                    const int s = Cfg::pad(x);
                    L0[s] = make_float2(qa.x - qbc.x, qa.y - qbc.y);      // row y of P_S
                    L1[s] = make_float2(qb.x - qac.x, qb.y - qac.y);      // row N-y of P_S, mirrored sequence
                    L2[s] = make_float2(A.x + B.x, A.y - B.y);            // row y of h_S
                }
            }
        }
    } else {
        // rows 0 and N/2: P_S(x; r) from (x, r) and (-x, r) in natural order, and h_S(x, 0) + i h_S(x, N/2) (both row
        // transforms are real, so they share one complex transform)
#pragma unroll 1
        for (uint32_t x = gt; x < N; x += GT) {
            const uint32_t xm = (N - x) & (N - 1), xp = (x - 1) & (N - 1);
            const float kxa = __ldg(kx_g + x), kxb = __ldg(kx_g + xm);
            float2 hs[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float ky = __ldg(kx_g + (i ? N / 2 : 0));
                const float2* fr = i ? F1 : F0;
                const float2* pr = i ? P1 : P0;
                const float* wr = i ? W1 : W0;
                const float4 A = propagate_record(ld2(fr + x), ld2(pr + (N - 1 - x)), ld1(wr + x), kxa, ky, time);
                const float4 B = propagate_record(ld2(fr + xm), ld2(pr + xp), ld1(wr + xm), kxb, ky, time);
                const float2 qa = rot_mul(A), qbc = rot_mul_conj(B);
                (i ? L1 : L0)[Cfg::pad(x)] = make_float2(qa.x - qbc.x, qa.y - qbc.y);
                hs[i] = make_float2(A.x + B.x, A.y - B.y);
            }
            L2[Cfg::pad(x)] = make_float2(hs[0].x - hs[1].y, hs[0].y + hs[1].x);
        }
    }
}

// (ptxas sizes registers for the thread count rounded up to whole groups of four warps: each of the SM's four
// sub-partitions has its own 16 K-register file, so 15 warps cost what 16 do -- the producer warp is free)
#ifndef OCEAN_ROWS_FOLD_MINB
#define OCEAN_ROWS_FOLD_MINB 5
#endif
template <int N, int P, int PAIRS, int C, int GROUPS, int SLOTS>
__global__ void __launch_bounds__(RowsCfg<N, P, PAIRS, C, GROUPS, SLOTS>::NTHREADS, GROUPS == 1 ? OCEAN_ROWS_FOLD_MINB : 1)
k_rows_p(const float2* __restrict__ h0_all, const float* __restrict__ omega_all, const float2* __restrict__ tw_g,
         const float* __restrict__ kx_g, float2* __restrict__ gp_all, float2* __restrict__ gh_all, float time,
         uint32_t first_tile, uint32_t n_units)
{
    using RC = RowsCfg<N, P, PAIRS, C, GROUPS, SLOTS>;
    using Cfg = typename RC::Line;
    using IL = Inter<N, C, Cfg::GS>;
    constexpr int T = Cfg::T, R1 = Cfg::R1, R2 = Cfg::R2, GT = RC::GT;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_next;
    __shared__ uint32_t s_unit[2][GROUPS];   // the group's next unit, double-buffered by iteration parity
    __shared__ uint32_t s_done[SLOTS > 0 ? SLOTS : 1];   // units whose phase A has drained each raw slot (monotonic)
    unsigned char* raw_base = smem_raw + RC::LINES_SMEM;
    uint64_t* full = reinterpret_cast<uint64_t*>(raw_base + size_t(RC::RING) * RC::SLOT_BYTES);   // [RING] raw rows landed (tx bytes)
    uint64_t* empty = full + RC::RING;                                                             // [RING] phase A done (GT arrivals)

    const int tid = threadIdx.x;
    pdl_launch_dependents();
    const uint32_t u_begin = uint32_t(uint64_t(blockIdx.x) * n_units / gridDim.x);
    const uint32_t u_end = uint32_t(uint64_t(blockIdx.x + 1) * n_units / gridDim.x);
    if (tid == 0) {
        s_next = u_begin + GROUPS;
        if constexpr (SLOTS > 0) {
            for (int i = 0; i < SLOTS; ++i) {
                s_done[i] = 0;
                ptx::mbar_init(full + i, 1);
                ptx::mbar_init(empty + i, GT);
            }
            ptx::fence_mbar_init();
        }
    }
    __syncthreads();

    if constexpr (SLOTS < 0) {
        if (tid >= GROUPS * GT) {
            // ================= prefetch warp =================
            // Pulls the raw rows of the units the groups will take next from DRAM into L2 (prefetch.global.L2), a few
            // units ahead of the ticket counter, so phase A's loads are L2 hits. No shared memory, no registers of the
            // compute warps, and nothing to wait for: a late prefetch only costs the latency it failed to hide.
            const uint32_t lane = tid & 31;
            uint32_t done = u_begin + GROUPS;
            while (done < u_end) {
                const uint32_t next = *reinterpret_cast<volatile uint32_t*>(&s_next);
                if (done < next + RC::AHEAD) {
                    const uint32_t tl = done / RC::UPT, bx = done % RC::UPT;
                    const char* h0 = reinterpret_cast<const char*>(h0_all + size_t(first_tile + tl) * N * N);
                    const char* om = reinterpret_cast<const char*>(omega_all + size_t(first_tile + tl) * N * N);
#pragma unroll 1
                    for (int p = 0; p < PAIRS; ++p) {
                        uint32_t hr[4], wr[2];
                        pair_rows<N>(bx * PAIRS + p, hr, wr);
                        for (int i = 0; i < 4; ++i)
                            for (uint32_t o = lane * 128; o < N * sizeof(float2); o += 32 * 128)
                                asm volatile(OCEAN_ROWS_PF_OP " [%0];" ::"l"(h0 + size_t(hr[i]) * N * sizeof(float2) + o));
                        for (int i = 0; i < 2; ++i)
                            for (uint32_t o = lane * 128; o < N * sizeof(float); o += 32 * 128)
                                asm volatile(OCEAN_ROWS_PF_OP " [%0];" ::"l"(om + size_t(wr[i]) * N * sizeof(float) + o));
                    }
                    ++done;
                } else {
                    __nanosleep(256);
                }
            }
            return;
        }
    }
    if constexpr (SLOTS > 0) {
        if (tid >= GROUPS * GT) {
            // ================= producer warp =================
            if ((tid & 31) == 0) {
                for (uint32_t k = 0; u_begin + k < u_end; ++k) {
                    const uint32_t slot = k % SLOTS, unit = u_begin + k;
                    const uint32_t tl = unit / RC::UPT, bx = unit % RC::UPT;
                    const float2* h0 = h0_all + size_t(first_tile + tl) * N * N;
                    const float* omega = omega_all + size_t(first_tile + tl) * N * N;
                    if (k >= SLOTS) {
                        ptx::mbar_wait_backoff(empty + slot, ((k / SLOTS) - 1) & 1);    // the group that had this slot is done reading
                        ptx::fence_proxy_async();
                    }
                    ptx::mbar_arrive_expect_tx(full + slot, RC::SLOT_BYTES);
                    unsigned char* dst = raw_base + size_t(slot) * RC::SLOT_BYTES;
#pragma unroll 1
                    for (int p = 0; p < PAIRS; ++p) {
                        uint32_t hr[4], wr[2];
                        pair_rows<N>(bx * PAIRS + p, hr, wr);
                        for (int i = 0; i < 4; ++i, dst += N * sizeof(float2))
                            ptx::bulk_g2s(dst, h0 + size_t(hr[i]) * N, N * sizeof(float2), full + slot);
                        for (int i = 0; i < 2; ++i, dst += N * sizeof(float))
                            ptx::bulk_g2s(dst, omega + size_t(wr[i]) * N, N * sizeof(float), full + slot);
                    }
                }
            }
            return;
        }
    }

    const int g = tid / GT, gt = tid % GT;
    float2* lines = reinterpret_cast<float2*>(smem_raw) + size_t(g) * RC::LINES * Cfg::LINE;
    auto group_bar = [&] { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(GT) : "memory"); };
    bool waited = false;

    // phase B roles
    const int f = gt / T;                 // line in the group
    const int pair = f / 3, seq = f % 3;
    const int k2 = gt % T;
    float2* line = lines + f * Cfg::LINE;

    uint32_t par = 0;
#pragma unroll 1
    for (uint32_t unit = u_begin + g; unit < u_end; par ^= 1u) {
        const uint32_t tl = unit / RC::UPT, bx = unit % RC::UPT;

        // ---- phase A: propagate + fold -> lines
        if constexpr (SLOTS > 0) {
            // The unit's slot is shared with the units SLOTS, 2 SLOTS, ... before it, which other groups consume. A
            // phase-parity wait alone cannot tell "my rows landed" from "the rows of two occupants ago landed", so
            // first wait on a monotonic count until every earlier occupant has drained the slot; from then on the
            // barrier is either pending on this unit's copy or completed by it.
            const uint32_t k = unit - u_begin, slot = k % SLOTS, occ = k / SLOTS;
            while (*reinterpret_cast<volatile uint32_t*>(&s_done[slot]) < occ) __nanosleep(32);
            ptx::mbar_wait(full + slot, occ & 1);
            const unsigned char* src = raw_base + size_t(slot) * RC::SLOT_BYTES;
#pragma unroll 1
            for (int p = 0; p < PAIRS; ++p, src += RC::PAIR_RAW) {
                const float2* R = reinterpret_cast<const float2*>(src);
                const float* W = reinterpret_cast<const float*>(src + 4 * N * sizeof(float2));
                float2* L0 = lines + (3 * p) * Cfg::LINE;
                fold_pair<Cfg, false>(bx * PAIRS + p, R, R + N, R + 2 * N, R + 3 * N, W, W + N, kx_g, time, L0, L0 + Cfg::LINE,
                                      L0 + 2 * Cfg::LINE, gt, GT, [](const float2* q) { return *q; }, [](const float* q) { return *q; });
            }
            ptx::fence_proxy_async();             // generic reads of the slot precede the bulk copy that refills it
            ptx::mbar_arrive(empty + slot);       // this thread's reads of the slot are done
        } else {
            const float2* __restrict__ h0 = h0_all + size_t(first_tile + tl) * N * N;
            const float* __restrict__ omega = omega_all + size_t(first_tile + tl) * N * N;
#pragma unroll 1
            for (int p = 0; p < PAIRS; ++p) {
                uint32_t hr[4], wr[2];
                pair_rows<N>(bx * PAIRS + p, hr, wr);
                float2* L0 = lines + (3 * p) * Cfg::LINE;
                fold_pair<Cfg, true>(bx * PAIRS + p, h0 + size_t(hr[0]) * N, h0 + size_t(hr[1]) * N, h0 + size_t(hr[2]) * N,
                                     h0 + size_t(hr[3]) * N, omega + size_t(wr[0]) * N, omega + size_t(wr[1]) * N, kx_g, time, L0,
                                     L0 + Cfg::LINE, L0 + 2 * Cfg::LINE, gt, GT, [](const float2* q) { return __ldg(q); },
                                     [](const float* q) { return __ldg(q); });
            }
        }
        // Next unit of this group, read by every thread after the unit's last barrier. Two slots: a warp that is slow
        // to read slot `par` after that barrier cannot be overtaken by the write of the next iteration (other slot);
        // the write after that one is two barriers away. (With one slot a starved warp read the following ticket
        // and a row went missing -- found by compute-sanitizer racecheck, seen as sporadic errors in multi-tile runs.)
        if (gt == 0) s_unit[par][g] = atomicAdd(&s_next, 1u);
        group_bar();
        if constexpr (SLOTS > 0) {
            if (gt == 0) *reinterpret_cast<volatile uint32_t*>(&s_done[(unit - u_begin) % SLOTS]) = (unit - u_begin) / SLOTS + 1;
        }

        // ---- phase B: one line transform per T threads
        const uint32_t j = bx * PAIRS + pair;
        const bool self_paired = (j == 0);
        float2 v[R1];
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) v[k1] = line[Cfg::pad(k1 * T + k2)];
        RegFft<R1, Cfg::PK>::run(v);
#pragma unroll
        for (int n1 = 0; n1 < R1; ++n1)          // in place: the slots this thread just read
            line[Cfg::pad(n1 * T + k2)] = n1 == 0 ? v[0] : cmul_tw(v[n1], __ldg(tw_g + n1 * T + k2));
        if constexpr (T <= 32) __syncwarp(); else group_bar();
        if (!waited) {                       // the previous frame's k_cols may still be reading the intermediate
            pdl_wait_prior_grid();
            waited = true;
        }

        // Destination row in the strip-major intermediate. Thread n1 owns columns n = n1 + R1 n2: strip n / C and
        // in-strip column n % C advance by a constant per n2 (R1 is a multiple of C), also for the mirrored
        // sequence, whose column (N - n) mod N runs backwards; only n = 0 maps to itself there.
        static_assert(R1 % C == 0, "pass-1 radix must cover whole strips");
        float2* __restrict__ gp = gp_all + size_t(tl) * IL::P_TILE;
        float2* __restrict__ gh = gh_all + size_t(tl) * IL::H_TILE;
        float2* dst;
        bool mirrored = false;
        if (seq == 0) dst = gp + IL::row_off(j);
        else if (seq == 1) { dst = gp + IL::row_off(self_paired ? N / 2 : N - j); mirrored = !self_paired; }
        else dst = gh + IL::row_off(j);
        const long strip_stride = seq == 2 ? IL::H_STRIP : IL::P_STRIP;
        if constexpr (Cfg::R3 > 1) {
            line_passes_23<Cfg>(line, k2, tw_g + N, [&] { group_bar(); }, [&](int n, float2 val) {
                const uint32_t col = mirrored ? ((N - n) & (N - 1)) : n;
                dst[long(col / C) * strip_stride + col % C] = val;
            });
        } else {
            const long step = (mirrored ? -long(R1 / C) : long(R1 / C)) * strip_stride;
#pragma unroll
            for (int i = 0; i < Cfg::SUB2; ++i) {
                const int n1 = k2 + R2 * i;
                float2 u[R2];
#pragma unroll
                for (int k = 0; k < R2; ++k) u[k] = line[Cfg::pad(n1 * R2 + k)];
                RegFft<R2, Cfg::PK>::run(u);
                const uint32_t col0 = mirrored ? ((N - n1) & (N - 1)) : n1;          // column of n2 = 0
                float2* q = dst + long(col0 / C) * strip_stride + col0 % C;
                if (mirrored && n1 == 0) {
                    q[0] = u[0];                                                       // n = 0 -> column 0
                    q += long(N / C) * strip_stride;                                   // n = R1 n2 -> column N - R1 n2
#pragma unroll
                    for (int n2 = 1; n2 < R2; ++n2) q[n2 * step] = u[n2];
                } else {
#pragma unroll
                    for (int n2 = 0; n2 < R2; ++n2) q[n2 * step] = u[n2];
                }
            }
        }
        group_bar();                          // lines are free for the next unit's phase A
        unit = s_unit[par][g];
    }
}

// ------------------------------------------------------------------------------------------
// k_cols: persistent, bulk-copy fed, warp-specialised
// ------------------------------------------------------------------------------------------

template <int N, int P, int C>
struct ColsCfg {
    using Line = LineCfg<N, P>;
    using IL = Inter<N, C, Line::GS>;
    static constexpr int T = Line::T;
    static constexpr int HC = C / 2;                  // packed height columns per strip
    static constexpr int NTP = C * T;                 // threads on the packed (dx, dz) columns
    static constexpr int NTH = HC * T;                // threads on the packed height columns
    static constexpr int NTHREADS = NTP + NTH + 32;   // + one producer warp
    static constexpr uint32_t P_BYTES = IL::P_STRIP * sizeof(float2);
    static constexpr uint32_t H_BYTES = IL::H_STRIP * sizeof(float2);
    static constexpr int LINE_H = HC == 2 ? Line::LINE_H2 : Line::LINE_H;   // line pitch of XH: conflict-free for HC lines per half-warp
    static constexpr uint32_t XH_BYTES = HC * LINE_H * sizeof(float2);
    // row of the height results HR (parked in a drained GP buffer); the shuffle-closed three-pass lines pad it
    __host__ __device__ static constexpr int hr_row(int m) { return (Line::R3 == 2 && OCEAN_SHFL_RADIX2) ? m + 4 * (m / 32) : m; }
    static_assert(sizeof(float) * (hr_row(N - 1) + 1) * C <= P_BYTES, "height results must fit in a drained GP buffer");
    static_assert(P_BYTES % 16 == 0 && H_BYTES % 16 == 0 && XH_BYTES % 16 == 0, "bulk copies need 16-byte granules");
    static constexpr uint32_t TW_BYTES = (N + Line::T) * sizeof(float2);   // pass-1 twiddles [R1][T] + pass-2 twiddles [R2][R3]
    static constexpr size_t SMEM = 2 * size_t(P_BYTES) + H_BYTES + XH_BYTES + TW_BYTES + 10 * sizeof(uint64_t);
};

// Work item = one strip of C columns of one tile. Per item:
//   producer warp : bulk-copies the strip of GH (1 copy) and of GP (1 copy) into shared memory
//   height warps  : build Z(y) = G(nA, y) + i G(nB, y) over the full column from the half-stored GH
//                   (G(n, N-y) = conj G(n, y)), transform, park the two real results per packed column in HR
//   packed warps  : transform the GP strip in place in shared memory, then write
//                   out(x, y) = (dx, height, dz, 0) * sign / 2  (correction.comp:29-34)
// GP is double-buffered so the next strip's copy overlaps this strip's transforms and stores; HR lives in
// the GP buffer the packed warps have just drained into registers, so the two roles only meet once per item.
// GENERAL = false is the product build (dense rows of N texels); GENERAL = true honours a per-tile row pitch and
// can accumulate a checksum of what it stores (tests: PDL on/off and GPU-count determinism).
template <int N, int P, int C, bool GENERAL>
__global__ void __launch_bounds__(ColsCfg<N, P, C>::NTHREADS, 1)
k_cols(const float2* __restrict__ gp_all, const float2* __restrict__ gh_all, const float2* __restrict__ tw_g,
       const OutDesc* __restrict__ out_tab, uint32_t first_tile, uint32_t n_items, unsigned long long* __restrict__ checksums,
       float* __restrict__ dx_plane)
{
    using CC = ColsCfg<N, P, C>;
    using Cfg = typename CC::Line;
    using IL = typename CC::IL;
    constexpr int T = Cfg::T, R1 = Cfg::R1, R2 = Cfg::R2, R3 = Cfg::R3, LINE = CC::LINE_H;
    constexpr int NTP = CC::NTP, NTH = CC::NTH, HC = CC::HC;
    constexpr int STRIPS = N / C;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* PB0 = reinterpret_cast<float2*>(smem_raw);
    float2* PB1 = reinterpret_cast<float2*>(smem_raw + CC::P_BYTES);
    float2* GB = reinterpret_cast<float2*>(smem_raw + 2 * CC::P_BYTES);
    float2* XH = reinterpret_cast<float2*>(smem_raw + 2 * CC::P_BYTES + CC::H_BYTES);
    float2* TW = reinterpret_cast<float2*>(smem_raw + 2 * CC::P_BYTES + CC::H_BYTES + CC::XH_BYTES);   // [R1][R2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + 2 * CC::P_BYTES + CC::H_BYTES + CC::XH_BYTES + CC::TW_BYTES);
    uint64_t* fullP = bars;        // [2] GP strip landed in PB[b]               (tx bytes)         producer -> packed
    uint64_t* drainedP = bars + 2; // [2] packed warps hold PB[b] in registers   (NTP arrivals)     packed -> height
    uint64_t* hrFree = bars + 4;   // [2] packed warps done with HR in PB[b]     (NTP arrivals)     packed -> producer
    uint64_t* fullG = bars + 6;    //     GH strip landed                        (tx bytes)         producer -> height
    uint64_t* emptyG = bars + 7;   //     height warps drained GB                (NTH arrivals)     height -> producer
    uint64_t* hrReady = bars + 8;  //     HR written into PB[b]                  (NTH arrivals)     height -> packed

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    pdl_launch_dependents();
    for (int i = tid; i < N + Cfg::T; i += CC::NTHREADS) TW[i] = tw_g[i];
    if (tid == 0) {
        ptx::mbar_init(fullP + 0, 1);
        ptx::mbar_init(fullP + 1, 1);
        ptx::mbar_init(drainedP + 0, NTP);
        ptx::mbar_init(drainedP + 1, NTP);
        ptx::mbar_init(hrFree + 0, NTP);
        ptx::mbar_init(hrFree + 1, NTP);
        ptx::mbar_init(fullG, 1);
        ptx::mbar_init(emptyG, NTH);
        ptx::mbar_init(hrReady, NTH);
        ptx::fence_mbar_init();
    }
    __syncthreads();

    if (tid >= NTP + NTH) {
        // ================= producer warp =================
        if (lane == 0) {
            pdl_wait_prior_grid();            // k_rows of this frame has completed and its stores are visible
            uint32_t it = 0;
            for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const uint32_t tl = item / STRIPS, strip = item % STRIPS;
                const float2* srcH = gh_all + size_t(tl) * IL::H_TILE + size_t(strip) * IL::H_STRIP;
                const float2* srcP = gp_all + size_t(tl) * IL::P_TILE + size_t(strip) * IL::P_STRIP;
                if (it >= 1) {
                    ptx::mbar_wait_backoff(emptyG, (it - 1) & 1);
                    // The height warps read GB through the generic proxy; the copy below writes it through the async
                    // proxy. The mbarrier orders generic accesses only: without this fence the copy can overtake
                    // reads that are still in flight (seen in round 2 as sporadically wrong height strips once a block
                    // processes four or more strips -- 1024^2 x 4+ tiles per launch; the intermediate was identical).
                    ptx::fence_proxy_async();
                }
                ptx::mbar_arrive_expect_tx(fullG, CC::H_BYTES);
                ptx::bulk_g2s(GB, srcH, CC::H_BYTES, fullG);
                const uint32_t b = it & 1;
                if (it >= 2) {
                    ptx::mbar_wait_backoff(hrFree + b, ((it >> 1) - 1) & 1);   // item it-2 is completely done with PB[b]
                    ptx::fence_proxy_async();
                }
                ptx::mbar_arrive_expect_tx(fullP + b, CC::P_BYTES);
                ptx::bulk_g2s(b ? PB1 : PB0, srcP, CC::P_BYTES, fullP + b);
            }
        }
    } else if (tid >= NTP) {
        // ================= height warps =================
        const int ht = tid - NTP;
        const int pc = ht % HC;               // packed column: real columns 2 pc and 2 pc + 1 of the strip
        const int k2 = ht / HC;
        float2* line = XH + pc * LINE;
        uint32_t it = 0;
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            ptx::mbar_wait(fullG, it & 1);
            float2 v[R1];
            const float4* G4 = reinterpret_cast<const float4*>(GB) + pc;   // (G(nA, .), G(nB, .)) adjacent: one 128-bit load
            constexpr int ROW4 = C / 2, GP4 = IL::GROUP_PITCH / 2;          // row / group pitch in float4
            if constexpr (R3 == 1) {
#pragma unroll
                for (int k1 = 0; k1 < R1; ++k1) {
                    // y = k1 R2 + k2; rows y < N/2 are stored, y > N/2 mirror N - y, rows 0 and N/2 share stored row 0
                    float4 g;
                    if (k1 < R1 / 2) {
                        g = G4[k1 * GP4 + k2 * ROW4];
                        v[k1] = (k1 == 0 && k2 == 0) ? make_float2(g.x, g.z) : make_float2(g.x - g.w, g.y + g.z);
                    } else {
                        // N - y = (R1 - k1 - 1) R2 + (R2 - k2) for k2 > 0, (R1 - k1) R2 for k2 = 0
                        const int grp = k2 ? R1 - k1 - 1 : (R1 - k1) % R1, row = k2 ? R2 - k2 : 0;
                        if (k1 == R1 / 2) {
                            g = G4[(k2 ? grp : 0) * GP4 + row * ROW4];
                            v[k1] = k2 ? make_float2(g.x + g.w, g.z - g.y) : make_float2(g.y, g.w);
                        } else {
                            g = G4[grp * GP4 + row * ROW4];
                            v[k1] = make_float2(g.x + g.w, g.z - g.y);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int k1 = 0; k1 < R1; ++k1) {
                    const uint32_t y = k1 * T + k2;                          // generic: any row grouping
                    const uint32_t yy = (y == N / 2) ? 0u : (y > N / 2 ? N - y : y);
                    const float4 g = G4[IL::row_off(yy) / 2];
                    if (y == 0) v[k1] = make_float2(g.x, g.z);
                    else if (y == N / 2) v[k1] = make_float2(g.y, g.w);
                    else if (y < N / 2) v[k1] = make_float2(g.x - g.w, g.y + g.z);
                    else v[k1] = make_float2(g.x + g.w, g.z - g.y);
                }
            }
            ptx::fence_proxy_async();          // this thread's generic reads of GB precede the next bulk copy into it
            ptx::mbar_arrive(emptyG);
            RegFft<R1, Cfg::PK>::run(v);
#pragma unroll
            for (int n1 = 1; n1 < R1; ++n1) v[n1] = cmul_tw(v[n1], TW[n1 * T + k2]);
#pragma unroll
            for (int n1 = 0; n1 < R1; ++n1) line[Cfg::pad(n1 * T + k2)] = v[n1];
            ptx::named_bar_sync<2, NTH>();
            const uint32_t b = it & 1;
            float* HR = reinterpret_cast<float*>(b ? PB1 : PB0);     // [N][C] height results
            if constexpr (R3 == 1) {
                float2 u[Cfg::SUB2][R2];
#pragma unroll
                for (int i = 0; i < Cfg::SUB2; ++i) {
                    const int n1 = k2 + R2 * i;
#pragma unroll
                    for (int k = 0; k < R2; ++k) u[i][k] = line[Cfg::pad(n1 * R2 + k)];
                }
                ptx::named_bar_sync<2, NTH>();                          // XH drained: the next item may overwrite it
#pragma unroll
                for (int i = 0; i < Cfg::SUB2; ++i) RegFft<R2, Cfg::PK>::run(u[i]);
                ptx::mbar_wait(drainedP + b, (it >> 1) & 1);             // the packed warps hold this item's PB in registers
#pragma unroll
                for (int i = 0; i < Cfg::SUB2; ++i) {
                    const int n1 = k2 + R2 * i;
#pragma unroll
                    for (int n2 = 0; n2 < R2; ++n2) {
                        const int m = n1 + R1 * n2;
                        *reinterpret_cast<float2*>(HR + m * C + 2 * pc) = make_float2(u[i][n2].x, u[i][n2].y);
                    }
                }
            } else {
                if constexpr (R3 == 2 && OCEAN_SHFL_RADIX2) {
                    // pass 2 in registers, closing radix-2 by shuffle between the threads k2 = 2g, 2g + 1 (HC lanes apart);
                    // the results wait in registers for the packed warps to drain PB
                    float2 r[R1];
                    int i = 0;
                    line_pass2_shfl<Cfg, HC>(k2, TW + N, [&](int p) { return line[Cfg::pad(p)]; },
                                             [&] { ptx::named_bar_sync<2, NTH>(); },     // XH drained: the next item may overwrite it
                                             [&](int, float2 val) { r[i++] = val; });
                    ptx::mbar_wait(drainedP + b, (it >> 1) & 1);
                    // rows m = g + 32 k3 + 64 q (+ N/2): HR rows get 4 pad rows per 32 (CC::hr_row) so that the two k3 halves
                    // of a half-warp fall into different banks
                    const int m0 = (k2 >> 1) + R1 * (k2 & 1);
#pragma unroll
                    for (int q = 0; q < R1 / 2; ++q) {
                        *reinterpret_cast<float2*>(HR + CC::hr_row(m0 + 2 * R1 * q) * C + 2 * pc) = r[2 * q];
                        *reinterpret_cast<float2*>(HR + CC::hr_row(m0 + 2 * R1 * q + R1 * R2) * C + 2 * pc) = r[2 * q + 1];
                    }
                } else {
                // pass 2 in XH, then wait for the packed warps before pass 3 streams its results into HR
                bool waited = false;
                line_passes_23<Cfg>(line, k2, TW + N, [&] {
                    ptx::named_bar_sync<2, NTH>();
                    ptx::mbar_wait(drainedP + b, (it >> 1) & 1);
                    waited = true;
                }, [&](int m, float2 val) { *reinterpret_cast<float2*>(HR + m * C + 2 * pc) = val; });
                (void)waited;
                ptx::named_bar_sync<2, NTH>();                          // XH drained: the next item may overwrite it
                }
            }
            ptx::fence_proxy_async();          // HR lives in PB[b], which the async proxy refills two items later
            ptx::mbar_arrive(hrReady);
        }
    } else {
        // ================= packed (dx, dz) warps =================
        const int c = tid % C;                // lanes run over columns first: a warp touches whole 64-byte rows
        const int k2 = tid / C;
        uint32_t it = 0;
        for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const uint32_t tl = item / STRIPS, n0 = (item % STRIPS) * C;
            const uint32_t b = it & 1;
            float2* PB = b ? PB1 : PB0;
            const OutDesc od = out_tab[first_tile + tl];          // in flight while the strip lands
            ptx::mbar_wait(fullP + b, (it >> 1) & 1);
            float2 v[R1];
            // position p = k1 T + k2 lives at row_off(p) + c; T is a multiple of the group size, so k1 strides whole groups
            float2* col = PB + IL::row_off(k2) + c;
            constexpr int K1_STRIDE = (T / Cfg::GS) * IL::GROUP_PITCH;
#pragma unroll
            for (int k1 = 0; k1 < R1; ++k1) v[k1] = col[k1 * K1_STRIDE];
            RegFft<R1, Cfg::PK>::run(v);
#pragma unroll
            for (int n1 = 0; n1 < R1; ++n1)
                col[n1 * K1_STRIDE] = n1 == 0 ? v[0] : cmul_tw(v[n1], TW[n1 * T + k2]);
            ptx::named_bar_sync<1, NTP>();
            float4* __restrict__ out = od.base + n0 + c;
            const size_t pitch = GENERAL ? size_t(od.pitch) : size_t(N);
            // optional dense copy of channel .x (what the normal map differentiates, ocean.frag:56-59): 4 B/pt here save
            // the consumer kernel 12 of the 16 B/pt it would otherwise fetch
            float* __restrict__ dxp = dx_plane ? dx_plane + size_t(first_tile + tl) * N * N + n0 + c : nullptr;
            unsigned long long csum = 0;
            auto emit = [&](uint32_t m, float dx, float hh, float dz) {
                // correction.comp:29 sign, times the 1/2 of the Hermitian fold
                const float sg = ((n0 + c + m) & 1u) ? 0.5f : -0.5f;
                const float4 t = make_float4(dx * sg, hh * sg, dz * sg, 0.0f);
                __stcs(out + size_t(m) * pitch, t);
                if (dxp) dxp[size_t(m) * N] = t.x;
                if constexpr (GENERAL)
                    csum += (unsigned long long)__float_as_uint(t.x) + __float_as_uint(t.y) * 3ull + __float_as_uint(t.z) * 5ull;
            };
            const float* HR = reinterpret_cast<const float*>(PB);
            if constexpr (R3 == 1) {
                float2 u[Cfg::SUB2][R2];
#pragma unroll
                for (int i = 0; i < Cfg::SUB2; ++i) {
                    const int n1 = k2 + R2 * i;
#pragma unroll
                    for (int k = 0; k < R2; ++k) u[i][k] = PB[n1 * IL::GROUP_PITCH + k * C + c];
                }
                ptx::mbar_arrive(drainedP + b);
#pragma unroll
                for (int i = 0; i < Cfg::SUB2; ++i) RegFft<R2, Cfg::PK>::run(u[i]);
                ptx::mbar_wait(hrReady, it & 1);
#pragma unroll
                for (int i = 0; i < Cfg::SUB2; ++i) {
                    const int n1 = k2 + R2 * i;
#pragma unroll
                    for (int n2 = 0; n2 < R2; ++n2) {
                        const uint32_t m = n1 + R1 * n2;
                        emit(m, u[i][n2].x, HR[m * C + c], u[i][n2].y);
                    }
                }
            } else {
                if constexpr (R3 == 2 && OCEAN_SHFL_RADIX2) {
                    // pass 2 in registers, closing radix-2 by shuffle between the threads k2 = 2g, 2g + 1 (C lanes apart)
                    float2 r[R1];
                    int i = 0;
                    line_pass2_shfl<Cfg, C>(k2, TW + N, [&](int p) { return PB[IL::row_off(p) + c]; },
                                            [&] { ptx::mbar_arrive(drainedP + b); },
                                            [&](int, float2 val) { r[i++] = val; });
                    ptx::mbar_wait(hrReady, it & 1);
                    const uint32_t m0 = (k2 >> 1) + R1 * (k2 & 1);
#pragma unroll
                    for (int q = 0; q < R1 / 2; ++q) {
                        const uint32_t ma = m0 + 2 * R1 * q, mb = ma + R1 * R2;
                        emit(ma, r[2 * q].x, HR[CC::hr_row(ma) * C + c], r[2 * q].y);
                        emit(mb, r[2 * q + 1].x, HR[CC::hr_row(mb) * C + c], r[2 * q + 1].y);
                    }
                } else {
                // pass 2 in place, pass 3 into registers, then the same exchange with the height warps
                const int g = k2 / R3, k3 = k2 % R3;
#pragma unroll
                for (int i = 0; i < Cfg::SUB2; ++i) {
                    const int n1 = g + R2 * i;
                    float2 u[R2];
#pragma unroll
                    for (int k = 0; k < R2; ++k) u[k] = PB[IL::row_off(n1 * T + k * R3 + k3) + c];
                    RegFft<R2, Cfg::PK>::run(u);
#pragma unroll
                    for (int n2 = 0; n2 < R2; ++n2)
                        PB[IL::row_off(n1 * T + n2 * R3 + k3) + c] = n2 == 0 ? u[0] : cmul_tw(u[n2], TW[N + n2 * R3 + k3]);
                }
                ptx::named_bar_sync<1, NTP>();
                float2 w[Cfg::SUB3][R3];
#pragma unroll
                for (int i = 0; i < Cfg::SUB3; ++i) {
                    const int q = k2 + T * i, n1 = q % R1, n2 = q / R1;
#pragma unroll
                    for (int k = 0; k < R3; ++k) w[i][k] = PB[IL::row_off(n1 * T + n2 * R3 + k) + c];
                    RegFft<R3, Cfg::PK>::run(w[i]);
                }
                ptx::mbar_arrive(drainedP + b);
                ptx::mbar_wait(hrReady, it & 1);
#pragma unroll
                for (int i = 0; i < Cfg::SUB3; ++i) {
                    const int q = k2 + T * i, n1 = q % R1, n2 = q / R1;
#pragma unroll
                    for (int n3 = 0; n3 < R3; ++n3) {
                        const uint32_t m = n1 + R1 * n2 + R1 * R2 * n3;
                        emit(m, w[i][n3].x, HR[m * C + c], w[i][n3].y);
                    }
                }
                }
            }
            ptx::fence_proxy_async();          // generic-proxy accesses to PB[b] precede the next bulk copy into it
            ptx::mbar_arrive(hrFree + b);
            if constexpr (GENERAL) {
                if (checksums) {
#pragma unroll
                    for (int o = 16; o; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
                    if (lane == 0) atomicAdd(checksums + tl, csum);
                }
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
    int num_sms = 0;
    int cols_blocks_per_sm = 1;  // persistent k_cols blocks resident per SM
    int rows_blocks_per_sm = 1;  // k_rows blocks resident per SM (occupancy)
    int pdl_mode = -1;           // programmatic dependent launch: -1 auto, 0 off, 1 on (env OCEAN_B200_PDL)
    // which row kernel (env OCEAN_B200_ROWS): 3 "tma" (default) = k_rows_t; 0 "staged" = k_rows; 1 "persistent" = k_rows_p
    // with its bulk-copy-fed ring; 2 "fold" = k_rows_p with one unit per block. Measured on B200 at N = 1024 x 8 tiles,
    // whole step: 99 / 107 / 113-131 / 118 us.
    int rows_mode = 3;
    float2* d_tw = nullptr;      // [R1][R2] inter-pass twiddles
    float* d_kx = nullptr;       // [N] wave numbers, propagate.comp:45-46,50-53
    float2* d_gp = nullptr;      // [tiles] strip-major packed (dx, dz) row-pass output
    float2* d_gh = nullptr;      // [tiles] strip-major height row-pass output (N/2 rows)
    float2* d_gp2 = nullptr;     // a second set (allocated on demand) for frames enqueued on the second lane, so that two
    float2* d_gh2 = nullptr;     // frames in flight on different streams never share an intermediate (ocean_update_overlapped)
    size_t gp_per_tile = 0, gh_per_tile = 0;   // float2 per tile
};

template <int N, int P, int PAIRS, int C, int MINB, int PPAIRS, int GROUPS, int SLOTS>
struct Launch {
    using Cfg = LineCfg<N, P>;
    using CC = ColsCfg<N, P, C>;
    using IL = typename CC::IL;
    using RC = RowsCfg<N, P, PPAIRS, C, GROUPS, SLOTS>;
    using RC1 = RowsCfg<N, P, PPAIRS, C, 1, 0>;            // one unit per block, hardware-scheduled
    static constexpr size_t smem_rows_t = size_t(2 * PAIRS) * RowSlot<N>::BYTES + 16 * PAIRS;   // + one mbarrier per row
    static constexpr size_t smem_rows = sizeof(float4) * 2 * PAIRS * (N + 1);

    static cudaError_t prepare(FusedPlan* p)
    {
        cudaError_t e = cudaFuncSetAttribute(k_rows<N, P, PAIRS, C, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_rows));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_rows_p<N, P, PPAIRS, C, GROUPS, SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(RC::SMEM));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_rows_p<N, P, PPAIRS, C, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(RC1::SMEM));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_rows_t<N, P, PAIRS, C, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_rows_t));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_cols<N, P, C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(CC::SMEM));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_cols<N, P, C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(CC::SMEM));
        if (e != cudaSuccess) return e;
        int per_sm = 1;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cols<N, P, C, false>, CC::NTHREADS, CC::SMEM);
        p->cols_blocks_per_sm = per_sm < 1 ? 1 : per_sm;
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rows<N, P, PAIRS, C, MINB>, 3 * PAIRS * Cfg::T, smem_rows);
        p->rows_blocks_per_sm = per_sm < 1 ? 1 : per_sm;
        return e;
    }
    static size_t gp_floats2_per_tile() { return IL::P_TILE; }
    static size_t gh_floats2_per_tile() { return IL::H_TILE; }

    static cudaError_t run(FusedPlan* p, const float2* h0, const float* omega, const OutDesc* out, float time,
                           uint32_t first_tile, uint32_t count, cudaStream_t s, cudaEvent_t* ev, bool general,
                           unsigned long long* checksums, float* dx_plane, int lane, cudaEvent_t cols_after)
    {
        if (ev) cudaEventRecord(ev[0], s);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        cudaLaunchConfig_t cfg{};
        cfg.stream = s;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const float2* tw = p->d_tw;
        const float* kx = p->d_kx;
        float2 *gp = lane ? p->d_gp2 : p->d_gp, *gh = lane ? p->d_gh2 : p->d_gh;
        cudaError_t e;
        static const bool debug_env = std::getenv("OCEAN_B200_DEBUG") != nullptr;
        const bool skip_rows = debug_env && std::getenv("OCEAN_B200_DEBUG_SKIP_ROWS") != nullptr;   // re-run k_cols on the same intermediate
        if (skip_rows) {
            e = cudaSuccess;
        } else if (p->rows_mode == 3) {
            // staged rows fed by bulk copies: every storing block waits on the prior grid itself, PDL is safe at any size
            attr[0].val.programmaticStreamSerializationAllowed = p->pdl_mode == 0 ? 0 : 1;
            cfg.gridDim = dim3(N / 2 / PAIRS, count);
            cfg.blockDim = dim3(3 * PAIRS * Cfg::T);
            cfg.dynamicSmemBytes = smem_rows_t;
            e = cudaLaunchKernelEx(&cfg, k_rows_t<N, P, PAIRS, C, MINB>, h0, omega, tw, kx, gp, gh, time, first_tile, 0u);
        } else if (p->rows_mode == 0) {
            // measured on B200 (N=1024): PDL gains 21% / 7% at 1 / 4 tiles per launch (it hides ramp and tail) and
            // loses 1.5-3.5% at 8-16 tiles, so it is used while the row grid is below four waves
            const bool pdl = p->pdl_mode == 1 || (p->pdl_mode < 0 && (N / 2 / PAIRS) * count < 4u * uint32_t(p->num_sms) * MINB);
            attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
            cfg.gridDim = dim3(N / 2 / PAIRS, count);
            cfg.blockDim = dim3(3 * PAIRS * Cfg::T);
            cfg.dynamicSmemBytes = smem_rows;
            const uint32_t resident = uint32_t(p->num_sms * p->rows_blocks_per_sm);
            e = cudaLaunchKernelEx(&cfg, k_rows<N, P, PAIRS, C, MINB>, h0, omega, tw, kx, gp, gh, time, first_tile, resident);
        } else if (p->rows_mode == 2) {
            // fold-at-source kernel, one unit per block (no tickets): the hardware block scheduler does the balancing
            attr[0].val.programmaticStreamSerializationAllowed = p->pdl_mode == 0 ? 0 : 1;
            const uint32_t units = count * uint32_t(RC1::UPT);
            cfg.gridDim = dim3(units);
            cfg.blockDim = dim3(RC1::NTHREADS);
            cfg.dynamicSmemBytes = RC1::SMEM;
            e = cudaLaunchKernelEx(&cfg, k_rows_p<N, P, PPAIRS, C, 1, 0>, h0, omega, tw, kx, gp, gh, time, first_tile, units);
        } else {
            // persistent rows: every storing thread waits on the prior grid itself, so PDL is safe at any size
            attr[0].val.programmaticStreamSerializationAllowed = p->pdl_mode == 0 ? 0 : 1;
            const uint32_t units = count * uint32_t(RC::UPT);
            cfg.gridDim = dim3(units < uint32_t(p->num_sms) ? units : uint32_t(p->num_sms));
            cfg.blockDim = dim3(RC::NTHREADS);
            cfg.dynamicSmemBytes = RC::SMEM;
            e = cudaLaunchKernelEx(&cfg, k_rows_p<N, P, PPAIRS, C, GROUPS, SLOTS>, h0, omega, tw, kx, gp, gh, time, first_tile, units);
        }
        if (e != cudaSuccess) return e;
        if (ev) cudaEventRecord(ev[1], s);
        // the column kernel writes the maps: order it behind the frame that wrote the same maps on the other lane. It is
        // then launched as an ordinary stream-ordered kernel: a programmatic launch is tied to the preceding KERNEL and
        // would start past the event wait that sits between the two.
        if (cols_after) {
            if ((e = cudaStreamWaitEvent(s, cols_after, 0)) != cudaSuccess) return e;
            attr[0].val.programmaticStreamSerializationAllowed = 0;
        }
        const uint32_t items = count * (N / C);
        const uint32_t slots = uint32_t(p->num_sms * p->cols_blocks_per_sm);
        cfg.gridDim = dim3(items < slots ? items : slots);
        cfg.blockDim = dim3(CC::NTHREADS);
        cfg.dynamicSmemBytes = CC::SMEM;
        const float2 *cgp = gp, *cgh = gh;
        e = general ? cudaLaunchKernelEx(&cfg, k_cols<N, P, C, true>, cgp, cgh, tw, out, first_tile, items, checksums, dx_plane)
                    : cudaLaunchKernelEx(&cfg, k_cols<N, P, C, false>, cgp, cgh, tw, out, first_tile, items, checksums, dx_plane);
        if (ev) cudaEventRecord(ev[2], s);
        return e;
    }
};

// Launch<N, P, legacy PAIRS, C, legacy MINB, persistent PAIRS, persistent GROUPS, raw-row ring SLOTS (0: plain loads)>
using L64 = Launch<64, 8, 4, 8, 4, 4, 4, 4>;       // small grids: 8 x 8 and 16 x 8 lines, several row pairs per unit
using L128 = Launch<128, 16, 4, 8, 4, 4, 4, 4>;
using L256 = Launch<256, 16, 2, 8, 4, 2, 5, 4>;
// N=512 (the reference's own size): one row pair per block (48 threads), 8 blocks/SM at 128 registers measured
// best of {1,2,4} pairs x {2..8} blocks (283 k frames/s batched, 122 k one tile per update)
#ifndef OCEAN_ROWS_PAIRS_512
#define OCEAN_ROWS_PAIRS_512 1
#endif
#ifndef OCEAN_ROWS_MINB_512
#define OCEAN_ROWS_MINB_512 8
#endif
#ifndef OCEAN_ROWS_GROUPS_512
#define OCEAN_ROWS_GROUPS_512 5
#endif
#ifndef OCEAN_ROWS_SLOTS_512
#define OCEAN_ROWS_SLOTS_512 2
#endif
using L512 = Launch<512, 32, OCEAN_ROWS_PAIRS_512, 8, OCEAN_ROWS_MINB_512, 2, OCEAN_ROWS_GROUPS_512, OCEAN_ROWS_SLOTS_512>;
#ifndef OCEAN_ROWS_GROUPS_2048
#define OCEAN_ROWS_GROUPS_2048 2
#endif
#ifndef OCEAN_ROWS_SLOTS_2048
#define OCEAN_ROWS_SLOTS_2048 1
#endif
#ifndef OCEAN_ROWS_MINB_2048
#define OCEAN_ROWS_MINB_2048 2
#endif
using L2048 = Launch<2048, 32, 1, 4, OCEAN_ROWS_MINB_2048, 1, OCEAN_ROWS_GROUPS_2048, OCEAN_ROWS_SLOTS_2048>;   // three-pass lines (32 x 32 x 2), strips of 4 columns
// k_rows at N=1024: 5 blocks/SM (128 registers, no spills) measured faster than 6 blocks/SM at 96 registers
// (68 vs 76 us per 8 tiles); overridable for A/B builds (scripts/ab_build.sh)
#ifndef OCEAN_ROWS_PAIRS_1024
#define OCEAN_ROWS_PAIRS_1024 1
#endif
#ifndef OCEAN_ROWS_MINB_1024
#define OCEAN_ROWS_MINB_1024 5
#endif
#ifndef OCEAN_STRIP_1024
#define OCEAN_STRIP_1024 8           // 4-column strips (two k_cols blocks per SM) measured 10 % slower
#endif
#ifndef OCEAN_ROWS_GROUPS_1024
#define OCEAN_ROWS_GROUPS_1024 5     // 480 threads at <= 128 registers
#endif
#ifndef OCEAN_ROWS_SLOTS_1024
#define OCEAN_ROWS_SLOTS_1024 2      // 2 x 40 KB of raw rows in flight beside the 5 x 25 KB of lines
#endif
using L1024 = Launch<1024, 32, OCEAN_ROWS_PAIRS_1024, OCEAN_STRIP_1024, OCEAN_ROWS_MINB_1024, 1, OCEAN_ROWS_GROUPS_1024, OCEAN_ROWS_SLOTS_1024>;

bool fused_supports(uint32_t n) { return n == 64 || n == 128 || n == 256 || n == 512 || n == 1024 || n == 2048; }

// pass-1 twiddles w_N^(n1 t) as [R1][T], then pass-2 twiddles w_T^(n2 k3) as [R2][R3]
template <class Cfg>
static std::vector<float2> make_twiddles()
{
    std::vector<float2> tw(Cfg::N + Cfg::T);
    for (int n1 = 0; n1 < Cfg::R1; ++n1)
        for (int t = 0; t < Cfg::T; ++t) {
            const double th = 2.0 * kPiD * double((n1 * t) % Cfg::N) / double(Cfg::N);
            tw[n1 * Cfg::T + t] = make_float2(float(std::cos(th)), float(std::sin(th)));
        }
    for (int n2 = 0; n2 < Cfg::R2; ++n2)
        for (int k3 = 0; k3 < Cfg::R3; ++k3) {
            const double th = 2.0 * kPiD * double((n2 * k3) % Cfg::T) / double(Cfg::T);
            tw[Cfg::N + n2 * Cfg::R3 + k3] = make_float2(float(std::cos(th)), float(std::sin(th)));
        }
    return tw;
}

cudaError_t fused_plan_create(FusedPlan** out, uint32_t n, uint32_t n_tiles, float domain_size, int device)
{
    *out = nullptr;
    if (!fused_supports(n)) return cudaErrorInvalidValue;
    FusedPlan* p = new (std::nothrow) FusedPlan;
    if (!p) return cudaErrorMemoryAllocation;
    p->n = n;
    p->n_tiles = n_tiles;
    p->domain_size = domain_size;
    if (const char* v = std::getenv("OCEAN_B200_PDL")) p->pdl_mode = v[0] == '1' ? 1 : (v[0] == '0' ? 0 : -1);
    if (const char* v = std::getenv("OCEAN_B200_ROWS")) p->rows_mode = v[0] == 'p' ? 1 : (v[0] == 'f' ? 2 : (v[0] == 's' ? 0 : 3));
    cudaError_t e;
    auto bail = [&](cudaError_t err) { fused_plan_destroy(p); return err; };
    if ((e = cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return bail(e);

    std::vector<float2> tw;
    switch (n) {
        case 64: tw = make_twiddles<L64::Cfg>(); break;
        case 128: tw = make_twiddles<L128::Cfg>(); break;
        case 256: tw = make_twiddles<L256::Cfg>(); break;
        case 512: tw = make_twiddles<L512::Cfg>(); break;
        case 1024: tw = make_twiddles<L1024::Cfg>(); break;
        default: tw = make_twiddles<L2048::Cfg>(); break;
    }
    if ((e = cudaMalloc(&p->d_tw, tw.size() * sizeof(float2))) != cudaSuccess) return bail(e);
    if ((e = cudaMemcpy(p->d_tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e);
    // the shader's fp32 arithmetic, op for op: k = pi * float(uint(2g - N - 1)) / domain_size
    std::vector<float> kx(n);
    for (uint32_t g = 0; g < n; ++g) {
        const uint32_t u = 2u * g - n - 1u;
        volatile float prod = kPi32 * float(u);
        kx[g] = prod / domain_size;
    }
    if ((e = cudaMalloc(&p->d_kx, n * sizeof(float))) != cudaSuccess) return bail(e);
    if ((e = cudaMemcpy(p->d_kx, kx.data(), n * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e);
    size_t gp_t, gh_t;
    switch (n) {
        case 64: gp_t = L64::gp_floats2_per_tile(); gh_t = L64::gh_floats2_per_tile(); e = L64::prepare(p); break;
        case 128: gp_t = L128::gp_floats2_per_tile(); gh_t = L128::gh_floats2_per_tile(); e = L128::prepare(p); break;
        case 256: gp_t = L256::gp_floats2_per_tile(); gh_t = L256::gh_floats2_per_tile(); e = L256::prepare(p); break;
        case 512: gp_t = L512::gp_floats2_per_tile(); gh_t = L512::gh_floats2_per_tile(); e = L512::prepare(p); break;
        case 1024: gp_t = L1024::gp_floats2_per_tile(); gh_t = L1024::gh_floats2_per_tile(); e = L1024::prepare(p); break;
        default: gp_t = L2048::gp_floats2_per_tile(); gh_t = L2048::gh_floats2_per_tile(); e = L2048::prepare(p); break;
    }
    if (e != cudaSuccess) return bail(e);
    p->gp_per_tile = gp_t;
    p->gh_per_tile = gh_t;
    // pad rows of the intermediate are never written by k_rows but are copied (and ignored) by k_cols: zero them once
    if ((e = cudaMalloc(&p->d_gp, size_t(n_tiles) * gp_t * sizeof(float2))) != cudaSuccess) return bail(e);
    if ((e = cudaMalloc(&p->d_gh, size_t(n_tiles) * gh_t * sizeof(float2))) != cudaSuccess) return bail(e);
    if ((e = cudaMemset(p->d_gp, 0, size_t(n_tiles) * gp_t * sizeof(float2))) != cudaSuccess) return bail(e);
    if ((e = cudaMemset(p->d_gh, 0, size_t(n_tiles) * gh_t * sizeof(float2))) != cudaSuccess) return bail(e);
    *out = p;
    return cudaSuccess;
}

// debug: the row-pass output of tile `tile` as raw float2 ranges (tests / hazard hunting only)
void fused_plan_intermediate(const FusedPlan* p, uint32_t tile, const float2** gp, size_t* gp_count, const float2** gh, size_t* gh_count)
{
    *gp = p->d_gp + size_t(tile) * p->gp_per_tile;
    *gh = p->d_gh + size_t(tile) * p->gh_per_tile;
    *gp_count = p->gp_per_tile;
    *gh_count = p->gh_per_tile;
}

void fused_plan_destroy(FusedPlan* p)
{
    if (!p) return;
    cudaFree(p->d_tw);
    cudaFree(p->d_kx);
    cudaFree(p->d_gp);
    cudaFree(p->d_gh);
    cudaFree(p->d_gp2);
    cudaFree(p->d_gh2);
    delete p;
}

cudaError_t fused_enqueue(FusedPlan* p, const float2* h0, const float* omega, const OutDesc* out, float time,
                          uint32_t first_tile, uint32_t count, cudaStream_t s, uint32_t* launches, cudaEvent_t* ev,
                          bool general, unsigned long long* checksums, float* dx_plane, int lane, cudaEvent_t cols_after)
{
    if (lane && !p->d_gp2) {
        // second intermediate set, same size as the first; pad rows are zeroed once like the first set's
        const size_t gpb = size_t(p->n_tiles) * p->gp_per_tile * sizeof(float2), ghb = size_t(p->n_tiles) * p->gh_per_tile * sizeof(float2);
        cudaError_t a = cudaMalloc(&p->d_gp2, gpb);
        if (a == cudaSuccess) a = cudaMalloc(&p->d_gh2, ghb);
        if (a == cudaSuccess) a = cudaMemset(p->d_gp2, 0, gpb);
        if (a == cudaSuccess) a = cudaMemset(p->d_gh2, 0, ghb);
        if (a != cudaSuccess) return a;
    }
    cudaError_t e;
    switch (p->n) {
        case 64: e = L64::run(p, h0, omega, out, time, first_tile, count, s, ev, general, checksums, dx_plane, lane, cols_after); break;
        case 128: e = L128::run(p, h0, omega, out, time, first_tile, count, s, ev, general, checksums, dx_plane, lane, cols_after); break;
        case 256: e = L256::run(p, h0, omega, out, time, first_tile, count, s, ev, general, checksums, dx_plane, lane, cols_after); break;
        case 512: e = L512::run(p, h0, omega, out, time, first_tile, count, s, ev, general, checksums, dx_plane, lane, cols_after); break;
        case 1024: e = L1024::run(p, h0, omega, out, time, first_tile, count, s, ev, general, checksums, dx_plane, lane, cols_after); break;
        case 2048: e = L2048::run(p, h0, omega, out, time, first_tile, count, s, ev, general, checksums, dx_plane, lane, cols_after); break;
        default: return cudaErrorInvalidValue;
    }
    if (launches) *launches = 2;
    return e;
}

}  // namespace ocean
