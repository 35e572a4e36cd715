// Kernels either side of the hot path: the consumer step of the displacement map
// (normal map, vertex displacement), output checksums, and the seeded spectrum generator.
#include <cuda_runtime.h>

#include "kernels.h"
#include "ocean_common.cuh"

namespace ocean {

// ------------------------------------------------------------------------------------------
// normal map: shader/ocean.frag:50-66 evaluated once per texel centre
// ------------------------------------------------------------------------------------------
// A warp owns 32 columns and walks ROWS rows downwards with a three-row window of channel .x in registers: per
// texel one new load (the row below), the x-neighbours come from the adjacent lanes by shuffle (the two edge
// lanes fetch theirs), the z-neighbours from the window. Stores are whole 512-byte row segments per warp.
// Sampler: Linear filter + Tile wrap (src/render.rs:397-398), so a one-texel textureOffset at a texel centre is
// exactly the wrapped neighbour texel.
constexpr int kNormalRows = 16;

// ocean.frag:64-66: na = normalize(-diff, (x1-x0)/height_scale, 0), nb = normalize(0, (z1-z0)/height_scale, diff),
// N = normalize(cross(na, nb)). cross(na, nb) = (dxv diff, diff^2, -diff dzv) / (|na| |nb|) with dxv, dzv the two
// scaled differences, and the positive factor diff / (|na| |nb|) drops out of the final normalisation:
// N = (dxv, diff, -dzv) / sqrt(dxv^2 + diff^2 + dzv^2) -- one reciprocal square root instead of three square roots
// and nine divisions (the literal form made this kernel ALU bound: ~150 instructions per texel).
__device__ __forceinline__ float4 normal_from_differences(float x0, float x1, float z0, float z1, float diff)
{
    const float inv_height_scale = 1.0f / 180.0f;                       // ocean.frag:19
    const float dxv = (x1 - x0) * inv_height_scale, dzv = (z1 - z0) * inv_height_scale;
    const float l2 = fmaf(dxv, dxv, fmaf(dzv, dzv, diff * diff));
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l2));           // l2 >= diff^2 > 0
    r = r * fmaf(-0.5f * l2, r * r, 1.5f);                              // one Newton step: ~1e-7 relative
    return make_float4(dxv * r, diff * r, -dzv * r, 0.0f);
}

__global__ void __launch_bounds__(256)
k_normal_map(const float4* __restrict__ disp, size_t disp_pitch, size_t disp_tile_stride, float4* __restrict__ nrm, uint32_t n)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t x = (blockIdx.x * (blockDim.x >> 5) + warp) * 32 + lane;
    if (x - lane >= n) return;                                          // whole warp out of range (n >= 32 here)
    const uint32_t m = n - 1;                                           // N is a power of two
    const float4* __restrict__ d = disp + size_t(blockIdx.z) * disp_tile_stride;
    float4* __restrict__ o = nrm + size_t(blockIdx.z) * n * n;
    const uint32_t y0 = blockIdx.y * kNormalRows;
    const float diff = 2.0f / float(n);                                 // :52 (dim is hard-coded 512 there)
    const uint32_t xl = (x - 1) & m, xr = (x + 1) & m;
    float up = __ldg(&d[x + disp_pitch * ((y0 - 1) & m)].x);
    float cur = __ldg(&d[x + disp_pitch * y0].x);
#pragma unroll 4
    for (uint32_t r = 0; r < kNormalRows; ++r) {
        const uint32_t y = y0 + r;
        const float down = __ldg(&d[x + disp_pitch * ((y + 1) & m)].x);
        float x0 = __shfl_up_sync(0xffffffffu, cur, 1), x1 = __shfl_down_sync(0xffffffffu, cur, 1);
        if (lane == 0) x0 = __ldg(&d[xl + disp_pitch * y].x);           // ocean.frag:56-59, channel .x
        if (lane == 31) x1 = __ldg(&d[xr + disp_pitch * y].x);
        __stcs(&o[x + size_t(n) * y], normal_from_differences(x0, x1, up, down, diff));
        up = cur;
        cur = down;
    }
}

// The same walk over the dense channel-.x plane k_cols writes on request: 4 B per texel read instead of 16.
__global__ void __launch_bounds__(256)
k_normal_map_plane(const float* __restrict__ dxp, float4* __restrict__ nrm, uint32_t n)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t x = (blockIdx.x * (blockDim.x >> 5) + warp) * 32 + lane;
    if (x - lane >= n) return;
    const uint32_t m = n - 1;
    const float* __restrict__ d = dxp + size_t(blockIdx.z) * n * n;
    float4* __restrict__ o = nrm + size_t(blockIdx.z) * n * n;
    const uint32_t y0 = blockIdx.y * kNormalRows;
    const float diff = 2.0f / float(n);
    // Everything the warp needs is requested before anything is computed: the kNormalRows + 2 rows of its columns and,
    // in the two edge lanes, the neighbour column of every row -- 34 independent loads in flight per thread instead of
    // a rolling window that waits for one row at a time (the kernel was latency bound at 0.64 of the copy rate).
    const bool edge = lane == 0 || lane == 31;
    const uint32_t xe = lane == 0 ? (x - 1) & m : (x + 1) & m;
    float col[kNormalRows + 2], side[kNormalRows];
#pragma unroll
    for (int r = 0; r < kNormalRows + 2; ++r) col[r] = __ldg(&d[x + size_t(n) * ((y0 + r - 1) & m)]);
#pragma unroll
    for (int r = 0; r < kNormalRows; ++r) side[r] = edge ? __ldg(&d[xe + size_t(n) * (y0 + r)]) : 0.f;
#pragma unroll
    for (int r = 0; r < kNormalRows; ++r) {
        float x0 = __shfl_up_sync(0xffffffffu, col[r + 1], 1), x1 = __shfl_down_sync(0xffffffffu, col[r + 1], 1);
        if (lane == 0) x0 = side[r];
        if (lane == 31) x1 = side[r];
        __stcs(&o[x + size_t(n) * (y0 + r)], normal_from_differences(x0, x1, col[r], col[r + 2], diff));
    }
}

cudaError_t launch_normal_map_plane(const float* dx_plane, float4* nrm, uint32_t n, uint32_t tiles, cudaStream_t s)
{
    if (n < 32 || n % kNormalRows) return cudaErrorInvalidValue;
    k_normal_map_plane<<<dim3((n + 255) / 256, n / kNormalRows, tiles), n < 256 ? n : 256, 0, s>>>(dx_plane, nrm, n);
    return cudaGetLastError();
}

// small grids (n < 32): one thread per texel
__global__ void __launch_bounds__(256)
k_normal_map_small(const float4* __restrict__ disp, size_t disp_pitch, size_t disp_tile_stride, float4* __restrict__ nrm, uint32_t n)
{
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= n) return;
    const float4* __restrict__ d = disp + size_t(blockIdx.z) * disp_tile_stride;
    const uint32_t m = n - 1;
    const float x0 = __ldg(&d[((x - 1) & m) + disp_pitch * y].x), x1 = __ldg(&d[((x + 1) & m) + disp_pitch * y].x);
    const float z0 = __ldg(&d[x + disp_pitch * ((y - 1) & m)].x), z1 = __ldg(&d[x + disp_pitch * ((y + 1) & m)].x);
    nrm[size_t(blockIdx.z) * n * n + x + size_t(n) * y] = normal_from_differences(x0, x1, z0, z1, 2.0f / float(n));
}

cudaError_t launch_normal_map(const float4* disp, size_t disp_pitch, size_t disp_tile_stride, float4* nrm, uint32_t n,
                              uint32_t tiles, cudaStream_t s)
{
    if (n >= 32 && n % kNormalRows == 0)
        k_normal_map<<<dim3((n + 255) / 256, n / kNormalRows, tiles), n < 256 ? n : 256, 0, s>>>(disp, disp_pitch, disp_tile_stride, nrm, n);
    else
        k_normal_map_small<<<dim3((n + 255) / 256, n, tiles), 256, 0, s>>>(disp, disp_pitch, disp_tile_stride, nrm, n);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// vertex displacement: shader/ocean.vert:21-25,29 for the reference's vertex grid
// ------------------------------------------------------------------------------------------
// Vertex (x, z) of a grid x grid patch: a_Pos = (x, 0, z), a_Uv = (x, z) / (grid - 1) in f32 (src/render.rs:498-506);
// d = texture(displacement_map, a_Uv) with the Linear / Tile sampler (Vulkan texel filtering: u * W - 0.5, floor,
// fraction, wrap); p_PosWorld = a_Pos + (d.x / 3.5, d.y / 3, d.z / 3.5) + (offset.x, 0, offset.y).
__global__ void __launch_bounds__(256)
k_displace_grid(const float4* __restrict__ disp, size_t pitch, uint32_t n, uint32_t grid, float off_x, float off_z,
                float* __restrict__ pos_world)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= grid * grid) return;
    const uint32_t vx = i % grid, vz = i / grid;
    const float fx = float(vx), fz = float(vz), den = float(grid - 1);
    const float u = __fdiv_rn(fx, den), v = __fdiv_rn(fz, den);
    const float su = __fsub_rn(__fmul_rn(u, float(n)), 0.5f), sv = __fsub_rn(__fmul_rn(v, float(n)), 0.5f);
    const float fu = floorf(su), fv = floorf(sv);
    const float a = su - fu, b = sv - fv;
    const int i0 = int(fu), j0 = int(fv);
    const uint32_t m = n - 1;
    const uint32_t x0 = uint32_t(i0) & m, x1 = uint32_t(i0 + 1) & m, y0 = uint32_t(j0) & m, y1 = uint32_t(j0 + 1) & m;
    const float4 t00 = __ldg(&disp[x0 + pitch * y0]), t10 = __ldg(&disp[x1 + pitch * y0]);
    const float4 t01 = __ldg(&disp[x0 + pitch * y1]), t11 = __ldg(&disp[x1 + pitch * y1]);
    auto lerp2 = [&](float p00, float p10, float p01, float p11) {
        const float top = p00 * (1.0f - a) + p10 * a, bot = p01 * (1.0f - a) + p11 * a;
        return top * (1.0f - b) + bot * b;
    };
    const float dx = lerp2(t00.x, t10.x, t01.x, t11.x), dy = lerp2(t00.y, t10.y, t01.y, t11.y), dz = lerp2(t00.z, t10.z, t01.z, t11.z);
    pos_world[3 * size_t(i) + 0] = fx + dx / 3.5f + off_x;              // ocean.vert:23-25
    pos_world[3 * size_t(i) + 1] = dy / 3.0f;                           // :22
    pos_world[3 * size_t(i) + 2] = fz + dz / 3.5f + off_z;
}

cudaError_t launch_displace_grid(const float4* disp, size_t pitch, uint32_t n, uint32_t grid, float off_x, float off_z,
                                 float* pos_world, cudaStream_t s)
{
    const uint32_t total = grid * grid;
    k_displace_grid<<<(total + 255) / 256, 256, 0, s>>>(disp, pitch, n, grid, off_x, off_z, pos_world);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// checksum of a displacement map: sum of the texels' bit patterns, mod 2^64 (order independent, so
// bit-identical outputs <=> identical sums whatever the thread / GPU layout)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_checksum(const float4* __restrict__ disp, size_t pitch, uint32_t n, unsigned long long* __restrict__ sum)
{
    unsigned long long acc = 0;
    const size_t total = size_t(n) * n;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const float4 t = __ldg(&disp[(i % n) + pitch * (i / n)]);
        acc += (unsigned long long)__float_as_uint(t.x) + __float_as_uint(t.y) * 3ull + __float_as_uint(t.z) * 5ull +
               __float_as_uint(t.w) * 7ull;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(sum, acc);
}

cudaError_t launch_checksum(const float4* disp, size_t pitch, uint32_t n, unsigned long long* sum, cudaStream_t s)
{
    const uint32_t blocks = uint32_t((size_t(n) * n + 1023) / 1024);
    k_checksum<<<blocks < 592 ? blocks : 592, 256, 0, s>>>(disp, pitch, n, sum);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// spectrum generator (the step before the path; the reference only ships its outputs data/*.bin)
// ------------------------------------------------------------------------------------------
// Philox-4x32-10 (Salmon et al., SC'11), counter = (idx, tile, 0, 0), key = (seed lo, seed hi).
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// omega = sqrt(g k tanh(k d)); h0 = (xi_r + i xi_i) sqrt(P(k) / 2), Phillips P(k) = A exp(-1/(k l)^2) / k^4 (khat.w)^2,
// w = (1, 0), l = V^2 / g, x0.07 against the wind; k on the half-sample grid 2 pi (i - N/2 - 1/2) / L that
// data/omega.bin follows. xi: Box-Muller on two 24-bit uniforms from the Philox words.
__global__ void __launch_bounds__(256)
k_generate_spectrum(float2* __restrict__ h0, float* __restrict__ omega, uint32_t n, uint32_t tile_id, uint2 seed,
                    float domain_size, float amplitude, float wind_speed, float gravity, float depth, uint32_t* __restrict__ words)
{
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * n) return;
    const uint32_t x = idx % n, y = idx / n;
    const uint4 r = philox4x32_10(make_uint4(idx, tile_id, 0u, 0u), seed);
    if (words) reinterpret_cast<uint4*>(words)[idx] = r;
    const float two_pi = 6.283185307179586f;
    const float kx = two_pi * (float(x) - 0.5f * float(n) - 0.5f) / domain_size;
    const float ky = two_pi * (float(y) - 0.5f * float(n) - 0.5f) / domain_size;
    const float k2 = kx * kx + ky * ky;
    const float k = sqrtf(k2);
    omega[idx] = sqrtf(gravity * k * tanhf(k * depth));
    const float ell = wind_speed * wind_speed / gravity;
    const float c = kx / k;
    float p = amplitude * expf(-1.0f / (k2 * ell * ell)) / (k2 * k2) * c * c;
    if (c < 0.0f) p *= 0.07f;
    const float u1 = (float(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = (float(r.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float rad = sqrtf(-2.0f * logf(u1)) * sqrtf(0.5f * p);
    float s, cs;
    sincosf(two_pi * u2, &s, &cs);
    h0[idx] = make_float2(rad * cs, rad * s);
}

cudaError_t launch_generate_spectrum(float2* h0, float* omega, uint32_t n, uint32_t tile_id, uint64_t seed, float domain_size,
                                     float amplitude, float wind_speed, float gravity, float depth, uint32_t* words, cudaStream_t s)
{
    const uint32_t total = n * n;
    k_generate_spectrum<<<(total + 255) / 256, 256, 0, s>>>(h0, omega, n, tile_id, make_uint2(uint32_t(seed), uint32_t(seed >> 32)),
                                                           domain_size, amplitude, wind_speed, gravity, depth, words);
    return cudaGetLastError();
}

}  // namespace ocean
