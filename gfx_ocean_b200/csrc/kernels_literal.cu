// The reference's own dataflow, recompiled for sm_100a: one CUDA kernel per compute
// shader, launched as Renderer::render() dispatches them (src/render.rs:1122-1287):
//   propagate [N/16, N/16] x 16x16          shader/propagate.comp
//   fft_row   [1, N] x N/2  (x3 fields)      shader/fft_row.comp
//   fft_col   [1, N] x N/2  (x3 fields)      shader/fft_col.comp
//   correction [N/16, N/16] x 16x16          shader/correction.comp
// generalised from the hard-coded 512 to a runtime power-of-two N. It moves 172 B per grid
// point per frame in 8 launches and evaluates one sincos per butterfly; it exists as the
// on-device A/B baseline for the fused path and as a per-stage parity aid
// (ocean_debug_spectra), not as the product.
#include "kernels.h"
#include "ocean_common.cuh"

namespace ocean {

__global__ void __launch_bounds__(256)
k_propagate_literal(const float2* __restrict__ h0, const float* __restrict__ omega, float time,
                    uint32_t n, float domain_size, float2* __restrict__ height_spec,
                    float2* __restrict__ dx_spec, float2* __restrict__ dz_spec)
{
    const uint32_t gx = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t gy = blockIdx.y * blockDim.y + threadIdx.y;
    if (gx >= n || gy >= n) return;
    const uint32_t index = gx + n * gy;                                 // propagate.comp:43
    const uint32_t index_neg = (n - gy - 1u) * n + n - gx - 1u;         // :48
    const float kx = wave_number(gx, n, domain_size);                   // :45-46,50-53
    const float ky = wave_number(gy, n, domain_size);
    const float2 h = propagate_point(h0[index], h0[index_neg], omega[index], time);  // :55-62
    const float2 kn = unit_wave_vector(kx, ky);                         // :64-67
    height_spec[index] = h;                                             // :69
    dx_spec[index] = cmul(make_float2(0.f, -kn.x), h);                  // :70
    dz_spec[index] = cmul(make_float2(0.f, -kn.y), h);                  // :71
}

// fft_row.comp:25-40 with `shared_row[2][512]` -> dynamic shared [2][N].
__device__ __forceinline__ void butterfly_literal(float2* shared_row, uint32_t n, uint32_t index,
                                                  uint32_t block_size, uint32_t src, uint32_t dst)
{
    const uint32_t k = index & (block_size - 1u);
    const float2 in0 = shared_row[src * n + index];
    const float2 in1 = shared_row[src * n + index + (n >> 1)];
    const float theta = kPi32 * float(k) / float(block_size);   // "not 2 * pi as stated in the paper"
    const float2 c = make_float2(cosf(theta), sinf(theta));
    const float2 temp = cmul(in1, c);
    const uint32_t dest = (index << 1) - k;
    shared_row[dst * n + dest] = cadd(in0, temp);
    shared_row[dst * n + dest + block_size] = csub(in0, temp);
}

// kColumn = false: fft_row.comp:44-63 (element j of line l at j + N*l);
// kColumn = true : fft_col.comp:44-63 (element j of line l at l + N*j).
template <bool kColumn>
__global__ void k_fft_literal(float2* __restrict__ fft_data, uint32_t n, uint32_t stages)
{
    extern __shared__ float2 shared_row[];
    const uint32_t t = threadIdx.x, line = blockIdx.y, half = n >> 1;
    const uint32_t i0 = kColumn ? line + n * t : t + n * line;
    const uint32_t i1 = kColumn ? line + n * (t + half) : t + half + n * line;
    shared_row[t] = fft_data[i0];
    shared_row[t + half] = fft_data[i1];
    __syncthreads();
    for (uint32_t i = 0; i < stages; ++i) {
        butterfly_literal(shared_row, n, t, 1u << i, i % 2u, (i + 1u) % 2u);
        __syncthreads();
    }
    const uint32_t r = stages % 2u;
    fft_data[i0] = shared_row[r * n + t];
    fft_data[i1] = shared_row[r * n + t + half];
}

__global__ void __launch_bounds__(256)
k_correction_literal(const float2* __restrict__ height, const float2* __restrict__ disp_x,
                     const float2* __restrict__ disp_z, uint32_t n, float4* __restrict__ out, size_t out_pitch)
{
    const uint32_t gx = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t gy = blockIdx.y * blockDim.y + threadIdx.y;
    if (gx >= n || gy >= n) return;
    const uint32_t index = gx + n * gy;                                  // correction.comp:25
    const float sign_mul = ((gx + gy) % 2u == 0u) ? -1.0f : 1.0f;        // :29
    out[gx + out_pitch * gy] = make_float4(disp_x[index].x * sign_mul, height[index].x * sign_mul,
                                           disp_z[index].x * sign_mul, 0.0f);          // :31-34 (imageStore at (gx, gy))
}

static uint32_t log2u(uint32_t n) { uint32_t s = 0; while ((1u << s) < n) ++s; return s; }

bool literal_supports(uint32_t n) { return n >= 2 && n <= 2048 && (n & (n - 1)) == 0; }

cudaError_t launch_propagate_literal(const float2* h0, const float* omega, float time, uint32_t n,
                                     float domain_size, float2* hs, float2* dx, float2* dz, cudaStream_t s)
{
    const dim3 block(16, 16), grid((n + 15) / 16, (n + 15) / 16);       // WORKGROUP_SIZE, src/render.rs:42
    k_propagate_literal<<<grid, block, 0, s>>>(h0, omega, time, n, domain_size, hs, dx, dz);
    return cudaGetLastError();
}

template <bool kColumn>
static cudaError_t launch_fft(float2* data, uint32_t n, cudaStream_t s)
{
    if (!literal_supports(n)) return cudaErrorInvalidValue;
    const size_t smem = 2 * size_t(n) * sizeof(float2);
    k_fft_literal<kColumn><<<dim3(1, n), n / 2, smem, s>>>(data, n, log2u(n));   // src/render.rs:1177,1229
    return cudaGetLastError();
}

cudaError_t launch_fft_row_literal(float2* data, uint32_t n, cudaStream_t s) { return launch_fft<false>(data, n, s); }
cudaError_t launch_fft_col_literal(float2* data, uint32_t n, cudaStream_t s) { return launch_fft<true>(data, n, s); }

cudaError_t launch_correction_literal(const float2* h, const float2* dx, const float2* dz, uint32_t n,
                                      float4* out, size_t out_pitch, cudaStream_t s)
{
    const dim3 block(16, 16), grid((n + 15) / 16, (n + 15) / 16);
    k_correction_literal<<<grid, block, 0, s>>>(h, dx, dz, n, out, out_pitch);
    return cudaGetLastError();
}

}  // namespace ocean
