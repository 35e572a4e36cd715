// Host-callable launchers of the ocean kernels. Internal to libocean_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ocean {

// ---- literal pipeline: the reference's 4-kernel / 8-dispatch dataflow (kernels_literal.cu)
cudaError_t launch_propagate_literal(const float2* h0, const float* omega, float time, uint32_t n,
                                     float domain_size, float2* height_spec, float2* dx_spec,
                                     float2* dz_spec, cudaStream_t s);
cudaError_t launch_fft_row_literal(float2* data, uint32_t n, cudaStream_t s);
cudaError_t launch_fft_col_literal(float2* data, uint32_t n, cudaStream_t s);
cudaError_t launch_correction_literal(const float2* height, const float2* dx, const float2* dz,
                                      uint32_t n, float4* out, cudaStream_t s);
bool literal_supports(uint32_t n);
// shader/ocean.frag:50-66 at texel centres; disp/out: [tiles][N][N] float4
cudaError_t launch_normal_map(const float4* disp, float4* nrm, uint32_t n, uint32_t tiles, cudaStream_t s);

// ---- fused pipeline: k_rows + k_cols (kernels_fused.cu)
struct FusedPlan;
bool fused_supports(uint32_t n);
cudaError_t fused_plan_create(FusedPlan** out, uint32_t n, uint32_t n_tiles, float domain_size, int device);
void fused_plan_destroy(FusedPlan* p);
// Enqueue one frame for tiles [first_tile, first_tile + count); *launches = kernels launched.
// If `ev` is non-null it holds 3 events recorded before, between and after the two kernels.
cudaError_t fused_enqueue(FusedPlan* p, const float2* h0, const float* omega, float4* out, float time,
                          uint32_t first_tile, uint32_t count, cudaStream_t s, uint32_t* launches,
                          cudaEvent_t* ev = nullptr);

}  // namespace ocean
