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
// out_pitch: row pitch of the output image in float4 texels
cudaError_t launch_correction_literal(const float2* height, const float2* dx, const float2* dz,
                                      uint32_t n, float4* out, size_t out_pitch, cudaStream_t s);
bool literal_supports(uint32_t n);

// ---- either side of the path (kernels_consumer.cu); pitches and strides in float4 texels
// shader/ocean.frag:50-66 at texel centres; nrm: dense [tiles][N][N] float4
cudaError_t launch_normal_map(const float4* disp, size_t disp_pitch, size_t disp_tile_stride, float4* nrm, uint32_t n,
                              uint32_t tiles, cudaStream_t s);
// same from the dense copy of channel .x that k_cols can write (float[tiles][N][N])
cudaError_t launch_normal_map_plane(const float* dx_plane, float4* nrm, uint32_t n, uint32_t tiles, cudaStream_t s);
// shader/ocean.vert:21-25,29 for a grid x grid vertex patch; pos_world: [grid*grid][3]
cudaError_t launch_displace_grid(const float4* disp, size_t pitch, uint32_t n, uint32_t grid, float off_x, float off_z,
                                 float* pos_world, cudaStream_t s);
// *sum += order-independent 64-bit checksum of one N x N map
cudaError_t launch_checksum(const float4* disp, size_t pitch, uint32_t n, unsigned long long* sum, cudaStream_t s);
// Philox-4x32-10 seeded finite-depth dispersion + Phillips spectrum for one tile; words (optional): the raw 4 x u32 per point
cudaError_t launch_generate_spectrum(float2* h0, float* omega, uint32_t n, uint32_t tile_id, uint64_t seed, float domain_size,
                                     float amplitude, float wind_speed, float gravity, float depth, uint32_t* words, cudaStream_t s);

// ---- fused pipeline: k_rows + k_cols (kernels_fused.cu)
struct FusedPlan;
bool fused_supports(uint32_t n);
cudaError_t fused_plan_create(FusedPlan** out, uint32_t n, uint32_t n_tiles, float domain_size, int device);
void fused_plan_destroy(FusedPlan* p);
void fused_plan_intermediate(const FusedPlan* p, uint32_t tile, const float2** gp, size_t* gp_count, const float2** gh, size_t* gh_count);
// Where tile t's displacement map goes: the context's own buffer or a caller-provided (e.g. imported Vulkan)
// allocation. One record per tile, in device memory.
struct OutDesc {
    float4* base;
    uint32_t pitch;      // row pitch in float4 texels (N when dense)
    uint32_t pad_;
};
// Enqueue one frame for tiles [first_tile, first_tile + count); *launches = kernels launched.
// out_tab: device array of OutDesc indexed by tile; `general` selects the k_cols build that honours row pitches
// != N and (checksums != nullptr) adds every tile's output checksum into checksums[tile - first_tile];
// dx_plane (optional): dense float[tile][y][x] copy of channel .x for the normal-map kernel;
// lane 0 / 1: which of the plan's two intermediate sets the frame uses (frames in flight on different streams must differ);
// cols_after (optional): an event the column kernel (which writes the maps) is ordered behind.
// If `ev` is non-null it holds 3 events recorded before, between and after the two kernels.
cudaError_t fused_enqueue(FusedPlan* p, const float2* h0, const float* omega, const OutDesc* out_tab, float time,
                          uint32_t first_tile, uint32_t count, cudaStream_t s, uint32_t* launches,
                          cudaEvent_t* ev = nullptr, bool general = false, unsigned long long* checksums = nullptr,
                          float* dx_plane = nullptr, int lane = 0, cudaEvent_t cols_after = nullptr);

}  // namespace ocean
