/*
 * ocean_b200.h -- C ABI of the B200-native ocean-surface synthesiser.
 *
 * Drop-in boundary for the per-frame compute path of gfx-rs/gfx-ocean
 *     propagate -> fft_row x3 -> fft_col x3 -> correction
 * which the reference records inline in Renderer::render()
 * (src/render.rs:1101-1310) through three operator holders:
 *     Propagation<B>  src/ocean.rs:15-177   (uniform PropagateLocals :8-13)
 *     Fft<B>          src/fft.rs:7-111
 *     Correction<B>   src/ocean.rs:184-328  (uniform CorrectionLocals :179-182)
 * The reference has no FFI for this path (its only extern "C" is the iOS
 * launcher, examples/ios/ios.rs:3-6); this header is what a Rust `extern "C"`
 * block would bind (bindings/rust/ocean.rs).
 *
 * Conventions
 *  - plain pointers and sizes only; every function returns an ocean_status
 *    (0 = OK, negative = error) and never throws or aborts across the boundary
 *    (the reference's `Result<_, Box<dyn Error>>` / unwrap() convention,
 *    src/fft.rs:19, src/render.rs:1078, mapped to codes + ocean_last_error()).
 *  - one context = one CUDA device + one stream; a context is not thread-safe
 *    (the reference drives everything from the winit thread, src/lib.rs:105-170).
 *  - grid point (x, y) lives at index x + N*y, x fastest (propagate.comp:43).
 *  - there is NO CPU fallback: without a CUDA device ocean_create fails with
 *    OCEAN_ERR_NO_DEVICE.
 */
#ifndef OCEAN_B200_H
#define OCEAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OCEAN_B200_ABI_VERSION 2

#if defined(__GNUC__)
#define OCEAN_API __attribute__((visibility("default")))
#else
#define OCEAN_API
#endif

typedef struct ocean_ctx ocean_ctx;

typedef enum ocean_status {
    OCEAN_OK = 0,
    OCEAN_ERR_INVALID_ARG = -1, /* null pointer, tile out of range, N not a supported power of two */
    OCEAN_ERR_NO_DEVICE = -2,   /* no CUDA device / device index out of range / not sm_100 */
    OCEAN_ERR_CUDA = -3,        /* a CUDA runtime call failed; see ocean_last_error */
    OCEAN_ERR_IO = -4,          /* bincode file missing, truncated or of the wrong size */
    OCEAN_ERR_NOT_READY = -5,   /* update/download before every tile has a spectrum */
    OCEAN_ERR_UNSUPPORTED = -6  /* option not available in this build */
} ocean_status;

/* Which kernels ocean_update enqueues. Both produce the same displacement map. */
typedef enum ocean_pipeline {
    /* The product path: two fused sm_100a kernels per frame (propagate + row
     * transforms; column transforms + sign correction + RGBA pack). */
    OCEAN_PIPELINE_FUSED = 0,
    /* The reference's own dataflow kept for A/B measurement: 4 kernels, 8
     * launches, radix-2 Stockham in shared memory, per-butterfly sincos
     * (src/render.rs:1122-1287 recompiled for sm_100a). */
    OCEAN_PIPELINE_LITERAL = 1
} ocean_pipeline;

/* Mirrors of the reference's uniform blocks (std430, 4-byte scalars). */
typedef struct ocean_propagate_locals { /* src/ocean.rs:8-13, propagate.comp:16-20 */
    float   time;
    int32_t resolution;
    float   domain_size;
} ocean_propagate_locals;

typedef struct ocean_correction_locals { /* src/ocean.rs:179-182, correction.comp:6-8 (never read there) */
    uint32_t resolution;
} ocean_correction_locals;

typedef struct ocean_config {
    uint32_t abi_version;   /* OCEAN_B200_ABI_VERSION */
    int32_t  cuda_device;   /* ordinal */
    uint32_t resolution;    /* N (reference: RESOLUTION = 512, src/render.rs:44). FUSED: 64, 128, 256, 512, 1024 or 2048;
                               LITERAL: any power of two in [8, 2048]; anything else fails with OCEAN_ERR_UNSUPPORTED
                               (a power of two in [8, 4096]) or OCEAN_ERR_INVALID_ARG */
    float    domain_size;   /* L (reference: DOMAIN_SIZE = 1000.0, src/render.rs:46); output-invariant */
    uint32_t n_tiles;       /* independent oceans held by this context, >= 1 */
    uint32_t pipeline;      /* ocean_pipeline */
    void*    stream;        /* cudaStream_t to enqueue on, or NULL to let the context create one. A caller-provided
                               stream must outlive the context: ocean_destroy drains it. */
    uint32_t flags;         /* OCEAN_FLAG_* */
} ocean_config;

#define OCEAN_FLAG_NONE          0u
#define OCEAN_FLAG_KEEP_SPECTRA  1u  /* LITERAL pipeline keeps post-propagate spectra for ocean_debug_spectra */
/* Two output buffers: ocean_update alternates between them, so frame n+1 is computed while ocean_download_all_async
 * reads frame n back on a separate copy stream (the reference keeps `frames_in_flight` = 3 command buffers over one
 * shared image, src/lib.rs:86). ocean_output_device then returns the buffer of the latest update. Such a context
 * updates all of its tiles together and takes no external outputs. */
#define OCEAN_FLAG_DOUBLE_BUFFER_OUTPUT 2u
/* The column kernel also writes a dense float copy of channel .x (dx) of every map it produces (+4 B per grid point),
 * which ocean_compute_normals then differentiates instead of gathering .x out of the RGBA texels: 20 instead of 32
 * bytes per grid point for the normal map. For contexts that call ocean_compute_normals every frame. */
#define OCEAN_FLAG_DX_PLANE 4u

/* Parameters of ocean_generate_spectrum. Defaults when NULL: amplitude 3e-8 (max|h0| ~ 1 at L = 1000, the scale of the
 * reference's data/spectrum.bin, |h0| <= 1.198), wind 30 m/s along +x, g 9.81, depth 100 m */
typedef struct ocean_spectrum_params {
    float amplitude;   /* Phillips constant A */
    float wind_speed;  /* V, m/s; l = V^2 / g */
    float gravity;
    float depth;       /* finite-depth dispersion omega = sqrt(g k tanh(k d)) */
} ocean_spectrum_params;

/* ---- lifetime (Propagation/Fft/Correction::init + Renderer::new buffers, src/render.rs:223-225,607-729) */
OCEAN_API int  ocean_create(ocean_ctx** out, int cuda_device, uint32_t resolution, float domain_size, uint32_t n_tiles);
OCEAN_API int  ocean_create_ex(ocean_ctx** out, const ocean_config* cfg);
OCEAN_API void ocean_destroy(ocean_ctx* ctx);            /* ::destroy(), src/fft.rs:102, src/ocean.rs:170,321 */

/* ---- inputs (staging upload of data/omega.bin + data/spectrum.bin, src/render.rs:742-931) */
/* h0_xy: N*N*2 floats (re, im interleaved); omega: N*N floats. Host pointers; copied before return. */
OCEAN_API int  ocean_set_spectrum(ocean_ctx* ctx, uint32_t tile, const float* h0_xy, const float* omega);
/* Same, device pointers on the context's device; copied on the context's stream. */
OCEAN_API int  ocean_set_spectrum_device(ocean_ctx* ctx, uint32_t tile, const float* d_h0_xy, const float* d_omega);
/* bincode Vec<f32> / Vec<[f32;2]> files as shipped by the reference (src/render.rs:769-771,808-810). */
OCEAN_API int  ocean_load_bincode(ocean_ctx* ctx, uint32_t tile, const char* omega_path, const char* spectrum_path);

/* Seeded generator of a tile's inputs ON THE DEVICE (the step before the path; the reference only ships its outputs
 * data/{omega,spectrum}.bin): omega = sqrt(g k tanh(k d)) and h0 = (xi_r + i xi_i) sqrt(P(k) / 2) with the Phillips spectrum
 * P(k) = A exp(-1/(k l)^2) / k^4 (khat . (1,0))^2 (x0.07 against the wind) on the half-sample grid
 * k = 2 pi (i - N/2 - 1/2) / L that data/omega.bin follows; xi from Philox-4x32-10 with counter (point index,
 * stream_id, 0, 0) and key = seed, Box-Muller on two 24-bit uniforms. h_words (optional, host, N*N*4 u32) receives
 * the raw Philox output per point. No host upload is involved. */
OCEAN_API int  ocean_generate_spectrum(ocean_ctx* ctx, uint32_t tile, uint64_t seed, uint32_t stream_id,
                                       const ocean_spectrum_params* params, uint32_t* h_words);
/* Copy a tile's inputs back to the host (N*N*2 and N*N floats). Waits for the stream. */
OCEAN_API int  ocean_get_spectrum(ocean_ctx* ctx, uint32_t tile, float* h0_xy, float* omega);

/* ---- the hot path (src/render.rs:1101-1310 steps 2-10) */
/* Enqueue one frame for every tile: writes PropagateLocals{time, resolution, domain_size}
 * and runs propagate -> 2-D inverse transform of (dx, height, dz) -> correction. Asynchronous. */
OCEAN_API int  ocean_update(ocean_ctx* ctx, float time);
/* Same for tiles [first_tile, first_tile + count). */
OCEAN_API int  ocean_update_tiles(ocean_ctx* ctx, float time, uint32_t first_tile, uint32_t count);
/* ocean_update_tiles through a CUDA graph: the frame's launches (with their programmatic dependency) are recorded
 * once per tile range and replayed with only `time` patched (FUSED pipeline, single-buffered contexts). */
OCEAN_API int  ocean_update_graph(ocean_ctx* ctx, float time, uint32_t first_tile, uint32_t count);
/* ocean_update_tiles with consecutive frames in flight together: calls alternate between two internal lanes (stream +
 * own row-pass intermediate), so the row kernel of one frame runs beside the column kernel of the previous one instead
 * of waiting for it. When the frame writes a map that a frame still in flight on the other lane wrote (its latest
 * or an older one), only this frame's column kernel is ordered behind that lane. Every other entry point (downloads, sync, plain updates, uploads, ...) first orders the
 * context's stream behind both lanes -- ocean_join does only that -- so results are those of the same calls made
 * through ocean_update_tiles, bit for bit; a caller that consumes the maps on its own stream calls ocean_join first.
 * FUSED pipeline, single-buffered contexts. */
OCEAN_API int  ocean_update_overlapped(ocean_ctx* ctx, float time, uint32_t first_tile, uint32_t count);
/* Enqueue, on the context's stream, a wait for the frames in flight on the lanes (no host blocking). */
OCEAN_API int  ocean_join(ocean_ctx* ctx);
/* Enqueue `n_frames` consecutive updates at times t0 + i*dt (host loop inside the library). */
OCEAN_API int  ocean_update_sequence(ocean_ctx* ctx, float t0, float dt, uint32_t n_frames);

/* Same as ocean_update_sequence, and h_sums[frame * n_tiles + tile] receives an order-independent 64-bit checksum
 * of every frame's displacement map, accumulated by the column kernel itself as it stores (no extra kernel between
 * consecutive frames, so the frames overlap exactly as in ocean_update_sequence). Waits for the stream. Two runs
 * produce equal sums iff every texel of every frame is bit-identical. */
OCEAN_API int  ocean_update_sequence_checksums(ocean_ctx* ctx, float t0, float dt, uint32_t n_frames, uint64_t* h_sums);
/* h_sums[tile] = the same checksum of the current displacement maps (separate reduction kernel). Waits. */
OCEAN_API int  ocean_output_checksums(ocean_ctx* ctx, uint64_t* h_sums);

/* ---- outputs (the RGBA32F displacement_map, src/render.rs:820-845; texel = (dx, height, dz, 0)) */
/* Renderer interop, CUDA half (src/render.rs:820-869 creates displacement_map + its views, :939 binds it): make the
 * kernels write tile's map straight into a caller-provided device allocation -- e.g. the linear VkImage / VkBuffer
 * memory the renderer exported and imported here with cudaImportExternalMemory -- with the given row pitch in bytes
 * (0 = dense, N*16). d_rgba must be 16-byte aligned device memory of the context's device that outlives its use;
 * NULL restores the context's own buffer. Takes effect for updates enqueued after the call. */
OCEAN_API int  ocean_set_output_device(ocean_ctx* ctx, uint32_t tile, float* d_rgba, size_t row_pitch_bytes);
/* Device pointer to tile's N*N*4 floats, row-major [y][x][4]; stable until ocean_destroy. */
OCEAN_API int  ocean_output_device(ocean_ctx* ctx, uint32_t tile, const float** d_rgba);
/* Copy tile's output to host memory (N*N*4 floats). Waits for the stream. */
OCEAN_API int  ocean_download(ocean_ctx* ctx, uint32_t tile, float* h_rgba);
/* Enqueue the copy only (h_rgba should be page-locked); pair with ocean_sync. */
OCEAN_API int  ocean_download_async(ocean_ctx* ctx, uint32_t tile, float* h_rgba);
/* Enqueue the read-back of EVERY tile into h_rgba_all (n_tiles*N*N*4 floats, page-locked). On a double-buffered
 * context the copy runs on the context's copy stream, behind the frame it reads and concurrently with later updates. */
OCEAN_API int  ocean_download_all_async(ocean_ctx* ctx, float* h_rgba_all);
/* Block until read-backs enqueued by ocean_download_all_async have landed: all of them (lag 0) or all but the
 * newest one (lag 1: frame n-1 is on the host while frame n is still in flight). */
OCEAN_API int  ocean_download_fence(ocean_ctx* ctx, uint32_t lag);
OCEAN_API int  ocean_sync(ocean_ctx* ctx);

/* ---- consumer step (what the renderer derives from the map) */
/* Enqueue the normal map of shader/ocean.frag:50-66 for tiles [first_tile, first_tile+count): central
 * differences of channel .x over wrapped neighbours (sampler Linear/Tile, src/render.rs:398), diff = 2/N,
 * height_scale = 180; texel = (N.x, N.y, N.z, 0). Reads the displacement map of the last ocean_update. */
OCEAN_API int  ocean_compute_normals(ocean_ctx* ctx, uint32_t first_tile, uint32_t count);
OCEAN_API int  ocean_normals_device(ocean_ctx* ctx, uint32_t tile, const float** d_nrm /* N*N*4 */);
OCEAN_API int  ocean_download_normals(ocean_ctx* ctx, uint32_t tile, float* h_nrm /* N*N*4 */);   /* waits */
/* The vertex shader's displaced grid, shader/ocean.vert:21-25,29: for the reference's grid x grid vertex patch
 * (a_Pos = (x, 0, z), a_Uv = (x, z) / (grid - 1), src/render.rs:498-506; HALF_RESOLUTION = 128, :45) sample the map
 * with the Linear / Tile sampler (src/render.rs:397-398) and return p_PosWorld = a_Pos + (d.x/3.5, d.y/3, d.z/3.5) +
 * (offset_x, 0, offset_z) -- the per-instance patch offsets of src/render.rs:540-551. pos_world: grid*grid*3 floats. */
OCEAN_API int  ocean_displace_grid(ocean_ctx* ctx, uint32_t tile, uint32_t grid, float offset_x, float offset_z, float* h_pos_world); /* waits */
OCEAN_API int  ocean_displace_grid_device(ocean_ctx* ctx, uint32_t tile, uint32_t grid, float offset_x, float offset_z, float* d_pos_world);

/* ---- measurement */
/* One update with CUDA events recorded on the context's stream around every kernel of the frame (FUSED: after an
 * unmeasured frame of the same kind, so that the intervals are kernel durations and not launch latency; the events
 * serialise the two kernels, which otherwise overlap under programmatic dependent launch). Waits, then writes each
 * kernel's duration in ms (FUSED: [row kernel, k_cols] over all tiles; LITERAL: the 8 dispatches of tile 0 in
 * reference order). *n_stages receives the count; capacity is stage_ms's length. */
OCEAN_API int  ocean_profile_update(ocean_ctx* ctx, float time, float* stage_ms, uint32_t capacity, uint32_t* n_stages);

/* ---- introspection */
/* Post-propagate spectra of a tile (N*N*2 floats each, host), recomputed at the time of the
 * last update with the standalone propagate kernel. Parity/debug only; waits for the stream. */
OCEAN_API int  ocean_debug_spectra(ocean_ctx* ctx, uint32_t tile, float* h, float* dx, float* dz);
/* The uniform blocks as last written by ocean_update (what the reference maps at src/render.rs:1101-1120). */
OCEAN_API int  ocean_get_locals(const ocean_ctx* ctx, ocean_propagate_locals* p, ocean_correction_locals* c);
OCEAN_API uint32_t ocean_resolution(const ocean_ctx* ctx);
OCEAN_API uint32_t ocean_n_tiles(const ocean_ctx* ctx);
/* Kernel launches enqueued by this context since creation (for gpu_launches accounting). */
OCEAN_API uint64_t ocean_launch_count(const ocean_ctx* ctx);
/* The stream the context enqueues on (cudaStream_t as void*). */
OCEAN_API void* ocean_stream(const ocean_ctx* ctx);
/* Algorithmic bytes one ocean_update moves: 76 * N*N * n_tiles. */
OCEAN_API uint64_t ocean_algorithmic_bytes_per_update(const ocean_ctx* ctx);
/* Message of the last error on this context ("" if none); ctx may be NULL for create-time errors. */
OCEAN_API const char* ocean_last_error(const ocean_ctx* ctx);
OCEAN_API const char* ocean_status_string(int status);
OCEAN_API uint32_t ocean_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* OCEAN_B200_H */
