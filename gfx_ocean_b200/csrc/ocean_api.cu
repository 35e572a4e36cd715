// C ABI of libocean_b200.so (include/ocean_b200.h): context, device memory, frame
// scheduling. Mirrors what Renderer::new / Renderer::render do for the compute half of
// the reference (src/render.rs:607-729 buffers, :742-931 upload, :1101-1310 per-frame
// record) with the operator holders of src/fft.rs and src/ocean.rs folded into one
// opaque context that owns every device buffer.
#include "../../include/ocean_b200.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "kernels.h"

struct ocean_ctx {
    int device = 0;
    uint32_t n = 0;
    float domain_size = 0.f;
    uint32_t n_tiles = 0;
    uint32_t pipeline = OCEAN_PIPELINE_FUSED;
    uint32_t flags = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;

    // initial_spec + omega (src/render.rs:608-670), one slab per tile
    float2* d_h0 = nullptr;
    float* d_omega = nullptr;
    // displacement_map stand-in (src/render.rs:820-845): linear float4[tile][y][x]
    float4* d_out = nullptr;
    // literal pipeline: dx_spec | dy_spec | dz_spec of the tile in flight (src/render.rs:608-646)
    float2* d_spec = nullptr;
    // fused pipeline: row-pass output, consumed by the column pass
    void* d_work = nullptr;
    size_t work_bytes = 0;
    ocean::FusedPlan* plan = nullptr;
    // normal map of the consumer step (lazy), float4[tile][y][x]
    float4* d_nrm = nullptr;
    // ocean_debug_spectra scratch (lazy)
    float2* d_dbg = nullptr;

    std::vector<uint8_t> loaded;
    ocean_propagate_locals plocals{0.f, 0, 0.f};
    ocean_correction_locals clocals{0};
    bool updated = false;
    uint64_t launches = 0;
    std::string err;
};

namespace {

thread_local std::string g_create_err;

size_t pts(const ocean_ctx* c) { return size_t(c->n) * c->n; }

int fail(ocean_ctx* c, int code, const std::string& msg)
{
    if (c) c->err = msg; else g_create_err = msg;
    return code;
}

int cuda_fail(ocean_ctx* c, cudaError_t e, const char* what)
{
    return fail(c, OCEAN_ERR_CUDA, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}

#define OCEAN_CUDA(ctx, call)                                          \
    do {                                                               \
        cudaError_t e_ = (call);                                       \
        if (e_ != cudaSuccess) return cuda_fail((ctx), e_, #call);     \
    } while (0)

bool is_pow2(uint32_t n) { return n && (n & (n - 1)) == 0; }

int check_tile(ocean_ctx* c, uint32_t tile)
{
    if (!c) return OCEAN_ERR_INVALID_ARG;
    if (tile >= c->n_tiles) return fail(c, OCEAN_ERR_INVALID_ARG, "tile index out of range");
    return OCEAN_OK;
}

// One frame of the reference's own dataflow for tile `t` (src/render.rs:1122-1287).
int enqueue_literal(ocean_ctx* c, uint32_t t, float time)
{
    const size_t np = pts(c);
    float2* dx = c->d_spec;            // desc_sets[0] = dx, [1] = dy (height), [2] = dz (src/render.rs:971-988)
    float2* dy = c->d_spec + np;
    float2* dz = c->d_spec + 2 * np;
    OCEAN_CUDA(c, ocean::launch_propagate_literal(c->d_h0 + t * np, c->d_omega + t * np, time, c->n,
                                                  c->domain_size, dy, dx, dz, c->stream));
    for (float2* f : {dx, dy, dz}) OCEAN_CUDA(c, ocean::launch_fft_row_literal(f, c->n, c->stream));
    for (float2* f : {dx, dy, dz}) OCEAN_CUDA(c, ocean::launch_fft_col_literal(f, c->n, c->stream));
    OCEAN_CUDA(c, ocean::launch_correction_literal(dy, dx, dz, c->n, c->d_out + t * np, c->stream));
    c->launches += 8;
    return OCEAN_OK;
}

}  // namespace

extern "C" {

uint32_t ocean_abi_version(void) { return OCEAN_B200_ABI_VERSION; }

const char* ocean_status_string(int s)
{
    switch (s) {
        case OCEAN_OK: return "OCEAN_OK";
        case OCEAN_ERR_INVALID_ARG: return "OCEAN_ERR_INVALID_ARG";
        case OCEAN_ERR_NO_DEVICE: return "OCEAN_ERR_NO_DEVICE";
        case OCEAN_ERR_CUDA: return "OCEAN_ERR_CUDA";
        case OCEAN_ERR_IO: return "OCEAN_ERR_IO";
        case OCEAN_ERR_NOT_READY: return "OCEAN_ERR_NOT_READY";
        case OCEAN_ERR_UNSUPPORTED: return "OCEAN_ERR_UNSUPPORTED";
        default: return "OCEAN_ERR_UNKNOWN";
    }
}

const char* ocean_last_error(const ocean_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int ocean_create(ocean_ctx** out, int cuda_device, uint32_t resolution, float domain_size, uint32_t n_tiles)
{
    ocean_config cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = OCEAN_B200_ABI_VERSION;
    cfg.cuda_device = cuda_device;
    cfg.resolution = resolution;
    cfg.domain_size = domain_size;
    cfg.n_tiles = n_tiles;
    cfg.pipeline = OCEAN_PIPELINE_FUSED;
    return ocean_create_ex(out, &cfg);
}

int ocean_create_ex(ocean_ctx** out, const ocean_config* cfg)
{
    g_create_err.clear();
    if (!out || !cfg) return fail(nullptr, OCEAN_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (cfg->abi_version != OCEAN_B200_ABI_VERSION)
        return fail(nullptr, OCEAN_ERR_INVALID_ARG, "ocean_config.abi_version mismatch");
    if (!is_pow2(cfg->resolution) || cfg->resolution < 8 || cfg->resolution > 4096)
        return fail(nullptr, OCEAN_ERR_INVALID_ARG, "resolution must be a power of two in [8, 4096]");
    if (cfg->n_tiles < 1) return fail(nullptr, OCEAN_ERR_INVALID_ARG, "n_tiles must be >= 1");
    if (!(cfg->domain_size > 0.f)) return fail(nullptr, OCEAN_ERR_INVALID_ARG, "domain_size must be > 0");
    if (cfg->pipeline != OCEAN_PIPELINE_FUSED && cfg->pipeline != OCEAN_PIPELINE_LITERAL)
        return fail(nullptr, OCEAN_ERR_INVALID_ARG, "unknown pipeline");
    if (cfg->pipeline == OCEAN_PIPELINE_LITERAL && !ocean::literal_supports(cfg->resolution))
        return fail(nullptr, OCEAN_ERR_UNSUPPORTED, "literal pipeline supports N <= 2048");
    if (cfg->pipeline == OCEAN_PIPELINE_FUSED && !ocean::fused_supports(cfg->resolution))
        return fail(nullptr, OCEAN_ERR_UNSUPPORTED, "fused pipeline does not support this resolution");

    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        (void)cudaGetLastError();
        return fail(nullptr, OCEAN_ERR_NO_DEVICE,
                    std::string("no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
    }
    if (cfg->cuda_device < 0 || cfg->cuda_device >= count)
        return fail(nullptr, OCEAN_ERR_NO_DEVICE, "cuda_device out of range");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, cfg->cuda_device)) != cudaSuccess)
        return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major != 10)
        return fail(nullptr, OCEAN_ERR_NO_DEVICE,
                    std::string("device '") + prop.name + "' is not sm_100 (kernels are built for sm_100a only)");

    ocean_ctx* c = new (std::nothrow) ocean_ctx;
    if (!c) return fail(nullptr, OCEAN_ERR_CUDA, "out of host memory");
    c->device = cfg->cuda_device;
    c->n = cfg->resolution;
    c->domain_size = cfg->domain_size;
    c->n_tiles = cfg->n_tiles;
    c->pipeline = cfg->pipeline;
    c->flags = cfg->flags;
    c->loaded.assign(c->n_tiles, 0);
    c->plocals = {0.f, int32_t(c->n), c->domain_size};
    c->clocals = {c->n};

    auto bail = [&](cudaError_t err, const char* what) {
        int rc = cuda_fail(nullptr, err, what);
        ocean_destroy(c);
        return rc;
    };
    if ((e = cudaSetDevice(c->device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    if (cfg->stream) {
        c->stream = static_cast<cudaStream_t>(cfg->stream);
    } else {
        if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess)
            return bail(e, "cudaStreamCreate");
        c->own_stream = true;
    }
    const size_t np = pts(c), nt = c->n_tiles;
    if ((e = cudaMalloc(&c->d_h0, nt * np * sizeof(float2))) != cudaSuccess) return bail(e, "cudaMalloc(h0)");
    if ((e = cudaMalloc(&c->d_omega, nt * np * sizeof(float))) != cudaSuccess) return bail(e, "cudaMalloc(omega)");
    if ((e = cudaMalloc(&c->d_out, nt * np * sizeof(float4))) != cudaSuccess) return bail(e, "cudaMalloc(out)");
    if (c->pipeline == OCEAN_PIPELINE_LITERAL) {
        if ((e = cudaMalloc(&c->d_spec, 3 * np * sizeof(float2))) != cudaSuccess) return bail(e, "cudaMalloc(spec)");
    } else {
        if ((e = ocean::fused_plan_create(&c->plan, c->n, c->n_tiles, c->domain_size, c->device)) != cudaSuccess)
            return bail(e, "fused_plan_create");
    }
    *out = c;
    return OCEAN_OK;
}

void ocean_destroy(ocean_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->plan) ocean::fused_plan_destroy(c->plan);
    cudaFree(c->d_h0);
    cudaFree(c->d_omega);
    cudaFree(c->d_out);
    cudaFree(c->d_spec);
    cudaFree(c->d_work);
    cudaFree(c->d_dbg);
    cudaFree(c->d_nrm);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    (void)cudaGetLastError();
    delete c;
}

int ocean_set_spectrum(ocean_ctx* c, uint32_t tile, const float* h0_xy, const float* omega)
{
    if (int rc = check_tile(c, tile)) return rc;
    if (!h0_xy || !omega) return fail(c, OCEAN_ERR_INVALID_ARG, "null spectrum pointer");
    OCEAN_CUDA(c, cudaSetDevice(c->device));
    const size_t np = pts(c);
    OCEAN_CUDA(c, cudaMemcpyAsync(c->d_h0 + tile * np, h0_xy, np * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
    OCEAN_CUDA(c, cudaMemcpyAsync(c->d_omega + tile * np, omega, np * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    OCEAN_CUDA(c, cudaStreamSynchronize(c->stream));   // caller may free its arrays on return
    c->loaded[tile] = 1;
    return OCEAN_OK;
}

int ocean_set_spectrum_device(ocean_ctx* c, uint32_t tile, const float* d_h0_xy, const float* d_omega)
{
    if (int rc = check_tile(c, tile)) return rc;
    if (!d_h0_xy || !d_omega) return fail(c, OCEAN_ERR_INVALID_ARG, "null spectrum pointer");
    OCEAN_CUDA(c, cudaSetDevice(c->device));
    const size_t np = pts(c);
    OCEAN_CUDA(c, cudaMemcpyAsync(c->d_h0 + tile * np, d_h0_xy, np * sizeof(float2), cudaMemcpyDeviceToDevice, c->stream));
    OCEAN_CUDA(c, cudaMemcpyAsync(c->d_omega + tile * np, d_omega, np * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    c->loaded[tile] = 1;
    return OCEAN_OK;
}

// bincode 1.3.1 Vec<T>: u64-LE element count, then the raw little-endian payload
// (decoded by the reference at src/render.rs:769-771 and :808-810).
static int read_bincode(ocean_ctx* c, const char* path, size_t elem_bytes, size_t want_elems, std::vector<float>& dst)
{
    FILE* f = std::fopen(path, "rb");
    if (!f) return fail(c, OCEAN_ERR_IO, std::string("cannot open ") + path);
    uint64_t count = 0;
    if (std::fread(&count, 8, 1, f) != 1) { std::fclose(f); return fail(c, OCEAN_ERR_IO, std::string(path) + ": no length prefix"); }
    if (count != want_elems) {
        std::fclose(f);
        return fail(c, OCEAN_ERR_IO, std::string(path) + ": holds " + std::to_string(count) + " elements, context needs " + std::to_string(want_elems));
    }
    dst.resize(want_elems * elem_bytes / sizeof(float));
    const size_t got = std::fread(dst.data(), elem_bytes, want_elems, f);
    const bool trailing = std::fgetc(f) != EOF;
    std::fclose(f);
    if (got != want_elems) return fail(c, OCEAN_ERR_IO, std::string(path) + ": truncated payload");
    if (trailing) return fail(c, OCEAN_ERR_IO, std::string(path) + ": trailing bytes after payload");
    return OCEAN_OK;
}

int ocean_load_bincode(ocean_ctx* c, uint32_t tile, const char* omega_path, const char* spectrum_path)
{
    if (int rc = check_tile(c, tile)) return rc;
    if (!omega_path || !spectrum_path) return fail(c, OCEAN_ERR_INVALID_ARG, "null path");
    std::vector<float> om, sp;
    if (int rc = read_bincode(c, omega_path, sizeof(float), pts(c), om)) return rc;
    if (int rc = read_bincode(c, spectrum_path, 2 * sizeof(float), pts(c), sp)) return rc;
    return ocean_set_spectrum(c, tile, sp.data(), om.data());
}

int ocean_update_tiles(ocean_ctx* c, float time, uint32_t first_tile, uint32_t count)
{
    if (!c) return OCEAN_ERR_INVALID_ARG;
    if (count == 0 || first_tile >= c->n_tiles || count > c->n_tiles - first_tile)
        return fail(c, OCEAN_ERR_INVALID_ARG, "tile range out of bounds");
    for (uint32_t t = first_tile; t < first_tile + count; ++t)
        if (!c->loaded[t]) return fail(c, OCEAN_ERR_NOT_READY, "tile " + std::to_string(t) + " has no spectrum");
    OCEAN_CUDA(c, cudaSetDevice(c->device));
    c->plocals = {time, int32_t(c->n), c->domain_size};      // src/render.rs:1101-1120
    c->clocals = {c->n};
    if (c->pipeline == OCEAN_PIPELINE_LITERAL) {
        for (uint32_t t = first_tile; t < first_tile + count; ++t)
            if (int rc = enqueue_literal(c, t, time)) return rc;
    } else {
        uint32_t nl = 0;
        OCEAN_CUDA(c, ocean::fused_enqueue(c->plan, c->d_h0, c->d_omega, c->d_out, time, first_tile, count, c->stream, &nl));
        c->launches += nl;
    }
    c->updated = true;
    return OCEAN_OK;
}

int ocean_update(ocean_ctx* c, float time)
{
    if (!c) return OCEAN_ERR_INVALID_ARG;
    return ocean_update_tiles(c, time, 0, c->n_tiles);
}

int ocean_update_sequence(ocean_ctx* c, float t0, float dt, uint32_t n_frames)
{
    if (!c) return OCEAN_ERR_INVALID_ARG;
    for (uint32_t i = 0; i < n_frames; ++i)
        if (int rc = ocean_update_tiles(c, t0 + dt * float(i), 0, c->n_tiles)) return rc;
    return OCEAN_OK;
}

int ocean_compute_normals(ocean_ctx* c, uint32_t first_tile, uint32_t count)
{
    if (!c) return OCEAN_ERR_INVALID_ARG;
    if (count == 0 || first_tile >= c->n_tiles || count > c->n_tiles - first_tile)
        return fail(c, OCEAN_ERR_INVALID_ARG, "tile range out of bounds");
    if (!c->updated) return fail(c, OCEAN_ERR_NOT_READY, "ocean_update has not been called");
    OCEAN_CUDA(c, cudaSetDevice(c->device));
    if (!c->d_nrm) OCEAN_CUDA(c, cudaMalloc(&c->d_nrm, size_t(c->n_tiles) * pts(c) * sizeof(float4)));
    OCEAN_CUDA(c, ocean::launch_normal_map(c->d_out + first_tile * pts(c), c->d_nrm + first_tile * pts(c), c->n, count, c->stream));
    c->launches += 1;
    return OCEAN_OK;
}

int ocean_normals_device(ocean_ctx* c, uint32_t tile, const float** d_nrm)
{
    if (int rc = check_tile(c, tile)) return rc;
    if (!d_nrm) return fail(c, OCEAN_ERR_INVALID_ARG, "null output pointer");
    if (!c->d_nrm) return fail(c, OCEAN_ERR_NOT_READY, "ocean_compute_normals has not been called");
    *d_nrm = reinterpret_cast<const float*>(c->d_nrm + tile * pts(c));
    return OCEAN_OK;
}

int ocean_download_normals(ocean_ctx* c, uint32_t tile, float* h_nrm)
{
    if (int rc = check_tile(c, tile)) return rc;
    if (!h_nrm) return fail(c, OCEAN_ERR_INVALID_ARG, "null output pointer");
    if (!c->d_nrm) return fail(c, OCEAN_ERR_NOT_READY, "ocean_compute_normals has not been called");
    OCEAN_CUDA(c, cudaSetDevice(c->device));
    OCEAN_CUDA(c, cudaMemcpyAsync(h_nrm, c->d_nrm + tile * pts(c), pts(c) * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    OCEAN_CUDA(c, cudaStreamSynchronize(c->stream));
    return OCEAN_OK;
}

int ocean_profile_update(ocean_ctx* c, float time, float* stage_ms, uint32_t capacity, uint32_t* n_stages)
{
    if (!c) return OCEAN_ERR_INVALID_ARG;
    if (!stage_ms || !n_stages) return fail(c, OCEAN_ERR_INVALID_ARG, "null output pointer");
    for (uint32_t t = 0; t < c->n_tiles; ++t)
        if (!c->loaded[t]) return fail(c, OCEAN_ERR_NOT_READY, "tile " + std::to_string(t) + " has no spectrum");
    OCEAN_CUDA(c, cudaSetDevice(c->device));
    c->plocals = {time, int32_t(c->n), c->domain_size};
    const uint32_t stages = c->pipeline == OCEAN_PIPELINE_FUSED ? 2u : 8u;
    if (capacity < stages) return fail(c, OCEAN_ERR_INVALID_ARG, "stage_ms too small");
    std::vector<cudaEvent_t> ev(stages + 1);
    for (auto& e : ev) OCEAN_CUDA(c, cudaEventCreate(&e));
    int rc = OCEAN_OK;
    if (c->pipeline == OCEAN_PIPELINE_FUSED) {
        uint32_t nl = 0;
        cudaError_t e = ocean::fused_enqueue(c->plan, c->d_h0, c->d_omega, c->d_out, time, 0, c->n_tiles, c->stream, &nl, ev.data());
        if (e != cudaSuccess) rc = cuda_fail(c, e, "fused_enqueue");
        c->launches += nl;
    } else {
        // tile 0 only, stage by stage in the reference's dispatch order (src/render.rs:1122-1287)
        const size_t np = pts(c);
        float2 *dx = c->d_spec, *dy = c->d_spec + np, *dz = c->d_spec + 2 * np;
        int k = 0;
        cudaEventRecord(ev[k++], c->stream);
        ocean::launch_propagate_literal(c->d_h0, c->d_omega, time, c->n, c->domain_size, dy, dx, dz, c->stream);
        cudaEventRecord(ev[k++], c->stream);
        for (float2* f : {dx, dy, dz}) { ocean::launch_fft_row_literal(f, c->n, c->stream); cudaEventRecord(ev[k++], c->stream); }
        for (float2* f : {dx, dy, dz}) { ocean::launch_fft_col_literal(f, c->n, c->stream); cudaEventRecord(ev[k++], c->stream); }
        ocean::launch_correction_literal(dy, dx, dz, c->n, c->d_out, c->stream);
        cudaEventRecord(ev[k++], c->stream);
        c->launches += 8;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = cuda_fail(c, e, "literal stages");
    }
    if (rc == OCEAN_OK) {
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = cuda_fail(c, e, "cudaStreamSynchronize");
    }
    if (rc == OCEAN_OK) {
        for (uint32_t i = 0; i < stages; ++i) cudaEventElapsedTime(&stage_ms[i], ev[i], ev[i + 1]);
        *n_stages = stages;
        c->updated = true;
    }
    for (auto& e : ev) cudaEventDestroy(e);
    return rc;
}

int ocean_output_device(ocean_ctx* c, uint32_t tile, const float** d_rgba)
{
    if (int rc = check_tile(c, tile)) return rc;
    if (!d_rgba) return fail(c, OCEAN_ERR_INVALID_ARG, "null output pointer");
    *d_rgba = reinterpret_cast<const float*>(c->d_out + tile * pts(c));
    return OCEAN_OK;
}

int ocean_download_async(ocean_ctx* c, uint32_t tile, float* h_rgba)
{
    if (int rc = check_tile(c, tile)) return rc;
    if (!h_rgba) return fail(c, OCEAN_ERR_INVALID_ARG, "null output pointer");
    if (!c->updated) return fail(c, OCEAN_ERR_NOT_READY, "ocean_update has not been called");
    OCEAN_CUDA(c, cudaSetDevice(c->device));
    OCEAN_CUDA(c, cudaMemcpyAsync(h_rgba, c->d_out + tile * pts(c), pts(c) * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    return OCEAN_OK;
}

int ocean_download(ocean_ctx* c, uint32_t tile, float* h_rgba)
{
    if (int rc = ocean_download_async(c, tile, h_rgba)) return rc;
    OCEAN_CUDA(c, cudaStreamSynchronize(c->stream));
    return OCEAN_OK;
}

int ocean_sync(ocean_ctx* c)
{
    if (!c) return OCEAN_ERR_INVALID_ARG;
    OCEAN_CUDA(c, cudaSetDevice(c->device));
    OCEAN_CUDA(c, cudaStreamSynchronize(c->stream));
    return OCEAN_OK;
}

int ocean_debug_spectra(ocean_ctx* c, uint32_t tile, float* h, float* dx, float* dz)
{
    if (int rc = check_tile(c, tile)) return rc;
    if (!h || !dx || !dz) return fail(c, OCEAN_ERR_INVALID_ARG, "null output pointer");
    if (!c->loaded[tile]) return fail(c, OCEAN_ERR_NOT_READY, "tile has no spectrum");
    OCEAN_CUDA(c, cudaSetDevice(c->device));
    const size_t np = pts(c);
    if (!c->d_dbg) OCEAN_CUDA(c, cudaMalloc(&c->d_dbg, 3 * np * sizeof(float2)));
    OCEAN_CUDA(c, ocean::launch_propagate_literal(c->d_h0 + tile * np, c->d_omega + tile * np, c->plocals.time, c->n,
                                                  c->domain_size, c->d_dbg, c->d_dbg + np, c->d_dbg + 2 * np, c->stream));
    c->launches += 1;
    float* dst[3] = {h, dx, dz};
    for (int i = 0; i < 3; ++i)
        OCEAN_CUDA(c, cudaMemcpyAsync(dst[i], c->d_dbg + i * np, np * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
    OCEAN_CUDA(c, cudaStreamSynchronize(c->stream));
    return OCEAN_OK;
}

int ocean_get_locals(const ocean_ctx* c, ocean_propagate_locals* p, ocean_correction_locals* cl)
{
    if (!c) return OCEAN_ERR_INVALID_ARG;
    if (p) *p = c->plocals;
    if (cl) *cl = c->clocals;
    return OCEAN_OK;
}

uint32_t ocean_resolution(const ocean_ctx* c) { return c ? c->n : 0; }
uint32_t ocean_n_tiles(const ocean_ctx* c) { return c ? c->n_tiles : 0; }
uint64_t ocean_launch_count(const ocean_ctx* c) { return c ? c->launches : 0; }
void* ocean_stream(const ocean_ctx* c) { return c ? static_cast<void*>(c->stream) : nullptr; }
uint64_t ocean_algorithmic_bytes_per_update(const ocean_ctx* c)
{
    return c ? 76ull * uint64_t(c->n) * c->n * c->n_tiles : 0;
}

}  // extern "C"
