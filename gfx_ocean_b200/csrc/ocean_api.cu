// C ABI of libocean_b200.so (include/ocean_b200.h): context, device memory, frame
// scheduling. Mirrors what Renderer::new / Renderer::render do for the compute half of
// the reference (src/render.rs:607-729 buffers, :742-931 upload, :1101-1310 per-frame
// record) with the operator holders of src/fft.rs and src/ocean.rs folded into one
// opaque context that owns every device buffer.
#include "../../include/ocean_b200.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <algorithm>
#include <vector>

#include "host/lane_order.hpp"
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
    // displacement_map stand-in (src/render.rs:820-845): linear float4[buffer][tile][y][x]; two buffers under
    // OCEAN_FLAG_DOUBLE_BUFFER_OUTPUT (frame n+1 is computed while frame n is being read back), else one
    float4* d_out = nullptr;
    uint32_t n_buffers = 1, cur = 0;
    // where each tile's map goes (own buffer or a caller-provided allocation): host mirror + one device table per buffer
    std::vector<ocean::OutDesc> out_tab[2];
    ocean::OutDesc* d_out_tab[2] = {nullptr, nullptr};
    bool any_pitched = false;           // some tile has a row pitch != N texels: k_cols runs its pitch-aware build
    // read-back pipeline (double-buffered contexts): copies run on their own stream
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_done[2] = {nullptr, nullptr};      // frame in buffer b is complete
    cudaEvent_t ev_copied[2] = {nullptr, nullptr};    // the last read-back of buffer b is complete
    bool copy_pending[2] = {false, false};
    // scratch: vertex positions of ocean_displace_grid, checksums
    float* d_vtx = nullptr;
    size_t vtx_floats = 0;
    unsigned long long* d_sums = nullptr;
    size_t sums_count = 0;
    // ocean_update_graph: one instantiated CUDA graph per (tile range, k_cols build)
    struct FrameGraph {
        uint32_t first, count;
        bool general;
        cudaGraph_t graph;
        cudaGraphExec_t exec;
        cudaGraphNode_t rows_node;
        cudaKernelNodeParams rows_params;
        std::vector<void*> kparams;
        float time;
    };
    std::vector<FrameGraph*> graphs;
    bool capturing = false;
    // ocean_update_overlapped: two lanes (stream + own intermediate set), frames alternate between them
    struct Lane {
        cudaStream_t stream = nullptr;
        cudaEvent_t done = nullptr;       // the lane's latest frame (and, by stream order, every earlier one) is complete
        uint32_t first = 0, count = 0;    // tile range of the latest frame (OCEAN_B200_DEBUG_LANES_LATEST only)
    } lanes[2];
    ocean::LaneOrder lane_order;          // which column kernels wait for the other lane (host/lane_order.hpp)
    cudaEvent_t ev_main = nullptr;        // work enqueued on the main stream that the lanes must see (inputs, output routing)
    bool main_dirty = true;               // the main stream saw activity since the lanes last synchronised with it
    uint32_t next_lane = 0;
    // literal pipeline: dx_spec | dy_spec | dz_spec of the tile in flight (src/render.rs:608-646)
    float2* d_spec = nullptr;
    // fused pipeline: row-pass output, consumed by the column pass
    void* d_work = nullptr;
    size_t work_bytes = 0;
    ocean::FusedPlan* plan = nullptr;
    // normal map of the consumer step (lazy), float4[tile][y][x]
    float4* d_nrm = nullptr;
    // OCEAN_FLAG_DX_PLANE: dense copy of channel .x per output buffer, float[buffer][tile][y][x], written by k_cols
    float* d_dxp = nullptr;
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

// Entry points run on the context's device and leave the caller's current device as they found it.
struct DeviceGuard {
    int prev = -1, dev;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int d) : dev(d)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) err = cudaSetDevice(dev);
    }
    ~DeviceGuard()
    {
        if (prev >= 0 && prev != dev) cudaSetDevice(prev);
    }
};
int join_lanes(ocean_ctx* c);
// Entry points that enqueue on (or synchronise with) the context's main stream: run on its device and first order the
// main stream behind any frame still in flight on the lanes of ocean_update_overlapped.
#define OCEAN_ON_DEVICE(ctx)                                             \
    DeviceGuard guard_((ctx)->device);                                   \
    if (guard_.err != cudaSuccess) return cuda_fail((ctx), guard_.err, "cudaSetDevice"); \
    (ctx)->main_dirty = true;                                            \
    if (int jrc_ = join_lanes(ctx)) return jrc_

// Everything but ocean_update_overlapped runs on the main stream: make it wait for frames still in flight on the lanes.
int join_lanes(ocean_ctx* c)
{
    for (int i = 0; i < 2; ++i)
        if (c->lane_order.busy[i]) OCEAN_CUDA(c, cudaStreamWaitEvent(c->stream, c->lanes[i].done, 0));
    c->lane_order.main_joined();          // (the lanes pick the main stream's work up again through ev_main, see below)
    return OCEAN_OK;
}

float4* tile_out(const ocean_ctx* c, uint32_t tile) { return c->out_tab[c->cur][tile].base; }
size_t tile_pitch(const ocean_ctx* c, uint32_t tile) { return c->out_tab[c->cur][tile].pitch; }

int upload_out_tab(ocean_ctx* c, uint32_t b)
{
    OCEAN_CUDA(c, cudaMemcpyAsync(c->d_out_tab[b], c->out_tab[b].data(), c->n_tiles * sizeof(ocean::OutDesc),
                                  cudaMemcpyHostToDevice, c->stream));
    OCEAN_CUDA(c, cudaStreamSynchronize(c->stream));     // the host mirror may change right after
    return OCEAN_OK;
}

int ensure_sums(ocean_ctx* c, size_t count)
{
    if (c->sums_count < count) {
        cudaFree(c->d_sums);
        c->d_sums = nullptr;
        c->sums_count = 0;
        OCEAN_CUDA(c, cudaMalloc(&c->d_sums, count * sizeof(unsigned long long)));
        c->sums_count = count;
    }
    return OCEAN_OK;
}

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
    OCEAN_CUDA(c, ocean::launch_correction_literal(dy, dx, dz, c->n, tile_out(c, t), tile_pitch(c, t), c->stream));
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
        return fail(nullptr, OCEAN_ERR_UNSUPPORTED, "fused pipeline supports N in {64, 128, 256, 512, 1024, 2048}");
    if (cfg->flags & ~(OCEAN_FLAG_KEEP_SPECTRA | OCEAN_FLAG_DOUBLE_BUFFER_OUTPUT | OCEAN_FLAG_DX_PLANE))
        return fail(nullptr, OCEAN_ERR_INVALID_ARG, "unknown flag bits");

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
    DeviceGuard guard(c->device);
    if (guard.err != cudaSuccess) return bail(guard.err, "cudaSetDevice");
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
    c->n_buffers = (c->flags & OCEAN_FLAG_DOUBLE_BUFFER_OUTPUT) ? 2 : 1;
    if ((e = cudaMalloc(&c->d_out, c->n_buffers * nt * np * sizeof(float4))) != cudaSuccess) return bail(e, "cudaMalloc(out)");
    for (uint32_t b = 0; b < c->n_buffers; ++b) {
        c->out_tab[b].resize(nt);
        for (size_t t = 0; t < nt; ++t) c->out_tab[b][t] = {c->d_out + (b * nt + t) * np, c->n, 0};
        if ((e = cudaMalloc(&c->d_out_tab[b], nt * sizeof(ocean::OutDesc))) != cudaSuccess) return bail(e, "cudaMalloc(out_tab)");
        if ((e = cudaMemcpy(c->d_out_tab[b], c->out_tab[b].data(), nt * sizeof(ocean::OutDesc), cudaMemcpyHostToDevice)) != cudaSuccess)
            return bail(e, "cudaMemcpy(out_tab)");
    }
    if ((c->flags & OCEAN_FLAG_DX_PLANE) && c->pipeline == OCEAN_PIPELINE_FUSED && c->n >= 32) {
        if ((e = cudaMalloc(&c->d_dxp, c->n_buffers * nt * np * sizeof(float))) != cudaSuccess) return bail(e, "cudaMalloc(dx plane)");
    }
    if (c->n_buffers == 2) {
        if ((e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate(copy)");
        for (int b = 0; b < 2; ++b) {
            if ((e = cudaEventCreateWithFlags(&c->ev_done[b], cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
            if ((e = cudaEventCreateWithFlags(&c->ev_copied[b], cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
        }
    }
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
    DeviceGuard guard(c->device);
    // an external stream must outlive the context (include/ocean_b200.h): it is drained here, not destroyed
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) {
        cudaStreamSynchronize(c->copy_stream);
        cudaStreamDestroy(c->copy_stream);
    }
    for (int b = 0; b < 2; ++b) {
        if (c->ev_done[b]) cudaEventDestroy(c->ev_done[b]);
        if (c->ev_copied[b]) cudaEventDestroy(c->ev_copied[b]);
        cudaFree(c->d_out_tab[b]);
    }
    for (auto& l : c->lanes) {
        if (l.stream) {
            cudaStreamSynchronize(l.stream);
            cudaStreamDestroy(l.stream);
        }
        if (l.done) cudaEventDestroy(l.done);
    }
    if (c->ev_main) cudaEventDestroy(c->ev_main);
    for (auto* g : c->graphs) {
        cudaGraphExecDestroy(g->exec);
        cudaGraphDestroy(g->graph);
        delete g;
    }
    cudaFree(c->d_vtx);
    cudaFree(c->d_sums);
    if (c->plan) ocean::fused_plan_destroy(c->plan);
    cudaFree(c->d_h0);
    cudaFree(c->d_omega);
    cudaFree(c->d_out);
    cudaFree(c->d_spec);
    cudaFree(c->d_work);
    cudaFree(c->d_dbg);
    cudaFree(c->d_nrm);
    cudaFree(c->d_dxp);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    (void)cudaGetLastError();
    delete c;
}

int ocean_set_spectrum(ocean_ctx* c, uint32_t tile, const float* h0_xy, const float* omega)
{
    if (int rc = check_tile(c, tile)) return rc;
    if (!h0_xy || !omega) return fail(c, OCEAN_ERR_INVALID_ARG, "null spectrum pointer");
    OCEAN_ON_DEVICE(c);
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
    OCEAN_ON_DEVICE(c);
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

}  // extern "C"

namespace {

// One frame for tiles [first_tile, first_tile + count); sums (optional, device): per-tile checksum accumulators.
// lane_stream / lane: enqueue on that stream with the plan's second intermediate set (ocean_update_overlapped).
int enqueue_frame(ocean_ctx* c, float time, uint32_t first_tile, uint32_t count, unsigned long long* sums, cudaEvent_t* ev,
                  cudaStream_t lane_stream = nullptr, int lane = 0, cudaEvent_t cols_after = nullptr)
{
    cudaStream_t st = lane_stream ? lane_stream : c->stream;
    c->plocals = {time, int32_t(c->n), c->domain_size};      // src/render.rs:1101-1120
    c->clocals = {c->n};
    if (c->n_buffers == 2) {
        // flip to the other output buffer; its previous contents may still be on their way to the host
        const uint32_t nb = c->cur ^ 1u;
        if (c->copy_pending[nb]) {
            OCEAN_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copied[nb], 0));
            c->copy_pending[nb] = false;
        }
        c->cur = nb;
    }
    if (c->pipeline == OCEAN_PIPELINE_LITERAL) {
        for (uint32_t t = first_tile; t < first_tile + count; ++t) {
            if (int rc = enqueue_literal(c, t, time)) return rc;
            if (sums) {
                OCEAN_CUDA(c, ocean::launch_checksum(tile_out(c, t), tile_pitch(c, t), c->n, sums + (t - first_tile), c->stream));
                c->launches += 1;
            }
        }
    } else {
        // tiles per kernel pair: the whole range by default; OCEAN_B200_BATCH=k splits it so that a batch's row-pass
        // output is still in L2 when its column pass reads it (experiment knob)
        static const uint32_t batch_env = [] { const char* v = std::getenv("OCEAN_B200_BATCH"); return v ? uint32_t(std::atoi(v)) : 0u; }();
        const uint32_t batch = (batch_env && !ev && !c->capturing) ? batch_env : count;   // a recorded frame is one kernel pair
        for (uint32_t t0 = first_tile; t0 < first_tile + count; t0 += batch) {
            const uint32_t cnt = t0 + batch <= first_tile + count ? batch : first_tile + count - t0;
            uint32_t nl = 0;
            OCEAN_CUDA(c, ocean::fused_enqueue(c->plan, c->d_h0, c->d_omega, c->d_out_tab[c->cur], time, t0, cnt, st,
                                               &nl, ev, c->any_pitched || sums != nullptr, sums ? sums + (t0 - first_tile) : nullptr,
                                               c->d_dxp ? c->d_dxp + size_t(c->cur) * c->n_tiles * pts(c) : nullptr, lane, cols_after));
            c->launches += nl;
        }
    }
    if (c->n_buffers == 2) OCEAN_CUDA(c, cudaEventRecord(c->ev_done[c->cur], c->stream));
    c->updated = true;
    return OCEAN_OK;
}

int check_range(ocean_ctx* c, uint32_t first_tile, uint32_t count, bool need_spectra)
{
    if (count == 0 || first_tile >= c->n_tiles || count > c->n_tiles - first_tile)
        return fail(c, OCEAN_ERR_INVALID_ARG, "tile range out of bounds");
    if (need_spectra)
        for (uint32_t t = first_tile; t < first_tile + count; ++t)
            if (!c->loaded[t]) return fail(c, OCEAN_ERR_NOT_READY, "tile " + std::to_string(t) + " has no spectrum");
    return OCEAN_OK;
}

}  // namespace

extern "C" {

int ocean_update_tiles(ocean_ctx* c, float time, uint32_t first_tile, uint32_t count)
{
    if (!c) return OCEAN_ERR_INVALID_ARG;
    if (int rc = check_range(c, first_tile, count, true)) return rc;
    if (c->n_buffers == 2 && count != c->n_tiles)
        return fail(c, OCEAN_ERR_UNSUPPORTED, "a double-buffered context updates all of its tiles together");
    OCEAN_ON_DEVICE(c);
    return enqueue_frame(c, time, first_tile, count, nullptr, nullptr);
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

// Index of `time` in the parameter list of k_rows / k_rows_p (kernels_fused.cu): the only per-frame argument.
static constexpr int kRowsTimeParam = 6, kRowsParamCount = 9;

int ocean_update_graph(ocean_ctx* c, float time, uint32_t first_tile, uint32_t count)
{
    if (!c) return OCEAN_ERR_INVALID_ARG;
    if (int rc = check_range(c, first_tile, count, true)) return rc;
    if (c->pipeline != OCEAN_PIPELINE_FUSED || c->n_buffers != 1)
        return fail(c, OCEAN_ERR_UNSUPPORTED, "graph replay is available for the fused pipeline on a single-buffered context");
    OCEAN_ON_DEVICE(c);
    ocean_ctx::FrameGraph* g = nullptr;
    for (auto* e : c->graphs)
        if (e->first == first_tile && e->count == count && e->general == c->any_pitched) g = e;
    if (!g) {
        // record the frame once (the two launches with their programmatic dependency), replay it afterwards
        g = new (std::nothrow) ocean_ctx::FrameGraph{first_tile, count, c->any_pitched, nullptr, nullptr, nullptr, {}, {}, time};
        if (!g) return fail(c, OCEAN_ERR_CUDA, "out of host memory");
        const uint64_t launches_before = c->launches;
        cudaError_t e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
        c->capturing = true;
        int rc = e == cudaSuccess ? enqueue_frame(c, time, first_tile, count, nullptr, nullptr) : cuda_fail(c, e, "cudaStreamBeginCapture");
        c->capturing = false;
        if (e == cudaSuccess) {
            e = cudaStreamEndCapture(c->stream, &g->graph);
            if (rc == OCEAN_OK && e != cudaSuccess) rc = cuda_fail(c, e, "cudaStreamEndCapture");
        }
        c->launches = launches_before;           // nothing ran yet
        size_t n_roots = 1;
        if (rc == OCEAN_OK && (e = cudaGraphInstantiate(&g->exec, g->graph, 0)) != cudaSuccess) rc = cuda_fail(c, e, "cudaGraphInstantiate");
        if (rc == OCEAN_OK && ((e = cudaGraphGetRootNodes(g->graph, &g->rows_node, &n_roots)) != cudaSuccess || n_roots != 1))
            rc = e != cudaSuccess ? cuda_fail(c, e, "cudaGraphGetRootNodes") : fail(c, OCEAN_ERR_CUDA, "unexpected frame graph shape");
        if (rc == OCEAN_OK && (e = cudaGraphKernelNodeGetParams(g->rows_node, &g->rows_params)) != cudaSuccess)
            rc = cuda_fail(c, e, "cudaGraphKernelNodeGetParams");
        if (rc != OCEAN_OK) {
            if (g->exec) cudaGraphExecDestroy(g->exec);
            if (g->graph) cudaGraphDestroy(g->graph);
            delete g;
            return rc;
        }
        g->kparams.assign(g->rows_params.kernelParams, g->rows_params.kernelParams + kRowsParamCount);
        g->kparams[kRowsTimeParam] = &g->time;
        g->rows_params.kernelParams = g->kparams.data();
        c->graphs.push_back(g);
    }
    g->time = time;
    OCEAN_CUDA(c, cudaGraphExecKernelNodeSetParams(g->exec, g->rows_node, &g->rows_params));
    OCEAN_CUDA(c, cudaGraphLaunch(g->exec, c->stream));
    c->plocals = {time, int32_t(c->n), c->domain_size};
    c->launches += 2;
    c->updated = true;
    return OCEAN_OK;
}

int ocean_update_overlapped(ocean_ctx* c, float time, uint32_t first_tile, uint32_t count)
{
    if (!c) return OCEAN_ERR_INVALID_ARG;
    if (int rc = check_range(c, first_tile, count, true)) return rc;
    if (c->pipeline != OCEAN_PIPELINE_FUSED || c->n_buffers != 1)
        return fail(c, OCEAN_ERR_UNSUPPORTED, "overlapped updates are available for the fused pipeline on a single-buffered context");
    DeviceGuard guard(c->device);
    if (guard.err != cudaSuccess) return cuda_fail(c, guard.err, "cudaSetDevice");
    if (!c->ev_main) {
        for (auto& l : c->lanes) {
            OCEAN_CUDA(c, cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
            OCEAN_CUDA(c, cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
        }
        OCEAN_CUDA(c, cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
    }
    c->lane_order.resize(c->n_tiles);
    const uint32_t li = c->next_lane;
    auto& L = c->lanes[li];
    auto& O = c->lanes[li ^ 1u];
    if (c->main_dirty) {
        // uploads, output routing and plain updates enqueued on the main stream come first, on both lanes
        OCEAN_CUDA(c, cudaEventRecord(c->ev_main, c->stream));
        OCEAN_CUDA(c, cudaStreamWaitEvent(c->lanes[0].stream, c->ev_main, 0));
        OCEAN_CUDA(c, cudaStreamWaitEvent(c->lanes[1].stream, c->ev_main, 0));
        c->main_dirty = false;
        // an entry point that touched the main stream joined both lanes first (OCEAN_ON_DEVICE), so the event just
        // recorded is behind every frame enqueued so far: both lanes are now ordered behind all of them
        c->lane_order.lanes_resumed();
    }
    // The frame in flight on the other lane uses the other intermediate set, so this frame's row kernel never has to
    // wait for it; only when both frames write the same maps is this frame's COLUMN kernel ordered behind that frame.
    // "That frame" is ANY frame of the other lane this lane is not ordered behind yet, not only its latest one
    // (host/lane_order.hpp). O.done is recorded behind all of them (stream order).
    static const bool latest_only = std::getenv("OCEAN_B200_DEBUG") && std::getenv("OCEAN_B200_DEBUG_LANES_LATEST");
    const bool other_busy = c->lane_order.busy[li ^ 1u];
    bool same_maps = c->lane_order.enqueue(int(li), first_tile, count);
    // hazard hunting: look at the other lane's LATEST frame only (what this entry point first shipped with), to show
    // that tests/test_gpu_features.py::test_overlapped_frame_returning_to_a_tile_of_an_older_frame_of_the_other_lane sees the gap
    if (latest_only) same_maps = other_busy && first_tile < O.first + O.count && O.first < first_tile + count;
    if (int rc = enqueue_frame(c, time, first_tile, count, nullptr, nullptr, L.stream, int(li), same_maps ? O.done : nullptr)) return rc;
    OCEAN_CUDA(c, cudaEventRecord(L.done, L.stream));
    L.first = first_tile;
    L.count = count;
    c->next_lane = li ^ 1u;
    return OCEAN_OK;
}

int ocean_join(ocean_ctx* c)
{
    if (!c) return OCEAN_ERR_INVALID_ARG;
    OCEAN_ON_DEVICE(c);
    return OCEAN_OK;
}

int ocean_update_sequence_checksums(ocean_ctx* c, float t0, float dt, uint32_t n_frames, uint64_t* h_sums)
{
    if (!c) return OCEAN_ERR_INVALID_ARG;
    if (!h_sums || n_frames == 0) return fail(c, OCEAN_ERR_INVALID_ARG, "null output pointer or no frames");
    if (int rc = check_range(c, 0, c->n_tiles, true)) return rc;
    OCEAN_ON_DEVICE(c);
    const size_t total = size_t(n_frames) * c->n_tiles;
    if (int rc = ensure_sums(c, total)) return rc;
    OCEAN_CUDA(c, cudaMemsetAsync(c->d_sums, 0, total * sizeof(unsigned long long), c->stream));
    // back-to-back frames exactly as ocean_update_sequence enqueues them (same launch attributes: the fused k_cols
    // accumulates the checksum of what it stores, so no extra kernel sits between a frame and the next one)
    for (uint32_t i = 0; i < n_frames; ++i)
        if (int rc = enqueue_frame(c, t0 + dt * float(i), 0, c->n_tiles, c->d_sums + size_t(i) * c->n_tiles, nullptr)) return rc;
    OCEAN_CUDA(c, cudaMemcpyAsync(h_sums, c->d_sums, total * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    OCEAN_CUDA(c, cudaStreamSynchronize(c->stream));
    return OCEAN_OK;
}

int ocean_output_checksums(ocean_ctx* c, uint64_t* h_sums)
{
    if (!c) return OCEAN_ERR_INVALID_ARG;
    if (!h_sums) return fail(c, OCEAN_ERR_INVALID_ARG, "null output pointer");
    if (!c->updated) return fail(c, OCEAN_ERR_NOT_READY, "ocean_update has not been called");
    OCEAN_ON_DEVICE(c);
    if (int rc = ensure_sums(c, c->n_tiles)) return rc;
    OCEAN_CUDA(c, cudaMemsetAsync(c->d_sums, 0, c->n_tiles * sizeof(unsigned long long), c->stream));
    const char* dbg = std::getenv("OCEAN_B200_DEBUG") ? std::getenv("OCEAN_B200_DEBUG_INTER") : nullptr;
    for (uint32_t t = 0; t < c->n_tiles; ++t) {
        if (dbg && c->plan) {
            // hazard hunting: checksum the row-pass output (gp or gh) instead of the displacement map
            const float2 *gp, *gh;
            size_t ngp, ngh;
            ocean::fused_plan_intermediate(c->plan, t, &gp, &ngp, &gh, &ngh);
            const float2* src = dbg[0] == 'h' ? gh : gp;
            const size_t n4 = (dbg[0] == 'h' ? ngh : ngp) / 2;         // float4 count
            const uint32_t side = 512;
            OCEAN_CUDA(c, ocean::launch_checksum(reinterpret_cast<const float4*>(src), side, side, c->d_sums + t, c->stream));
            (void)n4;
        } else {
            OCEAN_CUDA(c, ocean::launch_checksum(tile_out(c, t), tile_pitch(c, t), c->n, c->d_sums + t, c->stream));
        }
    }
    c->launches += c->n_tiles;
    OCEAN_CUDA(c, cudaMemcpyAsync(h_sums, c->d_sums, c->n_tiles * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    OCEAN_CUDA(c, cudaStreamSynchronize(c->stream));
    return OCEAN_OK;
}

int ocean_set_output_device(ocean_ctx* c, uint32_t tile, float* d_rgba, size_t row_pitch_bytes)
{
    if (int rc = check_tile(c, tile)) return rc;
    if (c->n_buffers == 2) return fail(c, OCEAN_ERR_UNSUPPORTED, "external outputs are not available on a double-buffered context");
    OCEAN_ON_DEVICE(c);
    ocean::OutDesc d;
    if (d_rgba) {
        if (row_pitch_bytes == 0) row_pitch_bytes = size_t(c->n) * sizeof(float4);
        if (row_pitch_bytes % sizeof(float4) || row_pitch_bytes < size_t(c->n) * sizeof(float4) || row_pitch_bytes / sizeof(float4) > 0xffffffffull)
            return fail(c, OCEAN_ERR_INVALID_ARG, "row pitch must be a multiple of 16 bytes and at least N * 16");
        if (reinterpret_cast<uintptr_t>(d_rgba) % sizeof(float4))
            return fail(c, OCEAN_ERR_INVALID_ARG, "output pointer must be 16-byte aligned");
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, d_rgba) != cudaSuccess || (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged) ||
            (attr.type == cudaMemoryTypeDevice && attr.device != c->device)) {
            (void)cudaGetLastError();
            return fail(c, OCEAN_ERR_INVALID_ARG, "output pointer is not device memory of the context's device");
        }
        d = {reinterpret_cast<float4*>(d_rgba), uint32_t(row_pitch_bytes / sizeof(float4)), 0};
    } else {
        d = {c->d_out + size_t(tile) * pts(c), c->n, 0};       // back to the context's own buffer
    }
    // frames already enqueued keep the old destination: the table is replaced in stream order
    c->out_tab[0][tile] = d;
    if (int rc = upload_out_tab(c, 0)) return rc;
    c->any_pitched = false;
    for (const auto& e : c->out_tab[0]) c->any_pitched |= (e.pitch != c->n);
    return OCEAN_OK;
}

int ocean_compute_normals(ocean_ctx* c, uint32_t first_tile, uint32_t count)
{
    if (!c) return OCEAN_ERR_INVALID_ARG;
    if (int rc = check_range(c, first_tile, count, false)) return rc;
    if (!c->updated) return fail(c, OCEAN_ERR_NOT_READY, "ocean_update has not been called");
    OCEAN_ON_DEVICE(c);
    if (!c->d_nrm) OCEAN_CUDA(c, cudaMalloc(&c->d_nrm, size_t(c->n_tiles) * pts(c) * sizeof(float4)));
    bool own = true;                       // all tiles of the range in the context's own dense buffer: one launch
    for (uint32_t t = first_tile; t < first_tile + count; ++t)
        own &= tile_out(c, t) == c->d_out + (size_t(c->cur) * c->n_tiles + t) * pts(c);
    if (c->d_dxp) {
        // the column kernel left a dense copy of channel .x beside the map (external outputs included): 4 B/pt to read
        OCEAN_CUDA(c, ocean::launch_normal_map_plane(c->d_dxp + (size_t(c->cur) * c->n_tiles + first_tile) * pts(c),
                                                     c->d_nrm + first_tile * pts(c), c->n, count, c->stream));
        c->launches += 1;
    } else if (own) {
        OCEAN_CUDA(c, ocean::launch_normal_map(tile_out(c, first_tile), c->n, pts(c), c->d_nrm + first_tile * pts(c), c->n, count, c->stream));
        c->launches += 1;
    } else {
        for (uint32_t t = first_tile; t < first_tile + count; ++t)
            OCEAN_CUDA(c, ocean::launch_normal_map(tile_out(c, t), tile_pitch(c, t), 0, c->d_nrm + t * pts(c), c->n, 1, c->stream));
        c->launches += count;
    }
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
    OCEAN_ON_DEVICE(c);
    OCEAN_CUDA(c, cudaMemcpyAsync(h_nrm, c->d_nrm + tile * pts(c), pts(c) * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    OCEAN_CUDA(c, cudaStreamSynchronize(c->stream));
    return OCEAN_OK;
}

int ocean_displace_grid_device(ocean_ctx* c, uint32_t tile, uint32_t grid, float offset_x, float offset_z, float* d_pos_world)
{
    if (int rc = check_tile(c, tile)) return rc;
    if (!d_pos_world) return fail(c, OCEAN_ERR_INVALID_ARG, "null output pointer");
    if (grid < 2 || grid > 16384) return fail(c, OCEAN_ERR_INVALID_ARG, "grid must be in [2, 16384]");
    if (!c->updated) return fail(c, OCEAN_ERR_NOT_READY, "ocean_update has not been called");
    OCEAN_ON_DEVICE(c);
    OCEAN_CUDA(c, ocean::launch_displace_grid(tile_out(c, tile), tile_pitch(c, tile), c->n, grid, offset_x, offset_z, d_pos_world, c->stream));
    c->launches += 1;
    return OCEAN_OK;
}

int ocean_displace_grid(ocean_ctx* c, uint32_t tile, uint32_t grid, float offset_x, float offset_z, float* h_pos_world)
{
    if (int rc = check_tile(c, tile)) return rc;
    if (!h_pos_world) return fail(c, OCEAN_ERR_INVALID_ARG, "null output pointer");
    if (grid < 2 || grid > 16384) return fail(c, OCEAN_ERR_INVALID_ARG, "grid must be in [2, 16384]");
    OCEAN_ON_DEVICE(c);
    const size_t need = size_t(grid) * grid * 3;
    if (c->vtx_floats < need) {
        cudaFree(c->d_vtx);
        c->d_vtx = nullptr;
        c->vtx_floats = 0;
        OCEAN_CUDA(c, cudaMalloc(&c->d_vtx, need * sizeof(float)));
        c->vtx_floats = need;
    }
    if (int rc = ocean_displace_grid_device(c, tile, grid, offset_x, offset_z, c->d_vtx)) return rc;
    OCEAN_CUDA(c, cudaMemcpyAsync(h_pos_world, c->d_vtx, need * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    OCEAN_CUDA(c, cudaStreamSynchronize(c->stream));
    return OCEAN_OK;
}

int ocean_generate_spectrum(ocean_ctx* c, uint32_t tile, uint64_t seed, uint32_t stream_id, const ocean_spectrum_params* params,
                            uint32_t* h_words)
{
    if (int rc = check_tile(c, tile)) return rc;
    ocean_spectrum_params p = {3.0e-8f, 30.0f, 9.81f, 100.0f};
    if (params) p = *params;
    if (!(p.amplitude > 0.f) || !(p.wind_speed > 0.f) || !(p.gravity > 0.f) || !(p.depth > 0.f))
        return fail(c, OCEAN_ERR_INVALID_ARG, "spectrum parameters must be positive");
    OCEAN_ON_DEVICE(c);
    const size_t np = pts(c);
    uint32_t* d_words = nullptr;
    if (h_words) OCEAN_CUDA(c, cudaMalloc(&d_words, np * 4 * sizeof(uint32_t)));
    cudaError_t e = ocean::launch_generate_spectrum(c->d_h0 + tile * np, c->d_omega + tile * np, c->n, stream_id, seed, c->domain_size,
                                                    p.amplitude, p.wind_speed, p.gravity, p.depth, d_words, c->stream);
    c->launches += 1;
    if (e == cudaSuccess && h_words) {
        e = cudaMemcpyAsync(h_words, d_words, np * 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    }
    cudaFree(d_words);
    if (e != cudaSuccess) return cuda_fail(c, e, "ocean_generate_spectrum");
    c->loaded[tile] = 1;
    return OCEAN_OK;
}

int ocean_get_spectrum(ocean_ctx* c, uint32_t tile, float* h0_xy, float* omega)
{
    if (int rc = check_tile(c, tile)) return rc;
    if (!h0_xy || !omega) return fail(c, OCEAN_ERR_INVALID_ARG, "null output pointer");
    if (!c->loaded[tile]) return fail(c, OCEAN_ERR_NOT_READY, "tile has no spectrum");
    OCEAN_ON_DEVICE(c);
    const size_t np = pts(c);
    OCEAN_CUDA(c, cudaMemcpyAsync(h0_xy, c->d_h0 + tile * np, np * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
    OCEAN_CUDA(c, cudaMemcpyAsync(omega, c->d_omega + tile * np, np * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    OCEAN_CUDA(c, cudaStreamSynchronize(c->stream));
    return OCEAN_OK;
}

int ocean_profile_update(ocean_ctx* c, float time, float* stage_ms, uint32_t capacity, uint32_t* n_stages)
{
    if (!c) return OCEAN_ERR_INVALID_ARG;
    if (!stage_ms || !n_stages) return fail(c, OCEAN_ERR_INVALID_ARG, "null output pointer");
    if (int rc = check_range(c, 0, c->n_tiles, true)) return rc;
    OCEAN_ON_DEVICE(c);
    const uint32_t stages = c->pipeline == OCEAN_PIPELINE_FUSED ? 2u : 8u;
    if (capacity < stages) return fail(c, OCEAN_ERR_INVALID_ARG, "stage_ms too small");
    std::vector<cudaEvent_t> ev(stages + 1, nullptr);
    int rc = OCEAN_OK;
    for (auto& e : ev) {
        const cudaError_t ce = cudaEventCreate(&e);
        if (ce != cudaSuccess) { rc = cuda_fail(c, ce, "cudaEventCreate"); break; }
    }
    if (rc == OCEAN_OK && c->pipeline == OCEAN_PIPELINE_FUSED) {
        // an unmeasured frame first: the measured kernels are already queued behind it when it finishes, so the
        // intervals between the events are kernel durations, not host launch latency after an idle stream
        rc = enqueue_frame(c, time, 0, c->n_tiles, nullptr, nullptr);
        if (rc == OCEAN_OK) rc = enqueue_frame(c, time, 0, c->n_tiles, nullptr, ev.data());
    } else if (rc == OCEAN_OK) {
        // LITERAL: tile 0 only, stage by stage in the reference's dispatch order (src/render.rs:1122-1287)
        c->plocals = {time, int32_t(c->n), c->domain_size};
        c->clocals = {c->n};
        const size_t np = pts(c);
        float2 *dx = c->d_spec, *dy = c->d_spec + np, *dz = c->d_spec + 2 * np;
        int k = 0;
        cudaError_t e = cudaEventRecord(ev[k++], c->stream);
        auto step = [&](cudaError_t launched) {
            if (e == cudaSuccess) e = launched;
            if (e == cudaSuccess) e = cudaEventRecord(ev[k], c->stream);
            ++k;
        };
        step(ocean::launch_propagate_literal(c->d_h0, c->d_omega, time, c->n, c->domain_size, dy, dx, dz, c->stream));
        for (float2* f : {dx, dy, dz}) step(ocean::launch_fft_row_literal(f, c->n, c->stream));
        for (float2* f : {dx, dy, dz}) step(ocean::launch_fft_col_literal(f, c->n, c->stream));
        step(ocean::launch_correction_literal(dy, dx, dz, c->n, tile_out(c, 0), tile_pitch(c, 0), c->stream));
        c->launches += 8;
        c->updated = true;
        if (e != cudaSuccess) rc = cuda_fail(c, e, "literal stages");
    }
    if (rc == OCEAN_OK) {
        const cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = cuda_fail(c, e, "cudaStreamSynchronize");
    }
    if (rc == OCEAN_OK) {
        for (uint32_t i = 0; i < stages; ++i) cudaEventElapsedTime(&stage_ms[i], ev[i], ev[i + 1]);
        *n_stages = stages;
    }
    for (auto& e : ev)
        if (e) cudaEventDestroy(e);
    return rc;
}

int ocean_output_device(ocean_ctx* c, uint32_t tile, const float** d_rgba)
{
    if (int rc = check_tile(c, tile)) return rc;
    if (!d_rgba) return fail(c, OCEAN_ERR_INVALID_ARG, "null output pointer");
    *d_rgba = reinterpret_cast<const float*>(tile_out(c, tile));
    return OCEAN_OK;
}

int ocean_download_async(ocean_ctx* c, uint32_t tile, float* h_rgba)
{
    if (int rc = check_tile(c, tile)) return rc;
    if (!h_rgba) return fail(c, OCEAN_ERR_INVALID_ARG, "null output pointer");
    if (!c->updated) return fail(c, OCEAN_ERR_NOT_READY, "ocean_update has not been called");
    OCEAN_ON_DEVICE(c);
    const size_t row = size_t(c->n) * sizeof(float4);
    OCEAN_CUDA(c, cudaMemcpy2DAsync(h_rgba, row, tile_out(c, tile), tile_pitch(c, tile) * sizeof(float4), row, c->n,
                                    cudaMemcpyDeviceToHost, c->stream));
    return OCEAN_OK;
}

int ocean_download(ocean_ctx* c, uint32_t tile, float* h_rgba)
{
    if (int rc = ocean_download_async(c, tile, h_rgba)) return rc;
    OCEAN_CUDA(c, cudaStreamSynchronize(c->stream));
    return OCEAN_OK;
}

int ocean_download_all_async(ocean_ctx* c, float* h_rgba_all)
{
    if (!c) return OCEAN_ERR_INVALID_ARG;
    if (!h_rgba_all) return fail(c, OCEAN_ERR_INVALID_ARG, "null output pointer");
    if (!c->updated) return fail(c, OCEAN_ERR_NOT_READY, "ocean_update has not been called");
    OCEAN_ON_DEVICE(c);
    if (c->n_buffers == 2) {
        // the copy runs on its own stream behind the frame that filled this buffer; the next ocean_update computes
        // into the other buffer meanwhile
        const uint32_t b = c->cur;
        OCEAN_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_done[b], 0));
        OCEAN_CUDA(c, cudaMemcpyAsync(h_rgba_all, c->d_out + size_t(b) * c->n_tiles * pts(c), size_t(c->n_tiles) * pts(c) * sizeof(float4),
                                      cudaMemcpyDeviceToHost, c->copy_stream));
        OCEAN_CUDA(c, cudaEventRecord(c->ev_copied[b], c->copy_stream));
        c->copy_pending[b] = true;
        return OCEAN_OK;
    }
    for (uint32_t t = 0; t < c->n_tiles; ++t)
        if (int rc = ocean_download_async(c, t, h_rgba_all + size_t(t) * pts(c) * 4)) return rc;
    return OCEAN_OK;
}

int ocean_download_fence(ocean_ctx* c, uint32_t lag)
{
    if (!c) return OCEAN_ERR_INVALID_ARG;
    OCEAN_ON_DEVICE(c);
    if (c->n_buffers == 2) {
        // lag 1: everything but the newest read-back is on the host; lag 0: all of them
        if (lag >= 1) {
            const uint32_t older = c->cur ^ 1u;
            if (c->copy_pending[older]) OCEAN_CUDA(c, cudaEventSynchronize(c->ev_copied[older]));
        } else {
            OCEAN_CUDA(c, cudaStreamSynchronize(c->copy_stream));
        }
        return OCEAN_OK;
    }
    OCEAN_CUDA(c, cudaStreamSynchronize(c->stream));
    return OCEAN_OK;
}

int ocean_sync(ocean_ctx* c)
{
    if (!c) return OCEAN_ERR_INVALID_ARG;
    OCEAN_ON_DEVICE(c);
    OCEAN_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->copy_stream) OCEAN_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    return OCEAN_OK;
}

int ocean_debug_spectra(ocean_ctx* c, uint32_t tile, float* h, float* dx, float* dz)
{
    if (int rc = check_tile(c, tile)) return rc;
    if (!h || !dx || !dz) return fail(c, OCEAN_ERR_INVALID_ARG, "null output pointer");
    if (!c->loaded[tile]) return fail(c, OCEAN_ERR_NOT_READY, "tile has no spectrum");
    OCEAN_ON_DEVICE(c);
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
