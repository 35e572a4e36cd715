// C++ host-side mirror of the reference's operator interface for the per-frame compute path,
// header-only, on top of the C ABI (include/ocean_b200.h). The reference is Rust and this image
// has no Rust toolchain, so the host side above the C ABI is C++ (and Python, gfx_ocean_b200/ocean.py);
// bindings/rust/ocean.rs carries the same facade as Rust source.
//
// Names follow the reference: PropagateLocals / CorrectionLocals (src/ocean.rs:8-13,179-182),
// constants of src/render.rs:42-46, and an `Ocean` with new/update/output/read_back in place of
// the Propagation<B>/Fft<B>/Correction<B> holders + the dispatch code of Renderer::render().
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/ocean_b200.h"

namespace ocean_b200 {

constexpr uint32_t WORKGROUP_SIZE = 16;                          // src/render.rs:42
constexpr uint32_t WORKGROUP_NUM = 32;                           // :43
constexpr uint32_t RESOLUTION = WORKGROUP_SIZE * WORKGROUP_NUM;  // :44
constexpr float DOMAIN_SIZE = 1000.0f;                           // :46

using PropagateLocals = ocean_propagate_locals;
using CorrectionLocals = ocean_correction_locals;

struct OceanError : std::runtime_error {
    int status;
    OceanError(int s, const std::string& m) : std::runtime_error(m + " [status " + std::to_string(s) + "]"), status(s) {}
};

class Ocean {
public:
    // Ocean::new(resolution, domain_size, &omega, &spectrum)
    Ocean(uint32_t resolution, float domain_size, const float* omega, const float* spectrum_xy, int device = 0)
    {
        check(nullptr, ocean_create(&ctx_, device, resolution, domain_size, 1));
        int rc = ocean_set_spectrum(ctx_, 0, spectrum_xy, omega);
        if (rc != OCEAN_OK) {
            std::string m = ocean_last_error(ctx_);
            ocean_destroy(ctx_);
            ctx_ = nullptr;
            throw OceanError(rc, m);
        }
    }
    // data/omega.bin + data/spectrum.bin as shipped (src/render.rs:769-771,808-810)
    static Ocean from_bincode(const std::string& omega_path, const std::string& spectrum_path,
                              uint32_t resolution = RESOLUTION, float domain_size = DOMAIN_SIZE, int device = 0)
    {
        Ocean o;
        check(nullptr, ocean_create(&o.ctx_, device, resolution, domain_size, 1));
        o.check(ocean_load_bincode(o.ctx_, 0, omega_path.c_str(), spectrum_path.c_str()));
        return o;
    }
    Ocean(Ocean&& other) noexcept : ctx_(other.ctx_) { other.ctx_ = nullptr; }
    Ocean(const Ocean&) = delete;
    Ocean& operator=(const Ocean&) = delete;
    ~Ocean() { ocean_destroy(ctx_); }      // ::destroy(self, device)

    void update(float time) { check(ocean_update(ctx_, time)); }                 // asynchronous
    // frames in flight (src/lib.rs:86): consecutive updates alternate between two internal lanes; same maps, bit for bit
    void update_overlapped(float time) { check(ocean_update_overlapped(ctx_, time, 0, 1)); }
    void join() { check(ocean_join(ctx_)); }   // order the context's stream behind the lanes (every other call does it too)
    void sync() { check(ocean_sync(ctx_)); }
    const float* output() const                                                   // N*N*4 floats on the device
    {
        const float* p = nullptr;
        const_cast<Ocean*>(this)->check(ocean_output_device(ctx_, 0, &p));
        return p;
    }
    // renderer interop: write the map into a caller-provided (e.g. imported Vulkan) allocation, pitch in bytes (0: dense)
    void set_output(float* d_rgba, size_t row_pitch_bytes = 0) { check(ocean_set_output_device(ctx_, 0, d_rgba, row_pitch_bytes)); }
    // shader/ocean.vert:21-25 for the grid x grid vertex patch (src/render.rs:498-506) -> p_PosWorld[grid*grid*3]
    std::vector<float> displace_grid(uint32_t grid, float offset_x, float offset_z)
    {
        std::vector<float> v(size_t(3) * grid * grid);
        check(ocean_displace_grid(ctx_, 0, grid, offset_x, offset_z, v.data()));
        return v;
    }
    uint64_t checksum()
    {
        uint64_t s = 0;
        check(ocean_output_checksums(ctx_, &s));
        return s;
    }
    void read_back(float* rgba) { check(ocean_download(ctx_, 0, rgba)); }
    std::vector<float> read_back()
    {
        std::vector<float> v(size_t(4) * resolution() * resolution());
        read_back(v.data());
        return v;
    }
    PropagateLocals propagate_locals() const
    {
        PropagateLocals p{};
        ocean_get_locals(ctx_, &p, nullptr);
        return p;
    }
    uint32_t resolution() const { return ocean_resolution(ctx_); }
    ocean_ctx* raw() { return ctx_; }

private:
    Ocean() = default;
    static void check(const ocean_ctx* c, int rc)
    {
        if (rc != OCEAN_OK) throw OceanError(rc, ocean_last_error(c));
    }
    void check(int rc) { check(ctx_, rc); }
    ocean_ctx* ctx_ = nullptr;
};

}  // namespace ocean_b200
