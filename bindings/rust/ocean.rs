//! Rust binding of libocean_b200.so (include/ocean_b200.h) -- SOURCE ONLY.
//!
//! This image has no Rust toolchain (no cargo/rustc), so this file is not compiled or tested
//! here; it is the binding a gfx-ocean maintainer would add as `src/ocean_b200.rs` next to
//! `src/ocean.rs` / `src/fft.rs`. It replaces steps 2-10 of `Renderer::render()`
//! (src/render.rs:1101-1310): `ocean.update(time)` instead of the propagate / fft_row x3 /
//! fft_col x3 / correction dispatches, and `ocean.output()` instead of `displacement_map`.
//!
//! build.rs:  println!("cargo:rustc-link-lib=dylib=ocean_b200");
//!            println!("cargo:rustc-link-search=native={}", env::var("OCEAN_B200_LIB_DIR")?);
#![allow(non_camel_case_types)]

use std::ffi::CStr;
use std::fmt;
use std::os::raw::{c_char, c_int, c_void};
use std::ptr;

#[repr(C)]
pub struct ocean_ctx {
    _private: [u8; 0],
}

/// Mirror of `PropagateLocals` (src/ocean.rs:8-13).
#[repr(C)]
#[derive(Debug, Clone, Copy, Default)]
pub struct PropagateLocals {
    pub time: f32,
    pub resolution: i32,
    pub domain_size: f32,
}

/// Mirror of `CorrectionLocals` (src/ocean.rs:179-182).
#[repr(C)]
#[derive(Debug, Clone, Copy, Default)]
pub struct CorrectionLocals {
    pub resolution: u32,
}

#[repr(C)]
pub struct ocean_config {
    pub abi_version: u32,
    pub cuda_device: i32,
    pub resolution: u32,
    pub domain_size: f32,
    pub n_tiles: u32,
    pub pipeline: u32,
    pub stream: *mut c_void,
    pub flags: u32,
}

pub const OCEAN_B200_ABI_VERSION: u32 = 2;
pub const OCEAN_PIPELINE_FUSED: u32 = 0;
pub const OCEAN_PIPELINE_LITERAL: u32 = 1;
pub const OCEAN_FLAG_KEEP_SPECTRA: u32 = 1;
pub const OCEAN_FLAG_DOUBLE_BUFFER_OUTPUT: u32 = 2;
pub const OCEAN_FLAG_DX_PLANE: u32 = 4;

/// Parameters of `ocean_generate_spectrum` (NULL: amplitude 3e-8, wind 30 m/s, g 9.81, depth 100 m).
#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct ocean_spectrum_params {
    pub amplitude: f32,
    pub wind_speed: f32,
    pub gravity: f32,
    pub depth: f32,
}

extern "C" {
    pub fn ocean_create(out: *mut *mut ocean_ctx, cuda_device: c_int, resolution: u32, domain_size: f32, n_tiles: u32) -> c_int;
    pub fn ocean_create_ex(out: *mut *mut ocean_ctx, cfg: *const ocean_config) -> c_int;
    pub fn ocean_destroy(ctx: *mut ocean_ctx);
    pub fn ocean_set_spectrum(ctx: *mut ocean_ctx, tile: u32, h0_xy: *const f32, omega: *const f32) -> c_int;
    pub fn ocean_set_spectrum_device(ctx: *mut ocean_ctx, tile: u32, d_h0_xy: *const f32, d_omega: *const f32) -> c_int;
    pub fn ocean_load_bincode(ctx: *mut ocean_ctx, tile: u32, omega_path: *const c_char, spectrum_path: *const c_char) -> c_int;
    pub fn ocean_generate_spectrum(ctx: *mut ocean_ctx, tile: u32, seed: u64, stream_id: u32, params: *const ocean_spectrum_params, h_words: *mut u32) -> c_int;
    pub fn ocean_get_spectrum(ctx: *mut ocean_ctx, tile: u32, h0_xy: *mut f32, omega: *mut f32) -> c_int;
    pub fn ocean_update(ctx: *mut ocean_ctx, time: f32) -> c_int;
    pub fn ocean_update_tiles(ctx: *mut ocean_ctx, time: f32, first_tile: u32, count: u32) -> c_int;
    pub fn ocean_update_graph(ctx: *mut ocean_ctx, time: f32, first_tile: u32, count: u32) -> c_int;
    pub fn ocean_update_overlapped(ctx: *mut ocean_ctx, time: f32, first_tile: u32, count: u32) -> c_int;
    pub fn ocean_join(ctx: *mut ocean_ctx) -> c_int;
    pub fn ocean_update_sequence(ctx: *mut ocean_ctx, t0: f32, dt: f32, n_frames: u32) -> c_int;
    pub fn ocean_update_sequence_checksums(ctx: *mut ocean_ctx, t0: f32, dt: f32, n_frames: u32, h_sums: *mut u64) -> c_int;
    pub fn ocean_output_checksums(ctx: *mut ocean_ctx, h_sums: *mut u64) -> c_int;
    pub fn ocean_set_output_device(ctx: *mut ocean_ctx, tile: u32, d_rgba: *mut f32, row_pitch_bytes: usize) -> c_int;
    pub fn ocean_displace_grid(ctx: *mut ocean_ctx, tile: u32, grid: u32, offset_x: f32, offset_z: f32, h_pos_world: *mut f32) -> c_int;
    pub fn ocean_displace_grid_device(ctx: *mut ocean_ctx, tile: u32, grid: u32, offset_x: f32, offset_z: f32, d_pos_world: *mut f32) -> c_int;
    pub fn ocean_download_all_async(ctx: *mut ocean_ctx, h_rgba_all: *mut f32) -> c_int;
    pub fn ocean_download_fence(ctx: *mut ocean_ctx, lag: u32) -> c_int;
    pub fn ocean_compute_normals(ctx: *mut ocean_ctx, first_tile: u32, count: u32) -> c_int;
    pub fn ocean_normals_device(ctx: *mut ocean_ctx, tile: u32, d_nrm: *mut *const f32) -> c_int;
    pub fn ocean_download_normals(ctx: *mut ocean_ctx, tile: u32, h_nrm: *mut f32) -> c_int;
    pub fn ocean_profile_update(ctx: *mut ocean_ctx, time: f32, stage_ms: *mut f32, capacity: u32, n_stages: *mut u32) -> c_int;
    pub fn ocean_output_device(ctx: *mut ocean_ctx, tile: u32, d_rgba: *mut *const f32) -> c_int;
    pub fn ocean_download(ctx: *mut ocean_ctx, tile: u32, h_rgba: *mut f32) -> c_int;
    pub fn ocean_download_async(ctx: *mut ocean_ctx, tile: u32, h_rgba: *mut f32) -> c_int;
    pub fn ocean_sync(ctx: *mut ocean_ctx) -> c_int;
    pub fn ocean_debug_spectra(ctx: *mut ocean_ctx, tile: u32, h: *mut f32, dx: *mut f32, dz: *mut f32) -> c_int;
    pub fn ocean_get_locals(ctx: *const ocean_ctx, p: *mut PropagateLocals, c: *mut CorrectionLocals) -> c_int;
    pub fn ocean_resolution(ctx: *const ocean_ctx) -> u32;
    pub fn ocean_n_tiles(ctx: *const ocean_ctx) -> u32;
    pub fn ocean_launch_count(ctx: *const ocean_ctx) -> u64;
    pub fn ocean_stream(ctx: *const ocean_ctx) -> *mut c_void;
    pub fn ocean_algorithmic_bytes_per_update(ctx: *const ocean_ctx) -> u64;
    pub fn ocean_last_error(ctx: *const ocean_ctx) -> *const c_char;
    pub fn ocean_status_string(status: c_int) -> *const c_char;
    pub fn ocean_abi_version() -> u32;
}

/// The reference returns `Result<_, Box<dyn Error>>` from `init` (src/fft.rs:19, src/ocean.rs:25,194).
#[derive(Debug)]
pub struct OceanError {
    pub status: i32,
    pub message: String,
}

impl fmt::Display for OceanError {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        write!(f, "{} [status {}]", self.message, self.status)
    }
}
impl std::error::Error for OceanError {}

/// `Ocean::new / update / output / read_back`: the facade BASELINE.json's north_star names.
pub struct Ocean {
    ctx: *mut ocean_ctx,
    resolution: u32,
}

impl Ocean {
    fn err(ctx: *const ocean_ctx, status: c_int) -> OceanError {
        let message = unsafe { CStr::from_ptr(ocean_last_error(ctx)).to_string_lossy().into_owned() };
        OceanError { status, message }
    }

    /// `omega`: N*N, `spectrum`: N*N `[re, im]` -- what `bincode::deserialize` yields for
    /// data/omega.bin and data/spectrum.bin (src/render.rs:769-771, 808-810).
    pub fn new(resolution: u32, domain_size: f32, omega: &[f32], spectrum: &[[f32; 2]]) -> Result<Ocean, OceanError> {
        let n2 = (resolution as usize) * (resolution as usize);
        if omega.len() != n2 || spectrum.len() != n2 {
            return Err(OceanError { status: -1, message: "omega/spectrum must hold resolution^2 elements".into() });
        }
        let mut ctx: *mut ocean_ctx = ptr::null_mut();
        let rc = unsafe { ocean_create(&mut ctx, 0, resolution, domain_size, 1) };
        if rc != 0 {
            return Err(Self::err(ptr::null(), rc));
        }
        let rc = unsafe { ocean_set_spectrum(ctx, 0, spectrum.as_ptr() as *const f32, omega.as_ptr()) };
        if rc != 0 {
            let e = Self::err(ctx, rc);
            unsafe { ocean_destroy(ctx) };
            return Err(e);
        }
        Ok(Ocean { ctx, resolution })
    }

    /// Enqueue one frame (asynchronous): PropagateLocals{time, resolution, domain_size} +
    /// propagate -> 2-D inverse FFT of (dx, height, dz) -> correction.
    pub fn update(&mut self, time: f32) -> Result<(), OceanError> {
        let rc = unsafe { ocean_update(self.ctx, time) };
        if rc != 0 { Err(Self::err(self.ctx, rc)) } else { Ok(()) }
    }

    /// `update` with frames in flight (src/lib.rs:86): consecutive calls alternate between two internal lanes, so the
    /// row kernel of one frame runs beside the column kernel of the previous one. Same maps, bit for bit; every
    /// other call (read_back, output, ...) first orders the context's stream behind both lanes.
    pub fn update_overlapped(&mut self, time: f32) -> Result<(), OceanError> {
        let rc = unsafe { ocean_update_overlapped(self.ctx, time, 0, 1) };
        if rc != 0 { Err(Self::err(self.ctx, rc)) } else { Ok(()) }
    }

    /// Device pointer to N*N RGBA32F texels (dx, height, dz, 0), row-major [y][x]; stable for the
    /// lifetime of `self`. Import it as the displacement texture (external memory), or `read_back`.
    pub fn output(&self) -> Result<*const [f32; 4], OceanError> {
        let mut p: *const f32 = ptr::null();
        let rc = unsafe { ocean_output_device(self.ctx, 0, &mut p) };
        if rc != 0 { Err(Self::err(self.ctx, rc)) } else { Ok(p as *const [f32; 4]) }
    }

    pub fn read_back(&self, dst: &mut [[f32; 4]]) -> Result<(), OceanError> {
        assert_eq!(dst.len(), (self.resolution as usize).pow(2));
        let rc = unsafe { ocean_download(self.ctx, 0, dst.as_mut_ptr() as *mut f32) };
        if rc != 0 { Err(Self::err(self.ctx, rc)) } else { Ok(()) }
    }

    /// Renderer interop (src/render.rs:820-869, :939): have the kernels write the displacement map straight into
    /// device memory the renderer owns -- the linear image / buffer it exported with VK_KHR_external_memory_fd and
    /// imported with cudaImportExternalMemory. `row_pitch_bytes` = the image's row pitch (0 = dense N*16).
    pub unsafe fn set_output(&mut self, d_rgba: *mut [f32; 4], row_pitch_bytes: usize) -> Result<(), OceanError> {
        let rc = ocean_set_output_device(self.ctx, 0, d_rgba as *mut f32, row_pitch_bytes);
        if rc != 0 { Err(Self::err(self.ctx, rc)) } else { Ok(()) }
    }

    /// p_PosWorld of shader/ocean.vert:21-25 for the `grid` x `grid` vertex patch of src/render.rs:498-506 with the
    /// patch offset of :540-551 (HALF_RESOLUTION = 128).
    pub fn displace_grid(&self, grid: u32, offset: [f32; 2], dst: &mut [[f32; 3]]) -> Result<(), OceanError> {
        assert_eq!(dst.len(), (grid as usize).pow(2));
        let rc = unsafe { ocean_displace_grid(self.ctx, 0, grid, offset[0], offset[1], dst.as_mut_ptr() as *mut f32) };
        if rc != 0 { Err(Self::err(self.ctx, rc)) } else { Ok(()) }
    }

    pub fn sync(&self) -> Result<(), OceanError> {
        let rc = unsafe { ocean_sync(self.ctx) };
        if rc != 0 { Err(Self::err(self.ctx, rc)) } else { Ok(()) }
    }
}

impl Drop for Ocean {
    /// `destroy(self, device)` of the operator holders (src/fft.rs:102, src/ocean.rs:170,321).
    fn drop(&mut self) {
        unsafe { ocean_destroy(self.ctx) };
    }
}
