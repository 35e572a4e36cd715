"""Host-side mirror of the reference's operator interface for the per-frame compute path.

The reference records the path inline in ``Renderer::render`` (src/render.rs:1101-1310) through
``Propagation<B>``/``Fft<B>``/``Correction<B>`` (src/ocean.rs, src/fft.rs: ``init``/``destroy``);
``Ocean`` is the facade BASELINE.json's north_star names (``Ocean::new / update / output``), a
thin wrapper over the C ABI (include/ocean_b200.h) with the same names, argument meaning and
error behaviour as bindings/rust/ocean.rs.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import (CorrectionLocals, OceanConfig, OceanError, PropagateLocals, SpectrumParams,  # noqa: F401
                   PIPELINE_FUSED, PIPELINE_LITERAL, FLAG_DOUBLE_BUFFER_OUTPUT, FLAG_DX_PLANE)

# src/render.rs:42-46
WORKGROUP_SIZE = 16
WORKGROUP_NUM = 32
RESOLUTION = WORKGROUP_SIZE * WORKGROUP_NUM
DOMAIN_SIZE = 1000.0


class Ocean:
    """One context = one CUDA device + one stream + ``n_tiles`` independent oceans."""

    def __init__(self, resolution: int = RESOLUTION, domain_size: float = DOMAIN_SIZE, n_tiles: int = 1,
                 device: int = 0, pipeline: int = PIPELINE_FUSED, stream: int | None = None, flags: int = 0):
        self._lib = _lib.load()
        self._ctx = C.c_void_p()
        cfg = OceanConfig(_lib.ABI_VERSION, device, resolution, domain_size, n_tiles, pipeline,
                          C.c_void_p(stream) if stream else None, flags)
        rc = self._lib.ocean_create_ex(C.byref(self._ctx), C.byref(cfg))
        if rc != _lib.OK:
            raise OceanError(rc, self._lib.ocean_last_error(None).decode() or self._lib.ocean_status_string(rc).decode())
        self.resolution, self.domain_size, self.n_tiles, self.pipeline = resolution, domain_size, n_tiles, pipeline

    # -- Ocean::new(resolution, domain_size, &omega, &spectrum) -----------------------------
    @classmethod
    def new(cls, resolution: int, domain_size: float, omega, spectrum, **kw) -> "Ocean":
        o = cls(resolution, domain_size, **kw)
        for t in range(o.n_tiles):
            o.set_spectrum(t, spectrum, omega)
        return o

    @classmethod
    def from_bincode(cls, omega_path: str, spectrum_path: str, resolution: int = RESOLUTION,
                     domain_size: float = DOMAIN_SIZE, **kw) -> "Ocean":
        o = cls(resolution, domain_size, **kw)
        for t in range(o.n_tiles):
            o.load_bincode(t, omega_path, spectrum_path)
        return o

    def _check(self, rc: int) -> None:
        if rc != _lib.OK:
            raise OceanError(rc, self._lib.ocean_last_error(self._ctx).decode() or
                             self._lib.ocean_status_string(rc).decode())

    def set_spectrum(self, tile: int, h0, omega) -> None:
        n = self.resolution
        h0 = np.ascontiguousarray(h0, np.float32)
        omega = np.ascontiguousarray(omega, np.float32)
        if h0.size != n * n * 2 or omega.size != n * n:
            raise OceanError(_lib.ERR_INVALID_ARG, f"spectrum must hold {n}x{n}x2 floats and omega {n}x{n}")
        self._check(self._lib.ocean_set_spectrum(self._ctx, tile, h0.ctypes.data, omega.ctypes.data))

    def set_spectrum_device(self, tile: int, d_h0: int, d_omega: int) -> None:
        self._check(self._lib.ocean_set_spectrum_device(self._ctx, tile, d_h0, d_omega))

    def generate_spectrum(self, tile: int, seed: int, stream_id: int | None = None, params: SpectrumParams | None = None,
                          want_words: bool = False):
        """Seeded Phillips / finite-depth inputs generated on the device (no upload). -> Philox words or None."""
        n = self.resolution
        words = np.empty((n, n, 4), np.uint32) if want_words else None
        self._check(self._lib.ocean_generate_spectrum(self._ctx, tile, seed, tile if stream_id is None else stream_id,
                                                      C.byref(params) if params is not None else None,
                                                      words.ctypes.data if want_words else None))
        return words

    def get_spectrum(self, tile: int = 0):
        n = self.resolution
        h0, om = np.empty((n, n, 2), np.float32), np.empty((n, n), np.float32)
        self._check(self._lib.ocean_get_spectrum(self._ctx, tile, h0.ctypes.data, om.ctypes.data))
        return h0, om

    def load_bincode(self, tile: int, omega_path: str, spectrum_path: str) -> None:
        self._check(self._lib.ocean_load_bincode(self._ctx, tile, omega_path.encode(), spectrum_path.encode()))

    # -- Ocean::update(&mut self, time: f32) ---------------------------------------------------
    def update(self, time: float) -> None:
        self._check(self._lib.ocean_update(self._ctx, time))

    def update_tiles(self, time: float, first_tile: int, count: int) -> None:
        self._check(self._lib.ocean_update_tiles(self._ctx, time, first_tile, count))

    def update_graph(self, time: float, first_tile: int = 0, count: int | None = None) -> None:
        """update_tiles through a recorded CUDA graph (replayed with `time` patched)."""
        self._check(self._lib.ocean_update_graph(self._ctx, time, first_tile, self.n_tiles - first_tile if count is None else count))

    def update_overlapped(self, time: float, first_tile: int = 0, count: int | None = None) -> None:
        """update_tiles on alternating internal lanes: frames of different tiles overlap on the device."""
        self._check(self._lib.ocean_update_overlapped(self._ctx, time, first_tile, self.n_tiles - first_tile if count is None else count))

    def join(self) -> None:
        """Order the context's stream behind the frames in flight on the lanes (no host blocking)."""
        self._check(self._lib.ocean_join(self._ctx))

    def update_sequence(self, t0: float, dt: float, n_frames: int) -> None:
        self._check(self._lib.ocean_update_sequence(self._ctx, t0, dt, n_frames))

    def profile_update(self, time: float) -> list[float]:
        """One update with CUDA events around every kernel; per-kernel durations in ms."""
        buf = (C.c_float * 8)()
        cnt = C.c_uint32()
        self._check(self._lib.ocean_profile_update(self._ctx, time, buf, 8, C.byref(cnt)))
        return [float(buf[i]) for i in range(cnt.value)]

    def update_sequence_checksums(self, t0: float, dt: float, n_frames: int) -> np.ndarray:
        """Back-to-back frames; -> uint64[n_frames, n_tiles] checksums of every frame's maps."""
        sums = np.zeros((n_frames, self.n_tiles), np.uint64)
        self._check(self._lib.ocean_update_sequence_checksums(self._ctx, t0, dt, n_frames, sums.ctypes.data))
        return sums

    def output_checksums(self) -> np.ndarray:
        sums = np.zeros(self.n_tiles, np.uint64)
        self._check(self._lib.ocean_output_checksums(self._ctx, sums.ctypes.data))
        return sums

    def set_output_device(self, tile: int, d_rgba: int | None, row_pitch_bytes: int = 0) -> None:
        """Renderer interop: write the tile's map into a caller-provided device allocation (None: own buffer)."""
        self._check(self._lib.ocean_set_output_device(self._ctx, tile, d_rgba, row_pitch_bytes))

    def read_back_all_async(self, host_ptr: int) -> None:
        self._check(self._lib.ocean_download_all_async(self._ctx, host_ptr))

    def download_fence(self, lag: int = 0) -> None:
        self._check(self._lib.ocean_download_fence(self._ctx, lag))

    def sync(self) -> None:
        self._check(self._lib.ocean_sync(self._ctx))

    # -- Ocean::output(&self) -> *const [f32; 4] -------------------------------------------------
    def output(self, tile: int = 0) -> int:
        """Device address of the tile's N*N RGBA32F texels (dx, height, dz, 0)."""
        p = C.c_void_p()
        self._check(self._lib.ocean_output_device(self._ctx, tile, C.byref(p)))
        return p.value

    # -- Ocean::read_back(&self, &mut [[f32; 4]]) --------------------------------------------------
    def read_back(self, tile: int = 0, out: np.ndarray | None = None) -> np.ndarray:
        n = self.resolution
        if out is None:
            out = np.empty((n, n, 4), np.float32)
        assert out.dtype == np.float32 and out.size == n * n * 4 and out.flags.c_contiguous
        self._check(self._lib.ocean_download(self._ctx, tile, out.ctypes.data))
        return out

    def read_back_async(self, tile: int, host_ptr: int) -> None:
        self._check(self._lib.ocean_download_async(self._ctx, tile, host_ptr))

    # -- consumer step: the normal map of shader/ocean.frag:50-66 ----------------------------------
    def compute_normals(self, first_tile: int = 0, count: int | None = None) -> None:
        self._check(self._lib.ocean_compute_normals(self._ctx, first_tile, self.n_tiles - first_tile if count is None else count))

    def normals(self, tile: int = 0) -> int:
        p = C.c_void_p()
        self._check(self._lib.ocean_normals_device(self._ctx, tile, C.byref(p)))
        return p.value

    def read_back_normals(self, tile: int = 0) -> np.ndarray:
        n = self.resolution
        out = np.empty((n, n, 4), np.float32)
        self._check(self._lib.ocean_download_normals(self._ctx, tile, out.ctypes.data))
        return out

    def displace_grid(self, grid: int = 128, offset=(0.0, 0.0), tile: int = 0) -> np.ndarray:
        """shader/ocean.vert:21-25 for the reference's vertex grid -> p_PosWorld[grid, grid, 3]."""
        out = np.empty((grid, grid, 3), np.float32)
        self._check(self._lib.ocean_displace_grid(self._ctx, tile, grid, offset[0], offset[1], out.ctypes.data))
        return out

    def debug_spectra(self, tile: int = 0):
        n = self.resolution
        outs = [np.empty((n, n, 2), np.float32) for _ in range(3)]
        self._check(self._lib.ocean_debug_spectra(self._ctx, tile, *[o.ctypes.data for o in outs]))
        return tuple(o[..., 0] + 1j * o[..., 1] for o in outs)

    def locals(self):
        p, c = PropagateLocals(), CorrectionLocals()
        self._check(self._lib.ocean_get_locals(self._ctx, C.byref(p), C.byref(c)))
        return p, c

    @property
    def launch_count(self) -> int:
        return int(self._lib.ocean_launch_count(self._ctx))

    @property
    def stream(self) -> int:
        return int(self._lib.ocean_stream(self._ctx) or 0)

    @property
    def algorithmic_bytes_per_update(self) -> int:
        return int(self._lib.ocean_algorithmic_bytes_per_update(self._ctx))

    def destroy(self) -> None:
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.ocean_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.destroy()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
