"""TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Python face of the CPU oracle for the gfx-ocean per-frame path
(propagate -> row iFFT x3 -> col iFFT x3 -> correction):

* ``COracle``     ctypes binding of ``oracle/libocean_oracle.so`` (ocean_oracle.c):
                  the literal f32 restatement and the parity-defining f64 one.
* ``*_np``        an independent numpy twin (vectorised, ``numpy.fft.ifft2 * N^2``
                  for the transform) used to cross-check the C restatement.

Reference files restated (relative to /root/reference): shader/propagate.comp:42-72,
shader/fft_row.comp:25-63, shader/fft_col.comp:44-63, shader/correction.comp:24-35,
src/render.rs:1122-1287.

PARITY PINNED BY EXECUTION OF THE REFERENCE'S OWN SHADER BINARIES: the reference holds no
golden vectors / tests for this path, so oracle/spv_exec.py interprets the SPIR-V modules the
reference ships and dispatches (shader/spv/*.spv) on the reference-owned inputs; the resulting
fixtures (tests/golden/spv_512.npz, generator tests/golden/make_spv_golden.py) agree with this
restatement to <= 6e-7 (tests/test_spv_pin.py; bar 1e-5). Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libocean_oracle.so")

PI32 = np.float32(3.1415926)  # propagate.comp:6 / fft_row.comp:5


def build(force: bool = False) -> str:
    """Compile the C oracle in place (``make -C oracle``)."""
    if force or not os.path.exists(_LIB) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB)
        for f in ("ocean_oracle.c", "ocean_oracle_impl.h", "Makefile")
    ):
        env = {k: v for k, v in os.environ.items() if k not in ("CC", "CFLAGS")}
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB


_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


class COracle:
    """ctypes binding of the C restatement. ``prec`` is 'f32' (literal) or 'f64'."""

    def __init__(self) -> None:
        self.lib = C.CDLL(build())
        for suf, rp in (("f32", _f32p), ("f64", _f64p)):
            f = getattr(self.lib, f"oracle_propagate_{suf}")
            f.argtypes = [_f32p, _f32p, C.c_float, C.c_int32, C.c_float, rp, rp, rp]
            f.restype = None
            for nm in ("fft_row", "fft_col"):
                f = getattr(self.lib, f"oracle_{nm}_{suf}")
                f.argtypes = [rp, C.c_uint32]
                f.restype = None
            f = getattr(self.lib, f"oracle_correction_{suf}")
            f.argtypes = [rp, rp, rp, C.c_uint32, rp]
            f.restype = None
            f = getattr(self.lib, f"oracle_normal_map_{suf}")
            f.argtypes = [rp, C.c_uint32, rp]
            f.restype = None
            f = getattr(self.lib, f"oracle_frame_{suf}")
            f.argtypes = [_f32p, _f32p, C.c_float, C.c_uint32, C.c_float, rp]
            f.restype = C.c_int
        self.lib.oracle_num_threads.restype = C.c_int
        self.lib.oracle_set_num_threads.argtypes = [C.c_int]
        self.lib.oracle_read_bincode.argtypes = [C.c_char_p, C.c_uint32, _f32p, C.c_uint64]
        self.lib.oracle_read_bincode.restype = C.c_int64

    @staticmethod
    def _dt(prec: str):
        return {"f32": np.float32, "f64": np.float64}[prec]

    def num_threads(self) -> int:
        return int(self.lib.oracle_num_threads())

    def set_num_threads(self, n: int) -> None:
        self.lib.oracle_set_num_threads(int(n))

    def propagate(self, h0, omega, time, n, domain_size=1000.0, prec="f64"):
        """-> (height_spec, disp_x_spec, disp_z_spec), each complex [N, N]."""
        dt = self._dt(prec)
        h0 = np.ascontiguousarray(h0, np.float32).reshape(n * n * 2)
        omega = np.ascontiguousarray(omega, np.float32).reshape(n * n)
        outs = [np.empty(n * n * 2, dt) for _ in range(3)]
        getattr(self.lib, f"oracle_propagate_{prec}")(h0, omega, float(time), int(n),
                                                      float(domain_size), *outs)
        ct = np.complex64 if prec == "f32" else np.complex128
        return tuple(o.view(ct).reshape(n, n) for o in outs)

    def fft_row(self, data, prec="f64"):
        return self._fft("fft_row", data, prec)

    def fft_col(self, data, prec="f64"):
        return self._fft("fft_col", data, prec)

    def _fft(self, name, data, prec):
        ct = np.complex64 if prec == "f32" else np.complex128
        a = np.array(data, dtype=ct, order="C", copy=True)
        n = a.shape[0]
        assert a.shape == (n, n)
        getattr(self.lib, f"oracle_{name}_{prec}")(a.view(self._dt(prec)).reshape(-1), n)
        return a

    def frame(self, h0, omega, time, n, domain_size=1000.0, prec="f64"):
        """One full frame -> out[N, N, 4] = (dx, height, dz, 0) * (-1 if (x+y) even else 1)."""
        dt = self._dt(prec)
        h0 = np.ascontiguousarray(h0, np.float32).reshape(n * n * 2)
        omega = np.ascontiguousarray(omega, np.float32).reshape(n * n)
        out = np.empty(n * n * 4, dt)
        rc = getattr(self.lib, f"oracle_frame_{prec}")(h0, omega, float(time), int(n),
                                                       float(domain_size), out)
        if rc != 0:
            raise ValueError(f"oracle_frame_{prec} failed rc={rc} (N must be a power of two)")
        return out.reshape(n, n, 4)

    def normal_map(self, disp, prec="f64"):
        """shader/ocean.frag:50-66 at texel centres: disp[N, N, 4] -> normals[N, N, 4] = (N.x, N.y, N.z, 0)."""
        dt = self._dt(prec)
        d = np.ascontiguousarray(disp, dt)
        n = d.shape[0]
        out = np.empty(n * n * 4, dt)
        getattr(self.lib, f"oracle_normal_map_{prec}")(d.reshape(-1), n, out)
        return out.reshape(n, n, 4)

    def read_bincode(self, path: str, elem_floats: int, capacity: int) -> np.ndarray:
        dst = np.empty(capacity * elem_floats, np.float32)
        cnt = self.lib.oracle_read_bincode(path.encode(), elem_floats, dst, capacity)
        if cnt < 0:
            raise IOError(f"oracle_read_bincode({path}) rc={cnt}")
        return dst[: cnt * elem_floats]


# ---------------------------------------------------------------------------
# numpy twin (independent formulation; f64 with the two fp32 rounding points)
# ---------------------------------------------------------------------------

def wave_vector_np(n: int, domain_size: float = 1000.0):
    """propagate.comp:45-46,50-53 -> (kx[N], ky[N]) as fp32 arrays.

    ``uint x = 2*gid - resolution - 1`` wraps mod 2^32 and is converted with an
    UNSIGNED int->float conversion (OpConvertUToF in the shipped SPIR-V)."""
    g = np.arange(n, dtype=np.int64)
    xu = (2 * g - n - 1) % (1 << 32)            # u32 wrap
    xf = xu.astype(np.float32)                  # round-to-nearest u32 -> f32
    k = (PI32 * xf).astype(np.float32) / np.float32(domain_size)
    k = k.astype(np.float32)
    return k, k.copy()


def propagate_np(h0, omega, time, n, domain_size=1000.0):
    """shader/propagate.comp:42-72 in f64 (phi and k rounded to fp32 first)."""
    h0 = np.asarray(h0, np.float32).reshape(n * n, 2)
    h0c = h0[:, 0].astype(np.float64) + 1j * h0[:, 1].astype(np.float64)
    om = np.asarray(omega, np.float32).reshape(n * n)
    phi = (om * np.float32(time)).astype(np.float32).astype(np.float64)   # :55, fp32 product
    e = np.cos(phi) + 1j * np.sin(phi)
    h = h0c * e + h0c[::-1] * np.conj(e)        # :48 index_neg == N*N-1-index: reversed array
    h = h.reshape(n, n)
    kx, ky = wave_vector_np(n, domain_size)
    KX = kx.astype(np.float64)[None, :].repeat(n, 0)
    KY = ky.astype(np.float64)[:, None].repeat(n, 1)
    ln = np.sqrt(KX * KX + KY * KY)
    ok = ln > 1.0e-10
    nx = np.where(ok, KX / np.where(ok, ln, 1.0), 0.0)
    nz = np.where(ok, KY / np.where(ok, ln, 1.0), 0.0)
    return h, (-1j * nx) * h, (-1j * nz) * h    # :69-71


def correction_np(height, disp_x, disp_z):
    """shader/correction.comp:24-35."""
    n = height.shape[0]
    yy, xx = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    sign = np.where((xx + yy) % 2 == 0, -1.0, 1.0)
    out = np.zeros((n, n, 4), np.float64)
    out[..., 0] = disp_x.real * sign
    out[..., 1] = height.real * sign
    out[..., 2] = disp_z.real * sign
    return out


def frame_np(h0, omega, time, n, domain_size=1000.0):
    """Exact-DFT twin: row pass then column pass == N^2 * ifft2 (unnormalised inverse DFT,
    fft_row.comp:25-40; true pi instead of the shader's 3.1415926 -- the gap between the
    two is measured in tests/test_oracle.py and is ~4e-7 of the field maximum)."""
    h, dx, dz = propagate_np(h0, omega, time, n, domain_size)
    s = float(n * n)
    return correction_np(np.fft.ifft2(h) * s, np.fft.ifft2(dx) * s, np.fft.ifft2(dz) * s)


def normal_map_np(disp):
    """Independent numpy formulation of shader/ocean.frag:50-66 at texel centres (Tile wrap)."""
    d = np.asarray(disp, np.float64)[..., 0]
    n = d.shape[0]
    diff, hs = 2.0 / n, 180.0
    dxv = (np.roll(d, -1, axis=1) - np.roll(d, 1, axis=1)) / hs
    dzv = (np.roll(d, -1, axis=0) - np.roll(d, 1, axis=0)) / hs
    na = np.stack([np.full_like(d, -diff), dxv, np.zeros_like(d)], -1)
    nb = np.stack([np.zeros_like(d), dzv, np.full_like(d, diff)], -1)
    na /= np.linalg.norm(na, axis=-1, keepdims=True)
    nb /= np.linalg.norm(nb, axis=-1, keepdims=True)
    c = np.cross(na, nb)
    c /= np.linalg.norm(c, axis=-1, keepdims=True)
    return np.concatenate([c, np.zeros(d.shape + (1,))], -1)


def sample_linear_tile_np(disp, u, v):
    """Vulkan linear filtering with REPEAT addressing (the reference's sampler: Filter::Linear,
    WrapMode::Tile, src/render.rs:397-398) of a [H, W, C] map at normalised coordinates (u, v)."""
    d = np.asarray(disp, np.float64)
    h, w = d.shape[:2]
    x = np.asarray(u, np.float64) * w - 0.5
    y = np.asarray(v, np.float64) * h - 0.5
    i0, j0 = np.floor(x), np.floor(y)
    a, b = (x - i0)[..., None], (y - j0)[..., None]
    i0, j0 = i0.astype(np.int64), j0.astype(np.int64)
    i1, j1 = (i0 + 1) % w, (j0 + 1) % h
    i0, j0 = i0 % w, j0 % h
    return (d[j0, i0] * (1 - a) + d[j0, i1] * a) * (1 - b) + (d[j1, i0] * (1 - a) + d[j1, i1] * a) * b


def displace_grid_np(disp, grid, offset=(0.0, 0.0)):
    """shader/ocean.vert:21-25,29 for the reference's vertex grid (src/render.rs:498-506: a_Pos = (x, 0, z),
    a_Uv = (x, z) / (grid - 1) computed in f32): p_PosWorld = a_Pos + (d.x/3.5, d.y/3, d.z/3.5) + (off.x, 0, off.y)
    with d = texture(displacement_map, a_Uv). -> [grid, grid, 3] float64."""
    g = np.arange(grid, dtype=np.float32)
    uvs = (g / np.float32(grid - 1)).astype(np.float64)
    x, z = np.meshgrid(g.astype(np.float64), g.astype(np.float64), indexing="xy")
    u, v = np.meshgrid(uvs, uvs, indexing="xy")
    d = sample_linear_tile_np(disp, u, v)
    return np.stack([x + d[..., 0] / 3.5 + offset[0], d[..., 1] / 3.0, z + d[..., 2] / 3.5 + offset[1]], -1)


def stockham_line_np(x: np.ndarray, pi=float(PI32)) -> np.ndarray:
    """fft_row.comp:25-40,51-59 for one line, vectorised over the N/2 'threads'."""
    n = x.shape[-1]
    half = n // 2
    stages = int(np.log2(n))
    src = np.array(x, np.complex128)
    idx = np.arange(half)
    for i in range(stages):
        bs = 1 << i
        k = idx & (bs - 1)
        w = np.exp(1j * (pi * k / bs))
        t = src[..., idx + half] * w
        dst = np.empty_like(src)
        dest = (idx << 1) - k
        dst[..., dest] = src[..., idx] + t
        dst[..., dest + bs] = src[..., idx] - t
        src = dst
    return src


def max_rel_err(out, ref):
    """SURVEY 8c tolerance metric: per channel max|out-ref| / max|ref| for (dx, h, dz)."""
    out = np.asarray(out, np.float64)
    ref = np.asarray(ref, np.float64)
    errs = []
    for c in range(3):
        den = np.abs(ref[..., c]).max()
        errs.append(float(np.abs(out[..., c] - ref[..., c]).max() / (den if den > 0 else 1.0)))
    return errs
