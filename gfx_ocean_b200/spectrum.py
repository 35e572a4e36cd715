"""Seeded synthetic inputs for grids the reference does not ship (N != 512).

The reference only carries a 512x512 ``data/omega.bin`` / ``data/spectrum.bin`` pair
(loaded at /root/reference/src/render.rs:769-771, 808-810); the program that made them
is not in the repository. SURVEY.md section 8a6/8d fitted their structure:

    omega[j, i]  = fl32(sqrt(g k tanh(k d))),  g = 9.81, d = 100,
                   k = 2 pi |(i - N/2 - 0.5, j - N/2 - 0.5)| / L        (half-sample grid)
    h0[j, i]     = (xi_r + i xi_i) sqrt(P(k) / 2),
                   P(k) = A exp(-1/(k l)^2) / k^4 (khat . what)^2,  what = (1, 0),
                   l = V^2 / g, V = 30 m/s, x0.07 where khat . what < 0,
                   A such that max|h0| = 1,  xi ~ N(0, 1) from default_rng(1234 + tile).

The hot path has no data-dependent control flow, so the distribution only matters for
the dynamic range seen by the parity tests.
"""
from __future__ import annotations

import numpy as np


def dispersion(n: int, domain_size: float = 1000.0, g: float = 9.81, depth: float = 100.0) -> np.ndarray:
    """Finite-depth dispersion table omega[N, N] (float32)."""
    c = np.arange(n, dtype=np.float64) - n / 2 - 0.5
    kx, ky = np.meshgrid(c, c, indexing="xy")
    k = 2.0 * np.pi * np.hypot(kx, ky) / domain_size
    return np.sqrt(g * k * np.tanh(k * depth)).astype(np.float32)


def phillips_h0(n: int, tile: int = 0, domain_size: float = 1000.0, wind_speed: float = 30.0,
                g: float = 9.81, seed_base: int = 1234) -> np.ndarray:
    """Initial spectrum h0[N, N, 2] (float32, re/im interleaved), max|h0| = 1."""
    c = np.arange(n, dtype=np.float64) - n / 2 - 0.5
    kx, ky = np.meshgrid(c, c, indexing="xy")
    kx = 2.0 * np.pi * kx / domain_size
    ky = 2.0 * np.pi * ky / domain_size
    k = np.hypot(kx, ky)                      # never 0 on the half-sample grid
    ell = wind_speed * wind_speed / g
    cosf = kx / k                             # khat . (1, 0)
    p = np.exp(-1.0 / (k * ell) ** 2) / k ** 4 * cosf * cosf
    p = np.where(cosf < 0.0, p * 0.07, p)
    rng = np.random.default_rng(seed_base + int(tile))
    xi = rng.standard_normal((n, n, 2))
    h0 = xi * np.sqrt(p / 2.0)[..., None]
    h0 /= np.hypot(h0[..., 0], h0[..., 1]).max()
    return h0.astype(np.float32)


def synthetic_tile(n: int, tile: int = 0, domain_size: float = 1000.0):
    """(h0[N,N,2] float32, omega[N,N] float32) for one tile."""
    return phillips_h0(n, tile, domain_size), dispersion(n, domain_size)
