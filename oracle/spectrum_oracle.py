"""TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the seeded spectrum generator (`ocean_generate_spectrum`). The reference
ships only the generator's OUTPUTS (data/omega.bin, data/spectrum.bin, loaded at src/render.rs:769-771,808-810);
their structure was fitted in SURVEY.md 8a6 (finite-depth dispersion on the half-sample grid, directional
Phillips-like spectrum), which is what this restates -- the random stream is ours: Philox-4x32-10 (Salmon, Moraes,
Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11), counter = (point index, stream id, 0, 0),
key = (seed lo, seed hi); known-answer vectors of the paper / Random123 are checked in tests/test_spectrum_gen.py.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox-4x32 with 10 rounds. Counters: uint32 arrays (or scalars); key: two python ints."""
    c = [np.asarray(v, np.uint64) & MASK for v in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return [v.astype(np.uint32) for v in c]


def generate_spectrum_np(n, seed, stream_id, domain_size=1000.0, amplitude=1.0, wind_speed=30.0, gravity=9.81, depth=100.0):
    """-> (h0[N, N, 2] float64, omega[N, N] float64, words[N, N, 4] uint32); f64 arithmetic on f32-rounded parameters."""
    idx = np.arange(n * n, dtype=np.uint32)
    r = philox4x32_10(idx, np.uint32(stream_id), 0, 0, seed & 0xFFFFFFFF, seed >> 32)
    words = np.stack(r, -1).reshape(n, n, 4)
    f = lambda v: float(np.float32(v))          # noqa: E731  the kernel receives f32 parameters
    L, A, V, g, d = f(domain_size), f(amplitude), f(wind_speed), f(gravity), f(depth)
    two_pi = float(np.float32(6.283185307179586))
    c = np.arange(n, dtype=np.float64) - 0.5 * n - 0.5
    kx, ky = np.meshgrid(two_pi * c / L, two_pi * c / L, indexing="xy")
    k2 = kx * kx + ky * ky
    k = np.sqrt(k2)
    omega = np.sqrt(g * k * np.tanh(k * d))
    ell = V * V / g
    cw = kx / k
    p = A * np.exp(-1.0 / (k2 * ell * ell)) / (k2 * k2) * cw * cw
    p = np.where(cw < 0.0, p * float(np.float32(0.07)), p)
    u1 = ((r[0] >> np.uint32(8)).astype(np.float64) + 0.5) / 16777216.0
    u2 = ((r[1] >> np.uint32(8)).astype(np.float64) + 0.5) / 16777216.0
    rad = (np.sqrt(-2.0 * np.log(u1)) * np.sqrt(0.5 * p.reshape(-1)))
    ang = two_pi * u2
    h0 = np.stack([rad * np.cos(ang), rad * np.sin(ang)], -1).reshape(n, n, 2)
    return h0, omega, words
