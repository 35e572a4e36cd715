"""Property tests (hypothesis) of the oracle on small grids: the structural facts the CUDA path relies on
(linearity, t-independence at omega = 0, Hermitian fold, sign pattern) hold for arbitrary inputs."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle.ocean_oracle import COracle, frame_np, max_rel_err, propagate_np

ORACLE = COracle()
sizes = st.sampled_from([8, 16, 32, 64])
seeds = st.integers(0, 2**31 - 1)
times = st.floats(-50.0, 5000.0, allow_nan=False, width=32)


def rand_inputs(n, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, n, 2)).astype(np.float32), rng.uniform(0.1, 5.0, (n, n)).astype(np.float32)


@settings(max_examples=25, deadline=None)
@given(sizes, seeds, times)
def test_c_oracle_equals_exact_dft_twin(n, seed, t):
    h0, w = rand_inputs(n, seed)
    a = ORACLE.frame(h0, w, t, n, prec="f64")
    assert max(max_rel_err(frame_np(h0, w, t, n), a)) < 5e-6          # literal pi32 vs true pi
    assert np.all(a[..., 3] == 0.0)


@settings(max_examples=20, deadline=None)
@given(sizes, seeds, times, st.floats(-3, 3, allow_nan=False), st.floats(-3, 3, allow_nan=False))
def test_frame_is_linear_in_the_spectrum(n, seed, t, alpha, beta):
    h0, w = rand_inputs(n, seed)
    g0, _ = rand_inputs(n, seed + 1)
    mix = (np.float32(alpha) * h0 + np.float32(beta) * g0).astype(np.float32)
    lhs = ORACLE.frame(mix, w, t, n, prec="f64")
    rhs = np.float64(np.float32(alpha)) * ORACLE.frame(h0, w, t, n, prec="f64") + \
        np.float64(np.float32(beta)) * ORACLE.frame(g0, w, t, n, prec="f64")
    assert np.abs(lhs - rhs).max() <= 1e-5 * max(1.0, np.abs(rhs).max())   # fp32 rounding of the mixed input


@settings(max_examples=15, deadline=None)
@given(sizes, seeds, times)
def test_zero_dispersion_makes_the_frame_time_independent(n, seed, t):
    h0, _ = rand_inputs(n, seed)
    w = np.zeros((n, n), np.float32)
    np.testing.assert_array_equal(ORACLE.frame(h0, w, t, n, prec="f64"), ORACLE.frame(h0, w, 0.0, n, prec="f64"))


@settings(max_examples=15, deadline=None)
@given(sizes, seeds, times)
def test_real_part_equals_transform_of_hermitian_fold(n, seed, t):
    """The identity the fused kernels are built on: Re ifft2(F) = ifft2((F + conj F(-k)) / 2)."""
    h0, w = rand_inputs(n, seed)
    h, dx, dz = propagate_np(h0, w, t, n)
    neg = (-np.arange(n)) % n
    for f in (h, dx, dz):
        fold = 0.5 * (f + np.conj(f[neg][:, neg]))
        full = np.fft.ifft2(f)
        np.testing.assert_allclose(np.fft.ifft2(fold).real, full.real, atol=1e-12 * max(1.0, np.abs(full).max()))
        assert np.abs(np.fft.ifft2(fold).imag).max() <= 1e-12 * max(1.0, np.abs(full).max())


@settings(max_examples=10, deadline=None)
@given(sizes, seeds)
def test_sign_pattern_is_a_half_period_shift(n, seed):
    """correction.comp:29: multiplying by -(-1)^(x+y) equals shifting the spectrum by N/2 in both axes."""
    h0, w = rand_inputs(n, seed)
    h, _, _ = propagate_np(h0, w, 1.0, n)
    out = ORACLE.frame(h0, w, 1.0, n, prec="f64")[..., 1]
    shifted = np.fft.ifft2(np.roll(h, (n // 2, n // 2), axis=(0, 1))) * n * n
    np.testing.assert_allclose(out, -shifted.real, atol=2e-6 * max(1.0, np.abs(out).max()))
