"""CPU tests of the oracle itself (no GPU): the C restatement against the committed golden
vectors, against the independent numpy twin, and against SURVEY.md 8c's probe values."""
import hashlib
import os

import numpy as np
import pytest

from oracle.ocean_oracle import (PI32, frame_np, max_rel_err, propagate_np,
                                 stockham_line_np, wave_vector_np)
from gfx_ocean_b200.spectrum import synthetic_tile

from conftest import GOLDEN


def test_input_fixtures_are_the_reference_files(golden_512):
    for f, sha in zip(("omega.bin", "spectrum.bin"), golden_512["input_sha256"]):
        got = hashlib.sha256(open(os.path.join(GOLDEN, "ref_data", f), "rb").read()).hexdigest()
        assert got == str(sha)
    assert str(golden_512["input_sha256"][0]).startswith("5e45d2654091ec55")
    assert str(golden_512["input_sha256"][1]).startswith("4b6d47959643ab4b")


def test_input_probe_values(ref_inputs):
    sp, om = ref_inputs                      # SURVEY 8c "inputs"
    assert om[0, 0] == pytest.approx(4.7286148, rel=1e-7)
    assert om[256, 256] == pytest.approx(0.1348451, rel=1e-6)
    assert sp[257, 260] == pytest.approx([-0.9845946, 0.6827147], rel=1e-6)
    assert np.all(sp[256:258, 256:258] == 0.0)


def test_pi_literal_is_not_fp32_pi():
    assert PI32 == np.float32(3.14159250259)
    assert PI32 != np.float32(np.pi)


def test_wave_vector_uint_wrap_quirk():
    """propagate.comp:45-46,50-53: u32 wrap + unsigned conversion for gx <= N/2."""
    kx, _ = wave_vector_np(512, 1000.0)
    assert kx[0] == np.float32(np.float32(PI32 * np.float32(4294966783)) / np.float32(1000))
    assert np.all(kx[:257] > 1.3e7)          # wrapped
    assert kx[257] == np.float32(PI32 * np.float32(1.0)) / np.float32(1000)
    assert np.all(np.diff(kx[257:]) > 0) and kx[511] < 1.61


@pytest.mark.parametrize("ti", [0, 1, 2, 3])
def test_c_oracle_matches_golden_512(oracle, ref_inputs, golden_512, ti):
    sp, om = ref_inputs
    t = float(golden_512["times"][ti])
    out = oracle.frame(sp, om, t, 512, prec="f64")
    py, px = golden_512["probe_y"], golden_512["probe_x"]
    np.testing.assert_allclose(out[py, px, :], golden_512[f"probe_out_{ti}"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(np.abs(out[..., :3]).sum(axis=(0, 1)), golden_512[f"sum_abs_{ti}"], rtol=1e-12)
    assert np.all(out[..., 3] == 0.0) and not np.any(np.signbit(out[..., 3]))


def test_survey_probe_values(oracle, ref_inputs):
    """SURVEY.md 8c probes (surveyor's exact-DFT oracle, so agreement is ~1e-6 of max)."""
    sp, om = ref_inputs
    want = {0.0: {(0, 0): (-3.082524, -1.326607, -0.321310), (17, 300): (0.275572, -0.677730, 1.497564),
                  (511, 511): (-3.017880, -1.183992, -0.466976)},
            1.0: {(0, 0): (-1.814249, -1.339759, -0.750584), (17, 300): (-0.649752, -1.637966, 0.004145)},
            37.5: {(0, 0): (-0.993247, 0.994430, 1.886184), (17, 300): (-0.071840, -2.282176, 1.024886)},
            600.0: {(0, 0): (-1.849456, -1.372785, -0.029064), (17, 300): (0.654969, -7.021063, 0.095426)}}
    sum_h = {0.0: 616059.936, 1.0: 622090.412, 37.5: 718689.483, 600.0: 699149.760}
    for t, pts in want.items():
        out = oracle.frame(sp, om, t, 512, prec="f64")
        for (y, x), v in pts.items():
            np.testing.assert_allclose(out[y, x, :3], v, atol=3e-5)
        assert np.abs(out[..., 1]).sum() == pytest.approx(sum_h[t], rel=2e-6)
    out0 = oracle.frame(sp, om, 0.0, 512, prec="f64")
    assert np.abs(out0[..., 1]).max() == pytest.approx(11.03755, rel=2e-6)


def test_propagate_quirk_cases(oracle, ref_inputs, golden_512):
    """dx = -i*h where only gx is wrapped (khat = (1, ~0)); dz = -i*h where only gy is."""
    sp, om = ref_inputs
    h, dx, dz = oracle.propagate(sp, om, 0.0, 512, prec="f64")
    np.testing.assert_allclose(h[300, 200], 3.6304e-4 - 5.7278e-4j, atol=2e-8)
    np.testing.assert_allclose(dx[300, 200], -1j * h[300, 200], rtol=1e-6)
    assert abs(dz[300, 200]) < 1e-9
    np.testing.assert_allclose(dz[100, 400], -1j * h[100, 400], rtol=1e-6)
    assert abs(dx[100, 400]) < 1e-9
    # both wrapped -> khat = (0.7071, 0.7071)
    np.testing.assert_allclose(dx[10, 20], -1j * h[10, 20] * np.sqrt(0.5), rtol=1e-6)
    qy, qx = golden_512["spec_y"], golden_512["spec_x"]
    np.testing.assert_allclose(h[qy, qx], golden_512["spec_h"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(dx[qy, qx], golden_512["spec_dx"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(dz[qy, qx], golden_512["spec_dz"], rtol=0, atol=1e-14)


def test_propagate_c_vs_numpy_twin(oracle, ref_inputs):
    sp, om = ref_inputs
    for t in (0.0, 37.5, 3600.0):
        hc = oracle.propagate(sp, om, t, 512, prec="f64")
        hn = propagate_np(sp, om, t, 512)
        for a, b in zip(hc, hn):
            assert np.abs(a - b).max() <= 1e-13 * max(1.0, np.abs(b).max())


def test_domain_size_does_not_change_output(oracle, ref_inputs):
    """khat is scale invariant: any L > 0 gives the same frame to ~1 ulp of fp32 k."""
    sp, om = ref_inputs
    a = oracle.frame(sp, om, 1.0, 512, domain_size=1000.0, prec="f64")
    b = oracle.frame(sp, om, 1.0, 512, domain_size=250.0, prec="f64")
    assert max(max_rel_err(b, a)) < 1e-6


@pytest.mark.parametrize("n", [2, 8, 64, 512, 1024, 2048])
def test_stockham_is_unnormalised_inverse_dft(oracle, n):
    """fft_row.comp:25-40 == N * ifft along x (natural order in/out); fft_col likewise along y.
    Literal pi32 twiddles vs exact pi differ by ~4e-7 of the maximum."""
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    r = oracle.fft_row(a, prec="f64")
    np.testing.assert_allclose(r, np.fft.ifft(a, axis=1) * n, atol=2e-6 * np.abs(r).max())
    c = oracle.fft_col(a, prec="f64")
    np.testing.assert_allclose(c, np.fft.ifft(a, axis=0) * n, atol=2e-6 * np.abs(c).max())
    # the numpy restatement of the same recurrence agrees to f64 rounding
    np.testing.assert_allclose(r, stockham_line_np(a), atol=1e-12 * np.abs(r).max())
    if n <= 64:   # with true pi the recurrence IS the exact DFT
        np.testing.assert_allclose(stockham_line_np(a, pi=np.pi), np.fft.ifft(a, axis=1) * n, atol=1e-12 * n)


def test_fft_linearity_and_impulse(oracle):
    n = 256
    rng = np.random.default_rng(3)
    a = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    b = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    lhs = oracle.fft_row(2.0 * a - 3.0 * b)
    rhs = 2.0 * oracle.fft_row(a) - 3.0 * oracle.fft_row(b)
    np.testing.assert_allclose(lhs, rhs, atol=1e-11 * np.abs(lhs).max())
    imp = np.zeros((n, n), complex); imp[:, 0] = 1.0
    np.testing.assert_allclose(oracle.fft_row(imp), np.ones((n, n)), atol=1e-13)


@pytest.mark.parametrize("t", [0.0, 1.0, 600.0])
def test_frame_f32_literal_and_exact_twin_within_tolerance(oracle, ref_inputs, t):
    """The literal fp32 restatement and the exact-pi numpy twin both sit ~1e-6 from the f64
    oracle: a 10x margin under the 1e-5 parity bar."""
    sp, om = ref_inputs
    ref = oracle.frame(sp, om, t, 512, prec="f64")
    assert max(max_rel_err(oracle.frame(sp, om, t, 512, prec="f32"), ref)) < 2e-6
    assert max(max_rel_err(frame_np(sp, om, t, 512), ref)) < 3e-6


def test_c_oracle_matches_golden_synth(oracle, golden_synth):
    for n in (1024, 2048):
        h0, w = synthetic_tile(n, 0)
        np.testing.assert_allclose([np.abs(h0.astype(np.float64)).sum(), w.astype(np.float64).sum()],
                                   golden_synth[f"n{n}_input_sums"], rtol=1e-9)
        py, px = golden_synth[f"n{n}_probe_y"], golden_synth[f"n{n}_probe_x"]
        for i, t in enumerate(golden_synth[f"n{n}_times"]):
            out = oracle.frame(h0, w, float(t), n, prec="f64")
            scale = golden_synth[f"n{n}_max_abs_{i}"].max()
            np.testing.assert_allclose(out[py, px, :], golden_synth[f"n{n}_probe_out_{i}"], rtol=0, atol=1e-9 * scale)
            np.testing.assert_allclose(np.abs(out[..., :3]).sum(axis=(0, 1)), golden_synth[f"n{n}_sum_abs_{i}"], rtol=1e-10)


def test_rejects_non_power_of_two(oracle):
    with pytest.raises(ValueError):
        oracle.frame(np.zeros((6, 6, 2), np.float32), np.zeros((6, 6), np.float32), 0.0, 6)


def test_smallest_grid(oracle):
    """N=2: one butterfly stage; compare with a hand-rolled DFT."""
    h0 = np.arange(8, dtype=np.float32).reshape(2, 2, 2) / 8
    w = np.array([[0.5, 1.0], [1.5, 2.0]], np.float32)
    out = oracle.frame(h0, w, 0.7, 2, prec="f64")
    assert max(max_rel_err(frame_np(h0, w, 0.7, 2), out)) < 1e-6


def test_normal_map_c_vs_numpy_twin(oracle, ref_inputs):
    """Consumer step (shader/ocean.frag:50-66 at texel centres): C restatement vs numpy formulation."""
    from oracle.ocean_oracle import normal_map_np
    sp, om = ref_inputs
    disp = oracle.frame(sp, om, 1.0, 512, prec="f64")
    nc = oracle.normal_map(disp, prec="f64")
    nn = normal_map_np(disp)
    np.testing.assert_allclose(nc, nn, atol=1e-13)
    np.testing.assert_allclose(np.linalg.norm(nc[..., :3], axis=-1), 1.0, atol=1e-12)
    assert np.all(nc[..., 3] == 0.0)
    # a flat map has the normal cross((-1,0,0),(0,0,1)) = (0, 1, 0)
    flat = oracle.normal_map(np.zeros((8, 8, 4)), prec="f64")
    np.testing.assert_allclose(flat[..., :3], np.broadcast_to([0.0, 1.0, 0.0], (8, 8, 3)), atol=1e-15)
    # fp32 literal restatement within 1e-6
    n32 = oracle.normal_map(disp.astype(np.float32), prec="f32")
    assert np.abs(n32 - nc).max() < 2e-6
