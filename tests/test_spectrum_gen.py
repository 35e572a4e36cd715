"""Spectrum generator: Philox-4x32-10 known answers on the CPU restatement, and the CUDA
generator against it (RNG words bit-exact, spectrum <= 1e-6 of its maximum)."""
import numpy as np
import pytest

from oracle.spectrum_oracle import generate_spectrum_np, philox4x32_10

# Random123's kat_vectors for philox4x32_10: (counter, key) -> output
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize("ctr,key,want", KAT)
def test_philox_known_answers(ctr, key, want):
    got = philox4x32_10(*ctr, *key)
    assert tuple(int(v) for v in got) == want


def test_generator_restatement_has_the_fitted_structure(ref_inputs):
    """Same dispersion table as the reference's data/omega.bin (fit of SURVEY.md 8a6) and unit-variance gaussians."""
    _, om_ref = ref_inputs
    h0, om, words = generate_spectrum_np(512, seed=1234, stream_id=0)
    assert np.abs(om - om_ref).max() <= 3e-5 * np.abs(om_ref).max()      # fits to 1.9e-5 (relative)
    assert words.shape == (512, 512, 4) and len(np.unique(words[..., 0])) > 260000
    h0b, _, _ = generate_spectrum_np(512, seed=1234, stream_id=1)
    assert np.abs(h0 - h0b).max() > 0          # another stream id: another tile


@pytest.mark.gpu
@pytest.mark.parametrize("n", [256, 1024])
def test_cuda_generator_matches_restatement(n):
    from gfx_ocean_b200 import Ocean, SpectrumParams
    seed = 0x1234_5678_9abc_def0
    p = SpectrumParams(3.5e-3, 31.0, 9.81, 80.0)
    with Ocean(n, 1000.0, n_tiles=2) as o:
        words = o.generate_spectrum(1, seed, stream_id=7, params=p, want_words=True)
        h0, om = o.get_spectrum(1)
        o.generate_spectrum(0, seed, stream_id=7, params=p)
        h0b, omb = o.get_spectrum(0)
    rh0, rom, rwords = generate_spectrum_np(n, seed, 7, amplitude=p.amplitude, wind_speed=p.wind_speed, gravity=p.gravity, depth=p.depth)
    np.testing.assert_array_equal(words, rwords)                       # the random stream: bit exact
    assert np.abs(om - rom).max() <= 1e-6 * np.abs(rom).max()
    assert np.abs(h0 - rh0).max() <= 1e-6 * np.abs(rh0).max()
    np.testing.assert_array_equal(h0, h0b)                            # same (seed, stream) -> same tile, any slot
    np.testing.assert_array_equal(om, omb)


@pytest.mark.gpu
def test_generated_tile_runs_through_the_path_and_matches_oracle(oracle):
    from gfx_ocean_b200 import Ocean
    from oracle.ocean_oracle import max_rel_err
    n = 512
    with Ocean(n, 1000.0) as o:
        o.generate_spectrum(0, seed=99)
        h0, om = o.get_spectrum(0)
        o.update(2.5)
        out = o.read_back()
    assert max(max_rel_err(out, oracle.frame(h0, om, 2.5, n, prec="f64"))) <= 1e-5
