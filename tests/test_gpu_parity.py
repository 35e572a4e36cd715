"""Parity of the CUDA path (through the C ABI) against the CPU oracle. Run on the B200 box:
    python -m pytest tests -m gpu

Bar (BASELINE.json north_star / SURVEY.md 8c): per channel max|out - ref| / max|ref| <= 1e-5
against the f64 oracle; the 4th channel is exactly +0.0.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from gfx_ocean_b200 import Ocean, OceanError, PIPELINE_FUSED, PIPELINE_LITERAL, _lib
from gfx_ocean_b200.spectrum import synthetic_tile
from oracle.ocean_oracle import max_rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-5            # the parity bar
OMEGA = os.path.join(GOLDEN, "ref_data", "omega.bin")
SPECTRUM = os.path.join(GOLDEN, "ref_data", "spectrum.bin")


def check_w_channel(out):
    assert np.all(out[..., 3] == 0.0) and not np.any(np.signbit(out[..., 3]))


@pytest.fixture(scope="module")
def shipped_fused():
    with Ocean.from_bincode(OMEGA, SPECTRUM, 512, 1000.0, pipeline=PIPELINE_FUSED) as o:
        yield o


@pytest.fixture(scope="module")
def shipped_literal():
    with Ocean.from_bincode(OMEGA, SPECTRUM, 512, 1000.0, pipeline=PIPELINE_LITERAL) as o:
        yield o


@pytest.mark.parametrize("t", [0.0, 1.0, 37.5, 600.0, 25000.0])   # 25000 s: phases > 1e5 rad (slow sincos path)
@pytest.mark.parametrize("which", ["fused", "literal"])
def test_shipped_data_matches_oracle(which, t, shipped_fused, shipped_literal, oracle, ref_inputs):
    """BASELINE.json config 2: the reference's own data/omega.bin + data/spectrum.bin at N=512."""
    o = shipped_fused if which == "fused" else shipped_literal
    sp, om = ref_inputs
    o.update(t)
    out = o.read_back()
    ref = oracle.frame(sp, om, t, 512, prec="f64")
    errs = max_rel_err(out, ref)
    print(f"{which} t={t}: max rel err (dx,h,dz) = {errs}")
    assert max(errs) <= TOL
    check_w_channel(out)
    p, c = o.locals()
    assert (p.time, p.resolution, p.domain_size, c.resolution) == (np.float32(t), 512, 1000.0, 512)


@pytest.mark.parametrize("ti", [0, 1, 2, 3])
def test_shipped_data_matches_committed_golden(ti, shipped_fused, golden_512):
    t = float(golden_512["times"][ti])
    shipped_fused.update(t)
    out = shipped_fused.read_back()
    py, px = golden_512["probe_y"], golden_512["probe_x"]
    scale = golden_512[f"max_abs_{ti}"]
    err = np.abs(out[py, px, :3] - golden_512[f"probe_out_{ti}"][:, :3]).max(axis=0) / scale
    assert err.max() <= TOL
    np.testing.assert_allclose(np.abs(out[..., :3].astype(np.float64)).sum(axis=(0, 1)), golden_512[f"sum_abs_{ti}"], rtol=2e-6)


def test_debug_spectra_match_oracle_propagate(shipped_fused, oracle, ref_inputs):
    """Post-propagate spectra incl. the u32-wrap quirk cases (propagate.comp:45-46)."""
    sp, om = ref_inputs
    for t in (0.0, 600.0):
        shipped_fused.update(t)
        h, dx, dz = shipped_fused.debug_spectra()
        rh, rdx, rdz = oracle.propagate(sp, om, t, 512, prec="f64")
        for a, b in ((h, rh), (dx, rdx), (dz, rdz)):
            assert np.abs(a - b).max() / np.abs(b).max() <= 2e-6
    np.testing.assert_allclose(dx[300, 200], -1j * h[300, 200], rtol=1e-6)
    assert abs(dz[300, 200]) < 1e-7 * abs(h[300, 200])          # khat.z ~ 2e-8 where only gx is wrapped


def test_fused_and_literal_agree(shipped_fused, shipped_literal):
    shipped_fused.update(12.25)
    shipped_literal.update(12.25)
    a, b = shipped_fused.read_back(), shipped_literal.read_back()
    assert max(max_rel_err(a, b)) <= 5e-6


@pytest.mark.parametrize("n", [64, 128, 256, 1024, 2048])
@pytest.mark.parametrize("t", [0.0, 1.0, 37.5])
def test_synthetic_tiles_match_oracle(n, t, oracle):
    """BASELINE.json configs 3/4: seeded synthetic grids at other resolutions
    (2048: three-pass lines, strips of 4 columns)."""
    h0, w = synthetic_tile(n, tile=3)
    with Ocean.new(n, 1000.0, w, h0) as o:
        o.update(t)
        out = o.read_back()
    ref = oracle.frame(h0, w, t, n, prec="f64")
    errs = max_rel_err(out, ref)
    print(f"N={n} t={t}: max rel err = {errs}")
    assert max(errs) <= TOL
    check_w_channel(out)


@pytest.mark.parametrize("t", [-12.5, 3.0e4, 1.0e6])
def test_negative_and_huge_times(t, oracle):
    """Negative time, and phases far beyond 1e5 rad where every point takes the Payne-Hanek sincos path."""
    h0, w = synthetic_tile(256, 7)
    with Ocean.new(256, 1000.0, w, h0) as o:
        o.update(t)
        out = o.read_back()
    assert max(max_rel_err(out, oracle.frame(h0, w, t, 256, prec="f64"))) <= TOL


def test_golden_synth_1024(golden_synth):
    h0, w = synthetic_tile(1024, 0)
    with Ocean.new(1024, 1000.0, w, h0) as o:
        for i, t in enumerate(golden_synth["n1024_times"]):
            o.update(float(t))
            out = o.read_back()
            py, px = golden_synth["n1024_probe_y"], golden_synth["n1024_probe_x"]
            err = np.abs(out[py, px, :3] - golden_synth[f"n1024_probe_out_{i}"][:, :3]).max(axis=0) / golden_synth[f"n1024_max_abs_{i}"]
            assert err.max() <= TOL


def test_golden_synth_2048_fused_and_literal(golden_synth):
    """BASELINE.json config 4 (2048 x 2048): both pipelines against the committed golden probes."""
    h0, w = synthetic_tile(2048, 0)
    py, px = golden_synth["n2048_probe_y"], golden_synth["n2048_probe_x"]
    for pipeline in (PIPELINE_FUSED, PIPELINE_LITERAL):
        with Ocean.new(2048, 1000.0, w, h0, pipeline=pipeline) as o:
            o.update(float(golden_synth["n2048_times"][0]))
            out = o.read_back()
        err = np.abs(out[py, px, :3] - golden_synth["n2048_probe_out_0"][:, :3]).max(axis=0) / golden_synth["n2048_max_abs_0"]
        assert err.max() <= TOL
        check_w_channel(out)


def test_tiles_are_independent_and_bit_identical_to_single_tile_runs():
    """the result of tile i must not depend on how tiles are batched."""
    n, t = 512, 3.5
    tiles = [synthetic_tile(n, i) for i in range(3)]
    with Ocean(n, n_tiles=3) as batch:
        for i, (h0, w) in enumerate(tiles):
            batch.set_spectrum(i, h0, w)
        batch.update(t)
        outs = [batch.read_back(i) for i in range(3)]
        batch.update_tiles(t, 1, 1)            # partial update leaves the result unchanged
        assert np.array_equal(batch.read_back(1), outs[1])
    assert not np.array_equal(outs[0], outs[1])
    for i, (h0, w) in enumerate(tiles):
        with Ocean.new(n, 1000.0, w, h0) as single:
            single.update(t)
            assert np.array_equal(single.read_back(), outs[i])


def test_repeatable_and_stateless_across_frames(shipped_fused):
    shipped_fused.update(5.0)
    a = shipped_fused.read_back().copy()
    shipped_fused.update(77.0)
    shipped_fused.update(5.0)
    assert np.array_equal(shipped_fused.read_back(), a)


def test_linearity_and_real_transform_roundtrip_1024(oracle):
    """Size-independent properties at the benchmark size: the frame is linear in h0, and the
    forward DFT of the sign-corrected height field returns the Hermitian part of the propagated
    height spectrum (encode -> decode round trip)."""
    n, t = 1024, 2.0
    h0, w = synthetic_tile(n, 1)
    g0, _ = synthetic_tile(n, 2)
    with Ocean(n) as o:
        o.set_spectrum(0, h0, w); o.update(t); a = o.read_back().astype(np.float64)
        o.set_spectrum(0, g0, w); o.update(t); b = o.read_back().astype(np.float64)
        o.set_spectrum(0, (2 * h0 - 3 * g0).astype(np.float32), w); o.update(t); c = o.read_back().astype(np.float64)
        h_spec, _, _ = o.debug_spectra()
    assert np.abs(c - (2 * a - 3 * b)).max() <= 2e-5 * np.abs(c).max()
    yy, xx = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    sign = np.where((xx + yy) % 2 == 0, -1.0, 1.0)
    hs = (2 * h0 - 3 * g0)  # spectra of the last set_spectrum
    spec_back = np.fft.fft2(c[..., 1] * sign) / (n * n)
    neg = (-np.arange(n)) % n
    herm = 0.5 * (h_spec + np.conj(h_spec[neg][:, neg]))
    assert np.abs(spec_back - herm).max() <= 2e-5 * np.abs(herm).max()
    del hs


def test_normal_map_matches_oracle(shipped_fused, oracle):
    """Consumer step: shader/ocean.frag:50-66 at texel centres, from the GPU's own map."""
    shipped_fused.update(2.5)
    disp = shipped_fused.read_back()
    shipped_fused.compute_normals()
    nrm = shipped_fused.read_back_normals()
    ref = oracle.normal_map(disp.astype(np.float64), prec="f64")
    assert np.abs(nrm - ref).max() <= 1e-5
    np.testing.assert_allclose(np.linalg.norm(nrm[..., :3], axis=-1), 1.0, atol=1e-5)
    assert np.all(nrm[..., 3] == 0.0)
    with Ocean(256) as o:
        with pytest.raises(OceanError) as e:
            o.compute_normals()
        assert e.value.status == _lib.ERR_NOT_READY


def test_error_behaviour(tmp_path):
    with Ocean(256) as o:
        with pytest.raises(OceanError) as e:
            o.update(0.0)                                 # no spectrum yet
        assert e.value.status == _lib.ERR_NOT_READY
        with pytest.raises(OceanError) as e:
            o.set_spectrum(1, np.zeros((256, 256, 2), np.float32), np.zeros((256, 256), np.float32))
        assert e.value.status == _lib.ERR_INVALID_ARG
        with pytest.raises(OceanError) as e:
            o.load_bincode(0, OMEGA, SPECTRUM)            # 512x512 files into a 256 context
        assert e.value.status == _lib.ERR_IO
        with pytest.raises(OceanError) as e:
            o.load_bincode(0, str(tmp_path / "missing.bin"), SPECTRUM)
        assert e.value.status == _lib.ERR_IO
    with pytest.raises(OceanError) as e:
        Ocean(512, device=99)
    assert e.value.status == _lib.ERR_NO_DEVICE


def test_unsupported_resolutions_fail_loudly():
    """The fused pipeline covers N in {64 .. 2048}; anything else is an error, never a silent fallback."""
    for n in (8, 32, 4096):
        with pytest.raises(OceanError) as e:
            Ocean(n)
        assert e.value.status == _lib.ERR_UNSUPPORTED
    with pytest.raises(OceanError) as e:
        Ocean(4096, pipeline=PIPELINE_LITERAL)
    assert e.value.status == _lib.ERR_UNSUPPORTED


def test_zero_spectrum_gives_zero_field():
    with Ocean.new(256, 1000.0, np.ones((256, 256), np.float32), np.zeros((256, 256, 2), np.float32)) as o:
        o.update(9.0)
        assert not o.read_back().any()


def test_two_contexts_interleaved_and_threaded(oracle):
    """Contexts are independent (own stream, own buffers): interleaving them, or driving them from two host
    threads, gives the same frames as running them alone (SURVEY.md 8b threading row)."""
    import threading
    n = 256
    data = [synthetic_tile(n, i) for i in (5, 6)]
    refs = [oracle.frame(h0, w, 4.0, n, prec="f64") for h0, w in data]
    ctxs = [Ocean.new(n, 1000.0, w, h0) for h0, w in data]
    try:
        for _ in range(3):                       # interleaved on one thread
            for c in ctxs:
                c.update(4.0)
        outs = [c.read_back() for c in ctxs]
        for o, r in zip(outs, refs):
            assert max(max_rel_err(o, r)) <= TOL
        results = [None, None]

        def work(i):
            for k in range(20):
                ctxs[i].update(float(k))
            ctxs[i].update(4.0)
            results[i] = ctxs[i].read_back()
        th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
        [t.start() for t in th]
        [t.join() for t in th]
        for o, r in zip(results, outs):
            assert np.array_equal(o, r)
    finally:
        for c in ctxs:
            c.destroy()


def test_external_stream_is_used():
    """ocean_config.stream: the context enqueues on the caller's stream (here a torch stream)."""
    import torch
    s = torch.cuda.Stream()
    h0, w = synthetic_tile(256, 1)
    with Ocean(256, stream=s.cuda_stream) as o:
        assert o.stream == s.cuda_stream
        o.set_spectrum(0, h0, w)
        ev = torch.cuda.Event()
        o.update(1.0)
        ev.record(s)
        ev.synchronize()
        a = o.read_back()
    with Ocean.new(256, 1000.0, w, h0) as o2:
        o2.update(1.0)
        assert np.array_equal(o2.read_back(), a)


def test_update_sequence_equals_the_last_update(shipped_fused):
    shipped_fused.update_sequence(0.0, 0.5, 9)            # t = 0, 0.5, ..., 4.0
    a = shipped_fused.read_back().copy()
    shipped_fused.update(4.0)
    assert np.array_equal(shipped_fused.read_back(), a)
    assert shipped_fused.locals()[0].time == 4.0


def test_create_destroy_does_not_leak_device_memory():
    import torch
    torch.cuda.synchronize()
    h0, w = synthetic_tile(512, 0)
    with Ocean.new(512, 1000.0, w, h0) as o:              # warm up lazy allocations of the runtime
        o.update(0.0); o.compute_normals(); o.debug_spectra(); o.sync()
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(20):
        with Ocean.new(512, 1000.0, w, h0, n_tiles=2) as o:
            o.update(1.0); o.compute_normals(); o.debug_spectra(); o.sync()
    free1 = torch.cuda.mem_get_info()[0]
    assert abs(free0 - free1) <= 8 << 20                  # allocator granularity, not a per-context leak


def test_soak_many_frames_stays_deterministic(shipped_fused):
    shipped_fused.update(2.0)
    a = shipped_fused.read_back().copy()
    shipped_fused.update_sequence(0.0, 0.016, 3000)
    shipped_fused.update(2.0)
    assert np.array_equal(shipped_fused.read_back(), a)


def test_fuzz_sizes_tiles_ranges_and_times(oracle):
    """Seeded sweep over resolution, tile count, partial tile ranges and time (both pipelines where legal)."""
    rng = np.random.default_rng(20261017)
    for case in range(14):
        n = int(rng.choice([64, 128, 256, 512, 1024]))
        n_tiles = int(rng.integers(1, 5))
        first = int(rng.integers(0, n_tiles))
        count = int(rng.integers(1, n_tiles - first + 1))
        t = float(np.float32(rng.uniform(-100.0, 2000.0)))
        pipeline = PIPELINE_LITERAL if case % 5 == 4 else PIPELINE_FUSED
        tiles = [synthetic_tile(n, 100 + case * 8 + i) for i in range(n_tiles)]
        with Ocean(n, 1000.0, n_tiles=n_tiles, pipeline=pipeline) as o:
            for i, (h0, w) in enumerate(tiles):
                o.set_spectrum(i, h0, w)
            o.update(0.25)                                   # every tile gets a frame ...
            o.update_tiles(t, first, count)                  # ... then only [first, first+count) moves to t
            for i, (h0, w) in enumerate(tiles):
                want_t = t if first <= i < first + count else 0.25
                err = max(max_rel_err(o.read_back(i), oracle.frame(h0, w, want_t, n, prec="f64")))
                assert err <= TOL, (case, n, n_tiles, first, count, t, i, err)


def test_launch_accounting(shipped_fused, shipped_literal):
    a = shipped_fused.launch_count; shipped_fused.update(1.0); assert shipped_fused.launch_count - a == 2
    b = shipped_literal.launch_count; shipped_literal.update(1.0); assert shipped_literal.launch_count - b == 8
    assert shipped_fused.algorithmic_bytes_per_update == 76 * 512 * 512
