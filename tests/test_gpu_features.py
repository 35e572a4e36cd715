"""Round-2 GPU tests: the benchmarked configuration, PDL on/off bit-identity of EVERY frame, output routing
(external / pitched buffers, double buffering), checksums / tile determinism, consumer kernels at all sizes."""
import os

import numpy as np
import pytest

from gfx_ocean_b200 import FLAG_DOUBLE_BUFFER_OUTPUT, FLAG_DX_PLANE, Ocean, OceanError, PIPELINE_LITERAL, _lib
from gfx_ocean_b200.spectrum import synthetic_tile
from oracle.ocean_oracle import displace_grid_np, max_rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


class env:
    """Set environment variables the library reads at context creation (OCEAN_B200_PDL / OCEAN_B200_ROWS)."""
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        for k, v in self.kv.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v

    def __exit__(self, *exc):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_benchmarked_configuration_every_tile_matches_oracle(oracle):
    """bench.py's workload: 1024^2 x 8 tiles in ONE ocean_update; every tile against the f64 oracle."""
    n, tiles, t = 1024, 8, 0.016 * 23
    data = [synthetic_tile(n, g) for g in range(tiles)]
    with Ocean(n, 1000.0, n_tiles=tiles) as o:
        for i, (h0, w) in enumerate(data):
            o.set_spectrum(i, h0, w)
        o.update(t)
        outs = [o.read_back(i) for i in range(tiles)]
        sums = o.output_checksums()
    for i, (h0, w) in enumerate(data):
        errs = max_rel_err(outs[i], oracle.frame(h0, w, t, n, prec="f64"))
        assert max(errs) <= TOL, (i, errs)
        assert np.all(outs[i][..., 3] == 0.0)
    assert len(set(int(s) for s in sums)) == tiles           # distinct tiles, distinct checksums


@pytest.mark.parametrize("n,tiles,frames", [(512, 1, 600), (1024, 1, 500), (1024, 3, 120), (256, 2, 300)])
def test_every_frame_is_bit_identical_with_and_without_pdl(n, tiles, frames):
    """The hazard programmatic dependent launch could open: k_rows of frame n+1 overwriting the intermediate while
    k_cols of frame n still reads it. The column kernel checksums what it stores, for EVERY frame of a back-to-back
    sequence; PDL on and PDL off must agree bit for bit, for the default row kernel and for its alternatives."""
    data = [synthetic_tile(n, g) for g in range(tiles)]
    sums = {}
    for name, kv in (("pdl1", dict(OCEAN_B200_PDL="1")), ("pdl0", dict(OCEAN_B200_PDL="0")),
                     ("staged_pdl1", dict(OCEAN_B200_PDL="1", OCEAN_B200_ROWS="staged")),
                     ("persistent_pdl1", dict(OCEAN_B200_PDL="1", OCEAN_B200_ROWS="persistent")),
                     ("fold_pdl0", dict(OCEAN_B200_PDL="0", OCEAN_B200_ROWS="fold"))):
        with env(**kv):
            with Ocean(n, 1000.0, n_tiles=tiles) as o:
                for i, (h0, w) in enumerate(data):
                    o.set_spectrum(i, h0, w)
                sums[name] = o.update_sequence_checksums(0.0, 0.016, frames)
                last = o.output_checksums()
        np.testing.assert_array_equal(sums[name][-1], last)       # in-kernel checksum == reduction kernel's
    np.testing.assert_array_equal(sums["pdl1"], sums["pdl0"])
    np.testing.assert_array_equal(sums["pdl1"], sums["staged_pdl1"])             # bulk-copy fed vs register-staged loads
    np.testing.assert_array_equal(sums["persistent_pdl1"], sums["fold_pdl0"])   # same arithmetic, two launch structures
    assert len(np.unique(sums["pdl0"][:, 0])) == frames            # the frames do differ from each other


@pytest.mark.parametrize("n", [256, 512, 1024, 2048])
def test_row_kernel_variants_agree_closely(n, oracle):
    """Staged rows (default, bulk-copy fed) vs fold at the source (persistent bulk-copy-fed / one unit per block):
    different summation order in the fold, same result to rounding; the two fold variants are bit-identical."""
    h0, w = synthetic_tile(n, 5)
    outs = []
    for kv in (dict(OCEAN_B200_ROWS="tma"), dict(OCEAN_B200_ROWS="persistent"), dict(OCEAN_B200_ROWS="fold")):
        with env(**kv):
            with Ocean.new(n, 1000.0, w, h0) as o:
                o.update(3.0)
                outs.append(o.read_back())
    assert max(max_rel_err(outs[0], outs[1])) <= 2e-6
    np.testing.assert_array_equal(outs[1], outs[2])
    assert max(max_rel_err(outs[1], oracle.frame(h0, w, 3.0, n, prec="f64"))) <= TOL


def test_tile_results_do_not_depend_on_batching_or_slot():
    """tile i's result is bit-identical whatever the tile count / position in the context."""
    n = 1024
    data = [synthetic_tile(n, g) for g in range(4)]
    with Ocean(n, 1000.0, n_tiles=4) as o:
        for i, (h0, w) in enumerate(data):
            o.set_spectrum(i, h0, w)
        o.update(1.25)
        batch = o.output_checksums()
    single = []
    for h0, w in data:
        with Ocean.new(n, 1000.0, w, h0) as o:
            o.update(1.25)
            single.append(int(o.output_checksums()[0]))
    assert [int(s) for s in batch] == single


@pytest.mark.parametrize("pipeline", ["fused", "literal"])
def test_external_pitched_output_buffer(pipeline):
    """Renderer interop, CUDA half: the kernels write into a caller-provided (pitched) allocation."""
    import torch
    n = 512
    h0, w = synthetic_tile(n, 1)
    kw = dict(pipeline=PIPELINE_LITERAL) if pipeline == "literal" else {}
    with Ocean.new(n, 1000.0, w, h0, **kw) as o:
        o.update(2.0)
        own = o.read_back()
        pitch_texels = n + 48
        ext = torch.full((n, pitch_texels, 4), -7.0, dtype=torch.float32, device="cuda")
        o.set_output_device(0, ext.data_ptr(), pitch_texels * 16)
        assert o.output() == ext.data_ptr()
        o.update(2.0)
        o.sync()
        got = ext.cpu().numpy()
        np.testing.assert_array_equal(got[:, :n], own)                # same bits, other destination
        assert np.all(got[:, n:] == -7.0)                            # the padding is never written
        np.testing.assert_array_equal(o.read_back(), own)            # download follows the routing (2-D copy)
        o.compute_normals()
        nrm_ext = o.read_back_normals()
        dense = torch.zeros((n, n, 4), dtype=torch.float32, device="cuda")
        o.set_output_device(0, dense.data_ptr(), 0)                   # dense external buffer: the product k_cols build
        o.update(2.0)
        o.sync()
        np.testing.assert_array_equal(dense.cpu().numpy(), own)
        o.set_output_device(0, None)                                  # back to the context's buffer
        o.update(2.0)
        np.testing.assert_array_equal(o.read_back(), own)
        o.compute_normals()
        np.testing.assert_array_equal(o.read_back_normals(), nrm_ext)
        with pytest.raises(OceanError):
            o.set_output_device(0, ext.data_ptr() + 4, 0)             # misaligned
        with pytest.raises(OceanError):
            o.set_output_device(0, ext.data_ptr(), n * 16 - 16)       # pitch too small


def test_double_buffered_readback_pipeline():
    import torch
    n, tiles, frames = 512, 3, 12
    data = [synthetic_tile(n, g) for g in range(tiles)]
    ref = []
    with Ocean(n, 1000.0, n_tiles=tiles) as o:
        for i, (h0, w) in enumerate(data):
            o.set_spectrum(i, h0, w)
        for f in range(frames):
            o.update(0.1 * f)
            ref.append(np.stack([o.read_back(i) for i in range(tiles)]))
    host = [torch.empty((tiles, n, n, 4), dtype=torch.float32, pin_memory=True) for _ in range(2)]
    with Ocean(n, 1000.0, n_tiles=tiles, flags=FLAG_DOUBLE_BUFFER_OUTPUT) as o:
        for i, (h0, w) in enumerate(data):
            o.set_spectrum(i, h0, w)
        for f in range(frames):
            o.update(0.1 * f)
            o.read_back_all_async(host[f % 2].data_ptr())
            o.download_fence(1)                                       # frame f-1 is on the host now
            if f >= 1:
                np.testing.assert_array_equal(host[(f - 1) % 2].numpy(), ref[f - 1])
        o.download_fence(0)
        np.testing.assert_array_equal(host[(frames - 1) % 2].numpy(), ref[-1])
        np.testing.assert_array_equal(o.read_back(1), ref[-1][1])     # per-tile download reads the current buffer
        with pytest.raises(OceanError):
            o.update_tiles(0.0, 0, 1)
        with pytest.raises(OceanError):
            o.set_output_device(0, 0)


@pytest.mark.parametrize("n", [64, 256, 1024, 2048])
def test_normal_map_all_sizes_and_tiles(n, oracle):
    tiles = 2 if n <= 1024 else 1
    data = [synthetic_tile(n, g + 20) for g in range(tiles)]
    with Ocean(n, 1000.0, n_tiles=tiles) as o:
        for i, (h0, w) in enumerate(data):
            o.set_spectrum(i, h0, w)
        o.update(0.75)
        o.compute_normals()
        for i in range(tiles):
            disp, nrm = o.read_back(i), o.read_back_normals(i)
            ref = oracle.normal_map(disp.astype(np.float64), prec="f64")
            assert np.abs(nrm - ref).max() <= 1e-5
            assert np.all(nrm[..., 3] == 0.0)


@pytest.mark.parametrize("n,grid", [(512, 128), (1024, 128), (256, 257)])
def test_displace_grid_matches_oracle(n, grid):
    h0, w = synthetic_tile(n, 4)
    with Ocean.new(n, 1000.0, w, h0) as o:
        o.update(1.5)
        disp = o.read_back()
        pw = o.displace_grid(grid, (127.0, -3.0))
    ref = displace_grid_np(disp, grid, (127.0, -3.0))
    assert np.abs(pw - ref).max() <= 1e-5 * np.abs(ref).max()


def test_graph_replay_is_bit_identical_to_plain_launches():
    n, tiles = 512, 3
    data = [synthetic_tile(n, g + 40) for g in range(tiles)]
    with Ocean(n, 1000.0, n_tiles=tiles) as o:
        for i, (h0, w) in enumerate(data):
            o.set_spectrum(i, h0, w)
        plain, graph = [], []
        for f in range(6):
            o.update(0.3 * f)
            plain.append(o.output_checksums())
        l0 = o.launch_count
        for f in range(6):
            o.update_graph(0.3 * f)
            graph.append(o.output_checksums())
        assert o.launch_count - l0 == 6 * (2 + tiles)          # 2 kernels per replayed frame (+ the checksum kernels)
        np.testing.assert_array_equal(np.array(plain), np.array(graph))
        o.update_tiles(0.9, 1, 2)
        a = o.output_checksums()
        o.update(0.0)
        o.update_graph(0.9, 1, 2)                               # another tile range: another recorded graph
        np.testing.assert_array_equal(o.output_checksums()[1:], a[1:])


@pytest.mark.parametrize("n,flags", [(1024, FLAG_DX_PLANE), (256, FLAG_DX_PLANE), (512, FLAG_DX_PLANE | FLAG_DOUBLE_BUFFER_OUTPUT)])
def test_dx_plane_normals_are_bit_identical(n, flags):
    """OCEAN_FLAG_DX_PLANE: the column kernel leaves a dense copy of channel .x; the normal map computed from it must
    equal the one gathered from the RGBA texels bit for bit, and the displacement map itself is unchanged."""
    tiles = 2
    data = [synthetic_tile(n, g + 60) for g in range(tiles)]
    res = []
    for fl in (0, flags):
        with Ocean(n, 1000.0, n_tiles=tiles, flags=fl) as o:
            for i, (h0, w) in enumerate(data):
                o.set_spectrum(i, h0, w)
            for t in (0.5, 1.5, 2.5):                      # three frames: both buffers of a double-buffered context
                o.update(t)
                o.compute_normals()
            res.append(([o.read_back(i) for i in range(tiles)], [o.read_back_normals(i) for i in range(tiles)]))
    for i in range(tiles):
        np.testing.assert_array_equal(res[0][0][i], res[1][0][i])
        np.testing.assert_array_equal(res[0][1][i], res[1][1][i])


def test_split_launches_are_bit_identical():
    """OCEAN_B200_BATCH splits an update into several kernel pairs that reuse one intermediate region; same bits."""
    n, tiles = 512, 5
    data = [synthetic_tile(n, g + 70) for g in range(tiles)]
    sums = []
    for kv in (dict(OCEAN_B200_BATCH=None), dict(OCEAN_B200_BATCH="2")):
        import subprocess, sys, json            # the knob is read once per process: run each setting in its own process
        code = ("import sys, json; sys.path.insert(0, %r); import numpy as np; from gfx_ocean_b200 import Ocean; "
                "from gfx_ocean_b200.spectrum import synthetic_tile; o = Ocean(%d, 1000.0, n_tiles=%d); "
                "[o.set_spectrum(i, *synthetic_tile(%d, i + 70)) for i in range(%d)]; "
                "s = o.update_sequence_checksums(0.0, 0.25, 6); print(json.dumps(s.tolist()))" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), n, tiles, n, tiles))
        env_ = dict(os.environ)
        env_.pop("OCEAN_B200_BATCH", None)
        if kv["OCEAN_B200_BATCH"]:
            env_["OCEAN_B200_BATCH"] = kv["OCEAN_B200_BATCH"]
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env_, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        sums.append(json.loads(r.stdout.strip().splitlines()[-1]))
    assert sums[0] == sums[1]
    del data


def test_overlapped_updates_match_plain_updates_bit_for_bit():
    """ocean_update_overlapped: frames of different tiles in flight on two lanes; same maps as ocean_update_tiles,
    also when ranges collide, when plain updates / uploads / read-backs are interleaved, and over many frames."""
    n, tiles = 512, 4
    data = [synthetic_tile(n, g + 80) for g in range(tiles)]
    rng = np.random.default_rng(5)
    plan = []                                        # (kind, time, first, count)
    for f in range(120):
        first = int(rng.integers(0, tiles))
        count = 1 if f % 3 else int(rng.integers(1, tiles - first + 1))
        plan.append(("plain" if f % 17 == 16 else "lane", 0.05 * f, first, count))
    results = []
    for mode in ("plain", "overlapped"):
        with Ocean(n, 1000.0, n_tiles=tiles) as o:
            for i, (h0, w) in enumerate(data):
                o.set_spectrum(i, h0, w)
            o.update(0.0)
            trace = []
            for step, (kind, t, first, count) in enumerate(plan):
                if mode == "plain" or kind == "plain":
                    o.update_tiles(t, first, count)
                else:
                    o.update_overlapped(t, first, count)
                if step % 11 == 10:
                    trace.append(o.output_checksums().copy())          # joins the lanes
                if step == 60:
                    o.set_spectrum(1, *data[3])                         # an upload in between: the lanes must see it
            trace.append(o.output_checksums().copy())
            trace.append(np.frombuffer(o.read_back(2).tobytes(), np.uint32).astype(np.uint64).sum())
            results.append(trace)
    for a, b in zip(*results):
        np.testing.assert_array_equal(a, b)


def test_overlapped_single_tile_rotation_is_faster_or_equal_and_correct(oracle):
    n, tiles = 1024, 8
    with Ocean(n, 1000.0, n_tiles=tiles) as o:
        for i in range(tiles):
            o.generate_spectrum(i, 77, stream_id=i)
        for f in range(64):
            o.update_overlapped(0.016 * f, f % tiles, 1)
        h0, w = o.get_spectrum(7)
        out = o.read_back(7)                                            # tile 7 was last updated at frame 63
    assert max(max_rel_err(out, oracle.frame(h0, w, 0.016 * 63, n, prec="f64"))) <= TOL


def test_overlapped_same_tiles_every_frame_matches_plain():
    """All tiles updated every frame through the lanes (the column kernel of frame n+1 is ordered behind frame n,
    the row kernel is not): per-frame results must equal the plain sequence."""
    n, tiles, frames = 1024, 3, 40
    data = [synthetic_tile(n, g + 90) for g in range(tiles)]
    with Ocean(n, 1000.0, n_tiles=tiles) as o:
        for i, (h0, w) in enumerate(data):
            o.set_spectrum(i, h0, w)
        plain = {}
        for f in range(frames):
            o.update(0.016 * f)
            if f % 7 == 6 or f == frames - 1:
                plain[f] = o.output_checksums().copy()
        got = {}
        for f in range(frames):
            o.update_overlapped(0.016 * f)                      # no host synchronisation between the frames
            if f % 7 == 6 or f == frames - 1:
                got[f] = o.output_checksums().copy()
        assert got.keys() == plain.keys()
        for f in got:
            np.testing.assert_array_equal(got[f], plain[f])


@pytest.mark.parametrize("n,tiles,recorded", [
    (64, 2, [0x492fdbd98f07, 0x498a7a9426d6]),
    (256, 2, [0x4654c563c6095, 0x46bbf0ef28c41]),
    (512, 3, [0x123d174b505765, 0x11675d731b6eb6, 0x11ce0c0c35c756]),
    (1024, 3, [0x46d6504a9e50f6, 0x46aee1b3966151, 0x46ee9539fd5459]),
    (2048, 2, [0x126bdd7f5985660, 0x120b26594278c21]),
])
def test_maps_are_bit_identical_to_the_recorded_round2_maps(n, tiles, recorded):
    """Checksums of the maps of device-generated tiles (seed 7, stream id = tile) after updates at t = 0, 0.1, 0.2,
    recorded on B200 before this round's kernel restructuring (batched phase A, per-row mbarriers, shuffle-closed
    radix-2 and packed butterflies at N = 2048, padded XH / HR): every one of those changes kept the maps bit for bit
    (scripts/san_target.py prints the same numbers)."""
    with Ocean(n, 1000.0, n_tiles=tiles) as o:
        for i in range(tiles):
            o.generate_spectrum(i, 7, stream_id=i)
        for f in range(3):
            o.update(0.1 * f)
        assert [int(s) for s in o.output_checksums()] == recorded


def test_overlapped_frame_returning_to_a_tile_of_an_older_frame_of_the_other_lane():
    """Lane 0: A (tiles 0..6, long), lane 1: B (tile 7), lane 0: C (tile 7), lane 1: D (tile 0). D returns to a tile that
    an OLDER frame of lane 0 wrote (lane 0's latest frame, C, does not touch it): D's column kernel must still be
    ordered behind A, or the long frame A finishes last and leaves its own, older, map in tile 0."""
    n, tiles = 1024, 8
    plan = [(0.5, 0, 7), (0.75, 7, 1), (1.0, 7, 1), (1.25, 0, 1)]
    sums = []
    for mode in ("plain", "overlapped"):
        with Ocean(n, 1000.0, n_tiles=tiles) as o:
            for i in range(tiles):
                o.generate_spectrum(i, 11, stream_id=i)
            o.update(0.0)
            o.sync()
            for rep in range(20):                                   # the same four frames, back to back
                for t, first, count in plan:
                    (o.update_tiles if mode == "plain" else o.update_overlapped)(t + 0.01 * rep, first, count)
            sums.append(o.output_checksums().copy())
    np.testing.assert_array_equal(sums[0], sums[1])
