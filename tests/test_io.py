"""bincode reader/writer and the synthetic spectrum generator (host logic, no GPU)."""
import os
import struct

import numpy as np
import pytest

from gfx_ocean_b200 import bincode, spectrum
from conftest import GOLDEN


def test_reference_files_decode(ref_inputs):
    sp, om = ref_inputs
    assert sp.shape == (512, 512, 2) and om.shape == (512, 512)
    assert os.path.getsize(os.path.join(GOLDEN, "ref_data", "omega.bin")) == 1048584
    assert os.path.getsize(os.path.join(GOLDEN, "ref_data", "spectrum.bin")) == 2097160
    assert 0.13484 < om.min() < 0.13486 and 4.7286 < om.max() < 4.7287
    assert np.isfinite(sp).all() and np.hypot(sp[..., 0], sp[..., 1]).max() < 1.2


def test_roundtrip(tmp_path):
    a = np.random.default_rng(0).standard_normal(64).astype(np.float32)
    b = np.random.default_rng(1).standard_normal((64, 2)).astype(np.float32)
    bincode.write_vec_f32(str(tmp_path / "a.bin"), a)
    bincode.write_vec_f32x2(str(tmp_path / "b.bin"), b)
    assert np.array_equal(bincode.read_vec_f32(str(tmp_path / "a.bin")), a)
    assert np.array_equal(bincode.read_vec_f32x2(str(tmp_path / "b.bin")), b)
    assert bincode.resolution_of(64) == 8


def test_rejects_bad_files(tmp_path):
    p = tmp_path / "short.bin"
    p.write_bytes(b"\x01\x02")
    with pytest.raises(ValueError):
        bincode.read_vec_f32(str(p))
    p.write_bytes(struct.pack("<Q", 5) + b"\0" * 16)      # length prefix / payload mismatch
    with pytest.raises(ValueError):
        bincode.read_vec_f32(str(p))
    p.write_bytes(struct.pack("<Q", 0))                   # empty vector is legal
    assert bincode.read_vec_f32(str(p)).size == 0
    with pytest.raises(ValueError):
        bincode.resolution_of(48)
    with pytest.raises(ValueError):
        bincode.resolution_of(36)                          # square but not a power of two


def test_dispersion_matches_shipped_omega(ref_inputs):
    """The synthetic dispersion formula reproduces data/omega.bin to ~2e-5."""
    _, om = ref_inputs
    w = spectrum.dispersion(512)
    assert np.abs(w - om).max() / om.max() < 3e-5


def test_synthetic_tiles_are_seeded_and_distinct():
    a, wa = spectrum.synthetic_tile(64, 0)
    b, _ = spectrum.synthetic_tile(64, 0)
    c, _ = spectrum.synthetic_tile(64, 1)
    assert a.dtype == np.float32 and a.shape == (64, 64, 2) and wa.shape == (64, 64)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert np.hypot(a[..., 0], a[..., 1]).max() == pytest.approx(1.0, rel=1e-6)
    assert np.isfinite(a).all() and (wa > 0).all()
