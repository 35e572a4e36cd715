import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.ocean_oracle import COracle
    return COracle()


@pytest.fixture(scope="session")
def ref_inputs():
    """The reference-owned 512x512 inputs (copies of /root/reference/data/*.bin)."""
    from gfx_ocean_b200.bincode import read_vec_f32, read_vec_f32x2
    om = read_vec_f32(os.path.join(GOLDEN, "ref_data", "omega.bin"))
    sp = read_vec_f32x2(os.path.join(GOLDEN, "ref_data", "spectrum.bin"))
    return sp.reshape(512, 512, 2), om.reshape(512, 512)


@pytest.fixture(scope="session")
def golden_512():
    return np.load(os.path.join(GOLDEN, "golden_512.npz"))


@pytest.fixture(scope="session")
def golden_synth():
    return np.load(os.path.join(GOLDEN, "golden_synth.npz"))
