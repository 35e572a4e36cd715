"""The oracle and the CUDA path pinned against the reference's OWN shipped SPIR-V, executed on the CPU.

tests/golden/spv_512.npz was produced by oracle/spv_exec.py interpreting /root/reference/shader/spv/*.spv
(the binaries the reference embeds and dispatches, src/fft.rs:20-25, src/ocean.rs:26-28,195-197,
src/render.rs:1122-1287) on data/omega.bin + data/spectrum.bin -- generator: tests/golden/make_spv_golden.py.
These are reference-authored results (not the self-authored oracle's), so agreement here is what turns
"faithful by inspection" into "pinned".

Bar: per channel max|a - b| / max|b| <= 1e-5 (north_star); measured oracle-vs-SPIR-V is <= 6e-7.
"""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle.ocean_oracle import max_rel_err

TOL = 1e-5
SPV_DIR = "/root/reference/shader/spv"
have_reference = os.path.isdir(SPV_DIR)


@pytest.fixture(scope="module")
def spv():
    return np.load(os.path.join(GOLDEN, "spv_512.npz"))


def _frame4(spv, ti):
    out = np.zeros((512, 512, 4), np.float32)
    out[..., :3] = spv["frames"][ti]
    return out


# ------------------------------------------------------------------------------------------------
# CPU: oracle vs the executed SPIR-V
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ti", [0, 1, 2, 3])
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_oracle_matches_executed_spirv(ti, prec, spv, oracle, ref_inputs):
    sp, om = ref_inputs
    t = float(spv["times"][ti])
    ref = _frame4(spv, ti)
    out = oracle.frame(sp, om, t, 512, prec=prec)
    errs = max_rel_err(out, ref)
    print(f"oracle {prec} vs SPIR-V t={t}: {errs}")
    assert max(errs) <= TOL
    assert max(errs) <= 2e-6          # regression guard: measured <= 6e-7 (f64), <= 3e-7 (f32)


def test_numpy_twin_matches_executed_spirv(spv, ref_inputs):
    from oracle.ocean_oracle import frame_np
    sp, om = ref_inputs
    out = frame_np(sp, om, 37.5, 512)
    assert max(max_rel_err(out, _frame4(spv, 2))) <= TOL


def test_propagate_matches_executed_spirv(spv, oracle, ref_inputs):
    """Post-propagate spectra (propagate.comp.spv alone): the u32 wrap + OpConvertUToF quirk is in here."""
    sp, om = ref_inputs
    h, dx, dz = oracle.propagate(sp, om, 1.0, 512, prec="f32")
    idx = spv["spectra_t1_index"]
    for name, a in (("h", h), ("dx", dx), ("dz", dz)):
        ref = spv[f"spectra_t1_{name}"]
        got = a.reshape(-1)[idx]
        got = np.stack([got.real, got.imag], -1)
        assert np.abs(got - ref).max() <= 3e-7 * max(np.abs(ref).max(), 1.0), name
    # the quirk itself, on reference-executed numbers: at (y=300, x=200) only gx wraps -> khat = (1, ~0)
    k = int(np.nonzero(idx == 300 * 512 + 200)[0][0])
    hh = spv["spectra_t1_h"][k]
    np.testing.assert_allclose(spv["spectra_t1_dx"][k], [hh[1], -hh[0]], rtol=1e-6)      # dx = -i h
    assert np.abs(spv["spectra_t1_dz"][k]).max() < 1e-6 * np.abs(hh).max()


def test_normal_map_oracle_matches_executed_fragment_shader(spv, oracle):
    """shader/spv/ocean.frag.spv's local `N` at texel centres vs the oracle's normal map (f1)."""
    nrm = oracle.normal_map(_frame4(spv, 1).astype(np.float64), prec="f64")
    assert np.abs(nrm[..., :3] - spv["normals_t1"]).max() <= 1e-6


def test_vertex_displacement_oracle_matches_executed_vertex_shader(spv):
    from oracle.ocean_oracle import displace_grid_np
    pw = displace_grid_np(_frame4(spv, 1), 128, tuple(spv["vertex_offset"]))
    ref = spv["vertex_t1"]
    assert np.abs(pw.reshape(-1, 3) - ref).max() <= 1e-5 * np.abs(ref).max()


@pytest.mark.skipif(not have_reference, reason="/root/reference is only present in the dev container")
def test_fixture_digests_and_live_interpreter(spv, ref_inputs):
    """Re-execute the shipped modules now: digests unchanged and one frame reproduces the fixture."""
    from oracle import spv_exec
    for name, digest in zip(spv["spv_names"], spv["spv_sha256"]):
        with open(os.path.join(SPV_DIR, str(name) + ".spv"), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == str(digest), name
    sp, om = ref_inputs
    img = spv_exec.run_reference_frame(SPV_DIR, sp, om, 1.0)
    np.testing.assert_allclose(img[..., :3], spv["frames"][1], rtol=0, atol=1e-6)
    assert np.all(img[..., 3] == 0.0)


@pytest.mark.skipif(not have_reference, reason="/root/reference is only present in the dev container")
def test_interpreter_opcode_census():
    """The interpreter implements every opcode the six modules contain (it raises on anything else) and
    the propagate module really converts with OpConvertUToF."""
    from oracle import spv_exec
    census = {}
    for m in ("propagate.comp", "fft_row.comp", "fft_col.comp", "correction.comp", "ocean.vert", "ocean.frag"):
        census[m] = spv_exec.load_module(os.path.join(SPV_DIR, m + ".spv")).opcode_census
    assert census["propagate.comp"]["ConvertUToF"] == 2 and "ConvertSToF" not in census["propagate.comp"]
    assert census["fft_row.comp"]["ControlBarrier"] == 2 and census["fft_col.comp"]["ControlBarrier"] == 2
    assert census["correction.comp"]["ImageWrite"] == 1 and census["correction.comp"]["UMod"] == 1


def test_interpreter_control_flow_and_masks():
    """A hand-assembled module: loop with a divergent trip count and a selection, checked against numpy."""
    from oracle import spv_exec
    from spv_asm import build_loop_module
    mod = spv_exec.Module(build_loop_module())
    n = 64
    buf = np.zeros(n, np.uint32)
    spv_exec.dispatch(mod, (n // 8, 1, 1), {(0, 0): [buf]})
    gid = np.arange(n, dtype=np.uint32)
    # out[i] = sum_{k < i % 5} (k*3)  +  (1000 if i even else 0)
    expect = np.array([sum(3 * k for k in range(i % 5)) + (1000 if i % 2 == 0 else 0) for i in gid], np.uint32)
    np.testing.assert_array_equal(buf, expect)


# ------------------------------------------------------------------------------------------------
# GPU: the CUDA path vs the executed SPIR-V
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("ti", [0, 1, 2, 3])
@pytest.mark.parametrize("which", ["fused", "literal"])
def test_cuda_matches_executed_spirv(which, ti, spv):
    from gfx_ocean_b200 import Ocean, PIPELINE_FUSED, PIPELINE_LITERAL
    om = os.path.join(GOLDEN, "ref_data", "omega.bin")
    sp = os.path.join(GOLDEN, "ref_data", "spectrum.bin")
    pipe = PIPELINE_FUSED if which == "fused" else PIPELINE_LITERAL
    with Ocean.from_bincode(om, sp, 512, 1000.0, pipeline=pipe) as o:
        o.update(float(spv["times"][ti]))
        out = o.read_back()
    errs = max_rel_err(out, _frame4(spv, ti))
    print(f"CUDA {which} vs SPIR-V t={float(spv['times'][ti])}: {errs}")
    assert max(errs) <= TOL
    assert np.all(out[..., 3] == 0.0) and not np.signbit(out[..., 3]).any()


@pytest.mark.gpu
def test_cuda_normals_and_vertices_match_executed_spirv(spv):
    from gfx_ocean_b200 import Ocean
    om = os.path.join(GOLDEN, "ref_data", "omega.bin")
    sp = os.path.join(GOLDEN, "ref_data", "spectrum.bin")
    with Ocean.from_bincode(om, sp, 512, 1000.0) as o:
        o.update(1.0)
        o.compute_normals()
        nrm = o.read_back_normals()
        pw = o.displace_grid(128, tuple(float(v) for v in spv["vertex_offset"]))
    assert np.abs(nrm[..., :3] - spv["normals_t1"]).max() <= 2e-5
    assert np.all(nrm[..., 3] == 0.0)
    ref = spv["vertex_t1"]
    assert np.abs(pw.reshape(-1, 3) - ref).max() <= 1e-5 * np.abs(ref).max()
