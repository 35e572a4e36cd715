"""CPU checks of the drop-in boundary: the shared library loads, exports every symbol that
include/ocean_b200.h declares, and fails loudly (no CPU fallback) without a GPU."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT
from gfx_ocean_b200 import _lib, build as _build_mod  # noqa: F401
from gfx_ocean_b200.build import build


@pytest.fixture(scope="module")
def lib():
    build()
    return _lib.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ocean_b200.h")).read()
    return sorted(set(re.findall(r"OCEAN_API\s+[\w\s\*]+?\b(ocean_\w+)\s*\(", src)))


def test_header_and_binding_agree():
    syms = header_symbols()
    assert len(syms) >= 20
    assert sorted(_lib.SIGNATURES) == syms


def test_library_exports_every_declared_symbol(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (ocean_\w+)", out))
    assert set(header_symbols()) <= exported
    # nothing but the ABI leaks out of the library
    assert all(s.startswith("ocean_") for s in re.findall(r" T (\w+)", out))


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_abi_version_and_status_strings(lib):
    assert lib.ocean_abi_version() == _lib.ABI_VERSION
    assert lib.ocean_status_string(0) == b"OCEAN_OK"
    assert lib.ocean_status_string(-2) == b"OCEAN_ERR_NO_DEVICE"
    assert lib.ocean_status_string(-99) == b"OCEAN_ERR_UNKNOWN"


def test_struct_layout_matches_reference_uniform_blocks():
    # PropagateLocals {f32 time, i32 resolution, f32 domain_size} at offsets 0/4/8 (src/ocean.rs:8-13)
    assert ctypes.sizeof(_lib.PropagateLocals) == 12
    assert [_lib.PropagateLocals.time.offset, _lib.PropagateLocals.resolution.offset,
            _lib.PropagateLocals.domain_size.offset] == [0, 4, 8]
    assert ctypes.sizeof(_lib.CorrectionLocals) == 4


def test_argument_validation_needs_no_gpu(lib):
    ctx = ctypes.c_void_p()
    for n in (0, 7, 500, 8192):
        rc = lib.ocean_create(ctypes.byref(ctx), 0, n, 1000.0, 1)
        assert rc == _lib.ERR_INVALID_ARG and not ctx.value
        assert b"power of two" in lib.ocean_last_error(None)
    assert lib.ocean_create(ctypes.byref(ctx), 0, 512, 1000.0, 0) == _lib.ERR_INVALID_ARG
    assert lib.ocean_create(ctypes.byref(ctx), 0, 512, -1.0, 1) == _lib.ERR_INVALID_ARG
    assert lib.ocean_create(None, 0, 512, 1000.0, 1) == _lib.ERR_INVALID_ARG
    cfg = _lib.OceanConfig(99, 0, 512, 1000.0, 1, 0, None, 0)
    assert lib.ocean_create_ex(ctypes.byref(ctx), ctypes.byref(cfg)) == _lib.ERR_INVALID_ARG
    # null-context calls are errors, not crashes
    assert lib.ocean_update(None, 0.0) == _lib.ERR_INVALID_ARG
    assert lib.ocean_sync(None) == _lib.ERR_INVALID_ARG
    assert lib.ocean_resolution(None) == 0
    lib.ocean_destroy(None)


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gfx_ocean_b200 import Ocean, OceanError
    with pytest.raises(OceanError) as ei:
        Ocean(512)
    assert ei.value.status == _lib.ERR_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)


def test_product_package_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may use oracle/."""
    pkg = os.path.join(ROOT, "gfx_ocean_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in os.path.relpath(dirpath, pkg).split(os.sep)[:1] and dirpath != pkg:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "ocean_oracle" not in txt and "oracle/" not in txt and "import oracle" not in txt, f


def test_header_is_plain_c99(tmp_path):
    """The boundary is a C ABI: the header compiles as C99 (no C++-isms), and every declared entry point links."""
    src = tmp_path / "use.c"
    calls = "\n".join(f"    p[{i}] = (void*){s};" for i, s in enumerate(header_symbols()))
    src.write_text('#include "ocean_b200.h"\n#include <stdio.h>\nint main(void) {\n    void* p[%d];\n%s\n'
                   '    ocean_spectrum_params sp = {3e-8f, 30.0f, 9.81f, 100.0f}; ocean_config cfg; (void)sp; (void)cfg;\n'
                   '    printf("%%u %%p\\n", ocean_abi_version(), p[0]);\n    return 0;\n}\n' % (len(header_symbols()), calls))
    exe = tmp_path / "use"
    build()
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-Wno-pedantic", "-I", os.path.join(ROOT, "include"),
                        str(src), "-o", str(exe), "-L", os.path.dirname(_lib.LIB_PATH), "-l:libocean_b200.so",
                        "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.split()[0] == str(_lib.ABI_VERSION)
