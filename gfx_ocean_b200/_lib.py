"""ctypes binding of libocean_b200.so -- the C ABI declared in include/ocean_b200.h.

There is no CPU fallback: if the library is missing, or there is no B200, every entry point
fails loudly (``OceanError``)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# OCEAN_B200_LIB selects an alternative build of the same library (A/B experiments, scripts/ab_build.sh)
LIB_PATH = os.environ.get("OCEAN_B200_LIB") or os.path.join(HERE, "libocean_b200.so")

ABI_VERSION = 2
OK, ERR_INVALID_ARG, ERR_NO_DEVICE, ERR_CUDA, ERR_IO, ERR_NOT_READY, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6
PIPELINE_FUSED, PIPELINE_LITERAL = 0, 1
FLAG_KEEP_SPECTRA, FLAG_DOUBLE_BUFFER_OUTPUT, FLAG_DX_PLANE = 1, 2, 4


class PropagateLocals(C.Structure):
    """Mirror of ``PropagateLocals`` (/root/reference/src/ocean.rs:8-13)."""
    _fields_ = [("time", C.c_float), ("resolution", C.c_int32), ("domain_size", C.c_float)]


class CorrectionLocals(C.Structure):
    """Mirror of ``CorrectionLocals`` (/root/reference/src/ocean.rs:179-182)."""
    _fields_ = [("resolution", C.c_uint32)]


class OceanConfig(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("cuda_device", C.c_int32), ("resolution", C.c_uint32),
                ("domain_size", C.c_float), ("n_tiles", C.c_uint32), ("pipeline", C.c_uint32),
                ("stream", C.c_void_p), ("flags", C.c_uint32)]


class SpectrumParams(C.Structure):
    _fields_ = [("amplitude", C.c_float), ("wind_speed", C.c_float), ("gravity", C.c_float), ("depth", C.c_float)]


class OceanError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{message} [status {status}]")
        self.status = status


# every symbol include/ocean_b200.h declares: name -> (restype, argtypes)
_F, _U32, _U64, _P = C.c_float, C.c_uint32, C.c_uint64, C.c_void_p
_FP = C.POINTER(C.c_float)
SIGNATURES = {
    "ocean_create": (C.c_int, [C.POINTER(_P), C.c_int, _U32, _F, _U32]),
    "ocean_create_ex": (C.c_int, [C.POINTER(_P), C.POINTER(OceanConfig)]),
    "ocean_destroy": (None, [_P]),
    "ocean_set_spectrum": (C.c_int, [_P, _U32, _P, _P]),
    "ocean_set_spectrum_device": (C.c_int, [_P, _U32, _P, _P]),
    "ocean_load_bincode": (C.c_int, [_P, _U32, C.c_char_p, C.c_char_p]),
    "ocean_update": (C.c_int, [_P, _F]),
    "ocean_update_tiles": (C.c_int, [_P, _F, _U32, _U32]),
    "ocean_update_sequence": (C.c_int, [_P, _F, _F, _U32]),
    "ocean_update_graph": (C.c_int, [_P, _F, _U32, _U32]),
    "ocean_update_overlapped": (C.c_int, [_P, _F, _U32, _U32]),
    "ocean_join": (C.c_int, [_P]),
    "ocean_output_device": (C.c_int, [_P, _U32, C.POINTER(_P)]),
    "ocean_download": (C.c_int, [_P, _U32, _P]),
    "ocean_download_async": (C.c_int, [_P, _U32, _P]),
    "ocean_download_all_async": (C.c_int, [_P, _P]),
    "ocean_download_fence": (C.c_int, [_P, _U32]),
    "ocean_sync": (C.c_int, [_P]),
    "ocean_set_output_device": (C.c_int, [_P, _U32, _P, C.c_size_t]),
    "ocean_update_sequence_checksums": (C.c_int, [_P, _F, _F, _U32, _P]),
    "ocean_output_checksums": (C.c_int, [_P, _P]),
    "ocean_displace_grid": (C.c_int, [_P, _U32, _U32, _F, _F, _P]),
    "ocean_displace_grid_device": (C.c_int, [_P, _U32, _U32, _F, _F, _P]),
    "ocean_generate_spectrum": (C.c_int, [_P, _U32, _U64, _U32, C.POINTER(SpectrumParams), _P]),
    "ocean_get_spectrum": (C.c_int, [_P, _U32, _P, _P]),
    "ocean_debug_spectra": (C.c_int, [_P, _U32, _P, _P, _P]),
    "ocean_compute_normals": (C.c_int, [_P, _U32, _U32]),
    "ocean_normals_device": (C.c_int, [_P, _U32, C.POINTER(_P)]),
    "ocean_download_normals": (C.c_int, [_P, _U32, _P]),
    "ocean_profile_update": (C.c_int, [_P, _F, _FP, _U32, C.POINTER(_U32)]),
    "ocean_get_locals": (C.c_int, [_P, C.POINTER(PropagateLocals), C.POINTER(CorrectionLocals)]),
    "ocean_resolution": (_U32, [_P]),
    "ocean_n_tiles": (_U32, [_P]),
    "ocean_launch_count": (_U64, [_P]),
    "ocean_stream": (_P, [_P]),
    "ocean_algorithmic_bytes_per_update": (_U64, [_P]),
    "ocean_last_error": (C.c_char_p, [_P]),
    "ocean_status_string": (C.c_char_p, [C.c_int]),
    "ocean_abi_version": (_U32, []),
}

_lib = None


def load() -> C.CDLL:
    """Load libocean_b200.so (built by ``python -m gfx_ocean_b200.build``) and bind every symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OceanError(ERR_UNSUPPORTED, f"{LIB_PATH} is not built (run `python -m gfx_ocean_b200.build`); "
                                          "there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype, fn.argtypes = res, args
    if lib.ocean_abi_version() != ABI_VERSION:
        raise OceanError(ERR_UNSUPPORTED, "libocean_b200.so ABI version mismatch")
    _lib = lib
    return lib
