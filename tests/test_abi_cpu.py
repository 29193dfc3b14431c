"""CPU: the C-ABI library loads and exports every symbol include/msr3d_b200.h declares."""
import ctypes
import os

from msr3d_b200 import _C


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_C.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_C.LIB_PATH)
    declared = _C.exported_symbols_in_header()
    assert len(declared) >= 12
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing


def test_abi_version_and_error_string():
    lib = _C.lib()
    assert lib.msr3d_abi_version() >= 1
    assert isinstance(lib.msr3d_last_error(), bytes)


def test_argument_validation_needs_no_gpu():
    # null pointers are rejected before any CUDA call
    lib = _C.lib()
    rc = lib.msr3d_furthest_point_sampling_f32(None, None, 1, 16, 4, None)
    assert rc == 1 and b"null" in lib.msr3d_last_error()
