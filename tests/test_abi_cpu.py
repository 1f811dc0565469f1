"""CPU-side checks of the C-ABI library: it builds, loads, and exports every symbol that
include/tocvp.h declares (no compute calls -- there is no GPU here)."""
import ctypes


def test_library_exports_all_declared_symbols():
    from textocvp_b200 import build, _lib
    build.build()
    lib = _lib.load()
    names = _lib.declared_symbols()
    assert "tocvp_gemm_f16" in names and "tocvp_init" in names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.tocvp_abi_version() == 1
