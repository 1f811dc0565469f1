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
    assert lib.tocvp_abi_version() == 3
    assert not [n for n in names if n.startswith("tocvp_set_") or "probe" in n], "no process-wide knobs in the product ABI"


def test_entry_points_reject_bad_arguments_without_a_device():
    """Error convention of the C ABI: argument checks come first, return a negative TOCVP_ERR_* code and
    leave a message in tocvp_last_error() -- no exception, no exit, no CUDA call needed to get there."""
    from textocvp_b200 import _lib
    lib = _lib.load()
    lib.tocvp_last_error.restype = ctypes.c_char_p
    null = ctypes.c_void_p(0)
    sz = ctypes.c_size_t
    calls = {
        "tocvp_savi_decode": (null, null, 0, null, null, null, null, sz(0), null, null, 0),
        "tocvp_savi_encode": (null, null, sz(0), 0, null, null, null, sz(0), null),
        "tocvp_slot_attention": (null, null, 0, sz(0), 0, 0, null, 0, null, 0, null, null, sz(0), null),
        "tocvp_slot_attention_seq": (null, null, 0, sz(0), sz(0), 0, 0, 0, 0, 0, null, null, sz(0), sz(0), null, null,
                                     sz(0), null),
        "tocvp_predictor_rollout": (null, null, sz(0), null, 0, 0, 0, 0, null, null, sz(0), null),
        "tocvp_frame_metrics": (null, null, sz(0), 0, 0, 0, 0, 0, 0, 0, null, null, null, null),
        "tocvp_mha_f16": (null, 0, null, null, 0, 0, 0, 0, 0, null, 0, null),
    }
    for name, args in calls.items():
        rc = getattr(lib, name)(*args)
        assert rc < 0, (name, rc)
        assert lib.tocvp_last_error(), name
    # workspace queries with null weights answer 0 instead of crashing
    lib.tocvp_savi_decode_workspace_bytes.restype = ctypes.c_size_t
    assert lib.tocvp_savi_decode_workspace_bytes(null, 4) == 0
    # tuning options travel with each call (tocvp_tuning); the library exports no process-wide setters
    assert not hasattr(lib, "tocvp_tuning.no_pdl") and not hasattr(lib, "tocvp_tuning.decode_mode")
    lib.tocvp_sizeof_tuning.restype = ctypes.c_size_t
    assert lib.tocvp_sizeof_tuning() == ctypes.sizeof(_lib.Tuning)
