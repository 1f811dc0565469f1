"""ctypes binding of libtocvp.so (the C ABI declared in include/tocvp.h).

There is deliberately no fallback: if the shared library is missing or the device is not an
sm_100 part, importing / initialising raises.
"""
from __future__ import annotations

import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TOCVP_LIB") or os.path.join(_HERE, "libtocvp.so")   # TOCVP_LIB: A/B of two builds (dev tools)
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "tocvp.h")

_lib = None
_inited_devices = set()


class Tuning(ctypes.Structure):
    """tocvp_tuning (include/tocvp.h): per-call kernel-selection options.  The library holds no tuning state; the modules
    attach a pointer to ONE caller-owned instance (``TUNING`` below) to every weights struct they pass, so tests and A/B
    tools flip a field here and the next call sees it.  All-zero = defaults."""
    _fields_ = [("gemm_mode", ctypes.c_int), ("gemm_no_wres", ctypes.c_int), ("conv_mode", ctypes.c_int),
                ("encode_mode", ctypes.c_int), ("decode_mode", ctypes.c_int), ("corrector_mode", ctypes.c_int),
                ("no_pdl", ctypes.c_int), ("no_tile_alternation", ctypes.c_int), ("reserved", ctypes.c_int * 8)]


TUNING = Tuning()


def tuning_ptr():
    return ctypes.c_void_p(ctypes.addressof(TUNING))


class TocvpError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TocvpError(f"{LIB_PATH} not found: build it with `python -m textocvp_b200.build` "
                             "(there is no CPU / PyTorch fallback for this path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.tocvp_last_error.restype = ctypes.c_char_p
        if _lib.tocvp_abi_version() != 3:
            raise TocvpError("libtocvp.so ABI version mismatch; rebuild")
        _lib.tocvp_sizeof_tuning.restype = ctypes.c_size_t
        if _lib.tocvp_sizeof_tuning() != ctypes.sizeof(Tuning):
            raise TocvpError("struct tocvp_tuning: C size != ctypes size")
    return _lib


def declared_symbols():
    """All function names declared in include/tocvp.h (used by the CPU-side export test)."""
    src = open(HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tocvp_[a-z0-9_]+)\s*\(", src)))


def check(rc: int):
    if rc != 0:
        msg = load().tocvp_last_error().decode()
        raise TocvpError(f"libtocvp error {rc}: {msg}")


def init(device) -> None:
    idx = torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    if idx not in _inited_devices:
        check(load().tocvp_init(idx))
        _inited_devices.add(idx)


def ptr(t):
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def call(name: str, *args):
    fn = getattr(load(), name)
    check(fn(*args))


_probe = None


def load_probe() -> ctypes.CDLL:
    """Test-only hardware probes (tests/native/libtocvp_probe.so); not part of the product library."""
    global _probe
    if _probe is None:
        path = os.path.join(os.path.dirname(_HERE), "tests", "native", "libtocvp_probe.so")
        if not os.path.exists(path):
            raise TocvpError(f"{path} not found: build it with `python -m textocvp_b200.build`")
        _probe = ctypes.CDLL(path)
    return _probe


c_int = ctypes.c_int
c_float = ctypes.c_float
c_size_t = ctypes.c_size_t
