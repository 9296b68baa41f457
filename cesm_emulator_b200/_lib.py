"""ctypes binding of libcesm_b200.so (the C ABI declared in include/cesm_b200.h).

There is deliberately no fallback: if the shared library is missing, or a call fails, the
caller gets an exception.  Importing this module does not need a GPU; calling a compute entry
point does.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_void_p
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libcesm_b200.so"

CESM_MAX_TAPS = 16


class IgemmArgs(Structure):
    """Mirror of `cesm_igemm_args` (include/cesm_b200.h)."""

    _fields_ = [
        ("a0", c_void_p), ("a1", c_void_p),
        ("c0", c_int32), ("c1", c_int32),
        ("n", c_int32), ("h", c_int32), ("w", c_int32),
        ("stride", c_int32),
        ("num_taps", c_int32),
        ("tap_dh", c_int32 * CESM_MAX_TAPS), ("tap_dw", c_int32 * CESM_MAX_TAPS),
        ("wt", c_void_p), ("cout", c_int32),
        ("oh", c_int32), ("ow", c_int32),
        ("out", c_void_p), ("out_fp32", c_int32), ("ldo", c_int32),
        ("out_h", c_int32), ("out_w", c_int32),
        ("o_sh", c_int32), ("o_sw", c_int32), ("o_h0", c_int32), ("o_w0", c_int32),
        ("bias", c_void_p), ("residual", c_void_p), ("ldr", c_int32),
    ]


class CesmError(RuntimeError):
    pass


_lib = None


def load() -> ctypes.CDLL:
    """Load the kernel library, building nothing: a missing .so is a hard error."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise CesmError(
            f"{LIB_PATH} not found: build it with `python -m cesm_emulator_b200.build` "
            "(there is no CPU or PyTorch fallback for the B200 kernels)")
    lib = ctypes.CDLL(str(LIB_PATH))
    lib.cesm_last_error.restype = c_char_p
    lib.cesm_version.restype = c_char_p
    _declare(lib)
    _lib = lib
    return lib


# name -> argtypes; every function returns int (cesm_status)
_SIGNATURES: dict[str, list] = {
    "cesm_igemm": [POINTER(IgemmArgs), c_void_p],
}


def _declare(lib: ctypes.CDLL) -> None:
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int


def exported_symbols() -> list[str]:
    return ["cesm_last_error", "cesm_version", *sorted(_SIGNATURES)]


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().cesm_last_error().decode("utf-8", "replace")
        raise CesmError(f"{what} failed (status {rc}): {msg}")


def call(name: str, *args) -> None:
    lib = load()
    check(getattr(lib, name)(*args), name)
