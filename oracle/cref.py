"""ctypes driver for the C oracle (oracle/c/check_ref.c) -- TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "_build", "libcheck_ref.so")
_lib = None


class _Batch(ctypes.Structure):
    _fields_ = [("n_requests", ctypes.c_uint64), ("max_actions", ctypes.c_uint32),
                ("now_unix_nanos", ctypes.c_int64), ("flags", ctypes.c_uint32),
                ("columns", ctypes.POINTER(ctypes.c_void_p)), ("column_bytes", ctypes.POINTER(ctypes.c_size_t)),
                ("n_columns", ctypes.c_uint32)]


def build(force=False):
    src = os.path.join(_ROOT, "oracle", "c", "check_ref.c")
    hdr = os.path.join(_ROOT, "include", "cerbos_b200_format.h")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["make", "-s", "-C", os.path.join(_ROOT, "oracle", "c")], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.cref_check.restype = ctypes.c_int
        _lib.cref_check.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(_Batch), ctypes.c_void_p,
                                    ctypes.c_int]
    return _lib


def check(blob: bytes, columns, n: int, max_actions: int, now_ns: int = 0, flags: int = 0, n_threads: int = 1):
    """Runs the C oracle. columns: list of C-contiguous numpy arrays (encode.py order).
    Returns uint8[n, max_actions] effects (1 ALLOW, 2 DENY, 0 padding)."""
    cols = [np.ascontiguousarray(c) for c in columns]
    ptrs = (ctypes.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
    sizes = (ctypes.c_size_t * len(cols))(*[c.nbytes for c in cols])
    b = _Batch(n, max_actions, now_ns, flags, ptrs, sizes, len(cols))
    out = np.zeros((n, max(max_actions, 1)), dtype=np.uint8)
    buf = ctypes.create_string_buffer(blob, len(blob))
    rc = lib().cref_check(buf, len(blob), ctypes.byref(b), out.ctypes.data, n_threads)
    if rc != 0:
        raise RuntimeError(f"cref_check failed: {rc}")
    return out
