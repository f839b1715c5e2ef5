"""Loads libhc_b200.so (the CUDA library behind include/hc_b200.h).  No CPU fallback: if the
library or a CUDA device is missing, this raises."""
from __future__ import annotations

import ctypes as C
import os

from .capi import CApi, Options, ResultsDesc, c_double_p, c_int32_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HC_B200_LIB", os.path.join(_HERE, "libhc_b200.so"))  # override: debugging builds only

_api: CApi | None = None


class Timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("kernel_ms", C.c_double), ("d2h_ms", C.c_double),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("grid", C.c_int32), ("block", C.c_int32), ("lanes", C.c_int32), ("slab_bytes", C.c_int64),
                ("devices", C.c_int32), ("engine", C.c_int32), ("handoff_paths", C.c_int64), ("handoff_ms", C.c_double)]


def load(device: int | None = None, devices: list[int] | None = None) -> CApi:
    """Returns the ctypes API bound to libhc_b200.so.  `device`: the one CUDA device this process drives (default:
    LOCAL_RANK); `devices`: a list -- every batch call is then split over them (hc_init_devices).  Calling it again
    with another device list re-initialises the library (handles created before must not be used afterwards)."""
    global _api
    if _api is not None and devices is not None:
        arr = (C.c_int32 * len(devices))(*devices)
        if _api.raw.hc_init_devices(arr, len(devices)) != 0:
            raise RuntimeError("hc_init_devices failed: " + _api.raw.hc_last_error().decode())
        return _api
    if _api is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        api = CApi(lib, "hc_")
        api.raw = lib
        lib.hc_last_error.restype = C.c_char_p
        lib.hc_init.restype, lib.hc_init.argtypes = C.c_int32, [C.c_int32]
        lib.hc_get_timing.restype, lib.hc_get_timing.argtypes = None, [C.POINTER(Timing)]
        lib.hc_dfma_peak.restype, lib.hc_dfma_peak.argtypes = C.c_double, [C.c_int32]
        lib.hc_resident_create.restype = C.c_void_p
        lib.hc_resident_create.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Options), C.c_int32, C.c_int64, c_double_p, c_double_p,
                                           c_double_p, c_double_p, c_double_p, c_int32_p, c_double_p, C.c_int32]
        lib.hc_resident_run.restype, lib.hc_resident_run.argtypes = C.c_int32, [C.c_void_p, C.POINTER(C.c_double)]
        lib.hc_resident_fetch.restype, lib.hc_resident_fetch.argtypes = C.c_int32, [C.c_void_p, C.POINTER(ResultsDesc)]
        lib.hc_resident_destroy.restype, lib.hc_resident_destroy.argtypes = None, [C.c_void_p]
        lib.hc_host_register.restype, lib.hc_host_register.argtypes = C.c_int32, [C.c_void_p, C.c_int64]
        lib.hc_host_unregister.restype, lib.hc_host_unregister.argtypes = C.c_int32, [C.c_void_p]
        lib.hc_init_devices.restype, lib.hc_init_devices.argtypes = C.c_int32, [C.POINTER(C.c_int32), C.c_int32]
        lib.hc_device_count.restype, lib.hc_device_count.argtypes = C.c_int32, []
        lib.hc_request_cancel.restype, lib.hc_request_cancel.argtypes = None, [C.c_int32]
        lib.hc_jit_prepare.restype, lib.hc_jit_prepare.argtypes = C.c_int32, [C.c_void_p, C.c_int32, c_double_p]
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        if devices is not None:
            arr = (C.c_int32 * len(devices))(*devices)
            rc = lib.hc_init_devices(arr, len(devices))
        else:
            rc = lib.hc_init(device)
        if rc != 0:
            raise RuntimeError("hc_init failed: " + lib.hc_last_error().decode())
        _api = api
    return _api


def pin(*arrays) -> list:
    """Page-locks caller-owned numpy arrays (hc_host_register) so that hc_track_batch's copies are DMA transfers.
    Returns the arrays that were registered; pass them to unpin() before they are freed."""
    raw = load().raw
    done = []
    for a in arrays:
        if a is None or a.nbytes == 0:
            continue
        if not a.flags.c_contiguous:
            raise ValueError("only contiguous arrays can be page-locked")
        if raw.hc_host_register(a.ctypes.data, a.nbytes) != 0:
            msg = raw.hc_last_error().decode()
            for b in done:   # all or nothing
                raw.hc_host_unregister(b.ctypes.data)
            raise RuntimeError("hc_host_register failed: " + msg)
        done.append(a)
    return done


def unpin(arrays) -> None:
    raw = load().raw
    for a in arrays:
        raw.hc_host_unregister(a.ctypes.data)


def last_error() -> str:
    return load().raw.hc_last_error().decode()


def timing() -> Timing:
    t = Timing()
    load().raw.hc_get_timing(C.byref(t))
    return t
