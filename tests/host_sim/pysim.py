"""Test-only loader of the host-compiled device code (see Makefile in this directory)."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    import hcb200
    subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(os.path.join(_HERE, "_build", "libhc_sim.so"))
    api = hcb200.capi.CApi(lib, "hc_")
    api.raw = lib
    lib.hc_last_error.restype = ctypes.c_char_p
    return api
