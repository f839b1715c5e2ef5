"""ORACLE (test infrastructure, NOT product code): builds and loads the CPU restatement.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force: bool = False) -> None:
    subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, stdout=subprocess.DEVNULL)


def load(fast: bool = False):
    """Returns a ``hcb200.capi.CApi`` bound to the oracle library (prefix ``orc_``)."""
    import hcb200
    name = "libhc_oracle_fast.so" if fast else "libhc_oracle.so"
    path = os.path.join(_HERE, "_build", name)
    if not os.path.exists(path):
        build()
    lib = ctypes.CDLL(path)
    api = hcb200.capi.CApi(lib, "orc_")
    api.raw = lib
    return api
