"""ORACLE (test infrastructure, NOT product code): builds and loads the CPU restatement.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force: bool = False) -> None:
    subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, stdout=subprocess.DEVNULL)


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown CPU"


def build_native() -> str:
    """The timing build for THIS host: -O3 -march=native (bench.py's CPU arms).  Compiled where it runs -- the GPU box's
    host CPU is not the build container's -- into a file named after the CPU model; falls back to the portable
    x86-64-v3 build if the compiler is missing."""
    import hashlib
    tag = hashlib.sha1(cpu_model().encode()).hexdigest()[:10]
    path = os.path.join(_HERE, "_build", f"libhc_oracle_native_{tag}.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp", ".h"))]
    if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        cmd = ["g++", "-std=c++17", "-O3", "-march=native", "-fPIC", "-pthread", "-Wno-misleading-indentation", "-shared", "-o", path,
               os.path.join(_HERE, "capi.cpp")]
        try:
            subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        except (OSError, subprocess.CalledProcessError):
            return os.path.join(_HERE, "_build", "libhc_oracle_fast.so")
    return path


def load(fast: bool = False, native: bool = False):
    """Returns a ``hcb200.capi.CApi`` bound to the oracle library (prefix ``orc_``).  fast: the timing build (FMA
    contraction allowed, -O3); native: that build compiled for the host it runs on."""
    import hcb200
    name = "libhc_oracle_fast.so" if fast else "libhc_oracle.so"
    path = os.path.join(_HERE, "_build", name)
    if not os.path.exists(path):
        build()
    if native:
        path = build_native()
    lib = ctypes.CDLL(path)
    api = hcb200.capi.CApi(lib, "orc_")
    api.raw = lib
    return api
