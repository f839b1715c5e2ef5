"""The C-ABI library loads and exports every symbol include/hc_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "hc_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hc_[a-z_0-9]+)\s*\(", txt)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for s in ("hc_init", "hc_system_create", "hc_homotopy_create", "hc_track_batch", "hc_polyhedral_track_batch",
              "hc_evaluate", "hc_evaluate_and_jacobian", "hc_taylor", "hc_options_default", "hc_last_error"):
        assert s in syms


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/hc_b200.h compiles as C99 (no C++ or torch types in the signatures) and a
    C translation unit can fill the structs and reference every entry point."""
    import subprocess
    src = tmp_path / "use.c"
    calls = "\n".join(f"    (void)&{s};" for s in declared_symbols())
    src.write_text('#include "hc_b200.h"\nint main(void) {\n    hc_options o; hc_results r; hc_program_desc p; hc_homotopy_desc h; hc_timing t;\n'
                   '    (void)o; (void)r; (void)p; (void)h; (void)t;\n' + calls + "\n    return 0;\n}\n")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-Wno-comment", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                   check=True)


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "homotopycontinuation.jl_b200", "libhc_b200.so"))
    for s in declared_symbols():
        assert hasattr(lib, s), s


def test_option_struct_layout_matches():
    """hc_options (C) == capi.Options (ctypes) == DevOptions (device): defaults round-trip by name."""
    import hcb200
    from hcb200 import capi
    lib = ctypes.CDLL(os.path.join(ROOT, "homotopycontinuation.jl_b200", "libhc_b200.so"))
    o = capi.Options()
    lib.hc_options_default(ctypes.byref(o))
    assert o.max_steps == 10000 and o.a == 0.125 and o.beta_tau == 0.4 and o.strict_beta_tau == 0.3
    assert o.endgame_start == 0.1 and o.max_endgame_steps == 2000 and o.max_winding_number == 6
    assert o.sing_cond == 1e14 and o.refine_steps == 3 and o.scale_min == 1e-4 and o.scale_max == 2.0 ** 511


def test_no_cpu_fallback_without_device():
    """Without a CUDA device hc_init must fail (this container has none; on the GPU box it succeeds)."""
    lib = ctypes.CDLL(os.path.join(ROOT, "homotopycontinuation.jl_b200", "libhc_b200.so"))
    lib.hc_last_error.restype = ctypes.c_char_p
    import torch
    rc = lib.hc_init(0)
    if torch.cuda.is_available():
        assert rc == 0
    else:
        assert rc != 0 and b"no CUDA device" in lib.hc_last_error()
