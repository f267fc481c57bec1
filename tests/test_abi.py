"""CPU-only: the C-ABI library loads and exports every symbol include/*.h declares; construction fails
loudly (no CPU fallback) when no GPU is present; the product never imports anything from oracle/."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^[A-Za-z_][A-Za-z0-9_ \*]*?\b([a-z_][a-z0-9_]*)\s*\(", src, flags=re.M)
    return sorted({n for n in names if n not in ("defined", "sizeof")})


def test_library_exports_every_declared_symbol():
    import lightning_b200 as L
    lib = L.load_library()
    funcs = declared_functions("cln_sigverify.h") + declared_functions("cln_dropin.h")
    assert len(funcs) >= 30, funcs
    missing = [f for f in funcs if not hasattr(lib, f)]
    assert not missing, missing


def test_key_sizes_without_gpu():
    import lightning_b200 as L
    lib = L.load_library()
    assert [lib.sv_key_size(k) for k in (0, 1, 2, 3)] == [33, 64, 32, 0]


def test_no_cpu_fallback():
    import torch
    import lightning_b200 as L
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    with pytest.raises(L.EngineError) as e:
        L.SigVerifier(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_missing_library_fails_loudly(monkeypatch):
    import lightning_b200.engine as E
    monkeypatch.setattr(E, "LIB_PATH", "/nonexistent/libcln_sigverify.so")
    with pytest.raises(E.EngineError):
        E.load_library()


def test_product_does_not_touch_the_oracle():
    """oracle/ is test infrastructure: nothing under lightning_b200/ or include/ may reference it."""
    out = subprocess.run(["grep", "-rIl", "-E", r"oracle/|libsecp_ref|libsecp_port|secp_port", os.path.join(ROOT, "lightning_b200"),
                          os.path.join(ROOT, "include")], capture_output=True, text=True).stdout.split()
    out = [o for o in out if not o.endswith(".so") and not o.endswith(".o") and "__pycache__" not in o]
    # build.py names the artefacts it can build for the tests; that is a build recipe, not a use
    assert all(os.path.basename(o) == "build.py" for o in out), out
    lib = os.path.join(ROOT, "lightning_b200", "libcln_sigverify.so")
    needed = subprocess.run(["ldd", lib], capture_output=True, text=True).stdout
    assert "secp" not in needed and "emul" not in needed
