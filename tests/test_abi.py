"""CPU: the C-ABI library loads and exports every symbol include/zs3b200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "zs3b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zs3_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from zs3_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build with `python -c 'import __graft_entry__ as g; g.build()'`"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_covers_header_and_reports_errors():
    from zs3_b200 import _lib
    lib = _lib.lib()
    bound = set(lib._zs3_declared)
    declared = set(_declared_symbols())
    assert declared <= bound, sorted(declared - bound)
    assert lib.zs3_abi_version() == 1
    # argument validation happens before any CUDA call: usable without a GPU
    rc = lib.zs3_conv_fprop(None, None)
    assert rc == -1 and b"null args" in lib.zs3_last_error()
    a = _lib.ConvArgs()
    a.num_segments = 9
    rc = lib.zs3_conv_fprop(ctypes.byref(a), None)
    assert rc == -1 and b"num_segments" in lib.zs3_last_error()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import importlib
    from zs3_b200 import _lib
    importlib.reload(_lib)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    import pytest
    with pytest.raises(_lib.Zs3NativeError):
        _lib.lib()
    importlib.reload(_lib)


def test_product_code_never_imports_the_oracle():
    bad = []
    for base in ("zs3_b200", "zs3"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if "zs3_oracle" in txt or "import oracle" in txt or "from oracle" in txt:
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
