"""CPU suite: the C-ABI library builds for sm_100a, loads, and exports every symbol the public header declares.
No kernel is launched here."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "de6d_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(de6d_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    syms = _declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "libde6d_b200.so does not export %s" % s


def test_python_prototypes_cover_header():
    from de6d_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == _declared_symbols()


def test_library_is_sm100a_and_torch_free(lib):
    import subprocess
    from de6d_b200 import _lib
    r = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert "sm_100a" in r.stdout, r.stdout[-500:]
    ldd = subprocess.run(["ldd", _lib.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "torch" not in ldd and "c10" not in ldd


def test_info_and_error_plumbing(lib):
    assert lib.de6d_version() == 100
    assert b"sm_100a" in lib.de6d_build_info()
    # argument validation happens before any CUDA call: usable without a GPU
    rc = lib.de6d_furthest_point_sampling(-1, 4, 2, None, None, None, None)
    assert rc == 1 and b"negative" in lib.de6d_last_error_string()
    rc = lib.de6d_nms_batched(1, 8, None, None, ctypes.c_float(0.1), 0, None, None, None, 0, None)
    assert rc == 1
    assert lib.de6d_nms_workspace_bytes(2, 512) >= 2 * 512 * 8 * 8


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from de6d_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    import pytest
    with pytest.raises(ImportError, match="no CPU or PyTorch fallback"):
        _lib.load()
