"""CPU-side checks of the drop-in boundary: the library builds, loads and exports every symbol that
include/bdetr.h declares, and the ctypes prototypes cover all of them (no compute calls here)."""
import os
import re

from util import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "bdetr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bdetr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from boosted_detr_b200 import _lib, build
    build.build_library()
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bdetr.h but not exported by libbdetr.so"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
    assert set(_lib.PROTOTYPES) == set(names)
    assert lib.bdetr_version() >= 100


def test_mode_and_error_text():
    from boosted_detr_b200 import _lib
    lib = _lib.load()
    assert lib.bdetr_set_mode(_lib.MODE_FP32) == 0 and lib.bdetr_get_mode() == _lib.MODE_FP32
    assert lib.bdetr_set_mode(77) == _lib.BDETR_E_UNSUPPORTED
    assert b"unknown mode" in lib.bdetr_last_error()
    assert lib.bdetr_lsap_smem_bytes(100, 300) == 300 * (3 * 8 + 5 * 4)


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, "boosted_detr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", ""), f"{f} must not reference oracle/"
