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


def test_host_side_queries_need_no_gpu():
    """Size / mode queries are host logic: they must answer on a GPU-less build host (no kernel is launched)."""
    from boosted_detr_b200 import _lib
    lib = _lib.load()
    # fp16 attention workspace: q16 | k16 | v16 halves for shapes the long-sequence kernel serves, 0 otherwise
    assert lib.bdetr_attention_f16_workspace_bytes(4, 8, 20020, 20020, 32) == 2 * 3 * 4 * 20020 * 256
    assert lib.bdetr_attention_f16_workspace_bytes(16, 8, 400, 400, 32) == 0          # config 2 keeps the TF32 kernels
    assert lib.bdetr_attention_f16_workspace_bytes(4, 8, 20020, 20020, 64) == 0          # head dim must be 32
    assert lib.bdetr_attention_f16_workspace_bytes(0, 8, 20020, 20020, 32) == 0
    # the three compute modes round-trip through set / get; BDETR_MODE_FP16 is the tensor-core mode plus the operand switch
    try:
        for mode in (_lib.MODE_TF32, _lib.MODE_FP16, _lib.MODE_FP32):
            assert lib.bdetr_set_mode(mode) == 0 and lib.bdetr_get_mode() == mode
            assert _lib.tc_mode() == (mode != _lib.MODE_FP32)
    finally:
        lib.bdetr_set_mode(_lib.MODE_FP32)
    assert lib.bdetr_cost_targets_bytes(256, 100, 82, 3) > 0 and lib.bdetr_cost_targets_bytes(0, 100, 82, 3) == 0
