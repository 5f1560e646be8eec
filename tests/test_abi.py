"""CPU: the C-ABI library loads and exports every symbol include/erlfill.h declares; struct layouts
match between the header (as compiled into the oracle-independent ctypes mirror) and the binding."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "erlfill.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(erlfill_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import erlfill_b200
    from erlfill_b200 import _lib
    assert os.path.exists(_lib.SO_PATH), "build with python erl-fill_b200/build.py"
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 7
    for n in names:
        assert hasattr(L, n), n
    assert sorted(_lib.EXPORTS) == names
    assert L.erlfill_version() == 100


def test_struct_layouts():
    from erlfill_b200 import _lib
    assert C.sizeof(_lib.Condition) == 104
    assert _lib.Condition.lo.offset == 24 and _lib.Condition.hi.offset == 64
    assert C.sizeof(_lib.EnvBatch) == 32 + 8 * 8
    assert C.sizeof(_lib.StepOut) == 7 * 8
    assert C.sizeof(_lib.Noise) == 24


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from erlfill_b200 import _lib
    with pytest.raises(_lib.ErlfillError):
        _lib.require_device()
    import erlfill_b200
    with pytest.raises(_lib.ErlfillError):
        erlfill_b200.VecFillEnv([erlfill_b200.conditions.SINGLE_25KG], 4, kind="single")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "erl-fill_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle_lib" not in txt and "import oracle" not in txt and "from oracle" not in txt, f
