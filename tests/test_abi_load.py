"""The C-ABI library loads and exports every symbol include/bgreat_b200.h declares (no GPU needed)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "bgreat_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bgr_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from bgreat_b200 import _abi as A
    L = A.lib()
    declared = _declared_symbols()
    assert declared, "no declarations found"
    for s in declared:
        assert hasattr(L, s), s
    assert sorted(A.SYMBOLS) == declared
    assert L.bgr_abi_version() == 2


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device context creation must fail loudly, never degrade to a host path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import bgreat_b200
    with pytest.raises(bgreat_b200.BgrError):
        bgreat_b200.Aligner(31)


def test_product_does_not_reference_oracle_or_emulator():
    pkg = os.path.join(ROOT, "bgreat_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle/" not in src and "liboracle" not in src and "bgr_emu" not in src, f
                assert "import oracle" not in src and "from oracle" not in src, f
