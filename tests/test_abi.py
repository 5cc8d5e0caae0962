"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/sc_b200.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "sc_b200.h")).read()
    return sorted(set(re.findall(r"\b(sc_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    import scb200

    L = scb200.load_library()
    names = _declared()
    assert set(names) == set(scb200.DECLARED_SYMBOLS)
    for n in names:
        assert hasattr(L, n), n


def test_struct_layout_matches_header():
    import scb200

    assert scb200.POSITION_DTYPE.itemsize == 544
    assert scb200.POSITION_DTYPE.fields["meta"][1] == 512 and scb200.POSITION_DTYPE.fields["n_hist"][1] == 540
    assert scb200.MOVE_DTYPE.itemsize == 4


def test_no_cpu_fallback(tmp_path):
    import torch

    import scb200

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(scb200.SCError) as ei:
        scb200.Engine(str(tmp_path / "missing.scw"), 0, scb200.SC_MODE_FP32, 8)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_does_not_import_oracle():
    # the product path must never route through oracle/
    for dirpath, _, files in os.walk(os.path.join(ROOT, "smart-chess-rust_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "chess_oracle" not in txt and "oracle/" not in txt and "import net" not in txt, f
