"""The C-ABI library loads on a CPU-only box and exports every symbol include/viddet_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

import viddet_b200
from viddet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "viddet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vd_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 14
    lib = ctypes.CDLL(_lib.SO_PATH)
    for n in names:
        assert hasattr(lib, n), "missing export " + n
        assert n in _lib.SIGNATURES, "no ctypes signature for " + n
    assert sorted(_lib.SIGNATURES) == names


def test_load_and_version():
    lib = viddet_b200.load()
    assert lib.vd_version() == 100
    assert lib.vd_last_error() is not None


def test_struct_layout_matches_header():
    # VdHeadScale: 3 ptr + 3 int + float + 6 float + 4 ptr ; natural alignment.  VdHeadParams: 12 x 4-byte scalars + 3 scales
    assert ctypes.sizeof(_lib.VdHeadScale) == 104
    assert ctypes.sizeof(_lib.VdHeadParams) == 48 + 3 * 104 + 8 + 7 * 8
    lib = viddet_b200.load()                         # the library's own sizeof of the compiled header
    assert lib.vd_sizeof(0) == ctypes.sizeof(_lib.VdHeadScale) and lib.vd_sizeof(1) == ctypes.sizeof(_lib.VdHeadParams)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "viddet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
