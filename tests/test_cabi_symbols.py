"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and
exports every function include/mchap_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "mchap_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mchb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    names = declared_functions()
    for required in ("mchb_create", "mchb_assemble_batch", "mchb_log_likelihood_batch",
                     "mchb_genotype_rank", "mchb_genotype_unrank", "mchb_mt19937_words"):
        assert required in names


def test_library_builds_loads_and_exports_every_symbol():
    from mchap_b200 import build, _lib

    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in declared_functions():
        assert hasattr(lib, name), "missing export: %s" % name
    assert sorted(_lib.SYMBOLS) == declared_functions()


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import mchap_b200

    with pytest.raises(mchap_b200.MchapB200Error):
        mchap_b200.Device(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mchap_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(base, f)).read()
                assert "oracle" not in src.replace("# oracle-free", ""), os.path.join(base, f)


def test_struct_layouts_match_the_header():
    from mchap_b200 import _lib

    assert ctypes.sizeof(_lib.AssembleItem) == 6 * 8 + 8 * 4 + 8
    assert ctypes.sizeof(_lib.ItemResult) == 24
    assert ctypes.sizeof(_lib.LlkItem) == 3 * 8 + 4 * 4
    assert _lib.AssembleItem.inbreeding.offset == 80
    assert ctypes.sizeof(_lib.TallyItem) == 3 * 8 + 6 * 4
    assert _lib.TallyItem.max_unique.offset == 44
    assert ctypes.sizeof(_lib.EncodeItem) == 5 * 8 + 4 * 4
