"""The C-ABI library loads and exports every symbol include/mfemad_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            txt = open(os.path.join(ROOT, "include", fn)).read()
            txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
            names |= set(re.findall(r"\b(madb_[a-z0-9_]+)\s*\(", txt))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    import mfem_ad_b200 as M
    L = ctypes.CDLL(M.LIB_PATH)
    decl = _declared()
    assert len(decl) >= 20
    for n in decl:
        assert hasattr(L, n), "libmadb.so does not export %s" % n
    assert M.lib().madb_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import mfem_ad_b200 as M
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(M.MadbError, match="no CPU fallback"):
        M.Context(0)


def test_product_does_not_import_oracle():
    # the oracle is test infrastructure: nothing under the package may reference it
    pkg = os.path.join(ROOT, "mfem-ad_b200")
    for dp, _, fns in os.walk(pkg):
        if os.path.basename(dp) == "build":
            continue
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert not re.search(r"import\s+oracle|from\s+oracle|liboracle|oracle[/.](oracle|py|cpp)|orc_", txt), \
                    "%s uses the oracle" % fn
