"""The C-ABI library loads and exports every symbol include/gpvb200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "gpvb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gpvb200_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_all_symbols():
    import __graft_entry__ as g
    g.build()
    from gpv1_b200 import _C
    L = _C.lib()
    names = _declared()
    assert len(names) >= 5
    for n in names:
        assert hasattr(L, n), f"{n} declared in gpvb200.h but not exported"
    assert L.gpvb200_version() >= 100
    assert L.gpvb200_gemm_desc_size() == ctypes.sizeof(_C.GemmDesc)


def test_no_cpu_fallback_in_product_package():
    """The product package must never import the oracle."""
    pkg = os.path.join(ROOT, "gpv-1_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
