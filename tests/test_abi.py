"""The C-ABI library loads and exports every symbol include/dip.h declares (no compute calls: no GPU needed)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    import dip_engine as de
    if not os.path.exists(de.LIB_PATH):
        de.build()
    lib = ctypes.CDLL(de.LIB_PATH)
    hdr = open(os.path.join(ROOT, "include", "dip.h")).read()
    declared = set(re.findall(r"\b(dip_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"dip_plan_bind"} - declared  # no-op; keeps the set explicit
    assert declared == set(de.ABI_SYMBOLS), declared ^ set(de.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert de.lib().dip_version() == 100


def test_workspace_query_needs_no_gpu():
    import dip_engine as de
    desc = de.NetDesc(32, 3, 5, 128, 4, 1, 1, 0)
    n = de.lib().dip_plan_workspace_bytes(ctypes.byref(desc), 512, 512)
    assert 2 ** 30 < n < 12 * 2 ** 30
    bad = de.NetDesc(32, 3, 5, 64, 4, 1, 1, 0)
    assert de.lib().dip_plan_workspace_bytes(ctypes.byref(bad), 512, 512) == 0
    assert b"128" in de.lib().dip_last_error()
