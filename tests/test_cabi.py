"""The C-ABI library loads and exports every symbol include/qcxms_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "qcxms_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qcxms_b200_\w+)\s*\(", text)))


def test_header_symbols_are_exported():
    import __graft_entry__ as ge
    ge.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "qcxms_b200", "libqcxms_b200.so"))
    names = _declared()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), n


def test_python_mirror_lists_the_same_entry_points():
    from qcxms_b200 import api
    assert sorted(api.EXPORTS) == _declared()
    assert api.lib().qcxms_b200_version().decode().startswith("qcxms_b200")
    # POD layouts agree with the header (6 int32 + 4 double / 6 int32 + 8 double)
    assert ctypes.sizeof(api.MdConfig) == 6 * 4 + 4 * 8 and ctypes.sizeof(api.MdResult) == 6 * 4 + 8 * 8
    assert (api.gfn1_xtb, api.gfn2_xtb, api.ipea1_xtb) == (1, 2, 11)     # reference src/tblite.f90:34-40


def test_product_never_touches_the_oracle():
    """A product path that routes through oracle/ would void every parity claim."""
    pkg = os.path.join(ROOT, "qcxms_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f)).read()
                for pat in ("import oracle", "from oracle", "pyoracle", "liboracle", "xtb_oracle", "md_oracle", "oracle/"):
                    assert pat not in src, (os.path.join(dp, f), pat)
