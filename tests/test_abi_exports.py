"""CPU: libagb200.so loads and exports every symbol include/agb200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "agb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"ag_status\s+(ag_\w+)\s*\(", src)))


def test_header_declares_something():
    syms = header_symbols()
    assert len(syms) >= 40 and "ag_gemm" in syms and "ag_scatter_add" in syms


def test_library_exports_every_header_symbol(agpkg):
    lib = ctypes.CDLL(agpkg._lib.LIB_PATH)
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_binding_covers_header(agpkg):
    assert sorted(agpkg._lib.EXPORTS) == header_symbols()


def test_uninitialised_calls_fail_loudly(agpkg):
    """Without ag_init (no GPU here) compute entry points return AG_ESTATE, never a CPU result."""
    lib = ctypes.CDLL(agpkg._lib.LIB_PATH)
    lib.ag_unary_fwd.restype = ctypes.c_int32
    st = lib.ag_unary_fwd(4, 0, None, None, ctypes.c_int64(4))
    assert st == agpkg._lib.ESTATE
    buf = ctypes.create_string_buffer(256)
    lib.ag_last_error(buf, ctypes.c_int64(256))
    assert b"ag_init" in buf.value


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "autograd.jl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dp, f)).read()
                assert "import oracle" not in s and "from oracle" not in s and "agref" not in s, f


def test_fused_instruction_set_agrees_between_host_and_kernel(agpkg):
    """The opcodes / operand modes / limits of the fused evaluator are written down three times (csrc/fuse.cu, the host
    code generator lazy.py, the numpy stand-in): they must agree."""
    import re
    import sys
    L = sys.modules["agb200.lazy"]
    src = open(os.path.join(ROOT, "autograd.jl_b200", "csrc", "fuse.cu")).read()
    enum = dict((k, int(v)) for k, v in re.findall(r"\bFZ_([A-Z_0-9]+) = (\d+)", src))
    for name in ("LD_IN", "LD_CONST", "LD_TMP", "ST_TMP", "UNARY", "BIN", "BIN_IN", "BIN_CONST", "UGRAD", "TERN", "STORE",
                 "REDUCE"):
        assert enum[name] == getattr(L, name), name
    for name in ("M_FULL", "M_DEVSCALAR", "M_IMM", "M_COLVEC", "M_ROWVEC"):
        assert enum[name] == getattr(L, name), name
    lim = dict((k, int(v)) for k, v in re.findall(r"FZ_(MAXCODE|MAXCONST|MAXIN|MAXOUT|SLOTS|DYN_MAXIN|DYN_MAXOUT) = (\d+)", src))
    assert (lim["MAXCODE"], lim["MAXCONST"], lim["MAXIN"], lim["MAXOUT"], lim["SLOTS"]) == \
        (L.MAXCODE, L.MAXCONST, L.MAXIN, L.MAXOUT, L.SLOTS)
    assert (lim["DYN_MAXIN"], lim["DYN_MAXOUT"]) == (L.DYN_MAXIN, L.DYN_MAXOUT)
    hdr = open(os.path.join(ROOT, "include", "agb200.h")).read()
    D = sys.modules["agb200.darray"]
    for name, code in D.U.items():                       # unary op ids of the host table == header enum
        m = re.search(r"AG_U_%s = (\d+)" % name.upper(), hdr)
        assert m and int(m.group(1)) == code, name
