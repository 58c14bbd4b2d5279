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


# ---- round 2: signatures, not only names -----------------------------------------------------------------------

def header_prototypes():
    """{name: [type class per parameter]} parsed from include/agb200.h; classes: i32 i64 u64 f64 ptr."""
    src = open(os.path.join(ROOT, "include", "agb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for name, params in re.findall(r"ag_status\s+(ag_\w+)\s*\(([^)]*)\)\s*;", src):
        params = params.strip()
        cls = []
        if params and params != "void":
            for p in params.split(","):
                p = " ".join(p.split())
                if "*" in p:
                    cls.append("ptr")
                elif re.search(r"\bint32_t\b", p):
                    cls.append("i32")
                elif re.search(r"\buint64_t\b", p):
                    cls.append("u64")
                elif re.search(r"\bint64_t\b", p):
                    cls.append("i64")
                elif re.search(r"\bdouble\b", p):
                    cls.append("f64")
                else:
                    raise AssertionError("unclassified parameter %r of %s" % (p, name))
        out[name] = cls
    return out


def _ctypes_class(t):
    if t in (ctypes.c_int32,):
        return "i32"
    if t in (ctypes.c_int64,):
        return "i64"
    if t in (ctypes.c_uint64,):
        return "u64"
    if t in (ctypes.c_double,):
        return "f64"
    if t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "_type_") and issubclass(t, ctypes._Pointer):
        return "ptr"
    raise AssertionError("unclassified ctypes type %r" % (t,))


def test_ctypes_signatures_match_the_header(agpkg):
    """Arity and C type of every parameter: _lib._SIGS (what the Python host passes) against the prototypes."""
    protos = header_prototypes()
    assert set(protos) == set(agpkg._lib._SIGS)
    for name, args in agpkg._lib._SIGS.items():
        assert [_ctypes_class(t) for t in args] == protos[name], name


_JL_CLASS = {"Int32": "i32", "Int64": "i64", "UInt64": "u64", "Float64": "f64"}


def julia_ccalls():
    """[(symbol, return type, [type class per argument], line)] for every ccall in julia/AutoGradB200.jl."""
    src = open(os.path.join(ROOT, "julia", "AutoGradB200.jl")).read()
    out = []
    for m in re.finditer(r"ccall\(\(:(ag_\w+), LIB\),\s*(\w+),\s*\(", src):
        i, depth = m.end(), 1
        while depth:                                    # the argument-type tuple, with nested Ptr{...} braces
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        tup = src[m.end():i - 1]
        parts, cur, br = [], "", 0
        for ch in tup:
            if ch == "," and br == 0:
                parts.append(cur.strip())
                cur = ""
            else:
                br += {"{": 1, "}": -1}.get(ch, 0)
                cur += ch
        if cur.strip():
            parts.append(cur.strip())
        cls = []
        for t in parts:
            if t.startswith(("Ptr{", "Ref{")):
                cls.append("ptr")
            else:
                cls.append(_JL_CLASS[t])
        out.append((m.group(1), m.group(2), cls, src.count("\n", 0, m.start()) + 1))
    return out


def test_julia_ccalls_match_the_header():
    """Every ccall of the Julia binding: the symbol exists in the header, returns Int32 (ag_status) and passes the
    prototype's number and kinds of arguments.  (The file cannot be executed here; this is its mechanical check.)"""
    protos = header_prototypes()
    calls = julia_ccalls()
    assert len(calls) >= 60
    for name, ret, cls, line in calls:
        assert name in protos, "julia/AutoGradB200.jl:%d calls %s, which include/agb200.h does not declare" % (line, name)
        assert ret == "Int32", (name, line)
        assert cls == protos[name], "julia/AutoGradB200.jl:%d %s: %s vs header %s" % (line, name, cls, protos[name])


def test_integration_table_names_existing_julia_methods():
    """Every `code` identifier in the 'Julia method' column of INTEGRATION.md's table is defined in the .jl."""
    jl = open(os.path.join(ROOT, "julia", "AutoGradB200.jl")).read()
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    rows = [r for r in md.splitlines() if r.startswith("|") and "`ag_" in r]
    assert len(rows) >= 15
    for r in rows:
        cols = [c.strip() for c in r.strip("|").split("|")]
        for ident in re.findall(r"`([A-Za-z_!@][\w!@.]*)", cols[1]):
            base = ident.split(".")[-1].lstrip("@")
            if base in ("Value", "DArray", "Arg", "Number", "Sparse", "Param", "Colon", "IArray", "typeof", "recording", "ids",
                        "dy", "y", "x", "f", "true", "false", "dims", "perm", "AutoGradB200", "AutoGrad"):
                continue
            assert re.search(r"(?<![\w])%s(?![\w!])" % re.escape(base), jl), "INTEGRATION.md names `%s`, not found in the .jl: %s" % (ident, r[:80])
        for sym in re.findall(r"`(ag_\w+)`", cols[2]):
            assert "(:%s, LIB)" % sym in jl, "INTEGRATION.md maps a row to %s but the .jl never calls it: %s" % (sym, r[:80])


def test_cabi_smoke_compiles_against_the_header(agpkg, tmp_path):
    """tests/cabi_smoke.c is plain C: it must compile with gcc against include/agb200.h and link with the library (the
    run itself needs a GPU: tests/test_gpu_round2.py::test_cabi_smoke_runs)."""
    import subprocess
    exe = str(tmp_path / "cabi_smoke")
    libdir = os.path.dirname(agpkg._lib.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "cabi_smoke.c"), "-o", exe, "-L", libdir, "-lagb200",
                        "-Wl,-rpath," + libdir, "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
