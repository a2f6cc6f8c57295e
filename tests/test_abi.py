"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/merzbild_b200.h
declares, the ctypes mirror binds exactly that list, and the product path fails loudly without a CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "merzbild_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mb_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(mb):
    L = C.CDLL(mb.LIB_PATH)
    names = _declared()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_ctypes_mirror_binds_the_header(mb):
    assert sorted(mb.SIGNATURES) == _declared()
    mb.lib()
    assert mb.lib().mb_version() >= 100


def test_struct_layouts(mb):
    assert C.sizeof(mb.Grid1D) == 56 and C.sizeof(mb.Walls1D) == 80 and C.sizeof(mb.Interaction) == 64 and C.sizeof(mb.OctreeParams) == 24


def test_host_helpers_match_oracle(mb, oracle):
    """Grid1DUniform / Interaction / sigma_g_w_max estimate are host arithmetic: identical to the oracle restatement."""
    g = mb.Grid1DUniform(5e-4, 50)
    o = oracle.grid_params(5e-4, 50)
    assert (g.dx, g.inv_dx, g.min_x, g.max_x) == (o["dx"], o["inv_dx"], o["min_x"], o["max_x"])
    m = oracle.MASS["Ar"]
    it = mb.make_interaction(m, m, 4.11e-10, 0.81, 273.0)
    ot = oracle.make_interaction(m, m, 4.11e-10, 0.81, 273.0)
    assert [it.m_r, it.mu1, it.mu2, it.vhs_d, it.vhs_o, it.vhs_Tref, it.vhs_muref, it.vhs_factor] == ot.tolist()
    assert mb.estimate_sigma_g_w_max(it, m, m, 300.0, 300.0, 1e10) == oracle.estimate_sigma_g_w_max(ot, m, m, 300.0, 300.0, 1e10)
    # slab partition == ChunkSplitters.chunks(1:nx; n): first nx mod n slabs one longer
    G = mb.Grid1DUniform(1.0, 10)
    sl = [G.slab(r, 4) for r in range(4)]
    assert [s.n_cells for s in sl] == [3, 3, 2, 2] and [s.cell_offset for s in sl] == [0, 3, 6, 8]


def test_no_cpu_fallback(mb):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(mb.MerzbildError) as e:
        mb.Context(0, 1)
    assert e.value.status == mb.MB_ERR_NO_DEVICE


def _julia_ccalls():
    """(symbol, n_argument_types) for every ccall of the Julia shim."""
    src = open(os.path.join(ROOT, "merzbild.jl_b200", "julia", "MerzbildB200.jl")).read()
    src = re.sub(r"#[^\n]*", "", src)
    out = {}
    for m in re.finditer(r"ccall\(\(:(mb_[A-Za-z0-9_]+),\s*libmb\),\s*[A-Za-z0-9_{}]+,\s*\(", src):
        i, depth = m.end(), 1
        while depth:  # the tuple of argument types, with nested Ptr{...} / NTuple{...}
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        body = src[m.end():i - 1]
        flat, d = "", 0
        for ch in body:
            d += {"{": 1, "}": -1}.get(ch, 0)
            flat += ch if not (ch == "," and d) else ";"
        n = len([t for t in flat.split(",") if t.strip()])
        out.setdefault(m.group(1), set()).add(n)
    return out


def test_julia_shim_binds_the_header(mb):
    """The Julia shim cannot run here (no Julia toolchain): check statically that it binds every symbol the header declares
    and that each ccall passes as many arguments as the ctypes mirror (which the GPU tests exercise)."""
    calls = _julia_ccalls()
    assert sorted(calls) == _declared()
    for name, arities in calls.items():
        assert arities == {len(mb.SIGNATURES[name][1])}, (name, arities, len(mb.SIGNATURES[name][1]))


def _julia_ccall_types():
    """(symbol -> list of argument-type tuples) for every ccall of the Julia shim."""
    src = open(os.path.join(ROOT, "merzbild.jl_b200", "julia", "MerzbildB200.jl")).read()
    src = re.sub(r"#[^\n]*", "", src)
    out = {}
    for m in re.finditer(r"ccall\(\(:(mb_[A-Za-z0-9_]+),\s*libmb\),\s*([A-Za-z0-9_{}]+),\s*\(", src):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        body = src[m.end():i - 1]
        flat, d = "", 0
        for ch in body:
            d += {"{": 1, "}": -1}.get(ch, 0)
            flat += ch if not (ch == "," and d) else ";"
        types = tuple(t.strip() for t in flat.split(",") if t.strip())
        out.setdefault(m.group(1), []).append((m.group(2), types))
    return out


def test_julia_shim_argument_types_match_the_ctypes_mirror(mb):
    """Position by position, every ccall of the Julia shim passes the same KIND of argument (pointer, Int64, Int32, UInt32, UInt64,
    Float64, Cint) as the ctypes signature the GPU tests exercise, and declares the same return kind."""
    def kind_py(t):
        if t in (C.c_void_p, C.c_char_p) or (isinstance(t, type) and issubclass(t, C._Pointer)):
            return "ptr"
        return {C.c_int64: "i64", C.c_int32: "i32", C.c_uint32: "u32", C.c_uint64: "u64", C.c_double: "f64", C.c_int: "i32"}[t]

    def kind_jl(t):
        if t.startswith("Ptr{") or t.startswith("Ref{") or t in ("Cstring",):
            return "ptr"
        return {"Int64": "i64", "Int32": "i32", "Cint": "i32", "UInt32": "u32", "UInt64": "u64", "Float64": "f64", "Cdouble": "f64"}[t]

    calls = _julia_ccall_types()
    assert sorted(calls) == _declared()
    for name, variants in calls.items():
        res, args = mb.SIGNATURES[name]
        want = [kind_py(t) for t in args]
        for ret, types in variants:
            got = [kind_jl(t) for t in types]
            assert got == want, (name, got, want)
            assert kind_jl(ret) == kind_py(res), (name, ret, res)
