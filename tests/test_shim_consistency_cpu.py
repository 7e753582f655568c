"""The Fortran interface module (shim/decipher_c_api.f90) against the C header it mirrors.

No Fortran compiler exists in the build image, so the module cannot be compiled here; what CAN be checked mechanically is
what a compiler would not catch anyway: that every `type, bind(C)` lists the same components in the same order and kinds as
the C struct, and that every `bind(C, name=...)` interface names an exported symbol with the same number of arguments,
scalars passed by `value` where C takes them by value and pointers where C takes pointers."""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HEADER = open(os.path.join(ROOT, "include", "decipher_b200.h")).read()
SHIM = open(os.path.join(ROOT, "shim", "decipher_c_api.f90")).read()


def c_structs():
    out = {}
    for m in re.finditer(r"typedef struct (\w+) \{(.*?)\}\s*\w+;", HEADER, re.S):
        body = re.sub(r"/\*.*?\*/", "", m.group(2), flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            mm = re.match(r"(const\s+)?(\w+)\s*(\*?)\s*(\w+)(\[.*\])?$", decl)
            assert mm, decl
            kind = "ptr" if mm.group(3) else {"int32_t": "i32", "int": "i32", "int64_t": "i64", "double": "f64"}[mm.group(2)]
            fields.append((kind, mm.group(4), bool(mm.group(5))))
        out[m.group(1)] = fields
    return out


def f_types():
    out = {}
    src = "\n".join(line.split("!")[0] for line in SHIM.splitlines())
    for m in re.finditer(r"type\s*,\s*bind\(C\)\s*::\s*(\w+)(.*?)end type", src, re.S | re.I):
        fields = []
        for line in m.group(2).splitlines():
            line = line.strip()
            if not line:
                continue
            mm = re.match(r"(integer\((\w+)\)|real\((\w+)\)|type\(c_ptr\))\s*::\s*(.*)$", line, re.I)
            assert mm, line
            kind = ("ptr" if mm.group(1).lower().startswith("type") else
                    {"c_int32_t": "i32", "c_int": "i32", "c_int64_t": "i64", "c_double": "f64"}[(mm.group(2) or mm.group(3)).lower()])
            for name in mm.group(4).split(","):
                name = name.strip()
                arr = "(" in name
                fields.append((kind, name.split("(")[0].strip(), arr))
        out[m.group(1)] = fields
    return out


def test_bind_c_types_mirror_the_c_structs():
    cs, fs = c_structs(), f_types()
    assert {"dcp_catchment_desc", "dcp_run_opts", "dcp_run_stats", "dcp_multi_stats"} <= set(fs)
    for name, fields in fs.items():
        assert name in cs, name
        assert fields == cs[name], (name, [a for a, b in zip(fields, cs[name]) if a != b][:3], len(fields), len(cs[name]))


def c_prototypes():
    text = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    out = {}
    for m in re.finditer(r"^(?:const\s+)?\w+\s*\*?\s*(dcp_\w+)\s*\(([^;{]*?)\)\s*;", text, re.M | re.S):
        args = " ".join(m.group(2).split())
        params = [] if args in ("void", "") else [a.strip() for a in args.split(",")]
        out[m.group(1)] = ["ptr" if "*" in a else "val" for a in params]
    return out


def f_interfaces():
    src = "\n".join(line.split("!")[0] for line in SHIM.splitlines())
    src = re.sub(r"&\s*\n\s*&?", " ", src)
    out = {}
    for m in re.finditer(r"function\s+(\w+)\s*\((.*?)\)\s*bind\(C,\s*name=\"(\w+)\"\)(.*?)end function", src, re.S | re.I):
        fname, args, cname, body = m.group(1), m.group(2), m.group(3), m.group(4)
        names = [a.strip() for a in args.split(",") if a.strip()]
        kinds = {}
        for line in body.splitlines():
            mm = re.match(r"\s*(.*?)::\s*(.*)$", line)
            if not mm or line.strip().lower().startswith("import"):
                continue
            spec = mm.group(1).lower()
            for n in mm.group(2).split(","):
                n = n.strip().split("(")[0]
                # a C pointer is either type(c_ptr), value or any dummy passed by reference (no value attribute)
                kinds[n] = "val" if ("value" in spec and "c_ptr" not in spec) else "ptr"
        assert fname == cname, (fname, cname)
        out[cname] = [kinds[n] for n in names]
    return out


def test_interfaces_name_exported_symbols_with_matching_arguments():
    protos, ifaces = c_prototypes(), f_interfaces()
    assert len(ifaces) >= 15
    lib_path = os.path.join(ROOT, "decipher_b200", "csrc", "build", "libdecipher_b200.so")
    lib = ctypes.CDLL(lib_path) if os.path.exists(lib_path) else None
    for name, kinds in ifaces.items():
        assert name in protos, f"{name} is not declared in include/decipher_b200.h"
        assert kinds == protos[name], (name, kinds, protos[name])
        if lib is not None:
            assert hasattr(lib, name), f"{name} is not exported by the library"


import pytest  # noqa: E402
import sys  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(HERE, "golden"))


def _ref_present():
    import make_f90_vectors
    return make_f90_vectors.reference_available()


@pytest.mark.skipif(not _ref_present(), reason="reference sources are not on this machine")
def test_shim_main_program_is_consistent_with_the_reference_modules():
    """shim/dyna_main_gpu.f90 (dyna_main.f90 with the Monte-Carlo loop replaced) run through the Fortran-subset translator
    together with the reference's modules and the interface module: every name it uses is declared or provided by a module,
    blocks balance, every call of a reference procedure (project, Tread_dyna, Inputs, mc_setup, genpar, file_open ...) passes
    as many arguments as the reference's definition takes.  Not a compiler, but the errors a first compile would throw."""
    import f90_translate as ft
    import make_f90_vectors as mk
    prog = ft.Program()
    for f in mk.FILES:
        prog.add_file(os.path.join(mk.REFERENCE, f))
    prog.add_file(os.path.join(ROOT, "shim", "decipher_c_api.f90"))
    prog.add_file(os.path.join(ROOT, "shim", "dyna_main_gpu.f90"))            # its `program main` replaces the reference's
    prog.finish()
    mine = [p for p in prog.procs.values() if p.src.endswith("dyna_main_gpu.f90")]
    assert mine and any(p.name == "main" for p in mine)
    for p in mine:
        assert not p.broken, (p.name, p.broken)
        try:
            ft.Gen(prog, p).gen_proc()
        except NotImplementedError as e:
            assert "host association" in str(e), (p.name, e)          # internal procedures reading the program's variables
    assert {"dcp_run_ensemble", "dcp_catchment_create"} <= prog.externals


@pytest.mark.skipif(not _ref_present(), reason="reference sources are not on this machine")
@pytest.mark.parametrize("name", ["p_tree", "p_missing_flow"])
def test_shim_main_program_executed_against_the_c_abi_call_sequence(name, tmp_path):
    """shim/dyna_main_gpu.f90 EXECUTED (translator + the reference's own modules) on a project directory.  Its three calls into
    the library are bound to stand-ins that rebuild the catchment from the bind(C) descriptor the Fortran code filled in and
    hand it to the oracle, so what is tested is the Fortran side of the boundary: the flattening of dyna_hru / route_riv, index
    and array-order conventions, one create + one run for the whole ensemble, mcpar in/out, and the per-member writer.  The
    flows and totals it writes must equal what the UNMODIFIED reference program writes on the same files (p_*.npz)."""
    import numpy as np
    sys.path.insert(0, HERE)
    import f90_cases
    import shim_exec
    from oracle_binding import oracle_run
    ns = shim_exec.load()
    root = str(tmp_path)
    f90_cases.write_project(name, root)
    r = shim_exec.run_main(ns, root, lambda cat, mc: oracle_run(cat, mc))
    v = np.load(os.path.join(HERE, "golden", "f90_exec", name + ".npz"), allow_pickle=False)
    assert r["n_create"] == 1 and r["n_run"] == 1                          # ONE call replaces the Monte-Carlo loop
    assert r["q"].tobytes() == v["q"].tobytes() and r["totals"].tobytes() == v["totals"].tobytes()
    assert r["flow_files"] == [str(f) for f in v["flow_files"]]
