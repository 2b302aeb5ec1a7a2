"""Text-level checks of the Fortran shim (cice4_b200/fortran/ice_dyn_evp_b200.F90) against the C header.

No Fortran compiler exists in this image, so the shim cannot be compiled here; what CAN be checked is that
its `bind(C)` derived types mirror the structs of include/evp_b200.h field for field (name, order and C
type), that every `bind(C, name=...)` interface names an exported entry point with the right number of
arguments, that the public names other CICE units import from `module ice_dyn_evp` are all there
(/root/reference/source/ice_init.F90:91,97; ice_step_mod.F90:575; ice_history.F90:1939), and that the
block structure of the file is balanced."""
import os
import re

from conftest import ROOT

SHIM = os.path.join(ROOT, "cice4_b200", "fortran", "ice_dyn_evp_b200.F90")
HEADER = os.path.join(ROOT, "include", "evp_b200.h")


def _strip_c_comments(src):
    return re.sub(r"/\*.*?\*/", "", src, flags=re.S)


def _header_structs():
    src = _strip_c_comments(open(HEADER).read())
    out = {}
    for body, name in re.findall(r"typedef struct \{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            m = re.match(r"(const\s+)?(double|int32_t|float)\s+(.*)", decl)
            assert m, decl
            ctype = m.group(2)
            for item in m.group(3).split(","):
                item = item.strip()
                ptr = item.startswith("*")
                fields.append((item.lstrip("*").strip(), "ptr" if ptr else ctype))
        out[name] = fields
    return out


def _join_continuations(lines):
    out, cur = [], ""
    for ln in lines:
        ln = ln.split("!")[0].rstrip() if not ln.lstrip().startswith("#") else ln.rstrip()
        if not ln.strip():
            continue
        if ln.rstrip().endswith("&"):
            cur += ln.rstrip()[:-1] + " "
            continue
        out.append(cur + ln)
        cur = ""
    return out


def _shim_types():
    lines = _join_continuations(open(SHIM).read().splitlines())
    out, name = {}, None
    for ln in lines:
        s = ln.strip()
        m = re.match(r"type\s*,\s*bind\(C\)\s*::\s*(\w+)", s, flags=re.I)
        if m:
            name = m.group(1)
            out[name] = []
            continue
        if name and re.match(r"end\s+type", s, flags=re.I):
            name = None
            continue
        if name:
            m = re.match(r"(integer\(c_int32_t\)|real\(c_double\)|real\(c_float\)|type\(c_ptr\))\s*::\s*(.*)", s, flags=re.I)
            assert m, s
            kind = {"integer(c_int32_t)": "int32_t", "real(c_double)": "double", "real(c_float)": "float",
                    "type(c_ptr)": "ptr"}[m.group(1).lower()]
            for item in m.group(2).split(","):
                out[name].append((item.strip(), kind))
    return out


def test_bind_c_types_match_header_field_for_field():
    hs, fs = _header_structs(), _shim_types()
    assert set(fs) <= set(hs) and {"evp_b200_dims", "evp_b200_params", "evp_b200_static_fields", "evp_b200_inputs",
                                    "evp_b200_state", "evp_b200_outputs"} <= set(fs)
    for name, ffields in fs.items():
        hfields = hs[name]
        assert [(n.lower(), k) for n, k in ffields] == [(n.lower(), k) for n, k in hfields], \
            f"{name}: the shim's bind(C) type differs from the header struct"


def test_interfaces_name_exported_entry_points():
    from cice4_b200 import evp as E
    src = _strip_c_comments(open(HEADER).read())
    lines = _join_continuations(open(SHIM).read().splitlines())
    bound = []
    for ln in lines:
        m = re.search(r"function\s+(\w+)\s*\((.*?)\)\s*bind\(C,\s*name='(\w+)'\)", ln, flags=re.I)
        if m:
            bound.append((m.group(3), len([a for a in m.group(2).split(",") if a.strip()])))
    assert len(bound) >= 12
    for cname, nargs in bound:
        assert cname in E.EXPORTS, f"{cname} is not an entry point of libevp_b200"
        m = re.search(r"\b" + cname + r"\s*\((.*?)\)\s*;", src, flags=re.S)
        assert m, cname
        cargs = [a for a in m.group(1).split(",") if a.strip() and a.strip() != "void"]
        assert len(cargs) == nargs, f"{cname}: {nargs} arguments in the shim, {len(cargs)} in the header"
    # what the shim needs for restart / history with a resident state
    names = {c for c, _ in bound}
    assert {"evp_b200_download_state", "evp_b200_invalidate_device_state", "evp_b200_principal_stress_n",
            "evp_b200_prep", "evp_b200_run", "evp_b200_init", "evp_b200_finalize"} <= names


def test_module_keeps_the_reference_public_names():
    txt = open(SHIM).read().lower()
    assert re.search(r"^\s*module ice_dyn_evp\s*$", txt, flags=re.M)
    for sub in ("init_evp", "evp", "principal_stress", "set_evp_parameters"):
        assert re.search(r"subroutine\s+" + sub + r"\b", txt), sub
    for var in ("kdyn", "ndte", "evp_damping", "yield_curve", "dragio", "cosw", "sinw"):
        assert re.search(r"::.*\b" + var + r"\b", txt), var
    # principal_stress is computed by the library, not on the host
    body = txt[txt.index("subroutine principal_stress"):txt.index("end subroutine principal_stress")]
    assert "evp_b200_principal_stress_n" in body and "sqrt" not in body


def test_block_structure_is_balanced():
    lines = [ln.lower() for ln in _join_continuations(open(SHIM).read().splitlines()) if not ln.lstrip().startswith("#")]

    def count(pat):
        return sum(1 for ln in lines if re.match(pat, ln.strip()))

    assert count(r"subroutine\s+\w+") == count(r"end\s+subroutine")
    assert count(r"(integer\(c_int\)|type\(c_ptr\)|integer\(c_int32_t\))\s+function\s+\w+") == count(r"end\s+function")
    assert count(r"type\s*,\s*bind\(c\)") == count(r"end\s+type")
    assert count(r"interface\s*$") == count(r"end\s+interface")
    assert count(r"module\s+ice_dyn_evp") == 1 and count(r"end\s+module") == 1
    assert count(r"do\s+\w+\s*=") == count(r"end\s*do")
    assert count(r"select\s+case") == count(r"end\s+select")
    ifs = sum(1 for ln in lines if re.match(r"if\s*\(.*\)\s*then\s*$", ln.strip()))
    assert ifs == count(r"end\s*if")
