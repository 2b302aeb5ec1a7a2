"""Name resolution for free-form Fortran 90 module sources -- the part of a compiler's front end that a text-level
test can reproduce without a Fortran compiler (none exists in this image).

`unresolved_names(path, search_dirs)` reads one module file, collects for every scope (the module head and each
procedure after `contains`) the names that are declared there, the names imported by `use M, only: ...` and -- for a
whole-module `use M` -- everything module M exports (M is looked up as <dir>/M.F90 in `search_dirs`, recursively
through M's own `use` statements), and returns the identifiers of every statement that resolve to none of these, to a
Fortran keyword / intrinsic or to an ISO_C_BINDING entity.  With `implicit none` each of them is a compile error.

It is deliberately conservative: preprocessor branches are all taken (union), a module's exports include what it
uses itself (Fortran re-exports used entities unless told otherwise), keyword arguments (`name=` inside parentheses)
and derived-type components (`%name`) are not names of the scope.  Test infrastructure only.
"""
import os
import re

KEYWORDS = set("""
tanh sinh cosh inquire flush rewind backspace exist opened number file status form access recl position action
if then else elseif endif end do enddo while call return stop continue select case default where elsewhere forall
subroutine function module program contains use only implicit none save private public interface import type
integer real logical character complex double precision kind len parameter dimension allocatable target pointer
intent in out inout optional value external intrinsic result recursive pure elemental bind c name
allocate deallocate allocated associated nullify stat write read print open close format unit fmt iostat advance
exit cycle go to goto data common equivalence namelist sequence
true false and or not eq ne lt le gt ge eqv neqv
abs sqrt sin cos tan asin acos atan atan2 exp log log10 min max mod modulo sign int nint real dble float aint anint
floor ceiling merge trim adjustl adjustr len_trim index size shape lbound ubound sum product minval maxval minloc
maxloc any all count pack unpack reshape transpose matmul dot_product spread huge tiny epsilon present char ichar
achar iachar transfer null cmplx aimag conjg dim dprod selected_real_kind selected_int_kind bit_size btest iand ior
ieor ishft not cpu_time system_clock date_and_time random_number random_seed
""".split())

ISO_C = set("""c_int c_int8_t c_int16_t c_int32_t c_int64_t c_long c_long_long c_size_t c_float c_double c_char c_bool
c_ptr c_funptr c_null_ptr c_null_funptr c_null_char c_loc c_funloc c_associated c_f_pointer c_f_procpointer
c_sizeof iso_c_binding""".split())

IDENT = re.compile(r"[a-z_]\w*")


def _strip(line):
    """remove string literals and the trailing comment of one source line"""
    out, q = [], None
    for ch in line:
        if q:
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
            out.append(" @ ")   # placeholder: keeps a line that holds only a string literal (continuations)
            continue
        if ch == "!":
            break
        out.append(ch)
    return "".join(out)


def statements(path):
    """logical statements of a free-form source, lower case, continuations joined, ';' split, cpp lines dropped"""
    out, cur = [], ""
    for raw in open(path, errors="replace").read().splitlines():
        if raw.lstrip().startswith("#"):
            continue
        s = _strip(raw).rstrip()
        if not s.strip():
            continue
        t = s.strip()
        if t.startswith("&"):
            t = t[1:]
        if t.endswith("&"):
            cur += t[:-1] + " "
            continue
        cur += t
        for part in cur.split(";"):
            if part.strip():
                out.append(part.strip().lower())
        cur = ""
    return out


def _split_top(s):
    """split at commas that are not inside parentheses"""
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return [p.strip() for p in parts if p.strip()]


def _entities(decl_rhs):
    """names declared by the right-hand side of `... :: a, b(3) = 1, c => null()`"""
    names = []
    for item in _split_top(decl_rhs):
        m = IDENT.match(item)
        if m:
            names.append(m.group(0))
    return names


DECL = re.compile(r"^(integer|real|logical|character|complex|double\s+precision|type\s*\()")
PROC = re.compile(r"^(?:(?:recursive|pure|elemental)\s+)*(?:(?:integer|real|logical|character|type|double\s+precision)"
                  r"\s*(?:\([^)]*\)|\*\s*\d+)?\s+)?(subroutine|function)\s+(\w+)\s*(?:\(([^)]*)\))?")
END_PROC = re.compile(r"^end\s*(subroutine|function)\b")


class Scope:
    def __init__(self, name):
        self.name = name
        self.declared = set()
        self.uses_all = []      # modules used without an only-list
        self.imported = set()   # names from only-lists
        self.body = []          # statements whose identifiers must resolve
        self.default_private = False
        self.public = set()     # `public :: a, b` / `, public ::` entities
        self.private = set()    # `private :: a, b` / `, private ::` entities
        self.args = []          # dummy arguments (procedures)
        self.optional = set()   # dummies declared optional
        self.target = set()     # entities declared with the TARGET or POINTER attribute


def _parse_use(st, scope):
    m = re.match(r"^use\s*(?:,\s*intrinsic\s*)?(?:::)?\s*(\w+)\s*(?:,\s*only\s*:\s*(.*))?$", st)
    if not m:
        return False
    mod, only = m.group(1), m.group(2)
    if only is None:
        scope.uses_all.append(mod)
    else:
        for item in _split_top(only):
            scope.imported.add(item.split("=>")[0].strip())
            if "=>" in item:
                pass
    return True


def parse_module(path):
    """-> (module scope, [procedure scopes]); statements inside interface blocks and derived-type definitions only
    contribute the names they declare at the enclosing level (procedure names, type names)"""
    sts = statements(path)
    mod = Scope("<module>")
    procs, cur = [], mod
    depth_iface = depth_type = 0
    iface_proc = 0
    for st in sts:
        if re.match(r"^module\s+\w+$", st) or re.match(r"^end\s*module", st) or st == "contains":
            continue
        if depth_iface:
            if re.match(r"^end\s*interface", st):
                depth_iface -= 1
                continue
            m = PROC.match(st)
            if m and not iface_proc:
                cur.declared.add(m.group(2))
                iface_proc = 1
            elif END_PROC.match(st) or re.match(r"^end$", st):
                iface_proc = 0
            elif re.match(r"^module\s+procedure\s+(.*)", st):
                pass
            continue
        if depth_type:
            if re.match(r"^end\s*type", st):
                depth_type -= 1
            continue
        m = re.match(r"^interface\s*(\w*)", st)
        if m and not st.startswith("interface_"):
            depth_iface += 1
            iface_proc = 0
            if m.group(1):
                cur.declared.add(m.group(1))
            continue
        m = re.match(r"^type\s*((?:,\s*[\w()\s]+)*)\s*(?:::)?\s*(\w+)$", st)
        if m and not st.startswith("type("):
            cur.declared.add(m.group(2))
            if "public" in m.group(1):
                cur.public.add(m.group(2))
            if "private" in m.group(1):
                cur.private.add(m.group(2))
            depth_type += 1
            continue
        m = PROC.match(st)
        if m and cur is mod:
            mod.declared.add(m.group(2))
            cur = Scope(m.group(2))
            cur.declared.add(m.group(2))
            if m.group(3):
                cur.args = [a.strip() for a in m.group(3).split(",") if a.strip()]
                cur.declared.update(cur.args)
            mr = re.search(r"result\s*\(\s*(\w+)\s*\)", st)
            if mr:
                cur.declared.add(mr.group(1))
            procs.append(cur)
            continue
        if END_PROC.match(st):
            cur = mod
            continue
        if _parse_use(st, cur):
            continue
        if st == "private":
            cur.default_private = True
            continue
        if st.startswith("implicit") or st == "save" or st == "public":
            continue
        m = re.match(r"^(public|private|save|external|intrinsic)\s*(?:::)?\s*(.*)$", st)
        if m and not DECL.match(st):
            if m.group(1) == "public":
                cur.public.update(_entities(m.group(2)))
            elif m.group(1) == "private":
                cur.private.update(_entities(m.group(2)))
            elif m.group(1) == "external":
                cur.declared.update(_entities(m.group(2)))
            continue
        if DECL.match(st):
            if "::" in st:
                lhs, rhs = st.split("::", 1)
                cur.declared.update(_entities(rhs))
                if re.search(r",\s*public\b", lhs):
                    cur.public.update(_entities(rhs))
                if re.search(r",\s*private\b", lhs):
                    cur.private.update(_entities(rhs))
                if re.search(r",\s*optional\b", lhs):
                    cur.optional.update(_entities(rhs))
                if re.search(r",\s*(target|pointer)\b", lhs):
                    cur.target.update(_entities(rhs))
                cur.body.append(lhs + " " + " ".join(x for x in re.findall(r"\(([^()]*)\)", rhs)))  # kinds, bounds
            else:   # old style: integer (kind=int_kind) i, j
                m2 = re.match(r"^(?:integer|real|logical|character|complex|double\s+precision)\s*(\([^)]*\))?\s*(.*)$", st)
                cur.declared.update(_entities(m2.group(2)))
                cur.body.append(m2.group(1) or "")
            continue
        m = re.match(r"^parameter\s*\((.*)\)$", st)
        if m:
            cur.body.append(m.group(1))
            continue
        if re.match(r"^namelist\s*/", st) or re.match(r"^(data|common|equivalence|format)\b", st) or re.match(r"^\d+\s+format", st):
            continue
        cur.body.append(st)
    return mod, procs


def identifiers(st):
    """identifiers of one statement that must resolve in its scope"""
    st = re.sub(r"%\s*\w+", "", st)                                  # derived-type components
    st = re.sub(r"\b\d+\.?\d*(?:[ed][+-]?\d+)?_(\w+)", r" \1 ", st)    # 1.0_dbl_kind -> dbl_kind
    st = re.sub(r"\.\d+(?:[ed][+-]?\d+)?_(\w+)", r" \1 ", st)
    st = re.sub(r"\b\d+\.?\d*[ed][+-]?\d+", " ", st)                  # 1.0e-11
    st = re.sub(r"\b\d+\b", " ", st)
    st = re.sub(r"\.(true|false|and|or|not|eq|ne|lt|le|gt|ge|eqv|neqv)\.", " ", st)
    out, depth = [], 0
    pos = 0
    for m in re.finditer(r"[()]|[a-z_]\w*", st):
        tok = m.group(0)
        if tok == "(":
            depth += 1
        elif tok == ")":
            depth -= 1
        else:
            rest = st[m.end():].lstrip()
            if depth > 0 and rest.startswith("=") and not rest.startswith("=="):
                continue                                              # keyword argument
            out.append(tok)
    return out


_EXPORT_CACHE = {}


def module_exports(name, search_dirs, _seen=None):
    """names a whole-module `use name` brings in (None: module file not found -> cannot be checked)"""
    key = (name, tuple(search_dirs))
    if key in _EXPORT_CACHE:
        return _EXPORT_CACHE[key]
    _seen = _seen or set()
    if name in _seen:
        return set()
    _seen.add(name)
    if name == "iso_c_binding":
        return set(ISO_C)
    path = None
    for d in search_dirs:
        for ext in (".F90", ".f90", ".F"):
            p = os.path.join(d, name + ext)
            if os.path.exists(p):
                path = p
                break
        if path:
            break
    if path is None:
        # case-insensitive file systems aside, module files of the reference carry the module's name
        _EXPORT_CACHE[key] = None
        return None
    mod, procs = parse_module(path)
    if mod.default_private:     # a bare `private`: only what is named public leaves the module
        names = set(mod.public)
    else:                       # default public: own entities and everything it uses itself, minus the private ones
        names = set(mod.declared) | set(mod.imported)
        for m in mod.uses_all:
            sub = module_exports(m, search_dirs, _seen)
            if sub:
                names |= sub
        names -= mod.private
    _EXPORT_CACHE[key] = names
    return names


def unresolved_names(path, search_dirs):
    """-> {scope name: sorted unresolved identifiers}, [modules that could not be found]"""
    mod, procs = parse_module(path)
    missing_modules = []

    def visible(scope):
        names = set(scope.declared) | set(scope.imported)
        for m in scope.uses_all:
            ex = module_exports(m, search_dirs)
            if ex is None:
                missing_modules.append(m)
            else:
                names |= ex
        return names

    base = visible(mod) | KEYWORDS | ISO_C
    bad = {}
    for scope in [mod] + procs:
        vis = base | (visible(scope) if scope is not mod else set())
        unres = set()
        for st in scope.body:
            for tok in identifiers(st):
                if tok not in vis:
                    unres.add(tok)
        if unres:
            bad[scope.name] = sorted(unres)
    return bad, sorted(set(missing_modules))


def calls(scope):
    """[(procedure name, number of actual arguments)] of the `call` statements and function references `name(...)`
    of a scope's statements (function references: only names given in `only_names`, see call_sites)"""
    out = []
    for st in scope.body:
        for m in re.finditer(r"\bcall\s+(\w+)\s*(\()?", st):
            name = m.group(1)
            if not m.group(2):
                out.append((name, 0))
                continue
            depth, k = 1, m.end()
            while k < len(st) and depth:
                depth += st[k] == "("
                depth -= st[k] == ")"
                k += 1
            out.append((name, len(_split_top(st[m.end():k - 1]))))
    return out


def function_refs(scope, names):
    """[(name, number of actual arguments)] for references `name(...)` to the functions in `names`"""
    out = []
    for st in scope.body:
        for m in re.finditer(r"\b(\w+)\s*\(", st):
            if m.group(1) not in names or st[:m.start()].rstrip().endswith("call"):
                continue
            depth, k = 1, m.end()
            while k < len(st) and depth:
                depth += st[k] == "("
                depth -= st[k] == ")"
                k += 1
            out.append((m.group(1), len(_split_top(st[m.end():k - 1]))))
    return out


def find_module_file(name, search_dirs):
    for d in search_dirs:
        for ext in (".F90", ".f90", ".F"):
            p = os.path.join(d, name + ext)
            if os.path.exists(p):
                return p
    return None


def procedure_signatures(path):
    """{procedure name: (number of dummies, number of optional dummies)} of a module's own procedures"""
    mod, procs = parse_module(path)
    return {p.name: (len(p.args), len(p.optional & set(p.args))) for p in procs}
