#!/usr/bin/env python
"""Syntax-level checks of the Fortran drop-in that need no Fortran compiler (the image has none).

1. Every `bind(C, name='...')` interface of fortran/m_gritas_b200.F90 matches its prototype in
   include/gritas_b200.h argument for argument: same count, same order, same C type class
   (handle by value / handle by reference, int by value, double by value, int32 array, double
   array, scalar double by reference, opaque byte buffer).
2. The module dependency graph of the drop-in has no cycle: m_gritas_grids (with
   fortran/m_gritas_grids_b200.inc substituted for its four routines) -> m_gritas_b200 ->
   { iso_c_binding, m_odsmeta }.  When the reference tree is present its m_gritas_grids.f is read
   for the modules it already uses.
3. Every procedure the include file calls is public in the shim, with the same number of arguments.
4. The replacement routines keep the reference's argument lists (checked against the reference
   source when /root/reference is present).
5. Every procedure of the shim has balanced if / do / select blocks and parentheses, and every identifier its
   executable statements and declarations use is a dummy argument, a local or module variable, a `use ... only` name, a
   procedure of the module or of its bind(C) interfaces, an iso_c_binding name or an intrinsic (what `implicit none`
   would reject).

Exit code 0 = all checks pass; prints one line per failure otherwise.
"""
from __future__ import annotations

import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "fortran", "m_gritas_b200.F90")
INC = os.path.join(ROOT, "fortran", "m_gritas_grids_b200.inc")
HEADER = os.path.join(ROOT, "include", "gritas_b200.h")
REF_GRIDS = "/root/reference/src/Components/gritas/m_gritas_grids.f"


def strip_fortran(text):
    """drop comments, join continuation lines, lower-case"""
    out, cur = [], ""
    for raw in text.splitlines():
        line = raw.split("!")[0].rstrip()
        if not line.strip():
            continue
        s = line.strip()
        if s.startswith("&"):
            s = s[1:].lstrip()
        if cur:
            cur += " " + s
        else:
            cur = s
        if cur.endswith("&"):
            cur = cur[:-1].rstrip()
            continue
        out.append(cur.lower())
        cur = ""
    if cur:
        out.append(cur.lower())
    return out


def strip_fixed_form(text):
    """fixed-form reference source: comment lines start with c/C/*/! in column 1, continuation = non-blank column 6"""
    out = []
    for raw in text.splitlines():
        if not raw.strip() or raw[0] in "cC*!":
            continue
        line = raw.split("!")[0].rstrip()
        if not line.strip():
            continue
        if len(line) > 5 and line[5] not in " 0" and line[:5].strip() == "":
            out[-1] += " " + line[6:].strip().lstrip("&").strip()
        else:
            s = line.strip()
            if s.startswith("&") and out:
                out[-1] += " " + s[1:].strip()
            else:
                out.append(s)
    return [l.rstrip("&").strip().lower() for l in out]


def split_args(s):
    """top-level commas only: achan(:,ichan_version) is one argument"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return [a for a in out if a]


# ---------------------------------------------------------------- C header
def c_prototypes():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int|void|uint64_t|const char \*|void \*)\s*(gritas_b200_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        name, args = m.group(2), " ".join(m.group(3).split())
        kinds = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                base = re.sub(r"\b[a-z_0-9]+$", "", a).strip() if not a.endswith("*") else a   # drop the parameter name
                base = base.replace("const ", "").strip()
                kinds.append(classify_c(base))
        protos[name] = kinds
    return protos


def classify_c(t):
    t = " ".join(t.split())
    return {
        "gritas_b200_handle": "handle", "gritas_b200_handle *": "handle*", "gritas_b200_xcov_handle": "handle",
        "gritas_b200_xcov_handle *": "handle*", "int": "int", "double": "double",
        "size_t": "size_t", "int32_t *": "int*", "int *": "int*", "double *": "double*", "void *": "bytes",
        "gritas_b200_ods *": "struct*", "double **": "double**",
    }.get(t, "?" + t)


# ---------------------------------------------------------------- Fortran interfaces
def fortran_interfaces(lines):
    """name -> list of kinds in dummy-argument order, for every bind(C) interface body"""
    res = {}
    i = 0
    while i < len(lines):
        m = re.match(r"(?:integer\(c_int\)\s+function|subroutine)\s+(\w+)\s*\((.*?)\)\s*bind\s*\(\s*c\s*,\s*name\s*=\s*'(\w+)'\s*\)", lines[i])
        if not m:
            i += 1
            continue
        dummies, cname = split_args(m.group(2)), m.group(3)
        decl = {}
        i += 1
        while i < len(lines) and not re.match(r"end\s+(function|subroutine)", lines[i]):
            d = re.match(r"(type\(c_ptr\)|integer\(c_int\)|real\(c_double\)|character\(kind=c_char\))\s*(?:,\s*([^:]*?))?\s*::\s*(.*)", lines[i])
            if d:
                typ, attrs, names = d.group(1), (d.group(2) or ""), d.group(3)
                for nm in re.findall(r"(\w+)\s*(\(\*\))?", names):
                    if nm[0]:
                        decl[nm[0]] = (typ, attrs, bool(nm[1]))
            i += 1
        kinds = []
        for a in dummies:
            if a not in decl:
                kinds.append("undeclared:" + a)
                continue
            typ, attrs, arr = decl[a]
            val = "value" in attrs
            if typ == "type(c_ptr)":
                kinds.append("handle" if val else "handle*")
            elif typ == "integer(c_int)":
                kinds.append("int" if val else "int*")          # scalar by reference == pointer to int
            elif typ == "real(c_double)":
                kinds.append("double" if val else "double*")
            else:
                kinds.append("bytes")
        res[cname] = kinds
    return res


def used_modules(lines):
    mods = set()
    for l in lines:
        m = re.match(r"use\s*(?:,\s*intrinsic\s*::)?\s*(\w+)", l)
        if m:
            mods.add(m.group(1))
    return mods


def public_procs(lines):
    pub = set()
    for l in lines:
        m = re.match(r"public\s*::\s*(.*)", l)
        if m:
            pub.update(split_args(m.group(1)))
    return pub


def subroutine_args(lines):
    res = {}
    for l in lines:
        m = re.match(r"subroutine\s+(\w+)\s*\((.*)\)\s*$", l)
        if m:
            res[m.group(1)] = split_args(m.group(2))
    return res


def calls(lines):
    res = []
    for l in lines:
        m = re.match(r"call\s+(\w+)\s*\((.*)\)\s*$", l)
        if m:
            res.append((m.group(1), split_args(m.group(2))))
    return res


def find_cycle(graph):
    seen, stack = {}, []

    def dfs(u):
        seen[u] = 1
        stack.append(u)
        for v in graph.get(u, ()):
            if seen.get(v) == 1:
                return stack[stack.index(v):] + [v]
            if v not in seen:
                c = dfs(v)
                if c:
                    return c
        stack.pop()
        seen[u] = 2
        return None
    for n in list(graph):
        if n not in seen:
            c = dfs(n)
            if c:
                return c
    return None


# ---------------------------------------------------------------------------------------------- check 5
KEYWORDS = set("""if then else elseif endif end do enddo call return subroutine function module contains use only implicit none
    integer real logical character type intent in out inout value optional save parameter allocatable dimension pointer target
    interface import public private bind c name result select case default exit cycle while where elsewhere allocate
    deallocate stat kind len intrinsic program stop print write read format continue go to goto associated""".split())
INTRINSIC_NAMES = set("""abs min max mod size present allocated trim len_trim sqrt exp log nint int dble real merge any all count sum
    product maxval minval reshape shape lbound ubound transfer huge tiny epsilon sign floor ceiling adjustl adjustr index iand ior
    ishft ieor not btest""".split())
ISO_C = set("""c_ptr c_int c_double c_float c_char c_size_t c_int32_t c_int64_t c_null_ptr c_null_char c_loc c_associated c_f_pointer
    c_funptr c_bool c_long c_short c_signed_char c_sizeof""".split())
DECL_START = re.compile(r"^(integer|real|logical|character|type\s*\(|double\s+precision|complex)")


def _names_declared(stmt):
    """names an entity declaration introduces"""
    rest = stmt.split("::", 1)[1] if "::" in stmt else re.sub(r"^(integer|real|logical|double\s+precision)\s*(\([^)]*\))?", "", stmt)
    out, depth, cur = [], 0, ""
    for ch in rest + ",":
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            m = re.match(r"\s*([a-z_][a-z0-9_]*)", cur)
            if m:
                out.append(m.group(1))
            cur = ""
        else:
            cur += ch
    return out


def _identifiers(stmt):
    """identifiers an executable statement refers to (no string contents, no component names, no keyword-argument names)"""
    stmt = re.sub(r"'[^']*'|\"[^\"]*\"", " ", stmt)
    stmt = re.sub(r"%\s*[a-z_][a-z0-9_]*", " ", stmt)                 # a%b: b is a component
    stmt = re.sub(r"\.(true|false|and|or|not|eq|ne|lt|le|gt|ge|eqv|neqv)\.", " ", stmt)
    stmt = re.sub(r"\b\d+(\.\d*)?([de][+-]?\d+)?(_[a-z_0-9]+)?\b", " ", stmt)      # numbers, with kind suffix
    stmt = re.sub(r"(?<=[(,])\s*[a-z_][a-z0-9_]*\s*=(?!=)", " ", stmt)              # keyword arguments: f(name=...)
    stmt = re.sub(r"\bcall\s+[a-z_][a-z0-9_]*", " ", stmt)             # an external subroutine needs no declaration
    return set(re.findall(r"[a-z_][a-z0-9_]*", stmt))


def lint_units(lines, extra_known=()):
    """block balance of every program unit and an `implicit none` lint of its executable statements"""
    errors = []
    module_names, procs = set(extra_known), set()
    # pass 1: module-level names and all procedure names (module procedures, interface bodies)
    depth_iface, in_unit = 0, 0
    for st in lines:
        if re.match(r"^interface\b", st):
            depth_iface += 1
        elif re.match(r"^end\s*interface", st):
            depth_iface -= 1
        m = re.match(r"^(?:[a-z_()0-9 ]*?\b)?(subroutine|function)\s+([a-z_][a-z0-9_]*)", st)
        if m and not st.startswith("end"):
            procs.add(m.group(2))
            if not depth_iface:
                in_unit += 1
            continue
        if re.match(r"^end\s*(subroutine|function)", st) and not depth_iface:
            in_unit -= 1
            continue
        if not in_unit and not depth_iface and DECL_START.match(st):
            module_names.update(_names_declared(st))
        m = re.match(r"^use\b.*\bonly\s*:\s*(.*)$", st)
        if m and not in_unit:
            module_names.update(x.split("=>")[0].strip() for x in m.group(1).split(","))
    # pass 2: every module procedure
    i, n = 0, len(lines)
    depth_iface = 0
    while i < n:
        st = lines[i]
        if re.match(r"^interface\b", st):
            depth_iface += 1
        elif re.match(r"^end\s*interface", st):
            depth_iface -= 1
        m = re.match(r"^(?:[a-z_()0-9 ]*?\b)?(subroutine|function)\s+([a-z_][a-z0-9_]*)\s*(?:\(([^)]*)\))?", st)
        if not m or st.startswith("end") or depth_iface:
            i += 1
            continue
        name = m.group(2)
        local = {a.strip() for a in (m.group(3) or "").split(",") if a.strip()} | {name}
        mres = re.search(r"result\s*\(\s*([a-z_][a-z0-9_]*)\s*\)", st)
        if mres:
            local.add(mres.group(1))
        stack = []
        j = i + 1
        inner_iface = 0
        while j < n and not re.match(r"^end\s*(subroutine|function)", lines[j]):
            s2 = lines[j]
            if re.match(r"^interface\b", s2):
                inner_iface += 1
            elif re.match(r"^end\s*interface", s2):
                inner_iface -= 1
            elif inner_iface:
                pass
            elif re.match(r"^use\b", s2):
                mo = re.match(r"^use\b.*\bonly\s*:\s*(.*)$", s2)
                if mo:
                    local.update(x.split("=>")[0].strip() for x in mo.group(1).split(","))
            elif DECL_START.match(s2) and "::" in s2 or re.match(r"^(integer|real|logical)\s+[a-z_]", s2):
                local.update(_names_declared(s2))
                # initialisers and bounds may refer to other names: check them too
                unknown = _identifiers(s2.split("::", 1)[1] if "::" in s2 else "") - local - module_names - procs - KEYWORDS - INTRINSIC_NAMES - ISO_C
                for u in sorted(unknown):
                    errors.append(f"{name}: `{u}` in a declaration is not declared ({s2[:70]})")
            elif s2.startswith(("implicit", "import")):
                pass
            else:
                # block structure
                if re.match(r"^(?:[a-z_0-9]+\s*:\s*)?if\s*\(.*\)\s*then$", s2):
                    stack.append("if")
                elif re.match(r"^(?:[a-z_0-9]+\s*:\s*)?do\b", s2):
                    stack.append("do")
                elif re.match(r"^select\s+case", s2):
                    stack.append("select")
                elif re.match(r"^end\s*if\b", s2):
                    if not stack or stack.pop() != "if":
                        errors.append(f"{name}: unbalanced `end if`")
                elif re.match(r"^end\s*do\b", s2):
                    if not stack or stack.pop() != "do":
                        errors.append(f"{name}: unbalanced `end do`")
                elif re.match(r"^end\s*select\b", s2):
                    if not stack or stack.pop() != "select":
                        errors.append(f"{name}: unbalanced `end select`")
                if s2.count("(") != s2.count(")"):
                    errors.append(f"{name}: unbalanced parentheses ({s2[:70]})")
                unknown = _identifiers(s2) - local - module_names - procs - KEYWORDS - INTRINSIC_NAMES - ISO_C
                for u in sorted(unknown):
                    errors.append(f"{name}: `{u}` is not declared ({s2[:70]})")
            j += 1
        if stack:
            errors.append(f"{name}: unterminated {stack}")
        if j >= n:
            errors.append(f"{name}: no end subroutine / end function")
        i = j + 1
    return errors


def run():
    errors = []
    shim = strip_fortran(open(SHIM).read())
    inc = strip_fortran(open(INC).read())
    protos = c_prototypes()
    ifaces = fortran_interfaces(shim)
    if len(ifaces) < 15:
        errors.append(f"only {len(ifaces)} bind(C) interfaces parsed from the shim")
    for name, kinds in ifaces.items():
        if name not in protos:
            errors.append(f"{name}: bound in the shim, not declared in include/gritas_b200.h")
            continue
        want = protos[name]
        if len(want) != len(kinds):
            errors.append(f"{name}: {len(kinds)} arguments in the shim, {len(want)} in the header")
            continue
        for pos, (a, b) in enumerate(zip(kinds, want)):
            ok = a == b or (a == "handle" and b == "bytes") or (a == "bytes" and b == "bytes")
            if not ok:
                errors.append(f"{name}: argument {pos + 1} is {a} in the shim, {b} in the header")
    # module graph
    graph = {"m_gritas_b200": used_modules(shim)}
    grids_uses = used_modules(inc)
    if os.path.exists(REF_GRIDS):
        grids_uses |= used_modules(strip_fixed_form(open(REF_GRIDS, errors="replace").read()))
    graph["m_gritas_grids"] = grids_uses
    if "m_gritas_grids" in graph["m_gritas_b200"]:
        errors.append("m_gritas_b200 uses m_gritas_grids (and m_gritas_grids uses m_gritas_b200): circular")
    if "m_gritas_b200" not in graph["m_gritas_grids"]:
        errors.append("the include file does not use m_gritas_b200")
    cyc = find_cycle(graph)
    if cyc:
        errors.append("module dependency cycle: " + " -> ".join(cyc))
    allowed = {"iso_c_binding", "m_odsmeta"}
    extra = graph["m_gritas_b200"] - allowed
    if extra:
        errors.append(f"the shim uses modules beyond iso_c_binding / m_odsmeta: {sorted(extra)}")
    # calls of the include file resolve to public shim procedures of the same arity
    pub = public_procs(shim)
    sargs = subroutine_args(shim)
    for name, args in calls(inc):
        if not name.startswith("b200_"):
            continue
        if name not in pub:
            errors.append(f"{name}: called by the include file, not public in the shim")
        elif name not in sargs:
            errors.append(f"{name}: public but not defined in the shim")
        elif len(sargs[name]) != len(args):
            errors.append(f"{name}: called with {len(args)} arguments, defined with {len(sargs[name])}")
    for name in pub:
        if name.startswith("b200_") and name not in sargs and name != "b200_peer_blob_bytes":
            errors.append(f"{name}: public but not defined")
    # the replacement routines keep the reference's dummy-argument lists
    if os.path.exists(REF_GRIDS):
        ref_args = subroutine_args(strip_fixed_form(open(REF_GRIDS, errors="replace").read()))
        for name, args in subroutine_args(inc).items():
            if name not in ref_args:
                errors.append(f"{name}: not a routine of the reference's m_gritas_grids.f")
            elif ref_args[name] != args:
                errors.append(f"{name}: dummy arguments differ from the reference: {args} vs {ref_args[name]}")
    # 5. block structure and an `implicit none` lint of every procedure of the shim
    errors += lint_units(shim)
    # ... and of the reference-side bodies, which see the module variables of m_gritas_grids (m_gritas_grids.f:21-72) and
    # the public procedures of the shim
    grids_names = {"im", "jm", "km", "nlevs", "nchan", "achan", "ichan_version", "plevs", "amiss", "myname"}
    errors += lint_units(inc, set(pub) | grids_names)
    return errors, ifaces, graph


if __name__ == "__main__":
    errs, ifaces, graph = run()
    print(f"{len(ifaces)} bind(C) interfaces checked against include/gritas_b200.h")
    print("module graph:", {k: sorted(v) for k, v in graph.items()})
    for e in errs:
        print("FAIL:", e)
    sys.exit(1 if errs else 0)
