#!/usr/bin/env python
"""Mechanical translation of a small Fortran subset to C -- TEST INFRASTRUCTURE, not product code.

Why: the reference (GEOS-ESM/GrITAS, Fortran) cannot be compiled in this image (no Fortran compiler),
so `oracle/gritas_oracle.c` -- a restatement written by hand -- had nothing of the reference's own to be checked against.
This script reads the reference's source files WHERE THEY LIE under /root/reference, translates the routines of the
hot path statement by statement by syntax rules alone (no knowledge of what they compute), and writes C into
oracle/_ref/ (git-ignored; no reference source is copied into the repository).  gcc compiles that into
oracle/_ref/libgritas_ref.so, `tests/test_oracle_vs_ref.py` demands that the hand-written oracle and this translation
agree bit for bit, and `tests/golden/make_ref_golden.py` freezes its outputs as golden vectors for the GPU tests.

Subset (what m_gritas_grids.f:217-296, 307-478, 490-656, 1019-1167, 1178-1341 use): fixed-form source with `&`
continuation lines and `!` comments; integer / real / logical declarations with explicit-shape or automatic arrays,
intent, optional; assignments to scalars, array elements and array sections (`a(i,j,1:nr)`, `a(:,:,:,1)`: a loop
nest over the section); do / labelled do / end do / continue, cycle, exit, go to; block and one-line if, else, else if;
call gritas_perror; return; print / format / allocatable declarations are dropped.  Expressions keep their tokens and
their order: `.and.` -> `&&`, `.le.` -> `<=`, `/=` -> `!=`, intrinsics nint / abs / min / max / sqrt / exp / log /
real / present map to C99 (`lround` is Fortran's NINT: half away from zero).  `real` is real*8 (the reference is
built with FREAL8, src/Components/gritas/CMakeLists.txt:3-4): literals and variables become C `double`, which also
gives the default-real literals their promoted value.  Integer / real mixing and left-to-right evaluation are the
same in C and Fortran, so nothing is re-associated.  Arrays keep Fortran's column-major, 1-based layout through
accessor macros built from the declared bounds.
"""
from __future__ import annotations

import re
import sys

INTRINSICS = {"nint", "abs", "min", "max", "sqrt", "exp", "log", "real", "present", "dble", "int", "mod", "any", "sum", "size",
              "f_streq"}
FUNCTIONS = {}      # translated functions: lower-case name -> (C name, [number of assumed-shape dimensions or 0 per argument])
RELOPS = {".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">=", ".eq.": "==", ".ne.": "!=", ".and.": "&&", ".or.": "||",
          ".not.": "!", ".true.": "1", ".false.": "0", "/=": "!=", "==": "==", "<=": "<=", ">=": ">=", "<": "<", ">": ">"}


# ---------------------------------------------------------------------------------------------- source lines
def strip_comment(line: str) -> str:
    out, q = [], None
    for ch in line:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        elif ch == "!":
            break
        else:
            out.append(ch)
    return "".join(out).rstrip()


def lower_outside_strings(s: str) -> str:
    out, q = [], None
    for ch in s:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        else:
            out.append(ch.lower())
    return "".join(out)


def logical_lines(text: str, first: int, last: int, free: bool = False):
    """(label or None, statement, first physical line number) of the source lines first..last (1-based, inclusive);
    fixed form (column-6 continuation) unless `free` (trailing `&`)."""
    res = []
    pending = False
    for no, raw in enumerate(text.split("\n"), 1):
        if no < first or no > last:
            continue
        if free:
            line = strip_comment(raw).strip()
            if not line:
                continue
            more = line.endswith("&")
            line = lower_outside_strings(line.rstrip("&").strip().lstrip("&").strip())
            if pending:
                lab, st, n0 = res[-1]
                res[-1] = (lab, st + " " + line, n0)
            else:
                res.append((None, line, no))
            pending = more
            continue
        if not raw.strip() or raw[0] in "cC*!":
            continue
        line = strip_comment(raw.replace("\t", "      "))
        if not line.strip():
            continue
        head, cont = line[:5], line[5:6]
        if cont not in ("", " ", "0") and not head.strip():       # continuation line
            if not res:
                raise ValueError(f"line {no}: continuation without a statement")
            lab, st, n0 = res[-1]
            res[-1] = (lab, st + " " + lower_outside_strings(line[6:].strip()), n0)
            continue
        label = head.strip() if head.strip().isdigit() else None
        body = line[6:] if label or not head.strip() else line    # free-ish lines that start before column 7
        res.append((label, lower_outside_strings(body.strip()), no))
    out = []
    for lab, st, no in res:
        parts = split_semicolons(normalise(st))
        for k, p_ in enumerate(parts):
            out.append((lab if k == 0 else None, p_, no))
    return out


def split_semicolons(st: str):
    parts, cur, q = [], [], None
    for ch in st:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur.append(ch)
        elif ch == ";":
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur).strip())
    return [p_ for p_ in parts if p_]


DOTOPS = "lt|le|gt|ge|eq|ne|and|or|not|true|false"


def normalise(st: str) -> str:
    """blanks are insignificant in fixed form: `.eq .31` is `.eq.` `31`; `a%b%c` names one array: a_b_c"""
    st = re.sub(rf"\.\s*({DOTOPS})\s*\.", r".\1.", st)
    st = re.sub(r"\s*%\s*", "_", st) if "'" not in st and '"' not in st else st
    st = re.sub(r"^elseif\b", "else if", st)
    st = re.sub(r"trim\s*\(\s*(\w+\s*\(\s*\w+\s*\))\s*\)\s*==\s*('[^']*')", r"f_streq(\1, \2)", st)
    return st


# ---------------------------------------------------------------------------------------------- tokens / expressions
TOKEN = re.compile(r"""\s*(?:
    (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*") |
    (?P<dotop>\.(?:lt|le|gt|ge|eq|ne|and|or|not|true|false)\.) |
    (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[ed][+-]?\d+)?) |
    (?P<id>[a-z_][a-z0-9_]*) |
    (?P<op>\*\*|//|==|/=|<=|>=|[-+*/(),:<>=])
)""", re.X)


def tokenize(s: str):
    pos, out = 0, []
    s = s.strip()
    while pos < len(s):
        m = TOKEN.match(s, pos)
        if not m or m.end() == pos:
            raise ValueError(f"cannot tokenize {s[pos:]!r} in {s!r}")
        kind = m.lastgroup
        out.append((kind, m.group(kind)))
        pos = m.end()
    return out


def split_top(tokens, sep=","):
    """split a token list at `sep` outside parentheses"""
    parts, cur, depth = [], [], 0
    for t in tokens:
        if t == ("op", "("):
            depth += 1
        elif t == ("op", ")"):
            depth -= 1
        if depth == 0 and t == ("op", sep):
            parts.append(cur)
            cur = []
        else:
            cur.append(t)
    parts.append(cur)
    return parts


def matching_paren(tokens, i):
    depth = 0
    for j in range(i, len(tokens)):
        if tokens[j] == ("op", "("):
            depth += 1
        elif tokens[j] == ("op", ")"):
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced parentheses")


class Scope:
    def __init__(self, globals_):
        self.vars = dict(globals_)      # name -> dict(type, dims, optional, arg)
        self.section_vars = None        # loop variables of the section assignment being translated
        self.where = None               # (mask tokens, negate) while inside a where / elsewhere block

    def is_array(self, name):
        v = self.vars.get(name)
        return bool(v and v["dims"])

    def extent(self, name, q):
        """C expression of the extent of dimension q (0-based) of array `name`"""
        return expr_c(self.vars[name]["dims"][q], self)

    def total(self, name):
        return " * ".join(f"(size_t)({self.extent(name, q)})" for q in range(len(self.vars[name]["dims"])))


def number_c(tok: str) -> str:
    t = tok.replace("d", "e")
    return t


def expr_c(tokens, sc: Scope) -> str:
    """tokens -> C, token by token; `**` with a small integer exponent becomes a product, array sections take the
    loop variables of the enclosing section assignment"""
    out = []
    i = 0
    while i < len(tokens):
        kind, val = tokens[i]
        if kind == "num":
            out.append(number_c(val))
        elif kind == "dotop":
            out.append(" " + RELOPS[val] + " ")
        elif kind == "str":
            out.append('"' + val[1:-1].replace('"', '\\"') + '"')
        elif kind == "op":
            if val == "**":
                # base = the last complete operand in `out`; exponent = next token (small integer literal only)
                nk, nv = tokens[i + 1]
                if nk != "num" or not nv.isdigit() or int(nv) > 4:
                    raise ValueError("only x**<small integer> is supported")
                base = out.pop()
                out.append("(" + " * ".join([base] * int(nv)) + ")")
                i += 1
            elif val in ("==", "/=", "<=", ">=", "<", ">"):
                out.append(" " + RELOPS[val] + " ")
            elif val == "//":
                raise ValueError("string concatenation outside gritas_perror")
            else:
                out.append(val)
        elif kind == "id":
            if i + 1 < len(tokens) and tokens[i + 1] == ("op", "("):
                j = matching_paren(tokens, i + 1)
                args = split_top(tokens[i + 2:j])
                if sc.is_array(val):
                    idx = []
                    sec = iter(sc.section_vars or [])
                    for a in args:
                        if any(t == ("op", ":") for t in a) and not any(t == ("op", "(") for t in a):
                            lo = [t for t in a[:a.index(("op", ":"))]]
                            v = next(sec)                     # section: the loop variable runs over lo..hi
                            idx.append(f"({expr_c(lo, sc) if lo else '1'}) + {v}")
                        elif any(t == ("op", ":") for t in a):
                            # vector subscript `x(indx(:))`: this dimension is a section too; the subscript array runs
                            # over the same loop variable
                            v = next(sec)
                            keep = sc.section_vars
                            sc.section_vars = [v]
                            try:
                                idx.append(expr_c(a, sc))
                            finally:
                                sc.section_vars = keep
                        else:
                            idx.append(expr_c(a, sc))
                    out.append(f"{val}({', '.join(idx)})")
                elif val in ("any", "sum", "size"):
                    nm = args[0][0][1]
                    if len(args[0]) != 1 or not sc.is_array(nm):
                        raise ValueError(f"{val}() of something other than a whole array")
                    if val == "size":
                        out.append(f"((int)({sc.total(nm)}))" if len(args) == 1 else f"({sc.extent(nm, int(args[1][0][1]) - 1)})")
                    else:
                        t = sc.vars[nm]["type"]
                        op = "r_ = r_ || " if val == "any" else "r_ += "
                        out.append(f"({{ {'int' if val == 'any' else t} r_ = 0; for (size_t z_ = 0; z_ < {sc.total(nm)}; ++z_) {op}{nm}[z_]; r_; }})")
                elif val in INTRINSICS:
                    a = [expr_c(x, sc) for x in args]
                    if val == "nint":
                        out.append(f"((int)lround({a[0]}))")
                    elif val == "abs":
                        out.append(f"F_ABS({a[0]})")
                    elif val in ("min", "max"):
                        e = a[0]
                        for x in a[1:]:
                            e = f"F_{val.upper()}({e}, {x})"
                        out.append(e)
                    elif val in ("sqrt", "exp", "log"):
                        out.append(f"{val}({a[0]})")
                    elif val in ("real", "dble"):
                        out.append(f"((double)({a[0]}))")
                    elif val == "int":
                        out.append(f"((int)({a[0]}))")
                    elif val == "mod":
                        out.append(f"(({a[0]}) % ({a[1]}))")
                    elif val == "present":
                        out.append(f"({a[0].strip('(*)')} != 0)")
                    elif val == "f_streq":
                        out.append(f"(strcmp({a[0]}, {a[1].strip()}) == 0)")
                    else:
                        raise ValueError(val)
                elif val in FUNCTIONS:
                    cname, kinds = FUNCTIONS[val]
                    cargs = []
                    for a, nd in zip(args, kinds):
                        nm = a[0][1]
                        if a[0][0] == "id" and sc.is_array(nm) and len(a) > 1:
                            sub = split_top(a[2:matching_paren(a, 1)])
                            secs = [q for q, x in enumerate(sub) if x == [("op", ":")]]
                            if secs != list(range(len(secs))):
                                raise ValueError("only leading whole-dimension sections can be passed on")
                            first = ", ".join("1" if x == [("op", ":")] else expr_c(x, sc) for x in sub)
                            cargs.append(f"&{nm}({first})")              # contiguous: the address of its first element ...
                            if nd:
                                cargs += [f"{sc.extent(nm, q)}" for q in secs]      # ... and, for an assumed-shape dummy, the extents
                        else:
                            cargs.append(expr_c(a, sc))
                    out.append(f"{cname}({', '.join(cargs)})")
                elif val in sc.vars and not sc.vars[val]["dims"]:
                    # declared like a scalar, used with arguments: an external function (the harness supplies it)
                    out.append(f"{val}_({', '.join(expr_c(x, sc) for x in args)})")
                else:
                    raise ValueError(f"call of unknown function or undeclared array {val!r}")
                i = j
            else:
                v = sc.vars.get(val)
                if v is None:
                    raise ValueError(f"undeclared name {val!r}")
                if v["dims"] and sc.section_vars is not None:          # whole array inside a section assignment: not used
                    raise ValueError(f"whole-array reference {val!r} is not supported")
                out.append(f"(*{val})" if (v.get("optional") or v.get("byref")) else val)
        i += 1
    return "".join(out)


# ---------------------------------------------------------------------------------------------- declarations
DECL = re.compile(r"^(integer|real|logical|character|double\s+precision)\b(.*)$")


def parse_decl(stmt):
    """-> (type, attrs, [(name, dims tokens list or None, init tokens or None)]) or None"""
    m = DECL.match(stmt)
    if not m:
        return None
    typ = {"integer": "int", "real": "double", "logical": "int", "character": "char", "double precision": "double"}[
        re.sub(r"\s+", " ", m.group(1))]
    rest = m.group(2).strip()
    if rest.startswith("(") and typ == "char":                           # character(len=*)
        rest = rest[rest.index(")") + 1:].strip()
    if typ == "double" and re.match(r"^\(\s*4\s*\)", rest):                # real(4)
        typ = "float"
        rest = rest[rest.index(")") + 1:].strip()
    attrs = []
    dim_attr = None
    if "::" in rest:
        a, rest = rest.split("::", 1)
        attrs, depth, cur = [], 0, ""
        for ch in a.strip(" ,") + ",":
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            if ch == "," and depth == 0:
                if cur.strip():
                    attrs.append(cur.strip())
                cur = ""
            else:
                cur += ch
        for x in attrs:
            if x.startswith("dimension"):
                t = tokenize(x[len("dimension"):])
                dim_attr = split_top(t[1:matching_paren(t, 0)])
        attrs = [re.sub(r"\s+", "", x) for x in attrs]
    toks = tokenize(rest)
    ents = []
    for part in split_top(toks):
        if not part:
            continue
        name = part[0][1]
        dims, init = None, None
        k = 1
        if k < len(part) and part[k] == ("op", "("):
            j = matching_paren(part, k)
            dims = split_top(part[k + 1:j])
            k = j + 1
        if k < len(part) and part[k] == ("op", "="):
            init = part[k + 1:]
        ents.append((name, dims if dims is not None else dim_attr, init))
    return typ, attrs, ents


# ---------------------------------------------------------------------------------------------- routines
class Translator:
    def __init__(self, text: str, globals_: dict):
        self.text = text
        self.globals = globals_
        self.out = []

    def find_routine(self, name):
        lines = self.text.split("\n")
        start = end = None
        pat = re.compile(rf"^\s*(subroutine|real\s+function)\s+{re.escape(name)}\b", re.I)
        pend = re.compile(rf"^\s*end(\s+(subroutine|function)(\s+{re.escape(name)})?)?\s*(!.*)?$", re.I)
        for no, l in enumerate(lines, 1):
            if start is None and pat.match(l):
                start = no
            elif start is not None and pend.match(l):
                end = no
                break
        if start is None or end is None:
            raise ValueError(f"routine {name} not found")
        return start, end

    def routine(self, name, cname):
        first, last = self.find_routine(name)
        ll = logical_lines(self.text, first, last)
        head = ll[0][1]
        m = re.match(r"(?:subroutine|real\s+function)\s+\w+\s*\((.*)\)\s*$", head)
        args = [a.strip() for a in m.group(1).split(",")] if m and m.group(1).strip() else []
        is_function = head.startswith("real")
        sc = Scope(self.globals)
        if is_function:                                                 # the result variable carries the function's name
            sc.vars[name.lower()] = {"type": "double", "dims": None, "optional": False, "arg": False}
        body, decl_order = [], []
        for label, st, no in ll[1:-1]:
            d = parse_decl(st)
            if d and not label:
                typ, attrs, ents = d
                if "parameter" in attrs or typ == "char":
                    continue                                            # myname_ strings: only used in messages
                for nm, dims, init in ents:
                    deferred = dims is not None and all(d == [("op", ":")] for d in dims)
                    if "allocatable" in attrs and not any(re.search(rf"\b{nm}\b", st2) for _, st2, _ in ll[1:-1] if not parse_decl(st2)):
                        continue                                        # declared, never touched
                    v = {"type": typ, "dims": dims, "optional": "optional" in attrs, "arg": nm in args,
                         "alloc": "allocatable" in attrs, "assumed": deferred and nm in args,
                         "byref": nm in args and dims is None and any(a_ in ("intent(out)", "intent(inout)") for a_ in attrs)}
                    if deferred:                                        # extents live in variables nm_n1, nm_n2, ...
                        v["dims"] = [[("id", f"{nm}_n{q + 1}")] for q in range(len(dims))]
                        for q in range(len(dims)):
                            sc.vars[f"{nm}_n{q + 1}"] = {"type": "int", "dims": None, "optional": False, "arg": False}
                    sc.vars[nm] = v
                    decl_order.append(nm)
                continue
            if st.startswith(("use ", "implicit ")):
                continue
            body.append((label, st, no))
        for a in args:
            if a not in sc.vars:
                raise ValueError(f"{name}: argument {a} is not declared")
        # a scalar argument the routine assigns to goes by reference (Fortran passes everything by reference)
        for label, st, no in body:
            m2 = re.match(r"^(?:if\s*\(.*\)\s*)?([a-z_][a-z0-9_]*)\s*=[^=]", st)
            if m2 and m2.group(1) in args and not sc.vars[m2.group(1)]["dims"]:
                sc.vars[m2.group(1)]["byref"] = True
        # signature: arrays, optional and assigned scalars by pointer, other scalars by value
        params = []
        for a in args:
            v = sc.vars[a]
            params.append(f"{v['type']} *{a}" if (v["dims"] or v["optional"] or v.get("byref")) else f"{v['type']} {a}")
            if v.get("assumed"):                                        # assumed shape: the extents follow the pointer
                params += [f"int {a}_n{q + 1}" for q in range(len(v["dims"]))]
        w = self.out.append
        w(f"/* {name}: translated from lines {first}-{last} */")
        w(f"{'double' if is_function else 'void'} {cname}({', '.join(params) or 'void'}) {{")
        if is_function:
            w(f"  double {name.lower()} = 0;")
        macros = []
        for nm in decl_order:
            v = sc.vars[nm]
            if not v["arg"]:
                if v.get("alloc"):
                    w(f"  {v['type']} *{nm} = 0; int " + ", ".join(f"{nm}_n{q + 1} = 0" for q in range(len(v["dims"]))) + ";")
                elif v["dims"]:
                    ext = " * ".join(f"({expr_c(d, sc)})" for d in v["dims"])
                    w(f"  {v['type']} {nm}_store[{ext}]; {v['type']} *{nm} = {nm}_store;")
                    w(f"  for (size_t z_ = 0; z_ < (size_t)({ext}); ++z_) {nm}_store[z_] = 0;   /* undefined in Fortran: zero here */")
                else:
                    w(f"  {v['type']} {nm} = 0;")
        for nm in decl_order:
            v = sc.vars[nm]
            if v["dims"]:
                macros.append(nm)
                w(self.accessor(nm, v["dims"], sc))
        if is_function:
            FUNCTIONS[name.lower()] = (cname, [len(sc.vars[a]["dims"]) if sc.vars[a].get("assumed") else 0 for a in args])
        self.statements(body, sc, ret=(name.lower() if is_function else None))
        w(f"  return {name.lower()};" if is_function else "  return;")
        for nm in macros:
            if nm in self.globals and self.globals[nm].get("define"):
                w(f"#undef {nm}")
                w(self.globals[nm]["define"])                           # a dummy of the same name shadowed the module's array
            else:
                w(f"#undef {nm}")
        w("}")
        w("")

    @staticmethod
    def accessor(nm, dims, sc):
        """column-major, 1-based: ((i1-1) + d1*((i2-1) + d2*((i3-1) + ...)))"""
        n = len(dims)
        idx = [f"i{q + 1}" for q in range(n)]
        e = f"((size_t)({idx[n - 1]}) - 1)"
        for q in range(n - 2, -1, -1):
            e = f"((size_t)({idx[q]}) - 1 + (size_t)({expr_c(dims[q], sc)}) * {e})"
        return f"#define {nm}({', '.join(idx)}) {nm}[{e}]"

    # ------------------------------------------------------------------------------------------ statements
    def statements(self, body, sc, ret=None):
        self.ret = ret
        w = self.out.append
        do_stack = []                 # label (or None) of every open do loop
        ind = "  "
        for label, st, no in body:
            closes = 0
            if label:
                # a labelled statement that terminates labelled do loops closes them after itself
                closes = sum(1 for d in do_stack if d == label)
            if label:
                w(f"{ind}L{label}: ;")
            c = self.statement(st, sc, do_stack, no)
            for line in c:
                w(ind + line)
            for _ in range(closes):
                do_stack.remove(label)
                w(ind + "}")
        if do_stack:
            raise ValueError(f"unterminated do loops: {do_stack}")

    def statement(self, st, sc, do_stack, no):
        if st in ("continue", ""):
            return []
        if st.startswith(("print", "format", "write")) or re.match(r"^\d+\s+format", st):
            return [f"/* line {no}: output statement dropped */"]
        if st == "return":
            return [f"return {self.ret};" if getattr(self, "ret", None) else "return;"]
        if st == "cycle":
            return ["continue;"]
        if st == "exit":
            return ["break;"]
        m = re.match(r"^go\s*to\s+(\d+)$", st)
        if m:
            return [f"goto L{m.group(1)};"]
        m = re.match(r"^call\s+(gritas_perror|die)\b", st)
        if m:
            return [f'gritas_ref_perror({no});']
        m = re.match(r"^call\s+(indexset|indexsort)\s*\((.*)\)$", st)
        if m:
            # m_MergeSorts (GMAO_mpeu, not in the tree): the harness supplies a stable ascending / descending index sort; the
            # key sequence, the key arrays and the direction are what is translated.  Array sections starting at 1 pass the array
            args = []
            typ = ""
            for a in split_top(tokenize(m.group(2))):
                if len(a) > 2 and a[1] == ("op", "=") and a[0][0] == "id":
                    a = a[2:]                                          # keyword argument: descend=.false.
                if a[0][0] == "id" and sc.is_array(a[0][1]):
                    if m.group(1) == "indexsort" and len(args) == 2:
                        typ = "_i" if sc.vars[a[0][1]]["type"] == "int" else "_d"
                    args.append(a[0][1])
                else:
                    args.append(expr_c(a, sc))
            return [f"{m.group(1)}{typ}_({', '.join(args)});"]
        m = re.match(r"^where\s*\((.*)\)$", st)
        if m:                                                  # masked section assignments until endwhere
            sc.where = (tokenize(m.group(1)), False)
            return []
        if st == "elsewhere":
            sc.where = (sc.where[0], True)
            return []
        if st in ("endwhere", "end where"):
            sc.where = None
            return []
        m = re.match(r"^allocate\s*\((.*)\)$", st)
        if m:
            lines = []
            for part in split_top(tokenize(m.group(1))):
                nm = part[0][1]
                v = sc.vars[nm]
                ext = split_top(part[2:matching_paren(part, 1)])
                for q, e in enumerate(ext):
                    lines.append(f"{nm}_n{q + 1} = {expr_c(e, sc)};")
                lines.append(f"{nm} = ({v['type']} *)calloc({sc.total(nm)} + 1, sizeof({v['type']}));   /* undefined in Fortran: zero here */")
            return lines
        m = re.match(r"^deallocate\s*\((.*)\)$", st)
        if m:
            return [f"free({x.strip()}); {x.strip()} = 0;" for x in m.group(1).split(",")]
        if st in ("end do", "enddo"):
            if not do_stack or do_stack[-1] is not None:
                raise ValueError(f"line {no}: end do without do")
            do_stack.pop()
            return ["}"]
        if st in ("end if", "endif"):
            return ["}"]
        if st == "else":
            return ["} else {"]
        toks = tokenize(st)
        if toks[0] == ("id", "do"):
            k = 1
            lab = None
            if toks[1][0] == "num":
                lab, k = toks[1][1], 2
            var = toks[k][1]
            assert toks[k + 1] == ("op", "="), st
            parts = split_top(toks[k + 2:])
            lo, hi = expr_c(parts[0], sc), expr_c(parts[1], sc)
            step = expr_c(parts[2], sc) if len(parts) > 2 else "1"
            do_stack.append(lab)
            # Fortran evaluates the bounds (and the stride) once; strides in the subset are positive
            if step != "1":
                return [f"for (int do_hi_{no} = ({var} = {lo}, {hi}), do_st_{no} = {step}; {var} <= do_hi_{no}; {var} += do_st_{no}) {{"]
            return [f"for (int do_hi_{no} = ({var} = {lo}, {hi}); {var} <= do_hi_{no}; ++{var}) {{"]
        if toks[0] == ("id", "else") and len(toks) > 1 and toks[1] == ("id", "if"):
            j = matching_paren(toks, 2)
            return [f"}} else if ({expr_c(toks[3:j], sc)}) {{"]
        if toks[0] == ("id", "if"):
            j = matching_paren(toks, 1)
            cond = expr_c(toks[2:j], sc)
            rest = toks[j + 1:]
            if rest == [("id", "then")]:
                return [f"if ({cond}) {{"]
            # one-line if: the statement text after the condition
            depth, pos = 0, None
            for idx, ch in enumerate(st):
                if ch == "(":
                    depth += 1
                elif ch == ")":
                    depth -= 1
                    if depth == 0:
                        pos = idx
                        break
            inner = self.statement(st[pos + 1:].strip(), sc, do_stack, no)
            return [f"if ({cond}) {{"] + ["  " + x for x in inner] + ["}"]
        # assignment
        depth, eq = 0, None
        for idx, t in enumerate(toks):
            if t == ("op", "("):
                depth += 1
            elif t == ("op", ")"):
                depth -= 1
            elif depth == 0 and t == ("op", "="):
                eq = idx
                break
        if eq is None:
            raise ValueError(f"line {no}: statement not in the subset: {st!r}")
        lhs, rhs = toks[:eq], toks[eq + 1:]
        if len(lhs) == 1 and sc.is_array(lhs[0][1]):
            # whole-array assignment: array = array of the same shape (element by element) or array = scalar
            nm = lhs[0][1]
            if len(rhs) == 1 and rhs[0][0] == "id" and sc.is_array(rhs[0][1]):
                return [f"for (size_t z_ = 0; z_ < {sc.total(nm)}; ++z_) {nm}[z_] = {rhs[0][1]}[z_];"]
            return [f"for (size_t z_ = 0; z_ < {sc.total(nm)}; ++z_) {nm}[z_] = {expr_c(rhs, sc)};"]
        # array section on the left: a loop nest, innermost = first dimension (Fortran's array element order)
        sections = len(lhs) > 1 and lhs[1] == ("op", "(") and any(
            any(t == ("op", ":") for t in a) for a in split_top(lhs[2:matching_paren(lhs, 1)]))
        if not sections:
            return [f"{expr_c(lhs, sc)} = {expr_c(rhs, sc)};"]
        # bounds of each section of the left-hand side (':' alone = the whole declared extent)
        name = lhs[0][1]
        dims = sc.vars[name]["dims"]
        args = split_top(lhs[2:matching_paren(lhs, 1)])
        loops, k = [], 0
        for q, a in enumerate(args):
            if any(t == ("op", ":") for t in a):
                c = a.index(("op", ":"))
                lo = expr_c(a[:c], sc) if a[:c] else "1"
                hi = expr_c(a[c + 1:], sc) if a[c + 1:] else expr_c(dims[q], sc)
                loops.append((f"s{k}_{no}", lo, hi))
                k += 1
        sc.section_vars = [v for v, _, _ in loops]
        try:
            assign = f"{expr_c(lhs, sc)} = {expr_c(rhs, sc)};"
            if sc.where:
                mask = expr_c(sc.where[0], sc)
                assign = f"if ({'!' if sc.where[1] else ''}({mask})) {{ {assign} }}"
        finally:
            sc.section_vars = None
        lines = []
        for v, lo, hi in reversed(loops):                       # last section outermost
            lines.append(f"for (int {v} = 0; {v} <= ({hi}) - ({lo}); ++{v}) {{")
        lines.append("  " + assign)
        lines += ["}"] * len(loops)
        return lines


PRELUDE = r"""/* GENERATED by oracle/f2c_subset.py from the reference's Fortran source -- do not edit, do not commit.
   Test infrastructure: a mechanical, statement-by-statement translation used to pin oracle/gritas_oracle.c. */
#include <math.h>
#include <setjmp.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#define F_ABS(x) _Generic((x), int: abs, long: labs, default: fabs)(x)
#define F_MIN(a, b) ((a) < (b) ? (a) : (b))
#define F_MAX(a, b) ((a) > (b) ? (a) : (b))
__thread jmp_buf gritas_ref_jmp;            /* per thread: bench.py times the routines on several host threads */
__thread int gritas_ref_error_line = 0;
static void gritas_ref_perror(int line) { gritas_ref_error_line = line; longjmp(gritas_ref_jmp, 7); }
"""


def module_globals(text, first, last):
    """module-level declarations (m_gritas_grids.f:21-72) -> C globals + the scope entries the routines see"""
    g, c = {}, []
    for label, st, no in logical_lines(text, first, last):
        d = parse_decl(st)
        if not d:
            continue
        typ, attrs, ents = d
        if typ == "char":
            for nm, dims, init in ents:
                if nm == "type_stats":                                # (/'mean','rms','stdv'/): compared after trim()
                    lits = [t[1][1:-1].strip() for t in init if t[0] == "str"]
                    c.append("static const char *" + nm + "[] = {" + ", ".join(f'"{x}"' for x in lits) + "};   /* line " + str(no) + " */")
                    c.append(f"#define {nm}(i1) {nm}[(size_t)(i1) - 1]")
                    g[nm] = {"type": "char", "dims": [[("num", str(len(lits)))]], "optional": False, "arg": False, "define": c[-1]}
            continue
        for nm, dims, init in ents:
            if "parameter" in attrs:
                sc = Scope(g)
                c.append(f"static const {typ} {nm} = {expr_c(init, sc)};   /* line {no} */")
                g[nm] = {"type": typ, "dims": None, "optional": False, "arg": False}
            elif "allocatable" in attrs:
                if nm == "plevs":
                    c.append(f"{typ} *{nm};   /* line {no}: allocatable, 1-based */")
                    c.append(f"#define {nm}(i1) {nm}[(size_t)(i1) - 1]")
                    g[nm] = {"type": typ, "dims": [[("id", "nlevs")]], "optional": False, "arg": False, "define": c[-1]}
                elif nm in ("lonindx", "regindx"):                    # (2, nslots) / (2, nregions)
                    c.append(f"{typ} *{nm}; int {nm}_n2;   /* line {no}: allocatable, 1-based, column-major */")
                    c.append(f"#define {nm}(i1, i2) {nm}[(size_t)(i1) - 1 + (size_t)2 * ((size_t)(i2) - 1)]")
                    g[nm] = {"type": typ, "dims": [[("num", "2")], [("id", f"{nm}_n2")]], "optional": False, "arg": False, "define": c[-1]}
                    g[f"{nm}_n2"] = {"type": "int", "dims": None, "optional": False, "arg": False}
                elif nm == "achan":                                   # (nchan, 2): channel index / channel number
                    c.append(f"{typ} *{nm}; int {nm}_n1;   /* line {no}: allocatable, 1-based, column-major */")
                    c.append(f"#define {nm}(i1, i2) {nm}[(size_t)(i1) - 1 + (size_t)({nm}_n1) * ((size_t)(i2) - 1)]")
                    g[nm] = {"type": typ, "dims": [[("id", f"{nm}_n1")], [("num", "2")]], "optional": False, "arg": False, "define": c[-1]}
                    g[f"{nm}_n1"] = {"type": "int", "dims": None, "optional": False, "arg": False}
            else:
                c.append(f"{typ} {nm};   /* line {no} */")
                g[nm] = {"type": typ, "dims": None, "optional": False, "arg": False}
    return g, c


def statements_range(text, first, last, globals_, cname, params=(), locals_=(), free=False):
    """a run of plain statements (no declarations) as a C function.  params / locals_: (ctype, name, dims or None, byref)
    -- the declarations the statements rely on, stated by the caller (they sit far away in the program unit)."""
    tr = Translator(text, globals_)
    sc = Scope(globals_)
    sig = []
    for typ, nm, dims, byref in params:
        d = [tokenize(x) for x in dims] if dims else None
        sc.vars[nm] = {"type": typ, "dims": d, "optional": False, "arg": True, "byref": bool(byref)}
        sig.append(f"{typ} *{nm}" if (dims or byref) else f"{typ} {nm}")
    tr.out.append(f"/* statements of lines {first}-{last} */")
    tr.out.append(f"void {cname}({', '.join(sig) or 'void'}) {{")
    for typ, nm, dims, _ in locals_:
        sc.vars[nm] = {"type": typ, "dims": None, "optional": False, "arg": False}
        tr.out.append(f"  {typ} {nm} = 0;")
    macros = []
    for typ, nm, dims, byref in params:
        if dims:
            macros.append(nm)
            if nm in globals_ and globals_[nm].get("define"):
                tr.out.append(f"#undef {nm}")
            tr.out.append(Translator.accessor(nm, sc.vars[nm]["dims"], sc))
    tr.statements(logical_lines(text, first, last, free), sc)
    for nm in macros:
        tr.out.append(f"#undef {nm}")
        if nm in globals_ and globals_[nm].get("define"):
            tr.out.append(globals_[nm]["define"])                       # a parameter of the same name shadowed the module's array
    tr.out.append("}")
    tr.out.append("")
    return tr.out


def find_line(text, pattern, after=1):
    for no, l in enumerate(text.split("\n"), 1):
        if no >= after and re.search(pattern, l, re.I):
            return no
    raise ValueError(f"pattern {pattern!r} not found")


def main(ref_root, dst):
    """ref_root: the reference tree (…/src); dst: the C file to write"""
    import os
    grids = open(os.path.join(ref_root, "Components/gritas/m_gritas_grids.f")).read()
    kxkt = open(os.path.join(ref_root, "Components/gritas/gritas_kxkt.f")).read()
    app = open(os.path.join(ref_root, "Applications/GrITAS_App/gritas.f")).read()

    g, decls = module_globals(grids, 21, 72)
    # m_odsmeta (GMAO_ods, not in the tree) supplies the table bounds: the harness sets them
    for nm in ("kxmax", "ktmax"):
        g[nm] = {"type": "int", "dims": None, "optional": False, "arg": False}
        decls.append(f"int {nm};   /* m_odsmeta */")
    out = [PRELUDE] + decls + ["int is_surface_(int kt);   /* gritas_kxkt.f:121-171 needs ktSurfAll of m_odsmeta: supplied by the harness */", ""]

    # m_gritas_grids.f: the two statements of init_ that define dLon / dLat, then the routines
    i0 = find_line(grids, r"^\s*dLon\s*=\s*360\.")
    out += ["/* ===== m_gritas_grids.f ===== */"] + statements_range(grids, i0, i0 + 1, g, "gritas_ref_mesh")
    tr = Translator(grids, g)
    for name in ("GrITAS_Pint_", "GrITAS_Accum_", "GrITAS_Norm_", "GrITAS_MinMax_", "GrITAS_3CH_Norm_",
                 "GrITAS_XCov_Accum_", "GrITAS_XCov_Norm_", "GrITAS_XCov_Prs_Accum_"):
        tr.routine(name, name.lower())
    tr.routine("ave_nomiss", "gritas_ave_nomiss_")
    tr.routine("scores2_", "gritas_scores2_")
    out += tr.out
    # the scorecard's latitude rows: the grid latitudes (:136-138) and the row search of init_ (:165-175)
    r0 = find_line(grids, r"^\s*glat\(j\) = LatMin")
    out += statements_range(grids, r0 - 1, r0 + 1, g, "gritas_glat_", params=[("double", "glat", ["jm"], False)],
                            locals_=[("int", "j", None, False)])
    r1 = find_line(grids, r"^\s*do j = 1, jm\s*$", find_line(grids, r"Wired regions"))
    r2 = find_line(grids, r"^\s*end do\s*$", r1)
    out += statements_range(grids, r1, r2, g, "gritas_region_rows_", params=[
        ("int", "i", None, False), ("int", "nreg", None, False), ("double", "glat", ["jm"], False),
        ("double", "regbounds", ["2", "nreg"], False), ("int", "regindx", ["2", "nreg"], False)],
        locals_=[("int", "j", None, False)])

    # m_gritas_masks.F90 (free form): the mesh of the mask grid (:152-153) and the loop of maskout_ (:164-183)
    masks = open(os.path.join(ref_root, "Components/gritas/m_gritas_masks.F90")).read()
    out += ["/* ===== m_gritas_masks.F90 ===== */"]
    gm = dict(g)
    for pat_ in (r"^\s*real\s*::\s*lonMin_", r"^\s*real\s*::\s*latMin_"):
        no = find_line(masks, pat_)
        typ, attrs, ents = parse_decl(logical_lines(masks, no, no, free=True)[0][1])
        nm, _, init = ents[0]
        out.append(f"static const {typ} {nm} = {expr_c(init, Scope(g))};   /* line {no} */")
        gm[nm] = {"type": typ, "dims": None, "optional": False, "arg": False}
    m0 = find_line(masks, r"^\s*dLon = 360\. / im")
    out += statements_range(masks, m0, m0 + 1, gm, "gritas_mask_mesh_", params=[
        ("int", "im", None, False), ("int", "jm", None, False), ("double", "dlon", None, True), ("double", "dlat", None, True)], free=True)
    m1 = find_line(masks, r"^\s*do n=1,nobs", m0)
    m2 = find_line(masks, r"^\s*enddo\s*$", m1)
    out += statements_range(masks, m1, m2, gm, "gritas_mask_loop_", params=[
        ("int", "nobs", None, False), ("int", "im", None, False), ("int", "jm", None, False), ("double", "dlon", None, False),
        ("double", "dlat", None, False), ("double", "lat", ["nobs"], False), ("double", "lon", ["nobs"], False),
        ("int", "qcx", ["nobs"], False), ("double", "maskfld", ["im", "jm"], False), ("int", "nexcl", None, True)],
        locals_=[("int", "n", None, False), ("int", "i", None, False), ("int", "j", None, False)], free=True)

    # gritas_kxkt.f: GrITAS_KxKt
    out += ["/* ===== gritas_kxkt.f ===== */"]
    tk = Translator(kxkt, g)
    tk.routine("GrITAS_KxKt", "gritas_kxkt_")
    out += tk.out

    # gritas.f: the residual selection chain (:562-704) and the QC / time / longitude filter loop (:712-735); their
    # declarations are at :226-257 of the program unit and are restated here as the functions' parameters
    out += ["/* ===== gritas.f ===== */"]
    a = find_line(app, r"^\s*if \( iopt \.eq\. 0 \.or\. iopt \.eq\. 1 \) then")
    stop = find_line(app, r"mod\(iopt,2\)\s*\.ne\.\s*0", a)
    b = max(no for no, l in enumerate(app.split("\n"), 1) if a < no < stop and re.match(r"^\s*end\s*if\s*$", l, re.I))
    col = lambda t, n: (t, n, ["nobs"], False)
    out += statements_range(app, a, b, g, "gritas_select_", params=[
        ("int", "iopt", None, False), ("int", "scale_res", None, False), ("int", "xcov", None, False),
        ("int", "pxcov", None, False), ("int", "tch", None, False), ("int", "self", None, False),
        ("int", "meanres", None, False), ("int", "nobs", None, False), ("int", "nr", None, False),
        ("int", "x_passive", None, False), ("int", "nstat", None, True), ("double", "del", ["nobs", "nr"], False),
        col("double", "ods_data_omf"), col("double", "ods_data_oma"), col("double", "ods_data_obs"),
        col("double", "ods_data_xvec"), col("double", "ods_data_xm"), col("int", "ods_data_qcexcl"),
        col("double", "mods_data_obs"), col("double", "mods_data_omf"), col("double", "mods_data_oma")],
        locals_=[("int", "iii", None, False)])
    f0 = find_line(app, r"^\s*do i = 1, nobs\s*$", stop)
    f1 = find_line(app, r"^\s*end do\s*$", f0)
    out += statements_range(app, f0, f1, g, "gritas_filter_", params=[
        ("int", "nobs", None, False), ("int", "ioptn", None, False), ("int", "t1", None, False), ("int", "t2", None, False),
        ("double", "lona", None, False), ("double", "lonb", None, False), ("int", "x_passive", None, False),
        ("int", "x_pre_bad", None, False), col("int", "ods_data_qcexcl"), col("int", "ods_data_time"),
        col("double", "ods_data_lon"), col("int", "ob_flag")],
        locals_=[("int", "i", None, False), ("int", "iampassive", None, False), ("int", "iwantpassive", None, False),
                 ("int", "timeselect", None, False), ("int", "lonselect", None, False)])
    # gritas.f:500-508: the eight index sorts before the accumulation (key order and direction; the sort itself is m_MergeSorts')
    s0 = find_line(app, r"^\s*call IndexSet\s*\(\s*nobs, is\s*\)")
    s1 = s0
    while re.match(r"^\s*call IndexSort", app.split("\n")[s1]):      # s1 is 1-based: index s1 is the following line
        s1 += 1
    out += ["void indexset_(int n, int *is); void indexsort_i_(int n, int *is, int *key, int descend); "
            "void indexsort_d_(int n, int *is, double *key, int descend);   /* m_MergeSorts: supplied by the harness */"]
    out += statements_range(app, s0, s1, g, "gritas_presort_", params=[
        ("int", "nobs", None, False), ("int", "is", ["nobs"], False), col("int", "ods_data_qcexcl"), col("double", "ods_data_lat"),
        col("double", "ods_data_lon"), col("double", "ods_data_lev"), col("int", "ods_data_time"), col("int", "ods_data_kt"),
        col("int", "ods_data_ks"), col("int", "ods_data_kx")])
    # gritas.f:738, right after the filter loop: omf / sigo for the Jo options
    j0 = find_line(app, r"if\(iopt==17 \.or\. iopt==19\)", f1)
    out += statements_range(app, j0, j0, g, "gritas_jo_scale_", params=[
        ("int", "iopt", None, False), ("int", "nobs", None, False), ("int", "nr", None, False),
        ("double", "del", ["nobs", "nr"], False)])
    open(dst, "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
