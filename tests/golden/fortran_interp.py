#!/usr/bin/env python3
"""A small interpreter for the Fortran subset the reference's host code is written in.

Purpose: tests/golden/make_golden_step.py EXECUTES THE REFERENCE'S OWN SOURCE TEXT -- the field set-up of the main
program (flux_calculator.F90 STEP 1.4-1.7), the prepare_* routines, the nine calculators of
flux_calculator_calculate.F90 and the time loop (STEP 2) -- instead of re-typing their loops.  No Fortran compiler
exists in this image; this is the closest thing to running the reference.  Only used to GENERATE golden files in the
build container (it reads /root/reference); never imported by the product or by the tests on the GPU box.

Supported: free-form source, `&` continuations, `;`, comments, cpp lines (skipped); declarations with
DIMENSION / PARAMETER / POINTER / ALLOCATABLE / initialisers / array constructors; derived types with pointer
components; IF / ELSEIF / ELSE, one-line IF, counted DO; assignment (scalar, whole array, section `(:)`), pointer
assignment `=>`, ALLOCATE, NULLIFY, CALL (internal subroutines by reference, external ones as Python callables);
expressions with Fortran precedence, character concatenation and blank-insensitive comparison, integer division,
`x**n` by repeated squaring; intrinsics TRIM, ASSOCIATED, PRESENT, MAX, MIN, ABS, SQRT, EXP, REAL, INT, SIZE.
WRITE / FLUSH / FORMAT are ignored except `WRITE(charvar,'(I0.2)') i` (internal write used for numtype).
Binary64 arithmetic throughout (the reference is built with -r8).
"""
import math
import re

# --------------------------------------------------------------------------------------------
# source -> logical statements
# --------------------------------------------------------------------------------------------

def _strip_comment(line):
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


def _split_semicolon(s):
    parts, cur, q = [], "", None
    for ch in s:
        if q:
            cur += ch
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur += ch
        elif ch == ";":
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return [p.strip() for p in parts if p.strip()]


def logical_lines(text):
    """[(first line number, statement)] with continuations joined and `;` split"""
    res, cur, start = [], "", None
    for no, raw in enumerate(text.splitlines(), 1):
        if raw.lstrip().startswith("#"):
            continue
        line = _strip_comment(raw).strip()
        if not line:
            continue
        if cur and line.startswith("&"):
            line = line[1:].lstrip()
        if start is None:
            start = no
        if line.endswith("&"):
            cur += line[:-1] + " "
            continue
        cur += line
        for part in _split_semicolon(cur):
            res.append((start, part))
        cur, start = "", None
    return res


# --------------------------------------------------------------------------------------------
# tokens and expressions
# --------------------------------------------------------------------------------------------
_TOK = re.compile(r"""\s*(?:
    (?P<real>(?:\d+\.\d*|\.\d+)(?:[eEdD][+-]?\d+)?|\d+[eEdD][+-]?\d+)
  | (?P<int>\d+)
  | (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
  | (?P<dot>\.(?:lt|le|gt|ge|eq|ne|and|or|not|true|false|eqv|neqv)\.)
  | (?P<name>[A-Za-z_][A-Za-z_0-9]*)
  | (?P<op>\*\*|//|==|/=|<=|>=|=>|[-+*/()<>,%:=\[\]])
)""", re.X | re.I)


def tokenize(s):
    pos, toks = 0, []
    s = s.rstrip()
    while pos < len(s):
        m = _TOK.match(s, pos)
        if not m or m.end() == pos:
            raise SyntaxError("cannot tokenize %r at %d" % (s, pos))
        pos = m.end()
        if m.group("real") is not None:
            toks.append(("real", float(m.group("real").lower().replace("d", "e"))))
        elif m.group("int") is not None:
            toks.append(("int", int(m.group("int"))))
        elif m.group("str") is not None:
            t = m.group("str")
            toks.append(("str", t[1:-1].replace(t[0] * 2, t[0])))
        elif m.group("dot"):
            d = m.group("dot").lower()
            if d == ".true.":
                toks.append(("bool", True))
            elif d == ".false.":
                toks.append(("bool", False))
            else:
                toks.append(("op", d))
        elif m.group("name"):
            toks.append(("name", m.group("name").lower()))
        else:
            toks.append(("op", m.group("op")))
    return toks


_REL = {".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">=", ".eq.": "==", ".ne.": "/="}


class Parser:
    """Fortran expression grammar (F2008 7.1.3): ** > * / > unary +- > + - > // > relational > .not. > .and. > .or."""

    def __init__(self, toks, i=0):
        self.t, self.i = toks, i

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else (None, None)

    def accept(self, kind, val=None):
        k, v = self.peek()
        if k == kind and (val is None or v == val):
            self.i += 1
            return True
        return False

    def expect(self, kind, val=None):
        k, v = self.peek()
        if k != kind or (val is not None and v != val):
            raise SyntaxError("expected %s %s, got %s %s in %r" % (kind, val, k, v, self.t))
        self.i += 1
        return v

    def expr(self):
        a = self.and_()
        while self.peek() == ("op", ".or."):
            self.i += 1
            a = ("or", a, self.and_())
        return a

    def and_(self):
        a = self.not_()
        while self.peek() == ("op", ".and."):
            self.i += 1
            a = ("and", a, self.not_())
        return a

    def not_(self):
        if self.peek() == ("op", ".not."):
            self.i += 1
            return ("not", self.not_())
        return self.rel()

    def rel(self):
        a = self.cat()
        k, v = self.peek()
        if k == "op" and (v in _REL or v in ("<", "<=", ">", ">=", "==", "/=")):
            self.i += 1
            return ("rel", _REL.get(v, v), a, self.cat())
        return a

    def cat(self):
        a = self.add()
        while self.peek() == ("op", "//"):
            self.i += 1
            a = ("cat", a, self.add())
        return a

    def add(self):
        k, v = self.peek()
        if k == "op" and v in ("+", "-"):
            self.i += 1
            a = self.mul()
            if v == "-":
                a = ("neg", a)
        else:
            a = self.mul()
        while True:
            k, v = self.peek()
            if k == "op" and v in ("+", "-"):
                self.i += 1
                a = ("bin", v, a, self.mul())
            else:
                return a

    def mul(self):
        a = self.power()
        while True:
            k, v = self.peek()
            if k == "op" and v in ("*", "/"):
                self.i += 1
                a = ("bin", v, a, self.power())
            else:
                return a

    def power(self):
        a = self.primary()
        if self.peek() == ("op", "**"):
            self.i += 1
            return ("pow", a, self.power())
        return a

    def primary(self):
        k, v = self.peek()
        if k in ("int", "real", "str", "bool"):
            self.i += 1
            return ("lit", v)
        if k == "op" and v == "(":
            self.i += 1
            e = self.expr()
            self.expect("op", ")")
            return ("paren", e)
        if k == "op" and v == "[":
            self.i += 1
            items = [self.expr()]
            while self.accept("op", ","):
                items.append(self.expr())
            self.expect("op", "]")
            return ("array", items)
        if k == "name":
            return self.designator()
        raise SyntaxError("unexpected token %s %s in %r" % (k, v, self.t))

    def designator(self):
        parts = [("name", self.expect("name"))]
        while True:
            if self.accept("op", "("):
                args = []
                if not self.accept("op", ")"):
                    while True:
                        if self.peek() == ("op", ":"):
                            self.i += 1
                            args.append(("colon",))
                        else:
                            args.append(self.expr())
                        if self.accept("op", ")"):
                            break
                        self.expect("op", ",")
                parts.append(("args", args))
            elif self.accept("op", "%"):
                parts.append(("comp", self.expect("name")))
            else:
                return ("des", parts)


def parse_expr(s):
    p = Parser(tokenize(s))
    e = p.expr()
    if p.i != len(p.t):
        raise SyntaxError("trailing tokens in %r" % s)
    return e


def split_top(s, sep=","):
    out, depth, cur, q = [], 0, "", None
    for ch in s:
        if q:
            cur += ch
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
        elif ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == sep and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [x.strip() for x in out]


def powi(x, m):
    """x**m for integer m by repeated squaring (what ifort / gfortran emit for small constant powers)"""
    n = abs(m)
    y = x if n % 2 else (1.0 if isinstance(x, float) else 1)
    n >>= 1
    while n:
        x = x * x
        if n % 2:
            y = y * x
        n >>= 1
    return 1.0 / y if m < 0 else y


# --------------------------------------------------------------------------------------------
# run-time objects
# --------------------------------------------------------------------------------------------
class FortranStop(Exception):
    """CALL mpi_finalize(1) / oasis_abort / STOP: the reference would end here"""


class FArr:
    """n-dimensional array with arbitrary lower bounds; elements in a flat Python list"""

    def __init__(self, bounds, fill):
        self.bounds = [(int(lo), int(hi)) for lo, hi in bounds]
        n = 1
        for lo, hi in self.bounds:
            n *= max(0, hi - lo + 1)
        self.data = [fill() if callable(fill) else fill for _ in range(n)]

    def flat(self, idx):
        if len(idx) != len(self.bounds):
            raise IndexError("rank mismatch: %r for bounds %r" % (idx, self.bounds))
        off, stride = 0, 1
        for (lo, hi), i in zip(self.bounds, idx):
            i = int(i)
            if i < lo or i > hi:
                raise IndexError("subscript %d out of bounds %d:%d" % (i, lo, hi))
            off += (i - lo) * stride
            stride *= hi - lo + 1
        return off

    def __len__(self):
        return len(self.data)


class FObj:
    def __init__(self, tname, comps):
        self.tname, self.c = tname, comps


class Alias:
    """a scalar dummy argument bound to the caller's storage"""

    def __init__(self, ref):
        self.ref = ref


class Ref:
    __slots__ = ("kind", "box", "key")

    def __init__(self, kind, box, key=None):
        self.kind, self.box, self.key = kind, box, key

    def get(self):
        if self.kind == "val":
            return self.box
        v = self.box.c[self.key] if self.kind == "comp" else (self.box.data[self.key] if self.kind == "elem" else self.box[self.key])
        if isinstance(v, Alias):
            return v.ref.get()
        return v

    def set(self, val):
        if self.kind == "val":
            raise RuntimeError("assignment to an expression")
        if self.kind == "comp":
            self.box.c[self.key] = val
        elif self.kind == "elem":
            self.box.data[self.key] = val
        else:
            cur = self.box.get(self.key)
            if isinstance(cur, Alias):
                cur.ref.set(val)
            else:
                self.box[self.key] = val


class Sub:
    def __init__(self, name, args, file, line):
        self.name, self.args, self.file, self.line = name, args, file, line
        self.decls = []      # raw declaration statements
        self.body = []       # nested statements


def _fstr_eq(a, b):
    return str(a).rstrip() == str(b).rstrip()


# --------------------------------------------------------------------------------------------
# the interpreter
# --------------------------------------------------------------------------------------------
_DECL = re.compile(r"^((?:integer|real|logical|character|double\s+precision)\b|type\s*\(\s*\w+\s*\))", re.I)


class Interp:
    def __init__(self):
        self.subs = {}
        self.types = {}        # type name -> [(component, decl dict)]
        self.glob = {}         # module variables of all loaded modules (one flat namespace, like `use` of everything)
        self.gdecl = {}
        self.externals = {}    # name -> python callable(interp, refs)
        self.log = []          # what the reference writes to its log file (w_unit)
        self.stack = []        # (subroutine, line of the call) for error reports

    # ---- loading ---------------------------------------------------------------------------
    def load_module(self, path):
        """types, module variables (with initialisers) and subroutines of one source file"""
        lines = logical_lines(open(path, encoding="utf-8", errors="replace").read())
        i, in_contains = 0, False
        while i < len(lines):
            no, st = lines[i]
            low = st.lower()
            m = re.match(r"type\s+(\w+)$", low)
            if m and not in_contains:
                comps = []
                i += 1
                while not re.match(r"end\s*type", lines[i][1], re.I):
                    for nm, d in self.parse_decl(lines[i][1]):
                        comps.append((nm, d))
                    i += 1
                self.types[m.group(1)] = comps
            elif re.match(r"contains$", low):
                in_contains = True
            elif re.match(r"subroutine\s+\w+", low):
                i = self.parse_sub(lines, i, path)
            elif low.startswith("enum"):
                val = 0
                i += 1
                while not re.match(r"end\s*enum", lines[i][1], re.I):
                    m2 = re.match(r"enumerator\s*::\s*(\w+)\s*(?:=\s*(.+))?$", lines[i][1], re.I)
                    if m2:
                        if m2.group(2):
                            val = int(self.eval(parse_expr(m2.group(2)), self.glob))
                        self.glob[m2.group(1).lower()] = val
                        val += 1
                    i += 1
            elif not in_contains and _DECL.match(st) and "::" in st:
                for nm, d in self.parse_decl(st):
                    self.gdecl[nm] = d
                    self.glob[nm] = self.instantiate(d, self.glob)
            i += 1

    def parse_sub(self, lines, i, path, kind="subroutine"):
        no, st = lines[i]
        m = re.match(r"subroutine\s+(\w+)\s*(?:\((.*)\))?\s*$", st, re.I)
        sub = Sub(m.group(1).lower(), [a.strip().lower() for a in (m.group(2) or "").split(",") if a.strip()], path, no)
        i += 1
        flat = []
        while not re.match(r"end\s*subroutine", lines[i][1], re.I):
            flat.append(lines[i])
            i += 1
        sub.decls = [s for s in flat if _DECL.match(s[1]) and "::" in s[1]]
        try:
            sub.body = self.nest([s for s in flat if not (_DECL.match(s[1]) and "::" in s[1])])
            self.subs[sub.name] = sub
        except SyntaxError as e:      # a routine outside the supported subset (I/O, MPI ...): callable only as an external
            self.unparsed = getattr(self, "unparsed", {})
            self.unparsed[sub.name] = str(e)
        return i

    def load_program(self, path):
        """main program: declarations -> sub.decls, all executable statements kept FLAT with line numbers"""
        lines = logical_lines(open(path, encoding="utf-8", errors="replace").read())
        sub = Sub("__main__", [], path, 1)
        sub.decls = [s for s in lines if _DECL.match(s[1]) and "::" in s[1]]
        sub.flat = [s for s in lines if not (_DECL.match(s[1]) and "::" in s[1])]
        self.subs["__main__"] = sub
        return sub

    # ---- declarations ----------------------------------------------------------------------
    def parse_decl(self, st):
        """'<type> [, attrs] :: a, b(3) = init' -> [(name, {base, tname, dims, init, pointer, alloc, param})]"""
        left, right = st.split("::", 1)
        m = _DECL.match(left.strip())
        base = m.group(1).lower()
        tname = None
        if base.startswith("type"):
            tname = re.match(r"type\s*\(\s*(\w+)\s*\)", base).group(1)
            base = "type"
        rest = left.strip()[m.end():]
        # skip a kind / len selector right after the type keyword
        rest = rest.lstrip()
        if rest.startswith("("):
            depth = 0
            for k, ch in enumerate(rest):
                depth += ch == "("
                depth -= ch == ")"
                if depth == 0:
                    rest = rest[k + 1:]
                    break
        attrs = [a.strip().lower() for a in split_top(rest) if a.strip()]
        dims = None
        for a in attrs:
            m2 = re.match(r"dimension\s*\((.*)\)$", a)
            if m2:
                dims = split_top(m2.group(1))
        d0 = dict(base=base, tname=tname, pointer="pointer" in attrs, alloc="allocatable" in attrs,
                  param="parameter" in attrs)
        out = []
        for ent in split_top(right):
            init = None
            if "=>" in ent:
                ent = ent.split("=>")[0]
            ent = ent.strip()
            m3 = re.match(r"(\w+)\s*", ent)
            if not m3:
                raise SyntaxError("cannot parse entity %r in %r" % (ent, st))
            nm = m3.group(1).lower()
            rest_e = ent[m3.end():]
            edims = dims
            if rest_e.startswith("("):
                depth = 0
                for k, ch in enumerate(rest_e):
                    depth += ch == "("
                    depth -= ch == ")"
                    if depth == 0:
                        break
                edims = split_top(rest_e[1:k])
                rest_e = rest_e[k + 1:].strip()
            if rest_e.startswith("*"):      # character*n
                rest_e = re.sub(r"^\*\s*\w+", "", rest_e).strip()
            if rest_e.startswith("="):
                init = parse_expr(rest_e[1:])
            d = dict(d0)
            d.update(dims=edims, init=init)
            out.append((nm, d))
        return out

    def default_scalar(self, d):
        if d["base"] == "type":
            comps = {}
            for nm, cd in self.types[d["tname"]]:
                comps[nm] = self.instantiate(cd, self.glob)
            return FObj(d["tname"], comps)
        return {"integer": 0, "real": float("nan"), "logical": False, "character": "", "double precision": float("nan")}[d["base"]]

    def instantiate(self, d, scope):
        """storage for a declared entity (None for pointers / deferred shapes until allocated or associated)"""
        if d["pointer"] or d["alloc"]:
            return None
        if d["dims"]:
            if any(x.strip() == ":" or x.strip().endswith(":") and x.strip() != ":" and False for x in d["dims"]):
                return None
            bounds = []
            for x in d["dims"]:
                if ":" in x:
                    lo, hi = x.split(":")
                    if not hi.strip():
                        return None      # assumed shape: a dummy argument
                    bounds.append((self.eval(parse_expr(lo), scope), self.eval(parse_expr(hi), scope)))
                else:
                    bounds.append((1, self.eval(parse_expr(x), scope)))
            arr = FArr(bounds, (lambda: self.default_scalar(d)) if d["base"] == "type" else self.default_scalar(d))
            if d["init"] is not None:
                v = self.eval(d["init"], scope)
                if isinstance(v, list):
                    if len(v) != len(arr.data):
                        raise ValueError("array constructor has %d elements, array %d" % (len(v), len(arr.data)))
                    arr.data = [self.coerce(d, x) for x in v]
                else:
                    arr.data = [self.coerce(d, v)] * len(arr.data)
            return arr
        if d["init"] is not None:
            return self.coerce(d, self.eval(d["init"], scope))
        return self.default_scalar(d)

    @staticmethod
    def coerce(d, v):
        if d["base"] in ("real", "double precision"):
            return float(v)
        if d["base"] == "integer":
            return int(v)
        return v

    # ---- statement structure ---------------------------------------------------------------
    def nest(self, flat):
        """flat [(line, text)] -> nested statement list"""
        root = []
        stack = [root]
        ifs = []      # open IF nodes
        for no, st in flat:
            low = st.lower()
            m = re.match(r"(?:\w+\s*:\s*)?if\s*\((.*)\)\s*then$", st, re.I | re.S)
            if m:
                node = ["if", [(parse_expr(m.group(1)), [])], None, no]
                stack[-1].append(node)
                ifs.append(node)
                stack.append(node[1][0][1])
                continue
            m = re.match(r"else\s*if\s*\((.*)\)\s*then$", st, re.I | re.S)
            if m:
                stack.pop()
                blk = []
                ifs[-1][1].append((parse_expr(m.group(1)), blk))
                stack.append(blk)
                continue
            if re.match(r"else$", low):
                stack.pop()
                ifs[-1][2] = []
                stack.append(ifs[-1][2])
                continue
            if re.match(r"end\s*if$", low):
                stack.pop()
                ifs.pop()
                continue
            m = re.match(r"do\s+(\w+)\s*=\s*(.+)$", st, re.I)
            if m:
                parts = split_top(m.group(2))
                node = ["do", m.group(1).lower(), [parse_expr(p) for p in parts], [], no]
                stack[-1].append(node)
                stack.append(node[3])
                continue
            if re.match(r"end\s*do$", low):
                stack.pop()
                continue
            m = re.match(r"if\s*\(", st, re.I)
            if m:      # one-line IF: find the matching parenthesis
                depth, k = 0, m.end() - 1
                for k in range(m.end() - 1, len(st)):
                    depth += st[k] == "("
                    depth -= st[k] == ")"
                    if depth == 0:
                        break
                cond, rest = st[m.end():k], st[k + 1:].strip()
                inner = self.nest([(no, rest)])
                stack[-1].append(["if", [(parse_expr(cond), inner)], None, no])
                continue
            stack[-1].append(self.simple(no, st))
        if len(stack) != 1:
            raise SyntaxError("unbalanced block structure")
        return root

    def simple(self, no, st):
        low = st.lower()
        m = re.match(r"call\s+(\w+)\s*(?:\((.*)\))?$", st, re.I | re.S)
        if m:
            args = [parse_expr(a) for a in split_top(m.group(2))] if m.group(2) and m.group(2).strip() else []
            return ["call", m.group(1).lower(), args, no]
        m = re.match(r"write\s*\(\s*(\w+)\s*\(\s*(\w+)\s*\)\s*,\s*'\(I0\.2\)'\s*\)\s*(\w+)$", st, re.I)
        if m:      # internal write: numtype(i) <- two-digit i
            return ["iwrite", parse_expr("%s(%s)" % (m.group(1), m.group(2))), parse_expr(m.group(3)), no]
        m = re.match(r"write\s*\(\s*w_unit\s*,\s*\*\s*\)\s*(.*)$", st, re.I | re.S)
        if m:      # list-directed write to the log file: kept as text (error / warning messages of the reference)
            try:
                return ["write", [parse_expr(a) for a in split_top(m.group(1))], no]
            except SyntaxError:
                return ["nop", no]
        if re.match(r"(write|print|format|flush|open|close|read|use|implicit|public|private|program|end\s*program|module|end\s*module|contains|namelist|return|continue)\b", low):
            return ["nop", no]
        if re.match(r"stop\b", low):
            return ["stop", no]
        m = re.match(r"allocate\s*\((.*)\)$", st, re.I | re.S)
        if m:
            return ["allocate", [parse_expr(a) for a in split_top(m.group(1))], no]
        m = re.match(r"nullify\s*\((.*)\)$", st, re.I | re.S)
        if m:
            return ["nullify", [parse_expr(a) for a in split_top(m.group(1))], no]
        # assignment / pointer assignment: split at the top-level '=' or '=>'
        toks = tokenize(st)
        depth = 0
        for k, (kk, vv) in enumerate(toks):
            if kk == "op" and vv in "([":
                depth += 1
            elif kk == "op" and vv in ")]":
                depth -= 1
            elif depth == 0 and kk == "op" and vv in ("=", "=>"):
                pl = Parser(toks[:k])
                lhs = pl.designator()
                pr = Parser(toks[k + 1:])
                rhs = pr.expr()
                if pl.i != k or pr.i != len(toks) - k - 1:
                    raise SyntaxError("cannot parse assignment %r" % st)
                return ["ptr" if vv == "=>" else "assign", lhs, rhs, no]
        raise SyntaxError("unsupported statement at line %d: %r" % (no, st))

    # ---- evaluation ------------------------------------------------------------------------
    def lookup_scope(self, name, scope):
        if name in scope:
            return scope
        if name in self.glob:
            return self.glob
        return None

    def ref(self, des, scope, create=False):
        """designator -> Ref (or ("section", arr) for whole arrays / (:) sections)"""
        parts = des[1]
        name = parts[0][1]
        sc = self.lookup_scope(name, scope)
        if sc is None:
            if create:
                sc = scope
                scope[name] = None
            else:
                raise NameError("unknown variable %s" % name)
        r = Ref("var", sc, name)
        for p in parts[1:]:
            v = r.get()
            if p[0] == "comp":
                if not isinstance(v, FObj):
                    raise TypeError("%%%s of a non-structure (%r) in %r" % (p[1], v, des))
                r = Ref("comp", v, p[1])
            else:
                args = p[1]
                if len(args) == 1 and args[0] == ("colon",):
                    continue      # x(:) designates the whole array
                if not isinstance(v, FArr):
                    raise TypeError("subscript of a non-array in %r (value %r)" % (des, v))
                idx = [self.eval(a, scope) for a in args]
                r = Ref("elem", v, v.flat(idx))
        return r

    INTRINSICS = ("trim", "associated", "present", "max", "min", "abs", "sqrt", "exp", "real", "int", "size", "adjustl", "len_trim",
                  "selected_real_kind")

    def eval(self, e, scope):
        k = e[0]
        if k == "lit":
            return e[1]
        if k == "paren":
            return self.eval(e[1], scope)
        if k == "array":
            return [self.eval(x, scope) for x in e[1]]
        if k == "des":
            parts = e[1]
            name = parts[0][1]
            if self.lookup_scope(name, scope) is None and name in self.INTRINSICS and len(parts) == 2 and parts[1][0] == "args":
                return self.intrinsic(name, parts[1][1], scope)
            return self.ref(e, scope).get()
        if k == "neg":
            return -self.eval(e[1], scope)
        if k == "bin":
            a, b = self.eval(e[2], scope), self.eval(e[3], scope)
            op = e[1]
            if op == "+":
                return a + b
            if op == "-":
                return a - b
            if op == "*":
                return a * b
            if isinstance(a, int) and isinstance(b, int) and not isinstance(a, bool):
                return int(a / b)      # integer division truncates towards zero
            return a / b
        if k == "pow":
            base = self.eval(e[1], scope)
            ex = self.eval(e[2], scope)
            if isinstance(ex, int):
                return powi(base, ex)
            return math.pow(float(base), float(ex))
        if k == "cat":
            return str(self.eval(e[1], scope)) + str(self.eval(e[2], scope))
        if k == "rel":
            a, b = self.eval(e[2], scope), self.eval(e[3], scope)
            op = e[1]
            if isinstance(a, str) or isinstance(b, str):
                eq = _fstr_eq(a, b)
                if op == "==":
                    return eq
                if op == "/=":
                    return not eq
                raise TypeError("ordering of strings")
            return {"<": a < b, "<=": a <= b, ">": a > b, ">=": a >= b, "==": a == b, "/=": a != b}[op]
        if k == "not":
            return not self.eval(e[1], scope)
        if k == "and":
            a, b = self.eval(e[1], scope), self.eval(e[2], scope)      # Fortran may evaluate both
            return bool(a) and bool(b)
        if k == "or":
            a, b = self.eval(e[1], scope), self.eval(e[2], scope)
            return bool(a) or bool(b)
        raise RuntimeError("bad expression node %r" % (e,))

    def intrinsic(self, name, args, scope):
        if name == "associated":
            return self.ref(args[0], scope).get() is not None
        if name == "present":
            nm = args[0][1][0][1]
            return nm in scope and scope[nm] is not ABSENT
        vals = [self.eval(a, scope) for a in args]
        if name in ("trim",):
            return str(vals[0]).rstrip()
        if name == "adjustl":
            return str(vals[0]).lstrip()
        if name == "len_trim":
            return len(str(vals[0]).rstrip())
        if name == "max":
            return max(vals)
        if name == "min":
            return min(vals)
        if name == "abs":
            return abs(vals[0])
        if name == "sqrt":
            return math.sqrt(vals[0])
        if name == "exp":
            return math.exp(vals[0])
        if name == "real":
            return float(vals[0])
        if name == "int":
            return int(vals[0])
        if name == "size":
            return len(vals[0].data)
        if name == "selected_real_kind":
            return 8 if vals[0] > 6 else 4
        raise NameError(name)

    # ---- execution -------------------------------------------------------------------------
    def run(self, block, scope):
        for st in block:
            k = st[0]
            if k == "assign":
                self.assign(st[1], self.eval(st[2], scope), scope)
            elif k == "if":
                done = False
                for cond, blk in st[1]:
                    if self.eval(cond, scope):
                        self.run(blk, scope)
                        done = True
                        break
                if not done and st[2] is not None:
                    self.run(st[2], scope)
            elif k == "do":
                lo = int(self.eval(st[2][0], scope))
                hi = int(self.eval(st[2][1], scope))
                step = int(self.eval(st[2][2], scope)) if len(st[2]) > 2 else 1
                r = self.ref(("des", [("name", st[1])]), scope, create=True)
                i = lo
                while (step > 0 and i <= hi) or (step < 0 and i >= hi):
                    r.set(i)
                    self.run(st[3], scope)
                    i += step
                r.set(i)
            elif k == "call":
                self.call(st[1], st[2], scope, st[3])
            elif k == "ptr":
                tgt = self.ref(st[2], scope).get()
                self.ref(st[1], scope).set(tgt)
            elif k == "allocate":
                for a in st[1]:
                    self.allocate(a, scope)
            elif k == "nullify":
                for a in st[1]:
                    self.ref(a, scope).set(None)
            elif k == "iwrite":
                self.ref(st[1], scope).set("%02d" % int(self.eval(st[2], scope)))
            elif k == "write":
                try:
                    items = [self.eval(x, scope) for x in st[1]]
                    self.log.append(" ".join(str(x).rstrip() if isinstance(x, str) else str(x) for x in items))
                except Exception:      # MINVAL(...) and the like: debug output only
                    pass
            elif k == "stop":
                raise FortranStop("STOP at line %d" % st[1])
            elif k == "nop":
                pass
            else:
                raise RuntimeError("bad statement %r" % (st,))

    def assign(self, des, val, scope):
        r = self.ref(des, scope, create=False)
        cur = r.get()
        if isinstance(cur, FArr):
            if isinstance(val, FArr):
                if len(val.data) != len(cur.data):
                    raise ValueError("array assignment with different sizes")
                cur.data[:] = list(val.data)
            elif isinstance(val, list):
                cur.data[:] = val
            else:
                if cur.data and isinstance(cur.data[0], float):
                    val = float(val)
                for i in range(len(cur.data)):
                    cur.data[i] = val
            return
        if isinstance(cur, float) and isinstance(val, int) and not isinstance(val, bool):
            val = float(val)
        if isinstance(cur, int) and not isinstance(cur, bool) and isinstance(val, float):
            val = int(val)
        r.set(val)

    def allocate(self, des, scope):
        """ALLOCATE(x%field(n)) / ALLOCATE(arr(n)): element type from the declaration of the last name/component"""
        parts = des[1]
        if parts[-1][0] != "args":
            raise SyntaxError("ALLOCATE without shape")
        shape = [self.eval(a, scope) for a in parts[-1][1]]
        target = ("des", parts[:-1])
        r = self.ref(target, scope)
        # find the declaration
        if len(parts) == 2:
            d = self.find_decl(parts[0][1], scope)
        else:
            owner = self.ref(("des", parts[:-2]), scope).get()
            d = dict(self.types[owner.tname])[parts[-2][1]]
        fill = (lambda: self.default_scalar(d)) if d["base"] == "type" else self.default_scalar(d)
        r.set(FArr([(1, n) for n in shape], fill))

    def find_decl(self, name, scope):
        d = scope.get("__decl__", {}).get(name)
        if d is None:
            d = self.gdecl.get(name)
        if d is None:
            raise NameError("no declaration for %s" % name)
        return d

    def new_scope(self, sub):
        scope = {"__decl__": {}}
        for no, st in sub.decls:
            for nm, d in self.parse_decl(st):
                scope["__decl__"][nm] = d
                if nm in sub.args:
                    continue
                try:
                    scope[nm] = self.instantiate(d, scope)
                except (NameError, KeyError):      # a type / constant of a library module (OASIS, MPI): never touched
                    scope[nm] = None
        return scope

    def call(self, name, args, scope, line=None):
        if name in self.externals:
            refs = []
            for a in args:
                try:
                    if a[0] == "des" and self.lookup_scope(a[1][0][1], scope) is not None:
                        refs.append(self.ref(a, scope))
                    else:
                        refs.append(Ref("val", self.eval(a, scope)))
                except NameError:      # e.g. MPI_COMM_WORLD: a name of a library this interpreter does not load
                    refs.append(Ref("val", None))
            return self.externals[name](self, refs)
        if name not in self.subs:
            raise NameError("call of unknown subroutine %s (line %s)" % (name, line))
        sub = self.subs[name]
        callee = self.new_scope(sub)
        for formal, a in zip(sub.args, args):
            if a[0] == "des" and self.lookup_scope(a[1][0][1], scope) is not None:
                r = self.ref(a, scope)
                v = r.get()
                if isinstance(v, (FArr, FObj)):
                    callee[formal] = v
                else:
                    callee[formal] = Alias(r)
            else:
                callee[formal] = self.eval(a, scope)
        for formal in sub.args[len(args):]:
            callee[formal] = ABSENT
        self.stack.append((name, line))
        self.run(sub.body, callee)
        self.stack.pop()

    def run_main_range(self, scope, first, last):
        """execute the statements of the main program whose first line lies in [first, last]"""
        sub = self.subs["__main__"]
        self.run(self.nest([s for s in sub.flat if first <= s[0] <= last]), scope)


ABSENT = object()
