#!/usr/bin/env python3
"""Generate tests/golden/flux_lib_golden.json by INTERPRETING THE REFERENCE'S FORTRAN SOURCE.

Runs only in the build container (needs /root/reference); the JSON it writes is committed and is
what travels to the GPU box.  No Fortran compiler exists in this image, so instead of executing a
reference binary this script

  1. parses every subroutine of /root/reference/src/flux_lib/**/*.F90 (declarations with their
     initialisers, assignments, IF/ELSE, PRESENT(), CALL forwarding) and evaluates the statements
     with binary64 arithmetic in Fortran precedence / left-to-right order (`-r8` semantics: every
     real literal is a double; `x**<integer literal>` is repeated squaring like the compilers emit;
     `x**<real>` is libm pow; exp/sqrt are libm),
  2. extracts the call-site argument wiring of flux_calculator_calculate.F90 (which idx_* variable
     feeds which dummy argument, per method string) and the calculator order of the time loop in
     flux_calculator.F90 by pattern matching on the source text,
  3. evaluates seeded inputs (plus edge cases) through 1+2 and stores inputs/outputs as C99 hex
     floats (bit exact).

Nothing here is transcribed from the C oracle or the CUDA kernels: the formulas come from the
reference text at run time.  Usage:  python tests/golden/make_golden.py [--ref /root/reference]
"""
import argparse
import hashlib
import json
import math
import os
import random
import re
import sys

# --------------------------------------------------------------------------------------------
# Fortran source handling
# --------------------------------------------------------------------------------------------

def strip_comment(line):
    out, in_s, q = [], False, ""
    for ch in line:
        if in_s:
            out.append(ch)
            if ch == q:
                in_s = False
        elif ch in "'\"":
            in_s, q = True, ch
            out.append(ch)
        elif ch == "!":
            break
        else:
            out.append(ch)
    return "".join(out).rstrip()


def logical_lines(text):
    """Join free-form continuation lines; return [(first_line_no, statement)]."""
    res, cur, start = [], "", None
    for no, raw in enumerate(text.splitlines(), 1):
        line = strip_comment(raw).strip()
        if not line or line.startswith("#"):
            continue
        if cur and line.startswith("&"):
            line = line[1:].lstrip()
        if start is None:
            start = no
        if line.endswith("&"):
            cur += line[:-1] + " "
            continue
        cur += line
        res.append((start, cur.strip()))
        cur, start = "", None
    return res


TOKEN = re.compile(r"""\s*(?:
    (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[eEdD][+-]?\d+)?)
  | (?P<dotop>\.(?:lt|le|gt|ge|eq|ne|and|or|not)\.)
  | (?P<name>[A-Za-z_][A-Za-z_0-9]*)
  | (?P<op>\*\*|==|/=|<=|>=|[-+*/()<>,%])
)""", re.X | re.I)


def tokenize(s):
    pos, toks = 0, []
    s = s.strip()
    while pos < len(s):
        m = TOKEN.match(s, pos)
        if not m:
            raise SyntaxError("cannot tokenize %r at %d" % (s, pos))
        pos = m.end()
        if m.group("num") is not None:
            t = m.group("num")
            is_int = re.fullmatch(r"\d+", t) is not None
            toks.append(("int", int(t)) if is_int else ("real", float(t.lower().replace("d", "e"))))
        elif m.group("dotop"):
            toks.append(("op", m.group("dotop").lower()))
        elif m.group("name"):
            toks.append(("name", m.group("name").lower()))
        else:
            toks.append(("op", m.group("op")))
    return toks


class Parser:
    """Fortran expression grammar (F2008 7.1.3 precedence): ** > * / > unary +- > binary +- > relational."""

    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else (None, None)

    def eat(self, kind=None, val=None):
        k, v = self.peek()
        if (kind and k != kind) or (val is not None and v != val):
            raise SyntaxError("expected %s %s got %s %s" % (kind, val, k, v))
        self.i += 1
        return v

    def expr(self):
        return self.rel()

    def rel(self):
        a = self.add()
        k, v = self.peek()
        if k == "op" and v in (".lt.", ".le.", ".gt.", ".ge.", ".eq.", ".ne.", "<", "<=", ">", ">=", "==", "/="):
            self.i += 1
            return ("rel", v, a, self.add())
        return a

    def add(self):
        k, v = self.peek()
        if k == "op" and v in "+-":
            self.i += 1
            a = ("neg", self.mul()) if v == "-" else self.mul()
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
        k, v = self.peek()
        if k == "op" and v == "**":
            self.i += 1
            return ("pow", a, self.power())      # right associative
        return a

    def primary(self):
        k, v = self.peek()
        if k in ("int", "real"):
            self.i += 1
            return (k, v)
        if k == "op" and v == "(":
            self.i += 1
            e = self.expr()
            self.eat("op", ")")
            return ("paren", e)
        if k == "name":
            self.i += 1
            k2, v2 = self.peek()
            if k2 == "op" and v2 == "(":
                self.i += 1
                args = []
                if self.peek() != ("op", ")"):
                    args.append(self.expr())
                    while self.peek() == ("op", ","):
                        self.i += 1
                        args.append(self.expr())
                self.eat("op", ")")
                return ("call", v, args)
            if k2 == "op" and v2 == "%":
                self.i += 1
                comp = self.eat("name")
                return ("comp", v, comp)
            return ("var", v)
        raise SyntaxError("unexpected token %s %s" % (k, v))


def parse_expr(s):
    p = Parser(tokenize(s))
    e = p.expr()
    if p.i != len(p.t):
        raise SyntaxError("trailing tokens in %r" % s)
    return e


def powi(x, m):
    """x**m for integer m by repeated squaring (what ifort/gfortran emit; libgcc __powidf2)."""
    n = abs(m)
    y = x if n % 2 else 1.0
    n >>= 1
    while n:
        x = x * x
        if n % 2:
            y = y * x
        n >>= 1
    return 1.0 / y if m < 0 else y


ABSENT = object()


class Routine:
    def __init__(self, name, args, file, line):
        self.name, self.args, self.file, self.line = name, args, file, line
        self.intent = {}          # arg -> 'in' / 'out'
        self.optional = set()
        self.inits = []           # (name, expr) for parameters and initialised locals
        self.body = []            # nested statements


class FluxLib:
    """All subroutines of flux_lib plus the default_values type."""

    def __init__(self, ref):
        self.ref = ref
        self.routines = {}
        self.defaults = {}
        self.files = []
        root = os.path.join(ref, "src", "flux_lib")
        for dp, _, fns in sorted(os.walk(root)):
            for fn in sorted(fns):
                if fn.endswith(".F90"):
                    self.load(os.path.join(dp, fn))

    def load(self, path):
        text = open(path, encoding="utf-8", errors="replace").read()
        self.files.append((os.path.relpath(path, self.ref), hashlib.sha256(text.encode()).hexdigest()))
        lines = logical_lines(text)
        i = 0
        in_type = False
        while i < len(lines):
            no, st = lines[i]
            low = st.lower()
            if re.match(r"type\s+default_values_type", low):
                in_type = True
            elif low.startswith("end type"):
                in_type = False
            elif in_type:
                m = re.match(r"real\s*\(\s*prec\s*\)\s*::\s*(\w+)\s*=\s*(.+)$", st, re.I)
                if m:
                    self.defaults[m.group(1).lower()] = self.eval(parse_expr(m.group(2)), {})
            m = re.match(r"subroutine\s+(\w+)\s*\((.*)\)\s*$", st, re.I)
            if m:
                r = Routine(m.group(1).lower(), [a.strip().lower() for a in m.group(2).split(",") if a.strip()],
                            os.path.relpath(path, self.ref), no)
                i = self.parse_routine(lines, i + 1, r)
                self.routines[r.name] = r
            i += 1

    def parse_routine(self, lines, i, r):
        stack = [r.body]
        while i < len(lines):
            no, st = lines[i]
            low = st.lower()
            if re.match(r"end\s*subroutine", low):
                return i
            m = re.match(r"real\s*\(\s*prec\s*\)\s*(.*?)::\s*(.+)$", st, re.I)
            if m:
                attrs, decl = m.group(1).lower(), m.group(2)
                for item in split_top(decl):
                    if "=" in item:
                        nm, ex = item.split("=", 1)
                        r.inits.append((nm.strip().lower(), parse_expr(ex)))
                    else:
                        nm = item.strip().lower()
                        im = re.search(r"intent\s*\(\s*(\w+)\s*\)", attrs)
                        if im:
                            r.intent[nm] = im.group(1)
                        if "optional" in attrs:
                            r.optional.add(nm)
                i += 1
                continue
            m = re.match(r"if\s*\((.*)\)\s*then$", st, re.I)
            if m:
                node = ["if", parse_expr(m.group(1)), [], []]
                stack[-1].append(node)
                stack.append(node[2])
                i += 1
                continue
            if re.match(r"else$", low):
                stack.pop()
                stack.append(stack[-1][-1][3])
                i += 1
                continue
            if re.match(r"end\s*if$", low):
                stack.pop()
                i += 1
                continue
            m = re.match(r"call\s+(\w+)\s*\((.*)\)$", st, re.I)
            if m:
                stack[-1].append(["call", m.group(1).lower(), [a.strip().lower() for a in split_top(m.group(2))]])
                i += 1
                continue
            m = re.match(r"(\w+)\s*=\s*(.+)$", st)
            if m and not re.match(r"(use|implicit|public|private|module|contains)\b", low):
                stack[-1].append(["assign", m.group(1).lower(), parse_expr(m.group(2)), no])
                i += 1
                continue
            i += 1
        raise SyntaxError("unterminated subroutine " + r.name)

    # ---- evaluation -------------------------------------------------------------------
    def eval(self, e, env):
        k = e[0]
        if k == "real":
            return e[1]
        if k == "int":
            return e[1]
        if k == "paren":
            return self.eval(e[1], env)
        if k == "var":
            v = env[e[1]]
            if v is ABSENT:
                raise RuntimeError("reference to absent optional " + e[1])
            return v
        if k == "comp":
            assert e[1] == "default_values", e
            return self.defaults[e[2]]
        if k == "neg":
            return -self.eval(e[1], env)
        if k == "bin":
            a, b = self.eval(e[2], env), self.eval(e[3], env)
            a, b = float(a), float(b)
            if e[1] == "+":
                return a + b
            if e[1] == "-":
                return a - b
            if e[1] == "*":
                return a * b
            return a / b
        if k == "pow":
            base = float(self.eval(e[1], env))
            ex = e[2]
            if ex[0] == "int":
                return powi(base, ex[1])
            return math.pow(base, float(self.eval(ex, env)))
        if k == "rel":
            a, b = float(self.eval(e[2], env)), float(self.eval(e[3], env))
            return {".lt.": a < b, "<": a < b, ".le.": a <= b, "<=": a <= b, ".gt.": a > b, ">": a > b,
                    ".ge.": a >= b, ">=": a >= b, ".eq.": a == b, "==": a == b, ".ne.": a != b, "/=": a != b}[e[1]]
        if k == "call":
            fn = e[1]
            if fn == "present":
                return env[e[2][0][1]] is not ABSENT
            args = [float(self.eval(a, env)) for a in e[2]]
            if fn == "exp":
                return math.exp(args[0])
            if fn == "sqrt":
                return math.sqrt(args[0])
            if fn == "max":
                return max(args)       # no NaNs in the golden inputs
            if fn == "min":
                return min(args)
            raise RuntimeError("unknown intrinsic " + fn)
        raise RuntimeError("bad node %r" % (e,))

    def run_block(self, block, env):
        for st in block:
            if st[0] == "assign":
                env[st[1]] = float(self.eval(st[2], env))
            elif st[0] == "if":
                self.run_block(st[2] if self.eval(st[1], env) else st[3], env)
            elif st[0] == "call":
                callee = self.routines[st[1]]
                vals = [env[a] for a in st[2]]
                outs = self.call(st[1], vals)
                for a, formal in zip(st[2], callee.args):
                    if callee.intent.get(formal) == "out":
                        env[a] = outs[formal]

    def call(self, name, actual):
        """actual: list of floats (None / missing trailing == absent OPTIONAL). Returns {out_name: value}."""
        r = self.routines[name]
        env = {}
        for idx, formal in enumerate(r.args):
            v = actual[idx] if idx < len(actual) else None
            if v is None:
                if formal in r.optional:
                    v = ABSENT
                elif r.intent.get(formal) == "out":
                    v = float("nan")
                else:
                    raise RuntimeError("missing argument %s of %s" % (formal, name))
            env[formal] = v
        for nm, ex in r.inits:
            env[nm] = float(self.eval(ex, env))
        self.run_block(r.body, env)
        return {a: env[a] for a in r.args if r.intent.get(a) == "out"}


def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


# --------------------------------------------------------------------------------------------
# call-site wiring and calculator order, extracted from the reference text
# --------------------------------------------------------------------------------------------

def extract_wiring(ref):
    path = os.path.join(ref, "src", "flux_calculator_calculate.F90")
    lines = logical_lines(open(path, encoding="utf-8", errors="replace").read())
    wiring, sub, method = {}, None, None
    for no, st in lines:
        m = re.match(r"subroutine\s+(\w+)", st, re.I)
        if m:
            sub, method = m.group(1).lower(), None
            continue
        m = re.search(r"trim\(method\)\s*==\s*'(\w+)'", st, re.I)
        if m:
            method = m.group(1)
        m = re.match(r"call\s+(\w+)\s*\((.*)\)$", st, re.I)
        if m and sub and m.group(1).lower() not in ("set", "get", "call_function"):
            args = []
            for a in split_top(m.group(2)):
                a = a.strip()
                mm = re.match(r"local_field\((\w+),\s*(\w+)\)%var\(idx_(\w+)\)%field\(j\)", a, re.I)
                if mm:
                    args.append({"type": mm.group(1), "grid": mm.group(2), "var": mm.group(3).upper()})
                else:
                    args.append({"var": a.lower()})
            wiring.setdefault(sub, {})[method or "*"] = {"routine": m.group(1).lower(), "args": args, "line": no}
    return wiring


def extract_step_sequence(ref):
    path = os.path.join(ref, "src", "flux_calculator.F90")
    seq, in_loop = [], False
    for no, st in logical_lines(open(path, encoding="utf-8", errors="replace").read()):
        if re.match(r"do\s+n_timestep", st, re.I):
            in_loop = True
        if in_loop:
            m = re.match(r"call\s+(calc_\w+|distribute_shortwave_radiation_flux)\s*\((.*)\)$", st, re.I)
            if m:
                args = [a.strip() for a in split_top(m.group(2))]
                grid = args[2] if m.group(1).lower() in ("calc_spec_vapor_surface", "calc_flux_momentum_east",
                                                         "calc_flux_momentum_north") else None
                seq.append({"calc": m.group(1).lower(), "grid": int(grid) if grid else None, "line": no})
        if in_loop and re.match(r"write\s*\(w_unit,\*\)\s*'Finished time loop", st, re.I):
            break
    return seq


# --------------------------------------------------------------------------------------------
# vectors
# --------------------------------------------------------------------------------------------

def hx(v):
    return float(v).hex()


def make_cells(rng, n):
    """Synthetic cells following SURVEY 8(d) distributions, plus hand-placed edge cases."""
    cells = []
    for k in range(n):
        tsur = rng.uniform(271.35, 303.15)
        c = dict(
            TSUR=tsur, TATM=tsur + rng.uniform(-5, 5), PSUR=rng.uniform(9.8e4, 1.04e5),
            QATM=rng.uniform(1e-3, 1.5e-2), UATM=max(-35, min(35, rng.gauss(0, 6))),
            VATM=max(-35, min(35, rng.gauss(0, 6))), AMOI=rng.uniform(8e-4, 2.5e-3),
            AMOM=rng.uniform(8e-4, 2.5e-3), CMOI=rng.uniform(8e-4, 2.5e-3), CHEA=rng.uniform(8e-4, 2.5e-3),
            CMOM=rng.uniform(8e-4, 2.5e-3), FICE=float(k % 2), RSDD=-rng.uniform(0, 900),
            ALBA=rng.uniform(0.05, 0.8), ALBE=rng.uniform(0.05, 0.8), FARE=rng.uniform(0, 1),
            CORR=rng.gauss(0, 1e-6),
        )
        c["PATM"] = c["PSUR"] - rng.uniform(100, 1500)
        cells.append(c)
    # Appendix D point
    cells[0].update(TSUR=283.15, PSUR=101325.0, TATM=281.15, PATM=100800.0, QATM=0.005, UATM=5.0, VATM=-3.0,
                    AMOI=1.2e-3, AMOM=1.3e-3, FICE=0.0)
    cells[1].update(cells[0]); cells[1]["FICE"] = 1.0
    cells[2].update(UATM=0.0, VATM=0.0)                       # calm: u_min clip, zero momentum
    cells[3].update(UATM=0.003, VATM=-0.004)                  # vel = 0.005 < u_min
    cells[4].update(UATM=11.0, VATM=0.0)                      # RCO drag threshold exactly
    cells[5].update(UATM=math.nextafter(11.0, 0.0), VATM=0.0)
    cells[6].update(UATM=12.0, VATM=5.0)                      # vel = 13
    cells[7]["TATM"] = cells[7]["TSUR"]                       # RCO stable branch with zero gradient
    cells[8].update(TSUR=281.15, TATM=283.15)                 # stable
    cells[9]["FICE"] = 0.35                                   # fractional ice (formula accepts it)
    cells[10]["CORR"] = 0.0
    cells[11].update(TSUR=243.15, FICE=1.0)                   # cold ice
    return cells


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "flux_lib_golden.json"))
    ap.add_argument("--cells", type=int, default=96)
    a = ap.parse_args()

    lib = FluxLib(a.ref)
    wiring = extract_wiring(a.ref)
    sequence = extract_step_sequence(a.ref)
    rng = random.Random(0x5EEDF1C5)
    cells = make_cells(rng, a.cells)

    out = {
        "generator": "tests/golden/make_golden.py (source-level interpreter of the reference Fortran)",
        "reference_files": lib.files,
        "default_values": {k: hx(v) for k, v in lib.defaults.items()},
        "routine_signatures": {n: {"args": r.args, "out": [x for x in r.args if r.intent.get(x) == "out"],
                                   "optional": sorted(r.optional, key=r.args.index), "file": r.file, "line": r.line}
                               for n, r in lib.routines.items()},
        "wiring": wiring,
        "step_sequence": sequence,
        "cells": [{k: hx(v) for k, v in c.items()} for c in cells],
    }

    # ---- level 0: every routine on its own dummy-argument order, defaults and overridden optionals
    lvl0 = {}
    for name, r in sorted(lib.routines.items()):
        ins = [x for x in r.args if r.intent.get(x) == "in" and x not in r.optional]
        opts = [x for x in r.args if x in r.optional]
        cases = []
        for k, c in enumerate(cells):
            # natural binding: each dummy gets the physically matching field
            nat = {"fraction_ice": c["FICE"], "pressure_surface": c["PSUR"], "pressure_atmos": c["PATM"],
                   "temperature_surface": c["TSUR"], "temperature_atmos": c["TATM"],
                   "specific_vapor_content_atmos": c["QATM"], "specific_vapor_content_surface": c["QATM"] * 1.5,
                   "diffusion_coefficient_moisture": c["AMOI"], "diffusion_coefficient_momentum": c["AMOM"],
                   "u_atmos": c["UATM"], "v_atmos": c["VATM"], "flux_mass_evap": c["CORR"] * 20.0 + 2e-5,
                   "flux_radiation_averaged": c["RSDD"], "albedo_averaged": c["ALBA"], "albedo_surface_type": c["ALBE"]}
            actual = []
            for formal in r.args:
                if r.intent.get(formal) == "out":
                    actual.append(None)
                elif formal in r.optional:
                    actual.append(None)
                else:
                    actual.append(nat[formal])
            res = lib.call(name, actual)
            case = {"in": [hx(nat[x]) for x in ins], "opt": None, "out": [hx(res[x]) for x in r.args if x in res]}
            cases.append(case)
            if opts and k % 4 == 0:    # overridden OPTIONAL constants (perturbed defaults)
                ov = {"heat_capacity_air_new": 1004.64, "u_min_evap_new": 0.02, "gas_constant_air_new": 287.058,
                      "gas_constant_vapor_new": 461.495, "latent_heat_sublimation_new": 2.834e6,
                      "latent_heat_vaporization_new": 2.5008e6, "stefan_boltzmann_constant_new": 5.670374419e-8}
                actual2 = [ov[f] if f in r.optional else v for f, v in zip(r.args, actual)]
                res2 = lib.call(name, actual2)
                cases.append({"in": case["in"], "opt": [hx(ov[x]) for x in opts],
                              "out": [hx(res2[x]) for x in r.args if x in res2]})
        lvl0[name] = {"in_names": ins, "opt_names": opts, "cases": cases}
    out["level0"] = lvl0

    # ---- level 1: per-cell chains through the call-site wiring, for the three formula sets
    def wired(calc, method, env, grid):
        w = wiring[calc][method]
        r = lib.routines[w["routine"]]
        actual, outs = [], []
        for formal, arg in zip(r.args, w["args"]):
            if r.intent.get(formal) == "out":
                actual.append(None)
                outs.append((formal, arg["var"]))
            else:
                actual.append(env[arg["var"]])
        res = lib.call(w["routine"], actual)
        return {var: res[formal] for formal, var in outs if var != "dummy"}

    chains = {}
    for fset, hl in (("CCLM", "water"), ("MOM5", "water"), ("RCO", "water"), ("CCLM", "ice")):
        rows = []
        for c in cells:
            env = dict(c)
            o = {}
            # time-loop order (flux_calculator.F90:972-991); t, u and v grids use the same cell data here
            o["QSUR"] = wired("calc_spec_vapor_surface", "CCLM", env, 1)["QSUR"]
            env["QSUR"] = o["QSUR"]
            o["MEVA_nobias"] = wired("calc_flux_mass_evap", fset, env, 1)["MEVA"]
            o["MEVA"] = o["MEVA_nobias"] + c["CORR"]            # flux_calculator_calculate.F90:114
            env["MEVA"] = o["MEVA"]
            o["HLAT"] = wired("calc_flux_heat_latent", hl, env, 1)["HLAT"]
            env["MEVA"] = o["MEVA_nobias"]
            o["HLAT_nobias"] = wired("calc_flux_heat_latent", hl, env, 1)["HLAT"]
            o["HSEN"] = wired("calc_flux_heat_sensible", fset, env, 1)["HSEN"]
            o["UMOM"] = wired("calc_flux_momentum_east", fset, env, 2)["UMOM"]
            o["VMOM"] = wired("calc_flux_momentum_north", fset, env, 3)["VMOM"]
            o["RBBR"] = wired("calc_flux_radiation_blackbody", "StBo", env, 1)["RBBR"]
            o["RSDR"] = wired("distribute_shortwave_radiation_flux", "*", env, 1)["RSDR"]
            rows.append({k: hx(v) for k, v in o.items()})
        chains["%s/%s" % (fset, hl)] = rows
    out["chains"] = chains

    with open(a.out, "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
        f.write("\n")
    print("wrote", a.out, "routines:", len(lib.routines), "cells:", len(cells))
    for calc, ms in wiring.items():
        for mth, w in ms.items():
            print("  %-38s %-6s -> %-32s %s" % (calc, mth, w["routine"], [x["var"] for x in w["args"]]))
    print("  sequence:", [(s["calc"], s["grid"]) for s in sequence])


if __name__ == "__main__":
    sys.exit(main())
