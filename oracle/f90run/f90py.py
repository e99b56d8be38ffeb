"""f90py -- a small Fortran-2008-subset -> Python translator (TEST INFRASTRUCTURE ONLY).

Purpose: the reference (AlexanderGSC/gmres) is 100 % Fortran and neither the build container nor the GPU box
has a Fortran compiler, so the reference cannot be *compiled* here.  This module lets it be *executed* anyway:
it reads the reference's own source files where they lie (/root/reference/src/**/*.f90, tests/*.f90),
translates them MECHANICALLY, statement by statement, into Python + numpy and runs the result.  No
algorithmic knowledge about GMRES / CG / BiCGSTAB lives here -- only the semantics of the Fortran
constructs the reference uses:

  modules / programs / contained subroutines, `use`, implicit none, private/public
  declarations: real(8), real, integer, logical, character(len=), procedure(iface), allocatable, intent,
                initialised module variables, explicit-shape and deferred-shape arrays
  allocate / deallocate, scalar / whole-array / array-section assignment (1-based, inclusive bounds,
  column-major), do / do concurrent / cycle / exit (loop variable defined as in Fortran after the loop),
  block if / else if / else, one-line if, call (scalars by reference, array sections as views,
  procedure arguments), stop, return, print / write / read(internal) in a simplified form
  intrinsics: size sqrt abs sign hypot norm2 dot_product matmul maxval minval sum real int dble max min mod
  OpenMP: directives are comments (the translation runs the reference with ONE thread, which is a legal
          execution of an OpenMP program); omp_lib's query functions are provided.

Arithmetic: IEEE binary64 through numpy scalars/arrays; default `real` (e.g. real(n), the literal 4.0) is
binary32 like gfortran's; integer division truncates; dot_product and the explicit reduction loops are
sequential sums.  What this does NOT reproduce is gfortran's FMA contraction under -O3 -march=native and
libgfortran's scaled norm2 -- differences at the 1e-16 relative level per operation, which is why the tests
that use the outputs compare to tolerance, not bit for bit (the reference itself is not bit-reproducible
across OpenMP thread counts).

Nothing in the product path (gmres_b200/, include/) imports this package; it is used by
tests/golden/make_reference_golden.py to generate committed golden vectors, and that script is the only
place that reads /root/reference.
"""
from __future__ import annotations

import re

# ----------------------------------------------------------------------------------------------------------
# 1. source -> logical statements
# ----------------------------------------------------------------------------------------------------------


def _strip_comment(line: str) -> str:
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


def _lower_outside_strings(s: str) -> str:
    out, q = [], None
    for ch in s:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        else:
            if ch in "'\"":
                q = ch
                out.append(ch)
            else:
                out.append(ch.lower())
    return "".join(out)


def _split_semicolons(s: str):
    parts, cur, q, depth = [], [], None, 0
    for ch in s:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur.append(ch)
        elif ch == "(":
            depth += 1
            cur.append(ch)
        elif ch == ")":
            depth -= 1
            cur.append(ch)
        elif ch == ";" and depth == 0:
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur))
    return [p.strip() for p in parts if p.strip()]


def logical_statements(text: str):
    """[(line_number, statement)] -- comments removed, continuations joined, ';' split, lower-cased."""
    stmts, buf, buf_line = [], "", 0
    for ln, raw in enumerate(text.splitlines(), 1):
        s = _strip_comment(raw).strip()
        if not s:
            continue
        if s.startswith("&"):
            s = s[1:].lstrip()
        if not buf:
            buf_line = ln
        if s.endswith("&"):
            buf += s[:-1].rstrip() + " "
            continue
        buf += s
        for part in _split_semicolons(buf):
            stmts.append((buf_line, _lower_outside_strings(part)))
        buf = ""
    if buf.strip():
        stmts.append((buf_line, _lower_outside_strings(buf)))
    return stmts


# ----------------------------------------------------------------------------------------------------------
# 2. expressions
# ----------------------------------------------------------------------------------------------------------
_DOT_OPS = {".and.": "and", ".or.": "or", ".not.": "not", ".true.": "True", ".false.": "False",
            ".eq.": "==", ".ne.": "/=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">=",
            ".eqv.": "eqv", ".neqv.": "neqv"}
_NUM = re.compile(r"(\d+\.\d*|\.\d+|\d+)([de][+-]?\d+)?(_\w+)?")
_NAME = re.compile(r"[a-z_]\w*")

INTRINSICS = {"size", "sqrt", "abs", "sign", "hypot", "norm2", "dot_product", "matmul", "maxval", "minval", "sum",
              "real", "int", "dble", "max", "min", "mod", "len", "trim", "adjustl", "command_argument_count",
              "omp_get_wtime", "omp_get_thread_num", "omp_get_num_threads", "omp_get_max_threads", "exp", "log",
              "sin", "cos", "nint", "floor", "transpose", "allocated", "epsilon", "huge", "tiny"}


class Tok:
    __slots__ = ("kind", "val")

    def __init__(self, kind, val):
        self.kind, self.val = kind, val

    def __repr__(self):
        return f"{self.kind}:{self.val}"


def tokenize(s: str):
    toks, i, n = [], 0, len(s)
    while i < n:
        ch = s[i]
        if ch.isspace():
            i += 1
            continue
        if ch in "'\"":
            j = i + 1
            while j < n and s[j] != ch:
                j += 1
            toks.append(Tok("str", s[i + 1:j]))
            i = j + 1
            continue
        if ch == ".":
            m = re.match(r"\.[a-z]+\.", s[i:])
            if m and m.group(0) in _DOT_OPS:
                toks.append(Tok("op", _DOT_OPS[m.group(0)]))
                i += len(m.group(0))
                continue
        if ch.isdigit() or (ch == "." and i + 1 < n and s[i + 1].isdigit()):
            m = _NUM.match(s, i)
            text = m.group(0)
            # "1.and." / "1.eq." : the dot belongs to the operator
            if "." in m.group(1) and m.group(1).endswith(".") and not m.group(2):
                m2 = re.match(r"\.[a-z]+\.", s[i + len(m.group(1)) - 1:])
                if m2 and m2.group(0) in _DOT_OPS:
                    text = m.group(1)[:-1]
            toks.append(Tok("num", text))
            i += len(text)
            continue
        m = _NAME.match(s, i)
        if m:
            toks.append(Tok("name", m.group(0)))
            i = m.end()
            continue
        for op in ("**", "//", "==", "/=", "<=", ">=", "=>", "(/", "/)"):
            if s.startswith(op, i):
                toks.append(Tok("op", op))
                i += 2
                break
        else:
            toks.append(Tok("op", ch))
            i += 1
    return toks


def _num_literal(text: str) -> str:
    t = text
    kind = None
    if "_" in t:
        t, kind = t.split("_", 1)
    if "d" in t:
        return f"_f8({t.replace('d', 'e')})"
    if "." in t or "e" in t:
        if kind in ("8", "dp", "real64"):
            return f"_f8({t})"
        return f"_f4({t})"            # default real = binary32, like gfortran
    return t                           # integer


class ExprParser:
    """Recursive descent over Fortran expression tokens; emits Python source."""

    def __init__(self, toks, scope):
        self.t, self.i, self.scope = toks, 0, scope

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else Tok("end", "")

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def accept(self, val):
        if self.peek().kind == "op" and self.peek().val == val:
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            raise SyntaxError(f"expected {val!r} at token {self.i}: {self.t}")

    # precedence climbing --------------------------------------------------------------------------------
    def expr(self):
        left = self.or_()
        while self.peek().kind == "op" and self.peek().val in ("eqv", "neqv"):
            op = self.next().val
            right = self.or_()
            left = f"(bool({left}) {'==' if op == 'eqv' else '!='} bool({right}))"
        return left

    def or_(self):
        left = self.and_()
        while self.accept("or"):
            left = f"({left} or {self.and_()})"
        return left

    def and_(self):
        left = self.not_()
        while self.accept("and"):
            left = f"({left} and {self.not_()})"
        return left

    def not_(self):
        if self.accept("not"):
            return f"(not {self.not_()})"
        return self.rel()

    def rel(self):
        left = self.concat()
        if self.peek().kind == "op" and self.peek().val in ("==", "/=", "<", "<=", ">", ">="):
            op = self.next().val
            right = self.concat()
            return f"({left} {'!=' if op == '/=' else op} {right})"
        return left

    def concat(self):
        left = self.add()
        while self.accept("//"):
            left = f"(str({left}) + str({self.add()}))"
        return left

    def add(self):
        if self.peek().kind == "op" and self.peek().val in ("+", "-"):
            op = self.next().val
            left = f"({op}{self.mul()})"
        else:
            left = self.mul()
        while self.peek().kind == "op" and self.peek().val in ("+", "-"):
            op = self.next().val
            left = f"({left} {op} {self.mul()})"
        return left

    def mul(self):
        left = self.pow_()
        while self.peek().kind == "op" and self.peek().val in ("*", "/"):
            op = self.next().val
            right = self.pow_()
            left = f"({left} * {right})" if op == "*" else f"_div({left}, {right})"
        return left

    def pow_(self):
        base = self.primary()
        if self.accept("**"):
            if self.peek().kind == "op" and self.peek().val in ("+", "-"):
                sgn = self.next().val
                ex = f"({sgn}{self.pow_()})"
            else:
                ex = self.pow_()          # right associative
            return f"_pow({base}, {ex})"
        return base

    def primary(self):
        tok = self.next()
        if tok.kind == "num":
            return _num_literal(tok.val)
        if tok.kind == "str":
            return repr(tok.val)
        if tok.kind == "op" and tok.val in ("True", "False"):
            return tok.val
        if tok.kind == "op" and tok.val == "(":
            e = self.expr()
            self.expect(")")
            return f"({e})"
        if tok.kind == "op" and tok.val in ("(/", "["):
            close = "/)" if tok.val == "(/" else "]"
            items = []
            while not self.accept(close):
                items.append(self.expr())
                self.accept(",")
            return f"_np.array([{', '.join(items)}])"
        if tok.kind == "name":
            name = tok.val
            if self.peek().kind == "op" and self.peek().val == "(":
                self.next()
                return self.name_with_args(name)
            return self.scope.pyname(name)
        raise SyntaxError(f"unexpected token {tok} in {self.t}")

    def subscripts(self):
        """after '(' : list of ('idx', e) / ('slice', lo, hi, step) until ')'."""
        subs = []
        if self.accept(")"):
            return subs
        while True:
            lo = hi = step = None
            is_slice = False
            if not (self.peek().kind == "op" and self.peek().val == ":"):
                lo = self.expr()
            if self.accept(":"):
                is_slice = True
                if not (self.peek().kind == "op" and self.peek().val in (",", ")", ":")):
                    hi = self.expr()
                if self.accept(":"):
                    step = self.expr()
            subs.append(("slice", lo, hi, step) if is_slice else ("idx", lo))
            if self.accept(")"):
                return subs
            self.expect(",")

    def name_with_args(self, name):
        sc = self.scope
        if sc.is_array(name):
            return sc.pyname(name) + "[" + index_code(self.subscripts()) + "]"
        if sc.is_character(name):          # substring: not used by the reference beyond whole variables
            subs = self.subscripts()
            (_, lo, hi, _), = subs
            return f"{sc.pyname(name)}[({lo or 1})-1:{hi or ''}]"
        subs = self.subscripts()
        args = []
        for s in subs:
            if s[0] != "idx":
                raise SyntaxError(f"slice in call to {name}")
            args.append(s[1])
        if name in INTRINSICS and not sc.is_procedure(name):
            return f"_i_{name}({', '.join(args)})"
        return f"{sc.pyname(name)}({', '.join(args)})"     # user function (none in the reference) / unknown


def _minus1(e: str) -> str:
    e = e.strip()
    if re.fullmatch(r"\d+", e):
        return str(int(e) - 1)
    return f"{e}-1" if re.fullmatch(r"\w+", e) else f"({e})-1"


def index_code(subs) -> str:
    parts = []
    for s in subs:
        if s[0] == "idx":
            parts.append(_minus1(s[1]))
        else:
            _, lo, hi, step = s
            if step is not None:
                parts.append(f"_fslice({lo or 'None'}, {hi or 'None'}, {step})")
            else:
                parts.append(f"{_minus1(lo) if lo else ''}:{hi if hi else ''}")
    return ", ".join(parts)


# ----------------------------------------------------------------------------------------------------------
# 3. program units, declarations
# ----------------------------------------------------------------------------------------------------------
class Var:
    def __init__(self, name, typ, rank=0, allocatable=False, intent=None, init=None, shape=None, iface=None,
                 parameter=False):
        self.name, self.typ, self.rank = name, typ, rank
        self.allocatable, self.intent, self.init = allocatable, intent, init
        self.shape, self.iface, self.parameter = shape, iface, parameter


class Sub:
    def __init__(self, name, args, line):
        self.name, self.args, self.line = name, args, line
        self.vars: dict[str, Var] = {}
        self.body = []          # [(line, stmt)]
        self.uses = []

    def out_positions(self):
        """dummy arguments handed back to the caller: scalars that may be modified, allocatable arrays."""
        out = []
        for k, a in enumerate(self.args):
            v = self.vars.get(a)
            if v is None or v.typ == "procedure":
                continue
            if v.rank == 0 and v.intent != "in":
                out.append(k)
            elif v.rank > 0 and v.allocatable:
                out.append(k)
        return out


class Unit:
    """module or program"""

    def __init__(self, kind, name):
        self.kind, self.name = kind, name
        self.uses, self.vars, self.subs, self.ifaces = [], {}, {}, {}
        self.body = []          # program executable statements
        self.public, self.default_private = set(), False


_DECL = re.compile(r"^(real\s*\(\s*(?:kind\s*=\s*)?8\s*\)|real\s*\(\s*(?:kind\s*=\s*)?4\s*\)|double\s+precision|real|integer|logical|"
                   r"character\s*(?:\([^)]*\))?|procedure\s*\(\s*\w+\s*\))(?![\w(])\s*(.*)$")


def _split_top(s: str, sep=","):
    parts, cur, depth, q = [], [], 0, None
    for ch in s:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur.append(ch)
        elif ch in "([":
            depth += 1
            cur.append(ch)
        elif ch in ")]":
            depth -= 1
            cur.append(ch)
        elif ch == sep and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    if "".join(cur).strip():
        parts.append("".join(cur).strip())
    return parts


def parse_decl(stmt: str):
    """-> list[Var] or None if `stmt` is not a type declaration."""
    m = _DECL.match(stmt)
    if not m:
        return None
    tword, rest = m.group(1), m.group(2)
    if tword.startswith("real") and re.search(r"\(\s*(kind\s*=\s*)?8", tword) or tword.startswith("double"):
        typ = "real8"
    elif tword.startswith("real"):
        typ = "real4"
    elif tword.startswith("integer"):
        typ = "integer"
    elif tword.startswith("logical"):
        typ = "logical"
    elif tword.startswith("character"):
        typ = "character"
    else:
        typ = "procedure"
    iface = re.search(r"\(\s*(\w+)\s*\)", tword).group(1) if typ == "procedure" else None
    attrs, names = "", rest
    if "::" in rest:
        attrs, names = rest.split("::", 1)
    elif rest.startswith(","):
        raise SyntaxError("attribute list without '::' : " + stmt)
    alist = [a.strip() for a in _split_top(attrs.strip().lstrip(","))] if attrs.strip() else []
    allocatable = "allocatable" in alist
    parameter = "parameter" in alist
    intent = None
    dim_attr = None
    for a in alist:
        mi = re.match(r"intent\s*\(\s*(\w+)\s*\)", a)
        if mi:
            intent = mi.group(1)
        md = re.match(r"dimension\s*\((.*)\)", a)
        if md:
            dim_attr = md.group(1)
    out = []
    for ent in _split_top(names):
        init = None
        if "=" in ent and "==" not in ent:
            ent, init = [x.strip() for x in ent.split("=", 1)]
        mm = re.match(r"^(\w+)\s*(?:\((.*)\))?$", ent.strip())
        if not mm:
            raise SyntaxError("cannot parse entity " + ent + " in " + stmt)
        name, dims = mm.group(1), mm.group(2) or dim_attr
        rank, shape = 0, None
        if dims is not None:
            d = _split_top(dims)
            rank = len(d)
            if not all(x.strip() == ":" for x in d):
                shape = d
        out.append(Var(name, typ, rank, allocatable, intent, init, shape, iface, parameter))
    return out


def parse_units(text: str):
    stmts = logical_statements(text)
    units, unit, sub, in_iface, iface_sub = [], None, None, False, None
    for ln, s in stmts:
        head = s.split("(")[0].split()
        w0 = head[0] if head else ""
        if in_iface:
            if re.match(r"^end\s*interface", s):
                in_iface = False
            elif w0 == "subroutine":
                m = re.match(r"subroutine\s+(\w+)\s*(?:\((.*)\))?", s)
                iface_sub = Sub(m.group(1), [a.strip() for a in _split_top(m.group(2) or "")], ln)
            elif re.match(r"^end\s*subroutine", s) or s == "end":
                unit.ifaces[iface_sub.name] = iface_sub
                iface_sub = None
            elif iface_sub is not None:
                d = parse_decl(s)
                if d:
                    for v in d:
                        iface_sub.vars[v.name] = v
            continue
        if sub is None and re.match(r"^(abstract\s+)?interface\b", s):
            in_iface = True
            continue
        if unit is None:
            m = re.match(r"^(module|program)\s+(\w+)$", s)
            if not m:
                raise SyntaxError(f"line {ln}: expected module/program, got {s!r}")
            unit = Unit(m.group(1), m.group(2))
            continue
        if sub is None:
            if re.match(r"^end\s*(module|program)(\s+\w+)?$", s) or s == "end":
                units.append(unit)
                unit = None
                continue
            if w0 == "use":
                unit.uses.append(re.match(r"use\s+(\w+)", s).group(1))
                continue
            if s.startswith("implicit") or s == "contains":
                continue
            if w0 == "private" and "::" not in s and len(s.split()) == 1:
                unit.default_private = True
                continue
            if w0 in ("public", "private"):
                names = s.split("::", 1)[1] if "::" in s else s.split(None, 1)[1]
                if w0 == "public":
                    unit.public.update(n.strip() for n in names.split(","))
                continue
            m = re.match(r"^subroutine\s+(\w+)\s*(?:\((.*)\))?$", s)
            if m:
                sub = Sub(m.group(1), [a.strip() for a in _split_top(m.group(2) or "")], ln)
                continue
            d = parse_decl(s)
            if d:
                for v in d:
                    unit.vars[v.name] = v
                continue
            if unit.kind == "program":
                unit.body.append((ln, s))
                continue
            raise SyntaxError(f"line {ln}: unexpected statement in module spec part: {s!r}")
        else:
            if re.match(r"^end\s*subroutine(\s+\w+)?$", s) or s == "end":
                unit.subs[sub.name] = sub
                sub = None
                continue
            if w0 == "use":
                sub.uses.append(re.match(r"use\s+(\w+)", s).group(1))
                continue
            if s.startswith("implicit"):
                continue
            d = parse_decl(s) if not sub.body else None
            if d:
                for v in d:
                    sub.vars[v.name] = v
                continue
            sub.body.append((ln, s))
    if unit is not None:
        units.append(unit)
    return units


# ----------------------------------------------------------------------------------------------------------
# 4. code generation
# ----------------------------------------------------------------------------------------------------------
class Scope:
    def __init__(self, world, unit, sub=None):
        self.world, self.unit, self.sub = world, unit, sub

    def lookup(self, name):
        if self.sub and name in self.sub.vars:
            return self.sub.vars[name]
        if name in self.unit.vars:
            return self.unit.vars[name]
        for u in self.world.visible_units(self.unit, self.sub):
            if name in u.vars:
                return u.vars[name]
        return None

    def is_array(self, name):
        v = self.lookup(name)
        return v is not None and v.rank > 0

    def is_character(self, name):
        v = self.lookup(name)
        return v is not None and v.typ == "character"

    def is_procedure(self, name):
        v = self.lookup(name)
        if v is not None and v.typ == "procedure":
            return True
        return self.world.find_sub(self.unit, self.sub, name) is not None

    def pyname(self, name):
        return name + "_" if name in _PY_RESERVED else name

    def signature_of(self, callee):
        """Sub describing the dummies of `callee` (a procedure dummy, a module procedure or None)."""
        v = self.lookup(callee)
        if v is not None and v.typ == "procedure":
            return self.world.find_iface(v.iface)
        return self.world.find_sub(self.unit, self.sub, callee)


_PY_RESERVED = {"lambda", "is", "in", "as", "def", "class", "from", "global", "pass", "del", "with", "yield", "try",
                "except", "raise", "assert", "import", "None", "True", "False", "print", "str", "int", "max", "min",
                "abs", "sum", "len", "iter", "id", "type", "all", "any", "not", "and", "or", "if", "else", "for",
                "while", "break", "continue", "return", "exec", "eval"}


def _zero(v: Var) -> str:
    return {"real8": "_f8(0.0)", "real4": "_f4(0.0)", "integer": "0", "logical": "False", "character": "''",
            "procedure": "None"}[v.typ]


def _dtype(v: Var) -> str:
    return {"real8": "_np.float64", "real4": "_np.float32", "integer": "_np.int64", "logical": "_np.bool_"}[v.typ]


class Emitter:
    def __init__(self, world, unit, sub):
        self.world, self.unit, self.sub = world, unit, sub
        self.scope = Scope(world, unit, sub)
        self.lines, self.ind = [], 1
        self.blocks = []        # stack of ('do', var, lo, hi, step) / ('if',)
        self.tmp = 0

    def emit(self, s):
        self.lines.append("    " * self.ind + s)

    def ex(self, text):
        p = ExprParser(tokenize(text), self.scope)
        e = p.expr()
        if p.peek().kind != "end":
            raise SyntaxError(f"trailing tokens in expression {text!r}: {p.t[p.i:]}")
        return e

    def designator(self, text):
        """Python assignment target for a Fortran variable / element / section."""
        text = text.strip()
        m = re.match(r"^(\w+)\s*\((.*)\)$", text)
        if m and self.scope.is_array(m.group(1)):
            p = ExprParser(tokenize("(" + m.group(2) + ")"), self.scope)
            p.next()
            return self.scope.pyname(m.group(1)) + "[" + index_code(p.subscripts()) + "]"
        if re.fullmatch(r"\w+", text):
            return self.scope.pyname(text)
        raise SyntaxError("not a designator: " + text)

    def is_designator(self, text):
        text = text.strip()
        if re.fullmatch(r"[a-z_]\w*", text):
            return self.scope.lookup(text) is not None
        m = re.match(r"^(\w+)\s*\((.*)\)$", text)
        return bool(m and self.scope.is_array(m.group(1)) and _balanced(m.group(2)))

    # ---- statements ------------------------------------------------------------------------------------
    def stmt(self, ln, s):
        try:
            self._stmt(ln, s)
        except Exception as e:
            raise SyntaxError(f"{self.unit.name}:{ln}: {s!r}: {e}") from e

    def ret_tuple(self):
        if self.sub is None:
            return ""
        outs = [self.scope.pyname(self.sub.args[k]) for k in self.sub.out_positions()]
        return "(" + "".join(o + ", " for o in outs) + ")"

    def _stmt(self, ln, s):
        w = re.match(r"[a-z_]\w*", s)
        w0 = w.group(0) if w else ""
        # --- block closers / openers
        if re.match(r"^end\s*do$", s):
            kind, var, hi, lo, st = self.blocks.pop()
            self.ind -= 1
            self.emit("else:")
            self.emit(f"    {var} = _after_loop({lo}, {hi}, {st})")
            return
        if re.match(r"^end\s*if$", s):
            self.blocks.pop()
            self.ind -= 1
            return
        m = re.match(r"^do\s+concurrent\s*\(\s*(\w+)\s*=\s*(.+?)\s*:\s*(.+?)\s*\)$", s)
        if m:
            return self.open_do(m.group(1), m.group(2), m.group(3), None)
        m = re.match(r"^do\s+(\w+)\s*=\s*(.+)$", s)
        if m:
            parts = _split_top(m.group(2))
            return self.open_do(m.group(1), parts[0], parts[1], parts[2] if len(parts) > 2 else None)
        if s == "do":
            raise SyntaxError("unbounded do not supported")
        m = re.match(r"^else\s*if\s*\((.*)\)\s*then$", s)
        if m:
            self.ind -= 1
            self.emit(f"elif {self.ex(m.group(1))}:")
            self.ind += 1
            self.emit("pass")
            return
        if s == "else":
            self.ind -= 1
            self.emit("else:")
            self.ind += 1
            self.emit("pass")
            return
        if w0 == "if":
            close = _match_paren(s, s.index("("))
            cond, rest = s[s.index("(") + 1:close], s[close + 1:].strip()
            if rest == "then":
                self.emit(f"if {self.ex(cond)}:")
                self.ind += 1
                self.emit("pass")
                self.blocks.append(("if", None, None, None, None))
            else:
                self.emit(f"if {self.ex(cond)}:")
                self.ind += 1
                self._stmt(ln, rest)
                self.ind -= 1
            return
        if s == "cycle":
            self.emit("continue")
            return
        if s == "exit":
            self.emit("break")
            return
        if s == "return":
            self.emit(f"return {self.ret_tuple()}")
            return
        if w0 == "stop":
            self.emit("raise _Stop()")
            return
        if w0 == "allocate":
            inner = s[s.index("(") + 1:_match_paren(s, s.index("("))]
            for ent in _split_top(inner):
                m = re.match(r"^(\w+)\s*\((.*)\)$", ent)
                v = self.scope.lookup(m.group(1))
                dims = ", ".join(self.ex(d) for d in _split_top(m.group(2)))
                self.emit(f"{self.scope.pyname(m.group(1))} = _alloc(({dims},), {_dtype(v)})")
            return
        if w0 == "deallocate":
            inner = s[s.index("(") + 1:_match_paren(s, s.index("("))]
            for ent in _split_top(inner):
                self.emit(f"{self.scope.pyname(ent.strip())} = None")
            return
        if w0 == "call":
            return self.call(s[4:].strip())
        if w0 == "print":
            items = _split_top(s[5:].strip())
            self.emit(f"_print({', '.join(self.ex(i) for i in items[1:])})")
            return
        if w0 == "write":
            close = _match_paren(s, s.index("("))
            ctl, items = _split_top(s[s.index("(") + 1:close]), _split_top(s[close + 1:].strip())
            fmt = ctl[1] if len(ctl) > 1 else "*"
            fmt = fmt.split("=", 1)[1].strip() if re.match(r"fmt\s*=", fmt) else fmt
            fmt_py = "None" if fmt == "*" else self.ex(fmt)
            self.emit(f"_write({fmt_py}, [{', '.join(self.ex(i) for i in items)}])")
            return
        if w0 == "read":
            close = _match_paren(s, s.index("("))
            ctl, items = _split_top(s[s.index("(") + 1:close]), _split_top(s[close + 1:].strip())
            for k, it in enumerate(items):
                v = self.scope.lookup(it.strip())
                self.emit(f"{self.designator(it)} = _read_item({self.ex(ctl[0])}, {k}, {v.typ!r})")
            return
        # --- assignment
        eq = _find_assign(s)
        if eq < 0:
            raise SyntaxError("unrecognised statement")
        lhs, rhs = s[:eq].strip(), s[eq + 1:].strip()
        r = self.ex(rhs)
        m = re.fullmatch(r"\w+", lhs)
        if m:
            v = self.scope.lookup(lhs)
            if v is None:
                raise SyntaxError(f"assignment to undeclared name {lhs} (implicit none)")
            py = self.scope.pyname(lhs)
            if v.rank > 0:
                if v.allocatable:
                    self.emit(f"{py} = _assign_whole({py}, {r}, {_dtype(v)})")
                else:
                    self.emit(f"{py}[...] = {r}")
            else:
                self.emit(f"{py} = _conv_{v.typ}({r})")
            return
        self.emit(f"{self.designator(lhs)} = {r}")

    def open_do(self, var, lo, hi, step):
        self.tmp += 1
        t = self.tmp
        self.emit(f"_lo{t}, _hi{t}, _st{t} = {self.ex(lo)}, {self.ex(hi)}, {self.ex(step) if step else 1}")
        self.emit(f"for {self.scope.pyname(var)} in _do_range(_lo{t}, _hi{t}, _st{t}):")
        self.ind += 1
        self.emit("pass")
        self.blocks.append(("do", self.scope.pyname(var), f"_hi{t}", f"_lo{t}", f"_st{t}"))

    def call(self, text):
        m = re.match(r"^(\w+)\s*(?:\((.*)\))?$", text)
        name, argtext = m.group(1), m.group(2) or ""
        actuals = _split_top(argtext)
        if name in _RUNTIME_SUBS and not self.scope.is_procedure(name):
            outs = _RUNTIME_SUBS[name]
            args = [self.ex(a) if k not in outs else "None" for k, a in enumerate(actuals)]
            callx = f"_s_{name}({', '.join(args)})"
            if outs:
                tg = [self.designator(actuals[k]) for k in outs if k < len(actuals)]
                self.emit(f"{', '.join(tg)}{',' if len(tg) == 1 else ''} = {callx}")
            else:
                self.emit(callx)
            return
        sig = self.scope.signature_of(name)
        if sig is None:
            raise SyntaxError(f"call to unknown procedure {name}")
        if len(actuals) != len(sig.args):
            raise SyntaxError(f"{name}: {len(actuals)} actual arguments for {len(sig.args)} dummies")
        args = []
        for a in actuals:
            a = a.strip()
            if re.fullmatch(r"[a-z_]\w*", a) and self.scope.lookup(a) is None and self.scope.is_procedure(a):
                args.append(self.scope.pyname(a))          # procedure passed as actual argument
            else:
                args.append(self.ex(a))
        outs = sig.out_positions()
        callx = f"{self.scope.pyname(name)}({', '.join(args)})"
        if not outs:
            self.emit(callx)
            return
        self.tmp += 1
        self.emit(f"_r{self.tmp} = {callx}")
        for j, k in enumerate(outs):
            if self.is_designator(actuals[k]):
                self.emit(f"{self.designator(actuals[k])} = _r{self.tmp}[{j}]")


# runtime subroutines with "out" positions
_RUNTIME_SUBS = {"get_command_argument": [1], "cpu_time": [0], "omp_set_num_threads": [], "omp_set_dynamic": [],
                 "system_clock": [0]}


def _balanced(s):
    d = 0
    for ch in s:
        d += ch == "("
        d -= ch == ")"
        if d < 0:
            return False
    return d == 0


def _match_paren(s, i):
    depth, q = 0, None
    for k in range(i, len(s)):
        ch = s[k]
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
            if depth == 0:
                return k
    raise SyntaxError("unbalanced parentheses in " + s)


def _find_assign(s):
    depth, q = 0, None
    for k, ch in enumerate(s):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        elif ch == "=" and depth == 0:
            if s[k + 1:k + 2] == "=" or s[k - 1:k] in ("=", "/", "<", ">"):
                continue
            return k
    return -1


class World:
    """all parsed units + their generated Python namespaces"""

    def __init__(self):
        self.units: dict[str, Unit] = {}
        self.ns: dict[str, dict] = {}
        self.source: dict[str, str] = {}

    def visible_units(self, unit, sub):
        names = list(unit.uses) + (list(sub.uses) if sub else [])
        return [self.units[n] for n in names if n in self.units]

    def find_sub(self, unit, sub, name):
        if name in unit.subs:
            return unit.subs[name]
        for u in self.visible_units(unit, sub):
            if name in u.subs and (not u.default_private or name in u.public):
                return u.subs[name]
        return None

    def find_iface(self, name):
        for u in self.units.values():
            if name in u.ifaces:
                return u.ifaces[name]
        raise KeyError("abstract interface " + name)

    # ---- code generation for one unit
    def generate(self, unit: Unit) -> str:
        out = [f"# generated from Fortran unit {unit.kind} {unit.name} by f90py -- do not edit"]
        for v in unit.vars.values():
            if v.rank == 0:
                sc = Scope(self, unit, None)
                init = ExprParser(tokenize(v.init), sc).expr() if v.init else _zero(v)
                out.append(f"{sc.pyname(v.name)} = _conv_{v.typ}({init})" if v.typ != "procedure" else f"{v.name} = None")
            else:
                out.append(f"{v.name} = None")
        for sub in unit.subs.values():
            out.extend(self.gen_sub(unit, sub))
        if unit.kind == "program":
            em = Emitter(self, unit, None)
            em.ind = 1
            out.append("def _main():")
            assigned = _assigned_names(unit.body)
            g = [n for n in sorted(assigned) if n in unit.vars]
            if g:
                out.append("    global " + ", ".join(Scope(self, unit).pyname(n) for n in g))
            out.append("    pass")
            for ln, s in unit.body:
                em.stmt(ln, s)
            out.extend(em.lines)
        return "\n".join(out) + "\n"

    def gen_sub(self, unit, sub):
        em = Emitter(self, unit, sub)
        sc = em.scope
        lines = [f"def {sc.pyname(sub.name)}({', '.join(sc.pyname(a) for a in sub.args)}):"]
        # names assigned but declared at unit level -> globals of the unit's namespace
        assigned = _assigned_names(sub.body)
        g = [n for n in sorted(assigned) if n not in sub.vars and n in unit.vars]
        if g:
            lines.append("    global " + ", ".join(sc.pyname(n) for n in g))
        for v in sub.vars.values():
            py = sc.pyname(v.name)
            if v.name in sub.args:
                if v.rank > 0 and v.allocatable and v.intent == "out":
                    lines.append(f"    {py} = None")
                continue
            if v.rank == 0:
                init = ExprParser(tokenize(v.init), sc).expr() if v.init else _zero(v)
                if v.typ != "procedure":
                    lines.append(f"    {py} = _conv_{v.typ}({init})")
            elif v.shape is not None and not v.allocatable:
                dims = ", ".join(ExprParser(tokenize(d), sc).expr() for d in v.shape)
                lines.append(f"    {py} = _alloc(({dims},), {_dtype(v)})")
            else:
                lines.append(f"    {py} = None")
        for ln, s in sub.body:
            em.stmt(ln, s)
        lines.extend(em.lines)
        lines.append(f"    return {em.ret_tuple()}")
        return lines

    # ---- loading
    def add_source(self, text: str, origin: str = "?"):
        for u in parse_units(text):
            self.units[u.name] = u
            self.source[u.name] = origin

    def build(self):
        from . import runtime
        done = set()

        def build_unit(name):
            if name in done or name not in self.units:
                return
            u = self.units[name]
            for dep in u.uses:
                build_unit(dep)
            for s in u.subs.values():
                for dep in s.uses:
                    build_unit(dep)
            ns = dict(runtime.NAMESPACE)
            for dep in list(u.uses) + [d for s in u.subs.values() for d in s.uses]:
                if dep in self.ns:
                    du = self.units[dep]
                    for k, val in self.ns[dep].items():
                        if k.startswith("_") or k in runtime.NAMESPACE:
                            continue
                        base = k[:-1] if k.endswith("_") and k[:-1] in _PY_RESERVED else k
                        if du.default_private and base not in du.public:
                            continue
                        ns[k] = val
            code = self.generate(u)
            ns["__source__"] = code
            exec(compile(code, f"<f90py:{name}>", "exec"), ns)
            self.ns[name] = ns
            done.add(name)

        for name in list(self.units):
            build_unit(name)
        return self

    def proc(self, module: str, name: str):
        return self.ns[module][name + "_" if name in _PY_RESERVED else name]

    def run_program(self, name: str, argv=()):
        from . import runtime
        runtime.set_argv(list(argv))
        runtime.clear_output()
        try:
            self.ns[name]["_main"]()
        except runtime._Stop:
            pass
        return runtime.get_output()


def _assigned_names(body):
    names = set()
    for _, s in body:
        t = s
        while True:
            m = re.match(r"^if\s*\(", t)
            if m:
                close = _match_paren(t, t.index("("))
                rest = t[close + 1:].strip()
                if rest and rest != "then":
                    t = rest
                    continue
            break
        k = _find_assign(t)
        if k > 0 and not re.match(r"^(do|if|call|allocate|write|print|read)\b", t):
            names.add(re.match(r"\w+", t.strip()).group(0))
        m = re.match(r"^do\s+(?:concurrent\s*\(\s*)?(\w+)\s*=", t)
        if m:
            names.add(m.group(1))
        if t.startswith("call "):
            for a in _split_top(t[t.index("(") + 1:_match_paren(t, t.index("("))]) if "(" in t else []:
                if re.fullmatch(r"[a-z_]\w*", a.strip()):
                    names.add(a.strip())
        if t.startswith("read"):
            close = _match_paren(t, t.index("("))
            for a in _split_top(t[close + 1:]):
                if re.fullmatch(r"[a-z_]\w*", a.strip()):
                    names.add(a.strip())
    return names


def load_reference(root: str, with_tests=False) -> World:
    """Parse and translate every module of the reference tree at `root` (and, optionally, its driver programs)."""
    import glob
    import os
    w = World()
    files = sorted(glob.glob(os.path.join(root, "src", "*.f90")) + glob.glob(os.path.join(root, "src", "*", "*.f90")))
    if with_tests:
        files += sorted(glob.glob(os.path.join(root, "tests", "*.f90")))
    for f in files:
        with open(f) as fh:
            w.add_source(fh.read(), os.path.relpath(f, root))
    return w.build()
