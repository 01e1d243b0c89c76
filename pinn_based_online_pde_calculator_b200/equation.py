"""Symbolic PDE front end: validate, parse and compile the user's equation string.

The reference accepts an ``equation`` string (pinn_app/software.py:627) and then
ignores it; the only definition of the language is the validation regex at
pinn_app/callbacks/input_validation.py:29-46 and the tooltip at
pinn_app/layout.py:114-121.  This module

* ``validate_reference(expr)`` -- hand-written recogniser with the same
  accept/reject set as that regex (returns True when INVALID, like the callback);
* ``parse(expr, extended=...)`` -- recursive-descent parser to an AST, either in
  the strict reference language or with the documented extensions (functions,
  ``pi``, bare ``t``/``z``, scientific notation, unary minus, nested parentheses,
  ``^``);
* ``compile_equation(expr, d_in, ...)`` -- AST -> stack bytecode for the device
  residual VM (csrc/pinn_common.h ``PinnOp``) plus the jet-channel structure
  (n1, n2, mix) the fused kernel must carry.
"""
from __future__ import annotations

import math
import string
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

# opcodes -- keep in sync with csrc/pinn_common.h
OP_CONST, OP_COORD, OP_JET, OP_AUX = 0, 1, 2, 3
OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG = 4, 5, 6, 7, 8
OP_POWI, OP_POWF = 9, 10
OP_SIN, OP_COS, OP_EXP, OP_LOG, OP_TANH, OP_SQRT = 11, 12, 13, 14, 15, 16
OP_STORE_AUX = 17  # aux program only: pop -> aux column arg
MAX_OPS, MAX_CONSTS, VM_STACK = 192, 48, 12

_FUNC_OPS = {"sin": OP_SIN, "cos": OP_COS, "exp": OP_EXP, "log": OP_LOG, "tanh": OP_TANH, "sqrt": OP_SQRT}
_LETTERS = set(string.ascii_lowercase)

# jet structures with a kernel instantiation (csrc/jet_configs.txt)
SUPPORTED_JETS = {1: [(1, 1, 0)], 2: [(2, 1, 0), (2, 2, 0), (2, 2, 1)], 3: [(3, 2, 0)]}
# (n1, 0, 2): one combined second-order channel  L = sum_i beta_i u_ii  (Laplacian-type operators)
SUPPORTED_LAP = {2: (2, 0, 2), 3: (3, 0, 2)}


class EquationError(ValueError):
    pass


# ------------------------------------------------------------------ tokenizer
@dataclass
class Tok:
    kind: str  # num, name, op, lpar, rpar, comma
    text: str
    value: float = 0.0


def _tokenize(expr: str, extended: bool) -> List[Tok]:
    s = "".join(expr.split())
    toks: List[Tok] = []
    i, n = 0, len(s)
    while i < n:
        ch = s[i]
        if ch.isdigit() or ch == ".":
            j = i
            while j < n and s[j].isdigit():
                j += 1
            if j < n and s[j] == ".":
                j += 1
                while j < n and s[j].isdigit():
                    j += 1
            if s[i:j] == ".":
                raise EquationError("lone '.'")
            if extended and j < n and s[j] in "eE":
                k = j + 1
                if k < n and s[k] in "+-":
                    k += 1
                if k < n and s[k].isdigit():
                    while k < n and s[k].isdigit():
                        k += 1
                    j = k
            toks.append(Tok("num", s[i:j], float(s[i:j])))
            i = j
        elif ch in _LETTERS or ch == "_":
            j = i
            while j < n and (s[j] in _LETTERS or s[j] == "_" or (extended and s[j].isdigit() and j > i)):
                j += 1
            toks.append(Tok("name", s[i:j]))
            i = j
        elif ch == "*":
            if i + 1 < n and s[i + 1] == "*":
                toks.append(Tok("op", "**"))
                i += 2
            else:
                toks.append(Tok("op", "*"))
                i += 1
        elif ch in "+-/":
            toks.append(Tok("op", ch))
            i += 1
        elif ch == "^" and extended:
            toks.append(Tok("op", "**"))
            i += 1
        elif ch == "(":
            toks.append(Tok("lpar", ch))
            i += 1
        elif ch == ")":
            toks.append(Tok("rpar", ch))
            i += 1
        elif ch == "," and extended:
            toks.append(Tok("comma", ch))
            i += 1
        else:
            raise EquationError(f"illegal character {ch!r}")
    return toks


# ------------------------------------------------------------------ reference validator
def _is_ref_var(name: str) -> bool:
    if name in ("x", "y", "u", "r"):
        return True
    if name.startswith("u_") and 1 <= len(name) - 2 <= 2 and all(c in _LETTERS for c in name[2:]):
        return True
    return False


def _split_ref_names(text: str) -> Optional[List[str]]:
    """The regex has no separators between tokens other than operators, so a
    maximal run of letters/underscores must be exactly one variable."""
    return [text] if _is_ref_var(text) else None


def validate_reference(expr: Optional[str]) -> bool:
    """Same contract as ``on_equation_change`` (input_validation.py:19-50):
    returns True when the expression is INVALID, False when valid or empty."""
    if not expr:
        return False
    try:
        toks = _tokenize(expr, extended=False)
    except EquationError:
        return True
    if not toks:
        return True  # whitespace only: the regex fails on the stripped empty string
    # grammar: atom (op atom)* ; atom = num | var | '(' token (op token)* ')'
    pos = 0

    def is_token(t: Tok) -> bool:
        return t.kind == "num" or (t.kind == "name" and _is_ref_var(t.text))

    def atom() -> bool:
        nonlocal pos
        if pos >= len(toks):
            return False
        t = toks[pos]
        if is_token(t):
            pos += 1
            return True
        if t.kind == "lpar":
            pos += 1
            if pos >= len(toks) or not is_token(toks[pos]):
                return False
            pos += 1
            while pos < len(toks) and toks[pos].kind == "op":
                pos += 1
                if pos >= len(toks) or not is_token(toks[pos]):
                    return False
                pos += 1
            if pos >= len(toks) or toks[pos].kind != "rpar":
                return False
            pos += 1
            return True
        return False

    if not atom():
        return True
    while pos < len(toks):
        if toks[pos].kind != "op":
            return True
        pos += 1
        if not atom():
            return True
    return False


# ------------------------------------------------------------------ AST
@dataclass
class Node:
    kind: str  # num, coord, u, du, aux, add, sub, mul, div, neg, pow, call
    value: float = 0.0
    name: str = ""
    args: Tuple["Node", ...] = ()


class _Parser:
    """Precedence climbing: + - < * / < unary - < ** (right assoc), as Python."""

    def __init__(self, toks: List[Tok], extended: bool):
        self.t, self.i, self.ext = toks, 0, extended

    def peek(self) -> Optional[Tok]:
        return self.t[self.i] if self.i < len(self.t) else None

    def take(self) -> Tok:
        tok = self.t[self.i]
        self.i += 1
        return tok

    def parse(self) -> Node:
        node = self.expr()
        if self.peek() is not None:
            raise EquationError(f"unexpected {self.peek().text!r}")
        return node

    def expr(self) -> Node:
        node = self.term()
        while (p := self.peek()) is not None and p.kind == "op" and p.text in "+-":
            op = self.take().text
            rhs = self.term()
            node = Node("add" if op == "+" else "sub", args=(node, rhs))
        return node

    def term(self) -> Node:
        node = self.unary()
        while (p := self.peek()) is not None and p.kind == "op" and p.text in ("*", "/"):
            op = self.take().text
            rhs = self.unary()
            node = Node("mul" if op == "*" else "div", args=(node, rhs))
        return node

    def unary(self) -> Node:
        p = self.peek()
        if p is not None and p.kind == "op" and p.text == "-":
            if not self.ext:
                raise EquationError("unary minus is not in the reference grammar")
            self.take()
            return Node("neg", args=(self.unary(),))
        if p is not None and p.kind == "op" and p.text == "+" and self.ext:
            self.take()
            return self.unary()
        return self.power()

    def power(self) -> Node:
        base = self.atom()
        p = self.peek()
        if p is not None and p.kind == "op" and p.text == "**":
            self.take()
            expo = self.unary() if self.ext else self.power()
            return Node("pow", args=(base, expo))
        return base

    def atom(self) -> Node:
        p = self.peek()
        if p is None:
            raise EquationError("unexpected end of expression")
        if p.kind == "num":
            self.take()
            return Node("num", value=p.value)
        if p.kind == "lpar":
            self.take()
            node = self.expr()
            q = self.peek()
            if q is None or q.kind != "rpar":
                raise EquationError("missing ')'")
            self.take()
            return node
        if p.kind == "name":
            self.take()
            name = p.text
            q = self.peek()
            if self.ext and name in _FUNC_OPS:
                if q is None or q.kind != "lpar":
                    raise EquationError(f"{name} needs an argument")
                self.take()
                arg = self.expr()
                q = self.peek()
                if q is None or q.kind != "rpar":
                    raise EquationError("missing ')'")
                self.take()
                return Node("call", name=name, args=(arg,))
            if self.ext and name == "pi":
                return Node("num", value=math.pi)
            if name == "u":
                return Node("u")
            if name.startswith("u_") and 1 <= len(name) - 2 <= 2:
                return Node("du", name=name[2:])
            if self.ext and name.startswith("aux") and name[3:].isdigit():
                return Node("aux", value=float(int(name[3:])))
            if name in ("x", "y", "r") or (self.ext and name in ("t", "z")):
                return Node("coord", name=name)
            raise EquationError(f"unknown name {name!r}")
        raise EquationError(f"unexpected {p.text!r}")


def parse(expr: str, extended: bool = True) -> Node:
    if not extended and validate_reference(expr):
        raise EquationError("expression is not in the reference grammar")
    toks = _tokenize(expr, extended)
    if not toks:
        raise EquationError("empty expression")
    return _Parser(toks, extended).parse()


# ------------------------------------------------------------------ compile
def coord_index(name: str, d_in: int) -> int:
    """Input column of a coordinate / derivative letter.  Column 0: x or r;
    column 1: y, or t when there are two inputs; column 2: t or z."""
    if name in ("x", "r"):
        return 0
    if name == "y":
        if d_in < 2:
            raise EquationError("'y' needs at least two inputs")
        return 1
    if name == "t":
        if d_in == 1:
            raise EquationError("'t' needs at least two inputs")
        return 1 if d_in == 2 else 2
    if name == "z":
        if d_in < 3:
            raise EquationError("'z' needs three inputs")
        return 2
    raise EquationError(f"no input column for {name!r}")


@dataclass
class CompiledEquation:
    expr: str
    d_in: int
    n1: int
    n2: int
    mix: int
    ops: List[int] = field(default_factory=list)
    consts: List[float] = field(default_factory=list)
    n_aux: int = 0            # total aux columns the residual program reads (user + hoisted)
    max_stack: int = 0
    n_aux_user: int = 0       # columns supplied by the caller (aux0..)
    aux_ops: List[int] = field(default_factory=list)  # program filling the hoisted columns from coords/user aux
    lap_beta: List[float] = field(default_factory=lambda: [0.0, 0.0, 0.0])  # mix == 2: constant coefficients ...
    lap_aux: List[int] = field(default_factory=lambda: [-1, -1, -1])        # ... or aux columns of beta_i

    @property
    def K(self) -> int:
        return 1 + self.n1 + self.n2 + (1 if self.mix else 0)

    def channel_names(self, names: Sequence[str] = ("x", "y", "t")) -> List[str]:
        out = ["u"] + [f"u_{names[i]}" for i in range(self.n1)] + [f"u_{names[i]}{names[i]}" for i in range(self.n2)]
        if self.mix == 1:
            out.append(f"u_{names[0]}{names[1]}")
        elif self.mix == 2:
            out.append("L[u]")
        return out


def _collect(node: Node, d_in: int, firsts: set, seconds: set, aux: set):
    if node.kind == "du":
        idx = tuple(sorted(coord_index(c, d_in) for c in node.name))
        if len(idx) == 1:
            firsts.add(idx[0])
        else:
            seconds.add(idx)
    elif node.kind == "aux":
        aux.add(int(node.value))
    for a in node.args:
        _collect(a, d_in, firsts, seconds, aux)


def choose_jets(d_in: int, firsts: set, seconds: set) -> Tuple[int, int, int]:
    need_n1 = max([i + 1 for i in firsts] + [max(p) + 1 for p in seconds] + [0])
    pure = [p[0] for p in seconds if p[0] == p[1]]
    mixed = [p for p in seconds if p[0] != p[1]]
    need_n2 = max([i + 1 for i in pure] + [0])
    for m in mixed:
        if m != (0, 1):
            raise EquationError("only the mixed derivative of inputs 0 and 1 is supported")
    need_mix = 1 if mixed else 0
    for n1, n2, mix in SUPPORTED_JETS[d_in]:
        if n1 >= need_n1 and n2 >= need_n2 and mix >= need_mix and (not need_mix or n2 >= 2):
            return n1, n2, mix
    raise EquationError(
        f"derivative set first={sorted(firsts)} second={sorted(seconds)} has no kernel instantiation for d_in={d_in}")


def _jet_free(n: Node) -> bool:
    return n.kind not in ("u", "du", "lap") and all(_jet_free(a) for a in n.args)


def _has_point_data(n: Node) -> bool:
    return n.kind in ("coord", "aux") or any(_has_point_data(a) for a in n.args)


def _add_terms(n: Node, sign: int, out: list):
    """Flatten a +/- chain into (sign, term) pairs."""
    if n.kind == "add":
        _add_terms(n.args[0], sign, out)
        _add_terms(n.args[1], sign, out)
    elif n.kind == "sub":
        _add_terms(n.args[0], sign, out)
        _add_terms(n.args[1], -sign, out)
    elif n.kind == "neg":
        _add_terms(n.args[0], -sign, out)
    else:
        out.append((sign, n))


def _chain(terms) -> Node:
    node = None
    for sg, t in terms:
        if node is None:
            node = t if sg > 0 else Node("neg", args=(t,))
        else:
            node = Node("add" if sg > 0 else "sub", args=(node, t))
    return node


def hoist_point_terms(ast: Node, n_user_aux: int):
    """Move every maximal sub-expression that does not depend on u or its derivatives (source
    terms, variable coefficients) out of the per-step residual program: it is evaluated once per
    point when the points are set and read back as an aux column.  Additive jet-free terms of one
    sum are merged into a single column."""
    hoisted: List[Node] = []

    def lift(n: Node) -> Node:
        idx = n_user_aux + len(hoisted)
        hoisted.append(n)
        return Node("aux", value=float(idx))

    def walk(n: Node) -> Node:
        if _jet_free(n):
            if _has_point_data(n) and n.kind not in ("coord", "aux"):
                return lift(n)
            return n
        if n.kind in ("add", "sub"):
            terms: list = []
            _add_terms(n, +1, terms)
            free = [(sg, t) for sg, t in terms if _jet_free(t)]
            dep = [(sg, walk(t)) for sg, t in terms if not _jet_free(t)]
            if free and any(_has_point_data(t) for _, t in free):
                dep.append((+1, lift(_chain(free))))
            else:
                dep.extend(free)
            return _chain(dep)
        return Node(n.kind, n.value, n.name, tuple(walk(a) for a in n.args))

    return walk(ast), hoisted


class _NotLinear(Exception):
    pass


def _has_second(n: Node) -> bool:
    return (n.kind == "du" and len(n.name) == 2) or any(_has_second(a) for a in n.args)


def _scale(coef: Optional[Node], by: Node, div: bool = False) -> Node:
    if coef is None:
        return Node("div", args=(Node("num", value=1.0), by)) if div else by
    return Node("div" if div else "mul", args=(coef, by) if div else (by, coef))


def linear_second_order(n: Node, d_in: int):
    """Decompose n = sum_i coef_i * u_ii + rest with jet-free coef_i and a rest free of second
    derivatives.  Returns ({dim: coef AST or None for 1}, rest AST or None) or raises _NotLinear."""
    if not _has_second(n):
        return {}, n
    k = n.kind
    if k == "du":
        idx = sorted(coord_index(c, d_in) for c in n.name)
        if idx[0] != idx[1]:
            raise _NotLinear
        return {idx[0]: None}, None
    if k in ("add", "sub"):
        da, ra = linear_second_order(n.args[0], d_in)
        db, rb = linear_second_order(n.args[1], d_in)
        if k == "sub":
            db = {i: Node("neg", args=(c if c is not None else Node("num", value=1.0),)) for i, c in db.items()}
            rb = None if rb is None else Node("neg", args=(rb,))
        out = dict(da)
        for i, c in db.items():
            if i in out:
                a = out[i] if out[i] is not None else Node("num", value=1.0)
                b = c if c is not None else Node("num", value=1.0)
                out[i] = Node("add", args=(a, b))
            else:
                out[i] = c
        rest = ra if rb is None else (rb if ra is None else Node("add", args=(ra, rb)))
        return out, rest
    if k == "neg":
        d, r = linear_second_order(n.args[0], d_in)
        return ({i: Node("neg", args=(c if c is not None else Node("num", value=1.0),)) for i, c in d.items()},
                None if r is None else Node("neg", args=(r,)))
    if k == "mul":
        a, b = n.args
        if _has_second(a) and _has_second(b):
            raise _NotLinear
        lin, other = (a, b) if _has_second(a) else (b, a)
        if not _jet_free(other):
            raise _NotLinear  # quasi-linear coefficients depend on the network output
        d, r = linear_second_order(lin, d_in)
        return ({i: _scale(c, other) for i, c in d.items()}, None if r is None else Node("mul", args=(other, r)))
    if k == "div":
        a, b = n.args
        if _has_second(b) or not _jet_free(b):
            raise _NotLinear
        d, r = linear_second_order(a, d_in)
        return ({i: _scale(c, b, div=True) for i, c in d.items()}, None if r is None else Node("div", args=(r, b)))
    raise _NotLinear


def compile_equation(expr: str, d_in: int = 2, extended: bool = True, hoist: bool = True,
                     combine_second: bool = True) -> CompiledEquation:
    if not isinstance(expr, str):
        raise EquationError(f"equation must be a string, got {type(expr).__name__}")
    try:
        ast = parse(expr, extended)
    except (OverflowError, ZeroDivisionError, TypeError, ValueError) as e:  # constant folding left the reals
        raise EquationError(f"constant sub-expression cannot be evaluated: {e}") from e
    firsts, seconds, aux = set(), set(), set()
    _collect(ast, d_in, firsts, seconds, aux)
    n_user = max(aux) + 1 if aux else 0
    # Laplacian-type operators: second derivatives enter only through L = sum_i beta_i(x) u_ii with
    # jet-free beta_i  ->  propagate ONE combined second-order channel (K = 1 + n1 + 1)
    lap_coefs = None
    if combine_second and hoist and d_in in SUPPORTED_LAP and len({p for p in seconds if p[0] == p[1]}) >= 2 \
            and all(p[0] == p[1] for p in seconds):
        try:
            coefs, rest = linear_second_order(ast, d_in)
            if len(coefs) >= 2:
                lap_coefs = coefs
                lap_node = Node("lap")
                ast = lap_node if rest is None else Node("add", args=(rest, lap_node))
        except _NotLinear:
            lap_coefs = None
    if lap_coefs is not None:
        n1, n2, mix = SUPPORTED_LAP[d_in]
    else:
        n1, n2, mix = choose_jets(d_in, firsts, seconds)
    hoisted: List[Node] = []
    if hoist:
        ast, hoisted = hoist_point_terms(ast, n_user)
    ce = CompiledEquation(expr, d_in, n1, n2, mix, n_aux=n_user + len(hoisted), n_aux_user=n_user)
    depth = 0
    target = ce.ops

    def const_index(v: float) -> int:
        if isinstance(v, complex) or not math.isfinite(float(v)):
            raise EquationError(f"constant sub-expression folds to {v!r}, not a finite real number")
        v = float(v)
        for i, c in enumerate(ce.consts):
            if c == v:
                return i
        ce.consts.append(v)
        if len(ce.consts) > MAX_CONSTS:
            raise EquationError("too many distinct constants")
        return len(ce.consts) - 1

    def emit(op: int, arg: int = 0, delta: int = 0):
        nonlocal depth
        target.append((op & 0xFF) | (int(arg) << 8))
        depth += delta
        ce.max_stack = max(ce.max_stack, depth)

    def const_value(n: Node) -> Optional[float]:
        """Fold constant sub-expressions (exponents like 2, (1/2), -1)."""
        if n.kind == "num":
            return n.value
        if n.kind == "neg":
            v = const_value(n.args[0])
            return None if v is None else -v
        if n.kind in ("add", "sub", "mul", "div", "pow"):
            a, b = const_value(n.args[0]), const_value(n.args[1])
            if a is None or b is None:
                return None
            try:
                return {"add": a + b, "sub": a - b, "mul": a * b, "div": a / b, "pow": a ** b}[n.kind]
            except (ZeroDivisionError, OverflowError, ValueError):
                return None
        return None

    def gen(n: Node):
        cv = const_value(n)
        if cv is not None:
            emit(OP_CONST, const_index(cv), +1)
            return
        k = n.kind
        if k == "coord":
            emit(OP_COORD, coord_index(n.name, d_in), +1)
        elif k == "u":
            emit(OP_JET, 0, +1)
        elif k == "lap":
            emit(OP_JET, 1 + n1, +1)
        elif k == "aux":
            emit(OP_AUX, int(n.value), +1)
        elif k == "du":
            idx = tuple(sorted(coord_index(c, d_in) for c in n.name))
            if len(idx) == 1:
                ch = 1 + idx[0]
            elif idx[0] == idx[1]:
                ch = 1 + n1 + idx[0]
            else:
                ch = 1 + n1 + n2
            emit(OP_JET, ch, +1)
        elif k in ("add", "sub", "mul", "div"):
            gen(n.args[0])
            gen(n.args[1])
            emit({"add": OP_ADD, "sub": OP_SUB, "mul": OP_MUL, "div": OP_DIV}[k], 0, -1)
        elif k == "neg":
            gen(n.args[0])
            emit(OP_NEG)
        elif k == "pow":
            e = const_value(n.args[1])
            if e is None:
                raise EquationError("exponent must be a constant")
            if float(e).is_integer() and 0 <= e <= 64:
                gen(n.args[0])
                emit(OP_POWI, int(e))
            elif float(e).is_integer() and -64 <= e < 0:
                emit(OP_CONST, const_index(1.0), +1)
                gen(n.args[0])
                emit(OP_POWI, int(-e))
                emit(OP_DIV, 0, -1)
            else:
                gen(n.args[0])
                emit(OP_POWF, const_index(e))
        elif k == "call":
            gen(n.args[0])
            emit(_FUNC_OPS[n.name])
        else:
            raise EquationError(f"cannot compile node {k}")

    gen(ast)
    # coefficients of the combined second-order channel: constants or extra hoisted columns
    if lap_coefs is not None:
        for dim, coef in lap_coefs.items():
            cv = 1.0 if coef is None else const_value(coef)
            if cv is not None:
                ce.lap_beta[dim] = float(cv)
            else:
                ce.lap_aux[dim] = n_user + len(hoisted)
                hoisted.append(coef)
        ce.n_aux = n_user + len(hoisted)
    target = ce.aux_ops
    for i, sub in enumerate(hoisted):
        depth = 0
        gen(sub)
        emit(OP_STORE_AUX, n_user + i, -1)
    if len(ce.ops) > MAX_OPS or len(ce.aux_ops) > MAX_OPS:
        raise EquationError(f"expression too long ({max(len(ce.ops), len(ce.aux_ops))} ops > {MAX_OPS})")
    if ce.max_stack > VM_STACK:
        raise EquationError(f"expression too deep (stack {ce.max_stack} > {VM_STACK})")
    return ce


# the hard-coded residual of the reference (software.py:296), in its own language
REFERENCE_POLAR_LAPLACE = "u_rr + 1/r*u_r + 1/(r**2)*u_tt"


def evaluate_host(ce: CompiledEquation, z, jets, aux=None):
    """Run the bytecode on the host with numpy (float64) -- used by tests to check
    the compiler, never by the training path."""
    import numpy as np

    if ce.aux_ops:
        full = np.zeros((z.shape[0], ce.n_aux))
        if ce.n_aux_user:
            full[:, :ce.n_aux_user] = aux[:, :ce.n_aux_user]
        aux = full
    st = []
    for w in list(ce.aux_ops) + list(ce.ops):
        op, arg = w & 0xFF, w >> 8
        if op == OP_CONST:
            st.append(np.full(z.shape[0], ce.consts[arg]))
        elif op == OP_COORD:
            st.append(z[:, arg].astype(np.float64))
        elif op == OP_JET:
            st.append(jets[:, arg].astype(np.float64))
        elif op == OP_AUX:
            st.append(aux[:, arg].astype(np.float64))
        elif op in (OP_ADD, OP_SUB, OP_MUL, OP_DIV):
            b = st.pop()
            a = st.pop()
            st.append({OP_ADD: a + b, OP_SUB: a - b, OP_MUL: a * b, OP_DIV: a / b}[op])
        elif op == OP_NEG:
            st.append(-st.pop())
        elif op == OP_STORE_AUX:
            aux[:, arg] = st.pop()
        elif op == OP_POWI:
            st.append(st.pop() ** arg)
        elif op == OP_POWF:
            st.append(st.pop() ** ce.consts[arg])
        else:
            f = {OP_SIN: np.sin, OP_COS: np.cos, OP_EXP: np.exp, OP_LOG: np.log, OP_TANH: np.tanh, OP_SQRT: np.sqrt}[op]
            st.append(f(st.pop()))
    assert len(st) == 1
    return st[0]
