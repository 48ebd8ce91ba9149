"""Safe translator for the C++ expression strings the reference passes to ``fenics.Expression``
(e.g. ``test_case_simulation_tumor_growth_2D_subdomains.py:39,64``): a small recursive-descent parser
for C arithmetic / comparison / logical / ternary expressions, evaluated with numpy over many points.
No ``eval``: only the grammar below is accepted.
"""
import math
import re

import numpy as np

_TOKEN = re.compile(r"\s*(?:(\d+\.\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|\d+(?:[eE][-+]?\d+)?)"
                    r"|([A-Za-z_][A-Za-z_0-9]*)|(<=|>=|==|!=|&&|\|\||[-+*/()<>?:,\[\]!]))")

_FUNCS = {
    "sqrt": np.sqrt, "exp": np.exp, "log": np.log, "sin": np.sin, "cos": np.cos, "tan": np.tan,
    "fabs": np.abs, "abs": np.abs, "pow": np.power, "tanh": np.tanh, "sinh": np.sinh, "cosh": np.cosh,
    "atan": np.arctan, "atan2": np.arctan2, "asin": np.arcsin, "acos": np.arccos, "floor": np.floor,
    "ceil": np.ceil, "fmin": np.minimum, "fmax": np.maximum, "min": np.minimum, "max": np.maximum,
    "erf": np.vectorize(math.erf),
}
_CONSTS = {"pi": math.pi, "DOLFIN_PI": math.pi, "DOLFIN_EPS": 3.0e-16, "M_PI": math.pi}


class CExprError(ValueError):
    pass


def tokenize(src):
    pos, out = 0, []
    src = src.strip()
    while pos < len(src):
        m = _TOKEN.match(src, pos)
        if not m:
            raise CExprError("cannot tokenize %r at %d" % (src, pos))
        pos = m.end()
        if m.group(1) is not None:
            out.append(("num", float(m.group(1))))
        elif m.group(2) is not None:
            out.append(("id", m.group(2)))
        else:
            out.append(("op", m.group(3)))
    return out


class _Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else ("end", None)

    def take(self, kind=None, val=None):
        k, v = self.peek()
        if (kind and k != kind) or (val is not None and v != val):
            raise CExprError("expected %s %s, got %s %s" % (kind, val, k, v))
        self.i += 1
        return v

    def accept(self, val):
        if self.peek() == ("op", val):
            self.i += 1
            return True
        return False

    # grammar: ternary > or > and > equality > relational > additive > multiplicative > unary > primary
    def ternary(self):
        c = self.logical_or()
        if self.accept("?"):
            a = self.ternary()
            self.take("op", ":")
            b = self.ternary()
            return ("?", c, a, b)
        return c

    def _binary(self, sub, ops):
        n = sub()
        while self.peek()[0] == "op" and self.peek()[1] in ops:
            op = self.take()
            n = (op, n, sub())
        return n

    def logical_or(self):
        return self._binary(self.logical_and, ("||",))

    def logical_and(self):
        return self._binary(self.equality, ("&&",))

    def equality(self):
        return self._binary(self.relational, ("==", "!="))

    def relational(self):
        return self._binary(self.additive, ("<", ">", "<=", ">="))

    def additive(self):
        return self._binary(self.multiplicative, ("+", "-"))

    def multiplicative(self):
        return self._binary(self.unary, ("*", "/"))

    def unary(self):
        if self.accept("-"):
            return ("neg", self.unary())
        if self.accept("+"):
            return self.unary()
        if self.accept("!"):
            return ("not", self.unary())
        return self.primary()

    def primary(self):
        k, v = self.peek()
        if k == "num":
            self.i += 1
            return ("num", v)
        if k == "id":
            self.i += 1
            if self.accept("("):
                args = []
                if not self.accept(")"):
                    args.append(self.ternary())
                    while self.accept(","):
                        args.append(self.ternary())
                    self.take("op", ")")
                return ("call", v, args)
            if self.accept("["):
                idx = self.ternary()
                self.take("op", "]")
                return ("index", v, idx)
            return ("var", v)
        if self.accept("("):
            n = self.ternary()
            self.take("op", ")")
            return n
        raise CExprError("unexpected token %s %s" % (k, v))


def parse(src):
    p = _Parser(tokenize(src))
    tree = p.ternary()
    if p.peek()[0] != "end":
        raise CExprError("trailing input in %r" % src)
    return tree


def evaluate(tree, x, params):
    """x: (n_points, dim) array; params: user parameters (floats). Returns (n_points,) array."""
    n = x.shape[0]

    def ev(node):
        op = node[0]
        if op == "num":
            return node[1]
        if op == "var":
            name = node[1]
            if name in params:
                return float(params[name])
            if name in _CONSTS:
                return _CONSTS[name]
            raise CExprError("unknown identifier %r" % name)
        if op == "index":
            if node[1] != "x":
                raise CExprError("only x[i] may be indexed")
            i = ev(node[2])
            return x[:, int(i)]
        if op == "call":
            f = _FUNCS.get(node[1])
            if f is None:
                raise CExprError("unknown function %r" % node[1])
            return f(*[ev(a) for a in node[2]])
        if op == "neg":
            return -ev(node[1])
        if op == "not":
            return np.logical_not(ev(node[1]))
        if op == "?":
            return np.where(ev(node[1]), ev(node[2]), ev(node[3]))
        a, b = ev(node[1]), ev(node[2])
        if op == "+": return a + b
        if op == "-": return a - b
        if op == "*": return a * b
        if op == "/": return a / b
        if op == "<": return a < b
        if op == ">": return a > b
        if op == "<=": return a <= b
        if op == ">=": return a >= b
        if op == "==": return a == b
        if op == "!=": return a != b
        if op == "&&": return np.logical_and(a, b)
        if op == "||": return np.logical_or(a, b)
        raise CExprError("bad node %r" % (op,))

    return np.broadcast_to(np.asarray(ev(tree), dtype=np.float64), (n,)).copy()
