"""TEST INFRASTRUCTURE -- evaluates the REFERENCE'S OWN KERNEL SOURCE TEXT with numpy.

GeoPhyInv's stencil kernels are plain text: `@parallel function ... @all(x) = @d_xi(p) * dxI ... end` bodies
(src/fdtd/advance_acou.jl, advance_elastic.jl, medium.jl, source.jl, gradient.jl), finite-difference macros
(src/fdtd/diff2D.jl, diff3D.jl, one block per `_fd_order`), `@eval`-templated `@parallel_indices` kernels (cpml.jl,
dirichlet.jl, boundary.jl), host functions that call them in a fixed order (`update_dstress!`, `update_v!`, ...), and
array shapes given as `get_mgrid` methods (src/fields.jl).  Julia is not installed in the build image, so the reference
cannot run here -- but its text can be READ: this module parses those files (a small recursive-descent parser for the
Julia subset they use) and evaluates the expressions on numpy arrays with Julia's semantics:

  * `@parallel` without ranges: every statement `@all(A) = rhs` / `@inn(A) = rhs` runs for all indices admitted by the
    `@within` macro of the same file (ParallelStencil wraps each statement in `if @within("@all", A) ... end`);
  * `@parallel ranges kernel(...)` / `@parallel_indices`: explicit index ranges, explicit `A[i + off, j]` references;
  * operators left-associative with Julia's precedence, no FMA contraction (numpy never fuses), Float32 arrays and
    Float32 `Data.Number` scalars; a store into a Float32 array rounds once (Julia `setindex!` converts);
  * float literals: `literals="f32"` retypes them to the kernel number type (what ParallelStencil's `@parallel` does with
    `@init_parallel_stencil(Threads, Float32, N)`, src/GeoPhyInv.jl:95-100), `literals="f64"` leaves them Float64 (plain
    Julia promotion) -- both are generated so that the oracle can be pinned in both modes.

Nothing here is derived from oracle/ or from the CUDA engine: index algebra, association order, sweep order and shapes
all come from the text under /root/reference.  Only tests/golden/from_reference.py imports this module.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

F32 = np.float32

# ====================================================================================================
# expression parser (Julia subset)
# ====================================================================================================
_TOK = re.compile(
    r"\s*(?:(?P<flt>\d+\.\d*(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)|(?P<int>\d+)|(?P<id>@?[A-Za-z_][A-Za-z_0-9!]*)|(?P<op>\$\(|\$|&&|<=|[-+*/()\[\],=:]))")


def tokenize(src: str):
    out, pos = [], 0
    src = src.strip()
    while pos < len(src):
        m = _TOK.match(src, pos)
        if not m:
            raise SyntaxError(f"cannot tokenize {src[pos:pos + 40]!r}")
        pos = m.end()
        kind = m.lastgroup
        out.append((kind, m.group(kind)))
    return out


class Parser:
    """expr := term (('+'|'-') term)* ; term := unary (('*'|'/') unary)* ; unary := '-' unary | postfix ;
    postfix := primary ('[' args ']')* ; primary := number | '$'? name ['(' args ')'] | '@'name '(' args ')' | '(' expr ')' | '$(' expr ')'"""

    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else (None, None)

    def take(self, val=None):
        k, v = self.peek()
        if val is not None and v != val:
            raise SyntaxError(f"expected {val!r}, got {v!r}")
        self.i += 1
        return k, v

    def args(self, close):
        a = []
        if self.peek()[1] != close:
            a.append(self.expr())
            while self.peek()[1] == ",":
                self.take()
                if self.peek()[1] == close:
                    break
                a.append(self.expr())
        self.take(close)
        return a

    def expr(self):
        n = self.term()
        while self.peek()[1] in ("+", "-"):
            op = self.take()[1]
            n = ("bin", op, n, self.term())
        return n

    def term(self):
        n = self.unary()
        while self.peek()[1] in ("*", "/"):
            op = self.take()[1]
            n = ("bin", op, n, self.unary())
        return n

    def unary(self):
        if self.peek()[1] == "-":
            self.take()
            return ("neg", self.unary())
        return self.postfix()

    def postfix(self):
        n = self.primary()
        while self.peek()[1] == "[":
            self.take()
            n = ("ref", n, self.args("]"))
        return n

    def primary(self):
        k, v = self.take()
        if k == "flt":
            return ("flt", float(v))
        if k == "int":
            return ("int", int(v))
        if v == "(":
            n = self.expr()
            self.take(")")
            return n
        if v == "$(":
            n = self.expr()
            self.take(")")
            return n
        if v == "$":
            k, v = self.take()
            assert k == "id"
            return ("var", v)
        if k == "id":
            if v.startswith("@"):
                self.take("(")
                return ("mac", v[1:], self.args(")"))
            if self.peek()[1] == "(":
                self.take()
                return ("call", v, self.args(")"))
            return ("var", v)
        raise SyntaxError(f"unexpected token {v!r}")


def parse_expr(src: str):
    p = Parser(tokenize(src))
    n = p.expr()
    if p.i != len(p.t):
        raise SyntaxError(f"trailing tokens in {src!r}: {p.t[p.i:]}")
    return n


def parse_assign(src: str):
    """`lhs = rhs` -> (lhs_ast, rhs_ast)"""
    depth = 0
    for i, ch in enumerate(src):
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        elif ch == "=" and depth == 0 and src[i - 1] not in "<>=!" and src[i + 1:i + 2] != "=":
            return parse_expr(src[:i]), parse_expr(src[i + 1:])
    raise SyntaxError(f"no assignment in {src!r}")


# ====================================================================================================
# source-text helpers
# ====================================================================================================
def strip_comments(text: str) -> str:
    text = re.sub(r"#=.*?=#", "", text, flags=re.S)
    return "\n".join(l.split("#")[0].rstrip() for l in text.splitlines())


def balanced(text: str, start: int, open_ch="(", close_ch=")") -> int:
    """index just after the bracket that closes the one at text[start]"""
    assert text[start] == open_ch
    d = 0
    for i in range(start, len(text)):
        if text[i] == open_ch:
            d += 1
        elif text[i] == close_ch:
            d -= 1
            if d == 0:
                return i + 1
    raise SyntaxError("unbalanced")


def split_top(s: str, sep=","):
    out, d, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            d += 1
        elif ch in ")]}":
            d -= 1
        if ch == sep and d == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [x.strip() for x in out]


def logical_statements(body: str) -> List[str]:
    """join continuation lines: a statement continues while brackets are open or the line ends with an operator / '='"""
    out, cur = [], ""
    for line in body.splitlines():
        line = line.strip()
        if not line:
            continue
        cur = (cur + " " + line).strip() if cur else line
        depth = cur.count("(") + cur.count("[") - cur.count(")") - cur.count("]")
        if depth > 0 or cur[-1] in "=+-*/,&" or cur.endswith("&&"):
            continue
        out.append(cur)
        cur = ""
    if cur:
        out.append(cur)
    return out


# ====================================================================================================
# finite-difference macros (diff2D.jl / diff3D.jl)
# ====================================================================================================
@dataclass
class Macros:
    ndims: int
    order: int
    inner: Dict[str, int]                      # izi -> offset K (izi = iz + K)
    body: Dict[str, tuple]                     # macro name -> AST with the argument bound to var "A"
    inn_shrink: int                            # @within("@inn", A): i <= size(A, d) - inn_shrink
    src: Dict[str, str] = field(default_factory=dict)


def parse_macros(text: str, order: int, ndims: int) -> Macros:
    text = strip_comments(text)
    cur: Optional[int] = None                  # order of the @static branch we are in (None = common code)
    lines = text.splitlines()
    inner: Dict[str, int] = {}
    body: Dict[str, tuple] = {}
    src: Dict[str, str] = {}
    shrink = None
    i = 0
    while i < len(lines):
        ln = lines[i]
        m = re.match(r"^(?:@static if|elseif)\s*\(_fd_order == (\d+)\)", ln)
        if m:
            cur = int(m.group(1))
        elif ln.strip() == "end" and not ln.startswith(" "):
            cur = None
        m = re.match(r"^\s*((?:i\wi,\s*)+i\wi)\s*=\s*(.*)$", ln)
        if m and cur == order:
            names = [x.strip() for x in m.group(1).split(",")]
            vals = re.findall(r":\(\$(i\w) \+ (\d+)\)", m.group(2))
            assert len(vals) == len(names), ln
            for nm, (base, k) in zip(names, vals):
                assert nm == base + "i"
                inner[nm] = int(k)
        m = re.match(r"^(\s*)macro (\w+)\((.*?)\)", ln)
        if m and re.search(r"\)\s*end\s*$", ln):
            m = None                               # one-line dummies: `macro d_yi(args...) end`
        if m:
            indent, name, margs = m.group(1), m.group(2), m.group(3)
            j = i
            while not re.match(r"^" + indent + r"end\s*$", lines[j]):
                j += 1
            block = "\n".join(lines[i:j + 1])
            if "args..." not in margs and (cur is None or cur == order):
                if name == "within":
                    b_all = block[block.index('macroname == "@all"'):block.index('macroname == "@inn"')]
                    b_inn = block[block.index('macroname == "@inn"'):block.index("error(")]
                    ks = re.findall(r"size\(\$A, \d\) - (\d+)", b_inn)
                    assert len(ks) == ndims and len(set(ks)) == 1, block
                    shrink = int(ks[0])
                    assert len(re.findall(r"size\(\$A, \d\)(?! -)", b_all)) == ndims, block
                else:
                    k = block.index("esc(") + 3
                    inside = block[k + 1:balanced(block, k) - 1].strip()          # :( ... )
                    assert inside.startswith(":(")
                    expr = inside[1:]
                    expr = expr[1:balanced(expr, 0) - 1]
                    body[name] = parse_expr(" ".join(expr.split()))
                    src[name] = " ".join(expr.split())
            i = j
        i += 1
    assert shrink is not None and inner, "macro file not understood"
    return Macros(ndims, order, inner, body, shrink, src)


# ====================================================================================================
# kernels
# ====================================================================================================
@dataclass
class Kernel:
    name: str
    params: List[str]
    nd_annot: Optional[int]                    # N of `first_arg::Data.Array{N}` when present
    stmts: List[Tuple[tuple, tuple]]           # (lhs, rhs) ASTs
    indices: Optional[List[str]] = None        # @parallel_indices kernels
    text: str = ""


def parse_parallel_kernels(text: str) -> List[Kernel]:
    text = strip_comments(text)
    out = []
    for m in re.finditer(r"@parallel(_indices)?\s*(\([^)]*\))?\s*function\s+([\w!]+)\s*\(", text):
        is_idx, idx, name = m.group(1), m.group(2), m.group(3)
        a0 = m.end() - 1
        a1 = balanced(text, a0)
        params_raw = split_top(text[a0 + 1:a1 - 1])
        nd = None
        params = []
        for p in params_raw:
            if "::" in p:
                p, ann = p.split("::")
                mm = re.search(r"Data\.Array\{(\d)\}", ann)
                if mm and not params:
                    nd = int(mm.group(1))
            params.append(p.strip())
        e = re.compile(r"\n\s*return\s*\n\s*end").search(text, a1)
        stmts = [parse_assign(s) for s in logical_statements(text[a1:e.start()])]
        out.append(Kernel(name, params, nd, stmts, [x.strip() for x in idx.strip("()").split(",") if x.strip()] if is_idx else None,
                          text[m.start():e.end()]))
    return out


class Evaluator:
    """evaluates kernel statements on numpy arrays"""

    def __init__(self, macros: Macros, literals: str):
        assert literals in ("f32", "f64")
        self.m = macros
        self.lit = F32 if literals == "f32" else np.float64
        self.ivars = ["iz", "iy", "ix"] if macros.ndims == 3 else ["iz", "ix"]

    # ---- index expressions -> (index variable | None, integer offset)
    def idx(self, n, env):
        k = n[0]
        if k == "int":
            return (None, n[1])
        if k == "var":
            v = n[1]
            if v in self.ivars:
                return (v, 0)
            if v in self.m.inner:
                return (v[:2], self.m.inner[v])
            val = env[v]
            assert isinstance(val, (int, np.integer)), f"index term {v} is not an integer"
            return (None, int(val))
        if k == "bin" and n[1] in "+-":
            (va, oa), (vb, ob) = self.idx(n[2], env), self.idx(n[3], env)
            assert not (va and vb)
            assert n[1] == "+" or vb is None
            return (va or vb, oa + ob if n[1] == "+" else oa - ob)
        raise SyntaxError(f"index expression {n}")

    def ref(self, arr, idxs, env, dom):
        """numpy view of arr[idxs] over the domain, shaped to broadcast against the full domain rank"""
        sl, present = [], []
        assert arr.ndim == len(idxs), (arr.shape, idxs)
        for d, ix in enumerate(idxs):
            v, off = self.idx(ix, env)
            if v is None:
                assert 1 <= off <= arr.shape[d], f"fixed index {off} out of bounds {arr.shape}"
                sl.append(off - 1)
            else:
                lo, hi = dom[v]
                assert lo + off >= 1 and hi + off <= arr.shape[d], f"index {v}{off:+d} over {lo}:{hi} leaves the array {arr.shape} (BoundsError in Julia)"
                sl.append(slice(lo - 1 + off, hi + off))
                present.append(v)
        view = arr[tuple(sl)]
        order = [v for v in dom]
        assert present == [v for v in order if v in present], "index variables out of order"
        shape = [(dom[v][1] - dom[v][0] + 1) if v in present else 1 for v in order]
        return view.reshape(shape)

    def ev(self, n, env, dom):
        k = n[0]
        if k == "flt":
            return self.lit(n[1])
        if k == "int":
            return n[1]                         # weak Python int: Float32 * Int stays Float32, like Julia
        if k == "var":
            v = env[n[1]]
            if isinstance(v, str):              # macro argument bound to an array NAME
                v = env[v]
            return v
        if k == "neg":
            return -self.ev(n[1], env, dom)
        if k == "bin":
            a, b = self.ev(n[2], env, dom), self.ev(n[3], env, dom)
            return a + b if n[1] == "+" else a - b if n[1] == "-" else a * b if n[1] == "*" else a / b
        if k == "ref":
            assert n[1][0] == "var"
            nm = n[1][1]
            arr = env[nm]
            if isinstance(arr, str):
                arr = env[arr]
            return self.ref(arr, n[2], env, dom)
        if k == "mac":
            assert len(n[2]) == 1 and n[2][0][0] == "var", n
            return self.ev(self.m.body[n[1]], {**env, "A": n[2][0][1]}, dom)
        raise SyntaxError(f"cannot evaluate {n}")

    # ---- statements
    def lhs_view(self, lhs, env, dom_given):
        """(array view to assign into, domain) for `@all(A)`, `@inn(A)` or `A[i, j]`"""
        if lhs[0] == "mac":
            name = lhs[2][0][1]
            arr = env[name]
            shrink = {"all": 0, "inn": self.m.inn_shrink}[lhs[1]]
            dom = {v: (1, arr.shape[d] - shrink) for d, v in enumerate(self.ivars)}
            if any(hi < 1 for _, hi in dom.values()):
                return None, dom
            return self.ev(self.m.body[lhs[1]], {**env, "A": name}, dom), dom
        assert lhs[0] == "ref"
        arr = env[lhs[1][1]]
        return self.ref(arr, lhs[2], env, dom_given), dom_given

    def run(self, kern: Kernel, args: list, ranges: Optional[List[Tuple[int, int]]] = None):
        assert len(args) == len(kern.params), (kern.name, len(args), len(kern.params))
        env = dict(zip(kern.params, args))
        dom_given = None
        saved = self.ivars
        if kern.indices is not None:
            assert ranges is not None and len(ranges) == len(kern.indices), (kern.name, ranges)
            dom_given = {v: r for v, r in zip(kern.indices, ranges)}
            if any(hi < lo for lo, hi in dom_given.values()):
                return
            self.ivars = list(kern.indices)      # an @parallel_indices kernel names its own index variables
        elif ranges is not None:
            raise ValueError("ranges given to an @parallel (non-indices) kernel")
        try:
            for lhs, rhs in kern.stmts:
                view, dom = self.lhs_view(lhs, env, dom_given)
                if view is None:
                    continue
                view[...] = self.ev(rhs, env, dom)   # the store rounds to the array's element type (Julia setindex! converts)
        finally:
            self.ivars = saved


# ====================================================================================================
# array shapes: get_mgrid methods of src/fields.jl
# ====================================================================================================
def parse_field_shapes(text: str, order: int) -> Dict[Tuple[str, str, int], List[str]]:
    """{(field, FdtdAcoustic|FdtdElastic, ndims): [length expression per axis]} from `get_mgrid(::f, ::attrib, mz, (my,) mx) = [...]`"""
    text = strip_comments(text)
    out = {}
    for m in re.finditer(r"get_mgrid\(::(\w+), ::(\w+), ([^)]*)\)\s*=\s*\[", text):
        fld, attrib, gargs = m.group(1), m.group(2), [a.strip() for a in m.group(3).split(",")]
        b0 = m.end() - 1
        inside = text[b0 + 1:balanced(text, b0, "[", "]") - 1]
        elems = split_top(inside)
        assert len(elems) == len(gargs), (fld, elems)
        lens = []
        for e, g in zip(elems, gargs):
            if e == g:
                lens.append(f"length({g})")
            else:
                mm = re.search(r"length\s*=\s*(.*?)\s*,?\s*\)\s*$", e, re.S)
                assert mm, e
                lens.append(" ".join(mm.group(1).split()))
        out[(fld, attrib, len(gargs))] = lens
    return out


def eval_length(expr: str, n: Dict[str, int], order: int) -> int:
    e = expr
    for g, v in n.items():
        e = e.replace(f"length({g})", str(v))
    e = re.sub(r"(\d)\(", r"\1*(", e)           # Julia juxtaposition 2(_fd_order - 1)
    e = e.replace("_fd_order", str(order))
    assert re.fullmatch(r"[\d\s()+*\-]+", e), e
    return int(eval(e))
