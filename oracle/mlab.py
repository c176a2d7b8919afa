"""mlab -- a minimal MATLAB-subset interpreter.  TEST INFRASTRUCTURE, never imported by the product.

Purpose: the reference (dangnq2501/Triple-Tensor-Decomposition-with-ADMM) is pure MATLAB and neither MATLAB nor
Octave exists in this image.  This interpreter executes the reference's OWN, UNMODIFIED `.m` sources where they lie
under /root/reference (fast_robust_triple_tensor/triple_decomp_ADMM.m, triple_decomp_ALS.m, triple_product.m,
unfold.m, buildF/G/H.m, soft_threshold.m, origin_triple_tensor/buildF/G/H.m, kronF/G/H.m, and the local function
`evaluate` of traffic_triple_comparison.m:194-202), so that statement order, index conventions, reshape/permute
orders, operator association and MATLAB's function-resolution order (local functions shadow files on the path,
SURVEY fact 2) come from the reference's source text instead of from a hand restatement.  It is used
  * by tests/golden/make_ref_golden.py to generate the committed fixtures tests/golden/ref_m/*.npz
    (outputs of the reference source with randn shadowed so that A0, B0, C0 are injected, triple_decomp_ADMM.m:23),
  * by tests/test_reference_pin.py to pin oracle/tritd_oracle.py (and, on the GPU, the CUDA path) to those outputs.
What it does NOT pin: MathWorks' built-ins.  pinv (LAPACK SVD + the documented cutoff max(size)*eps(norm)),
mtimes (dgemm), norm (dnrm2) are numpy/OpenBLAS/LAPACK here; they agree with MATLAB's to summation-order noise.

Language subset: functions (multiple outputs, `~` placeholders, local functions, path lookup), for / while / if /
elseif / else / switch / break / continue / return, real double and logical N-d arrays in column-major order,
structs (field get / set), strings, matrix literals with MATLAB's whitespace rules, ranges, `end` in indices,
linear / subscript / logical indexing and indexed assignment, operators + - * / ^ .* ./ .^ ' .' comparison
& | && || ~, command-syntax `clear`.  Everything else raises MlabError (no silent guess).
"""
from __future__ import annotations

import os
import re

import numpy as np


class MlabError(Exception):
    pass


class _Break(Exception):
    pass


class _Continue(Exception):
    pass


class _Return(Exception):
    pass


KEYWORDS = {"function", "for", "while", "if", "elseif", "else", "end", "switch", "case", "otherwise", "break",
            "continue", "return"}

# ---------------------------------------------------------------------------------------------------------------
# tokenizer
# ---------------------------------------------------------------------------------------------------------------
_NUM = re.compile(r"(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)")
_ID = re.compile(r"[A-Za-z_]\w*")
_OPS = ["...", "==", "~=", "<=", ">=", "&&", "||", ".*", "./", ".^", ".'", "+", "-", "*", "/", "^", "<", ">", "=", "&",
        "|", "~", "(", ")", "[", "]", "{", "}", ",", ";", ":", ".", "'", "@"]


class Tok:
    __slots__ = ("kind", "val", "ws_before", "ws_after", "line")

    def __init__(self, kind, val, ws_before, line):
        self.kind, self.val, self.ws_before, self.ws_after, self.line = kind, val, ws_before, False, line

    def __repr__(self):
        return f"{self.kind}:{self.val!r}@{self.line}"


def tokenize(src: str):
    toks = []
    i, n, line = 0, len(src), 1
    depth = 0            # bracket depth [ ] { } (quote disambiguation needs it)
    ws = True
    while i < n:
        c = src[i]
        if c in " \t\r":
            i += 1
            ws = True
            if toks:
                toks[-1].ws_after = True
            continue
        if c == "%":
            # block comment %{ ... %} on their own lines
            ls = src.rfind("\n", 0, i) + 1
            le = src.find("\n", i)
            le = n if le < 0 else le
            if src[ls:le].strip() == "%{":
                end = re.compile(r"^[ \t]*%\}[ \t]*$", re.M).search(src, le)
                if not end:
                    raise MlabError("unterminated block comment")
                line += src.count("\n", i, end.end())
                i = end.end()
                continue
            i = le
            continue
        if src.startswith("...", i):
            le = src.find("\n", i)
            i = n if le < 0 else le + 1
            line += 1
            ws = True
            continue
        if c == "\n":
            toks.append(Tok("nl", "\n", ws, line))
            i += 1
            line += 1
            ws = True
            continue
        prev = toks[-1] if toks else None
        if c == "'" or c == '"':
            is_transpose = (c == "'" and prev is not None and
                            (prev.kind in ("id", "num") or prev.val in (")", "]", "}", "'", ".'") or
                             (prev.kind == "kw" and prev.val == "end")) and not (depth > 0 and ws))
            if not is_transpose:
                j = i + 1
                buf = []
                while True:
                    if j >= n or src[j] == "\n":
                        raise MlabError(f"line {line}: unterminated string")
                    if src[j] == c:
                        if j + 1 < n and src[j + 1] == c:
                            buf.append(c)
                            j += 2
                            continue
                        break
                    buf.append(src[j])
                    j += 1
                toks.append(Tok("str", "".join(buf), ws, line))
                i = j + 1
                ws = False
                continue
        m = _NUM.match(src, i)
        if m and (c.isdigit() or (c == "." and i + 1 < n and src[i + 1].isdigit())):
            # "1." followed by an operator char of an element-wise op (1.*x) belongs to the operator
            text = m.group(0)
            if text.endswith(".") and i + len(text) < n and src[i + len(text)] in "*/^'":
                text = text[:-1]
            toks.append(Tok("num", float(text), ws, line))
            i += len(text)
            ws = False
            continue
        m = _ID.match(src, i)
        if m:
            w = m.group(0)
            toks.append(Tok("kw" if w in KEYWORDS else "id", w, ws, line))
            i = m.end()
            ws = False
            continue
        for op in _OPS:
            if src.startswith(op, i):
                if op in "[{":
                    depth += 1
                elif op in "]}":
                    depth -= 1
                toks.append(Tok("op", op, ws, line))
                i += len(op)
                ws = False
                break
        else:
            raise MlabError(f"line {line}: unexpected character {c!r}")
    toks.append(Tok("nl", "\n", True, line))
    toks.append(Tok("eof", None, True, line))
    return toks


# ---------------------------------------------------------------------------------------------------------------
# parser -> tuples
# ---------------------------------------------------------------------------------------------------------------
class Parser:
    def __init__(self, toks, fname="<src>"):
        self.t, self.p, self.fname = toks, 0, fname
        self.mat = [False]        # stack: inside a matrix literal (whitespace separates elements)?
        self.idx = [False]        # stack: inside an index / call argument list (`end` is a value)?

    # -- token helpers
    def peek(self, k=0):
        return self.t[self.p + k]

    def next(self):
        tk = self.t[self.p]
        self.p += 1
        return tk

    def at(self, val, kind=None):
        tk = self.peek()
        return tk.val == val and (kind is None or tk.kind == kind) and tk.kind != "str"

    def accept(self, val):
        if self.at(val):
            return self.next()
        return None

    def expect(self, val):
        if not self.at(val):
            tk = self.peek()
            raise MlabError(f"{self.fname}:{tk.line}: expected {val!r}, found {tk.val!r}")
        return self.next()

    def skip_nl(self):
        while self.peek().kind == "nl" or self.at(";") or self.at(","):
            self.next()

    # -- file
    def parse_file(self):
        """-> (script_statements, [function definitions in file order])"""
        funcs, script = [], []
        self.skip_nl()
        while self.peek().kind != "eof":
            if self.at("function", "kw"):
                funcs.append(self.parse_function())
            else:
                script.append(self.parse_statement())
            self.skip_nl()
        return script, funcs

    def parse_function(self):
        self.expect("function")
        outs = []
        # forms: function name(...) | function out = name(...) | function [o1, o2] = name(...)
        if self.at("["):
            self.next()
            while not self.at("]"):
                if self.accept(","):
                    continue
                outs.append(self.next().val)
            self.next()
            self.expect("=")
            name = self.next().val
        else:
            name = self.next().val
            if self.accept("="):
                outs = [name]
                name = self.next().val
        params = []
        if self.accept("("):
            while not self.at(")"):
                if self.accept(","):
                    continue
                tk = self.next()
                params.append("~" if tk.val == "~" else tk.val)
            self.next()
        body = self.parse_block(("end", "function"))
        if self.at("end", "kw"):
            self.next()
        return ("function", name, params, outs, body)

    def parse_block(self, stops):
        out = []
        self.skip_nl()
        while True:
            tk = self.peek()
            if tk.kind == "eof" or (tk.kind == "kw" and tk.val in stops):
                return out
            out.append(self.parse_statement())
            self.skip_nl()

    def end_stmt(self):
        """consume the statement terminator; returns True when output is suppressed (';')"""
        if self.accept(";"):
            return True
        if self.accept(",") or self.peek().kind in ("nl", "eof"):
            return False
        tk = self.peek()
        raise MlabError(f"{self.fname}:{tk.line}: unexpected {tk.val!r} at end of statement")

    def parse_statement(self):
        tk = self.peek()
        line = tk.line
        if tk.kind == "kw":
            if tk.val == "for":
                self.next()
                paren = self.accept("(")
                var = self.next().val
                self.expect("=")
                rng = self.parse_expr()
                if paren:
                    self.expect(")")
                body = self.parse_block(("end",))
                self.expect("end")
                return ("for", var, rng, body, line)
            if tk.val == "while":
                self.next()
                cond = self.parse_expr()
                body = self.parse_block(("end",))
                self.expect("end")
                return ("while", cond, body, line)
            if tk.val == "if":
                self.next()
                clauses, els = [], None
                cond = self.parse_expr()
                body = self.parse_block(("end", "elseif", "else"))
                clauses.append((cond, body))
                while True:
                    if self.at("elseif", "kw"):
                        self.next()
                        cond = self.parse_expr()
                        clauses.append((cond, self.parse_block(("end", "elseif", "else"))))
                    elif self.at("else", "kw"):
                        self.next()
                        els = self.parse_block(("end",))
                    else:
                        break
                self.expect("end")
                return ("if", clauses, els, line)
            if tk.val == "switch":
                self.next()
                subj = self.parse_expr()
                self.skip_nl()
                cases, other = [], None
                while self.at("case", "kw"):
                    self.next()
                    val = self.parse_expr()
                    cases.append((val, self.parse_block(("end", "case", "otherwise"))))
                if self.at("otherwise", "kw"):
                    self.next()
                    other = self.parse_block(("end",))
                self.expect("end")
                return ("switch", subj, cases, other, line)
            if tk.val in ("break", "continue", "return"):
                self.next()
                self.end_stmt()
                return (tk.val, line)
            raise MlabError(f"{self.fname}:{line}: unexpected keyword {tk.val!r}")
        # command syntax: clear x y / close all / clc
        if tk.kind == "id" and tk.val in ("clear", "clc", "close", "hold", "format", "warning") and \
                (self.peek(1).kind in ("id", "nl", "eof") or self.peek(1).val == ";") and not self.peek(1).val == "=":
            self.next()
            names = []
            while self.peek().kind == "id":
                names.append(self.next().val)
            self.end_stmt()
            return ("command", tk.val, names, line)
        # multi-assignment [a, ~, c] = f(...)
        if tk.val == "[" and tk.kind == "op":
            j, depth = self.p, 0
            while True:
                v = self.t[j]
                if v.kind in ("nl", "eof"):
                    break
                if v.kind == "op" and v.val == "[":
                    depth += 1
                elif v.kind == "op" and v.val == "]":
                    depth -= 1
                    if depth == 0:
                        break
                j += 1
            if self.t[j].val == "]" and self.t[j + 1].kind == "op" and self.t[j + 1].val == "=":
                self.next()
                lhs = []
                while not self.at("]"):
                    if self.accept(","):
                        continue
                    if self.accept("~"):
                        lhs.append(None)
                    else:
                        lhs.append(self.parse_lvalue())
                self.next()
                self.expect("=")
                rhs = self.parse_expr()
                self.end_stmt()
                return ("massign", lhs, rhs, line)
        # assignment or expression statement
        start = self.p
        if tk.kind == "id":
            try:
                lv = self.parse_lvalue()
                if self.at("=") and self.peek().kind == "op":
                    self.next()
                    rhs = self.parse_expr()
                    quiet = self.end_stmt()
                    return ("assign", lv, rhs, quiet, line)
            except MlabError:
                pass
            self.p = start
        e = self.parse_expr()
        quiet = self.end_stmt()
        return ("expr", e, quiet, line)

    def parse_lvalue(self):
        name = self.next()
        if name.kind != "id":
            raise MlabError(f"{self.fname}:{name.line}: bad assignment target {name.val!r}")
        lv = ("var", name.val)
        while True:
            if self.at("(") and not (self.mat[-1] and self.peek().ws_before):
                self.next()
                args = self.parse_args(")")
                lv = ("index", lv, args)
            elif self.at(".") and self.peek(1).kind == "id":
                self.next()
                lv = ("field", lv, self.next().val)
            else:
                return lv

    def parse_args(self, close):
        self.mat.append(False)
        self.idx.append(True)
        args = []
        while self.peek().kind == "nl":
            self.next()
        while not self.at(close):
            if self.at(":") and (self.peek(1).val in (",", close)) and self.peek(1).kind == "op":
                self.next()
                args.append(("colon_all",))
            else:
                args.append(self.parse_expr())
            if not self.accept(","):
                break
        self.expect(close)
        self.mat.pop()
        self.idx.pop()
        return args

    # -- expressions, lowest precedence first
    def parse_expr(self):
        return self.parse_oror()

    def _binary_loop(self, sub, ops):
        left = sub()
        while True:
            tk = self.peek()
            if tk.kind == "op" and tk.val in ops and not self._starts_new_element(tk):
                self.next()
                right = sub()
                left = ("bin", tk.val, left, right)
            else:
                return left

    def _starts_new_element(self, tk):
        # inside [ ]: "a -b" is two elements, "a - b" and "a-b" are one
        return self.mat[-1] and tk.val in ("+", "-") and tk.ws_before and not tk.ws_after

    def parse_oror(self):
        left = self.parse_andand()
        while self.at("||"):
            self.next()
            left = ("oror", left, self.parse_andand())
        return left

    def parse_andand(self):
        left = self.parse_or()
        while self.at("&&"):
            self.next()
            left = ("andand", left, self.parse_or())
        return left

    def parse_or(self):
        return self._binary_loop(self.parse_and, ("|",))

    def parse_and(self):
        return self._binary_loop(self.parse_cmp, ("&",))

    def parse_cmp(self):
        return self._binary_loop(self.parse_range, ("==", "~=", "<", "<=", ">", ">="))

    def parse_range(self):
        first = self.parse_add()
        if self.at(":") and self.peek().kind == "op" and not (self.idx[-1] and self.peek(1).val in (",", ")")):
            self.next()
            second = self.parse_add()
            if self.at(":") and self.peek().kind == "op":
                self.next()
                third = self.parse_add()
                return ("range", first, second, third)
            return ("range", first, None, second)
        return first

    def parse_add(self):
        return self._binary_loop(self.parse_mul, ("+", "-"))

    def parse_mul(self):
        return self._binary_loop(self.parse_unary, ("*", "/", ".*", "./"))

    def parse_unary(self):
        tk = self.peek()
        if tk.kind == "op" and tk.val in ("-", "+", "~"):
            self.next()
            operand = self.parse_unary()
            return ("un", tk.val, operand)
        return self.parse_power()

    def parse_power(self):
        base = self.parse_postfix()
        while self.peek().kind == "op" and self.peek().val in ("^", ".^"):
            op = self.next().val
            tk = self.peek()
            if tk.kind == "op" and tk.val in ("-", "+", "~"):       # 2^-1
                self.next()
                exp = ("un", tk.val, self.parse_power_operand())
            else:
                exp = self.parse_power_operand()
            base = ("bin", op, base, exp)
        return base

    def parse_power_operand(self):
        return self.parse_postfix()

    def parse_postfix(self):
        e = self.parse_primary()
        while True:
            tk = self.peek()
            if tk.kind != "op":
                return e
            if tk.val == "(" and not (self.mat[-1] and tk.ws_before):
                self.next()
                e = ("call", e, self.parse_args(")"))
            elif tk.val == "." and self.peek(1).kind == "id" and not tk.ws_before:
                self.next()
                e = ("getfield", e, self.next().val)
            elif tk.val in ("'", ".'") and not (self.mat[-1] and tk.ws_before):
                self.next()
                e = ("transpose", e)
            else:
                return e

    def parse_primary(self):
        tk = self.next()
        if tk.kind == "num":
            return ("num", tk.val)
        if tk.kind == "str":
            return ("str", tk.val)
        if tk.kind == "id":
            return ("name", tk.val)
        if tk.kind == "kw" and tk.val == "end" and self.idx[-1]:
            return ("end",)
        if tk.kind == "op" and tk.val == "(":
            self.mat.append(False)
            self.idx.append(False)
            e = self.parse_expr()
            self.expect(")")
            self.mat.pop()
            self.idx.pop()
            return ("paren", e)
        if tk.kind == "op" and tk.val == "[":
            self.mat.append(True)
            self.idx.append(self.idx[-1])
            rows, row = [], []
            while True:
                if self.at("]"):
                    self.next()
                    break
                if self.accept(";") or self.peek().kind == "nl":
                    if self.peek().kind == "nl":
                        self.next()
                    if row:
                        rows.append(row)
                        row = []
                    continue
                if self.accept(","):
                    continue
                row.append(self.parse_expr())
            if row:
                rows.append(row)
            self.mat.pop()
            self.idx.pop()
            return ("matrix", rows)
        if tk.kind == "op" and tk.val == ":":
            return ("colon_all",)
        raise MlabError(f"{self.fname}:{tk.line}: unexpected token {tk.val!r}")


# ---------------------------------------------------------------------------------------------------------------
# values
# ---------------------------------------------------------------------------------------------------------------
def mat(x):
    """numeric / logical value -> F-ordered ndarray with ndim >= 2 and no trailing singleton dims beyond 2"""
    a = np.asarray(x)
    if a.dtype != np.bool_ and a.dtype != np.float64:
        a = a.astype(np.float64)
    if a.ndim == 0:
        a = a.reshape(1, 1)
    elif a.ndim == 1:
        a = a.reshape(1, -1)
    while a.ndim > 2 and a.shape[-1] == 1:
        a = a.reshape(a.shape[:-1], order="F")
    if not a.flags.f_contiguous:
        a = np.asfortranarray(a)
    return a


def is_scalar(a):
    return isinstance(a, np.ndarray) and a.size == 1


def scalar(a, what="value"):
    if isinstance(a, (int, float, bool)):
        return float(a)
    if isinstance(a, np.ndarray) and a.size == 1:
        return float(a.reshape(-1)[0])
    raise MlabError(f"{what} must be a scalar")


def dims_of(a, n):
    """size vector of `a` padded with ones / merged to exactly n dims (MATLAB's [d1,..,dn] = size(a))"""
    s = list(a.shape)
    if n >= len(s):
        return s + [1] * (n - len(s))
    return s[:n - 1] + [int(np.prod(s[n - 1:]))]


def broadcast2(a, b):
    nd = max(a.ndim, b.ndim)
    a2 = a.reshape(tuple(a.shape) + (1,) * (nd - a.ndim), order="F")
    b2 = b.reshape(tuple(b.shape) + (1,) * (nd - b.ndim), order="F")
    for x, y in zip(a2.shape, b2.shape):
        if x != y and x != 1 and y != 1:
            raise MlabError(f"Arrays have incompatible sizes for this operation: {a.shape} vs {b.shape}")
    return a2, b2


def num(a):
    return a.astype(np.float64) if a.dtype == np.bool_ else a


def matlab_pinv(A):
    """pinv(A): SVD-based, singular values <= max(size(A)) * eps(norm(A)) are treated as zero (MATLAB doc / pinv.m)."""
    A = num(A)
    U, s, Vh = np.linalg.svd(A, full_matrices=False)
    if s.size == 0:
        return mat(np.zeros((A.shape[1], A.shape[0])))
    tol = max(A.shape) * np.spacing(s[0])
    r = int(np.sum(s > tol))
    V = Vh.T[:, :r]
    return mat((V * (1.0 / s[:r])) @ U[:, :r].T)


# ---------------------------------------------------------------------------------------------------------------
# interpreter
# ---------------------------------------------------------------------------------------------------------------
class FileUnit:
    def __init__(self, path, script, funcs):
        self.path = path
        self.script = script
        self.funcs = {f[1]: f for f in funcs}
        self.main = funcs[0] if funcs else None


class Interp:
    """interp = Interp(path=[dirs...]);  outs = interp.call('triple_decomp_ADMM', [D, r, opts], nargout=5)"""

    def __init__(self, path=(), overrides=None, echo=False):
        self.path = list(path)
        self.units = {}            # file path -> FileUnit
        self.overrides = dict(overrides or {})   # name -> python callable(interp, args, nargout) -> list of values
        self.out = []              # fprintf / disp output
        self.echo = echo
        self.calls = []            # (function name, file) of every user function called: resolution evidence
        self.builtins = _make_builtins()

    # -- loading
    def load(self, fpath, functions_only=False):
        """parse a .m file; functions_only: skip the script part of a script file (everything before its first
        `function` line) and keep only its local functions, e.g. evaluate() of traffic_triple_comparison.m"""
        fpath = os.path.abspath(fpath)
        if fpath not in self.units:
            with open(fpath, encoding="utf-8") as f:
                src = f.read()
            if functions_only:
                m = re.search(r"^[ \t]*function\b", src, re.M)
                if not m:
                    raise MlabError(f"{fpath}: no function definitions")
                src = "\n" * src.count("\n", 0, m.start()) + src[m.start():]      # keeps the line numbers
            script, funcs = Parser(tokenize(src), os.path.basename(fpath)).parse_file()
            self.units[fpath] = FileUnit(fpath, script, funcs)
        return self.units[fpath]

    def find_on_path(self, name):
        for d in self.path:
            fp = os.path.join(d, name + ".m")
            if os.path.isfile(fp):
                return self.load(fp)
        return None

    # -- calling
    def call(self, name, args, nargout=1, unit=None):
        args = [self.to_value(a) for a in args]
        res = self.call_named(name, args, nargout, unit)
        return res

    def to_value(self, a):
        if isinstance(a, dict):
            return {k: self.to_value(v) for k, v in a.items()}
        if isinstance(a, str):
            return a
        return mat(a)

    def call_named(self, name, args, nargout, unit):
        if name in self.overrides:
            return self.overrides[name](self, args, nargout)
        if unit is not None and name in unit.funcs:                 # local functions shadow the path
            return self.call_function(unit, unit.funcs[name], args, nargout)
        u = self.find_on_path(name)
        if u is not None and u.main is not None:
            return self.call_function(u, u.main, args, nargout)
        if name in self.builtins:
            return self.builtins[name](self, args, nargout)
        raise MlabError(f"Unrecognized function or variable '{name}'.")

    def call_function(self, unit, fdef, args, nargout):
        _, name, params, outs, body = fdef
        if len(args) > len(params):
            raise MlabError(f"{name}: too many input arguments")
        self.calls.append((name, unit.path))
        ws = {}
        for pn, a in zip(params, args):
            if pn != "~":
                ws[pn] = a
        ws["__nargin__"] = len(args)
        ws["__nargout__"] = nargout
        try:
            self.exec_block(body, ws, unit)
        except _Return:
            pass
        res = []
        for o in outs[:max(nargout, 1)]:
            if o not in ws:
                if len(res) < nargout:
                    raise MlabError(f"Output argument \"{o}\" (and possibly others) not assigned a value in {name}.")
                break
            res.append(ws[o])
        return res

    # -- statements
    def exec_block(self, stmts, ws, unit):
        for s in stmts:
            self.exec_stmt(s, ws, unit)

    def exec_stmt(self, s, ws, unit):
        k = s[0]
        if k == "assign":
            _, lv, rhs, _quiet, _line = s
            v = self.eval_multi(rhs, ws, unit, 1)
            if not v:
                raise MlabError(f"line {_line}: right-hand side produced no value")
            self.assign(lv, v[0], ws, unit)
        elif k == "massign":
            _, lhs, rhs, _line = s
            vals = self.eval_multi(rhs, ws, unit, len(lhs))
            if len(vals) < len([x for x in lhs if x is not None and False]) or len(vals) < self._needed(lhs):
                raise MlabError(f"line {_line}: too many output arguments requested")
            for lv, v in zip(lhs, vals):
                if lv is not None:
                    self.assign(lv, v, ws, unit)
        elif k == "expr":
            _, e, _quiet, _line = s
            vals = self.eval_multi(e, ws, unit, 0)
            if vals:
                ws["ans"] = vals[0]
        elif k == "for":
            _, var, rng, body, _line = s
            r = self.eval(rng, ws, unit)
            r2 = r.reshape(r.shape[0], -1, order="F")
            for c in range(r2.shape[1]):
                ws[var] = mat(r2[:, c].copy()) if r2.shape[0] != 1 else mat(r2[0, c])
                try:
                    self.exec_block(body, ws, unit)
                except _Break:
                    break
                except _Continue:
                    continue
        elif k == "while":
            _, cond, body, _line = s
            while self.truth(self.eval(cond, ws, unit)):
                try:
                    self.exec_block(body, ws, unit)
                except _Break:
                    break
                except _Continue:
                    continue
        elif k == "if":
            _, clauses, els, _line = s
            for cond, body in clauses:
                if self.truth(self.eval(cond, ws, unit)):
                    self.exec_block(body, ws, unit)
                    return
            if els is not None:
                self.exec_block(els, ws, unit)
        elif k == "switch":
            _, subj, cases, other, _line = s
            v = self.eval(subj, ws, unit)
            for cv, body in cases:
                c = self.eval(cv, ws, unit)
                same = (v == c) if isinstance(v, str) or isinstance(c, str) else (scalar(v) == scalar(c))
                if same:
                    self.exec_block(body, ws, unit)
                    return
            if other is not None:
                self.exec_block(other, ws, unit)
        elif k == "break":
            raise _Break()
        elif k == "continue":
            raise _Continue()
        elif k == "return":
            raise _Return()
        elif k == "command":
            _, cmd, names, _line = s
            if cmd == "clear":
                for nme in names:
                    ws.pop(nme, None)
        else:
            raise MlabError(f"unknown statement {k}")

    @staticmethod
    def _needed(lhs):
        last = 0
        for i, lv in enumerate(lhs):
            if lv is not None:
                last = i + 1
        return last

    def truth(self, v):
        if isinstance(v, str):
            return len(v) > 0
        return v.size > 0 and bool(np.all(v != 0))

    def assign(self, lv, v, ws, unit):
        kind = lv[0]
        if kind == "var":
            ws[lv[1]] = v
        elif kind == "field":
            base = lv[1]
            if base[0] != "var":
                raise MlabError("nested struct assignment is not supported")
            st = ws.get(base[1])
            st = dict(st) if isinstance(st, dict) else {}
            st[lv[2]] = v
            ws[base[1]] = st
        elif kind == "index":
            base = lv[1]
            if base[0] != "var":
                raise MlabError("indexed assignment into a struct field is not supported")
            if base[1] not in ws:
                raise MlabError(f"indexed assignment creating '{base[1]}' is not supported (array growth)")
            arr = ws[base[1]]
            if not isinstance(arr, np.ndarray):
                raise MlabError("indexed assignment target is not an array")
            ws[base[1]] = self.index_assign(arr, lv[2], v, ws, unit)
        else:
            raise MlabError("bad assignment target")

    # -- indexing
    def eval_index_args(self, arr, args, ws, unit):
        """-> list of index specs (None = ':' | int ndarray 0-based with original shape | bool mask)"""
        n = len(args)
        shp = dims_of(arr, n) if n > 1 else [arr.size]
        out = []
        for pos, a in enumerate(args):
            if a[0] == "colon_all":
                out.append(None)
                continue
            v = self.eval(a, ws, unit, end_ctx=(shp[pos],))
            if isinstance(v, str):
                raise MlabError("character indices are not supported")
            if v.dtype == np.bool_:
                out.append(v)
            else:
                iv = np.rint(v).astype(np.int64)
                if np.any(np.abs(v - iv) > 0) or np.any(iv < 1):
                    raise MlabError("Array indices must be positive integers or logical values.")
                out.append(iv - 1)
        return out, shp

    def index_get(self, arr, args, ws, unit):
        specs, shp = self.eval_index_args(arr, args, ws, unit)
        if len(specs) == 0:
            return arr
        if len(specs) == 1:
            sp = specs[0]
            flat = arr.reshape(-1, order="F")
            if sp is None:
                return mat(flat.reshape(-1, 1).copy())
            if sp.dtype == np.bool_:
                if sp.size > arr.size:
                    raise MlabError("logical index exceeds the array")
                idx = np.flatnonzero(sp.reshape(-1, order="F"))
                res = flat[idx]
                return mat(res.reshape(1, -1) if (arr.ndim == 2 and arr.shape[0] == 1 and arr.shape[1] != 1) else res.reshape(-1, 1))
            if sp.size and sp.max() >= arr.size:
                raise MlabError("Index exceeds the number of array elements.")
            res = flat[sp.reshape(-1, order="F")]
            arr_is_vec = arr.ndim == 2 and 1 in arr.shape
            idx_is_vec = sp.ndim == 2 and 1 in sp.shape
            if arr_is_vec and idx_is_vec:
                return mat(res.reshape(1, -1) if arr.shape[0] == 1 and arr.shape[1] != 1 or (arr.size == 1 and sp.shape[0] == 1) else res.reshape(-1, 1))
            return mat(res.reshape(sp.shape, order="F"))
        a = arr.reshape(shp, order="F")
        idx = []
        for d, sp in enumerate(specs):
            if sp is None:
                idx.append(np.arange(shp[d]))
            elif sp.dtype == np.bool_:
                idx.append(np.flatnonzero(sp.reshape(-1, order="F")))
            else:
                if sp.size and sp.max() >= shp[d]:
                    raise MlabError(f"Index in position {d + 1} exceeds array bounds.")
                idx.append(sp.reshape(-1, order="F"))
        return mat(a[np.ix_(*idx)])

    def index_assign(self, arr, args, v, ws, unit):
        specs, shp = self.eval_index_args(arr, args, ws, unit)
        if isinstance(v, str):
            raise MlabError("assigning strings into arrays is not supported")
        out = arr.copy(order="F")
        if v.dtype != np.bool_ and out.dtype == np.bool_:
            out = out.astype(np.float64)
        if len(specs) == 1:
            sp = specs[0]
            flat = out.reshape(-1, order="F")
            assert np.shares_memory(flat, out)
            if sp is None:
                idx = np.arange(out.size)
            elif sp.dtype == np.bool_:
                idx = np.flatnonzero(sp.reshape(-1, order="F"))
            else:
                idx = sp.reshape(-1, order="F")
                if idx.size and idx.max() >= out.size:
                    raise MlabError("indexed assignment beyond the array (growth) is not supported")
            if v.size == 1:
                flat[idx] = v.reshape(-1)[0]
            elif v.size == idx.size:
                flat[idx] = v.reshape(-1, order="F")
            else:
                raise MlabError("Unable to perform assignment because the left and right sides have a different number of elements.")
            return out
        a = out.reshape(shp, order="F")
        assert np.shares_memory(a, out)
        idx = []
        for d, sp in enumerate(specs):
            if sp is None:
                idx.append(np.arange(shp[d]))
            elif sp.dtype == np.bool_:
                idx.append(np.flatnonzero(sp.reshape(-1, order="F")))
            else:
                if sp.size and sp.max() >= shp[d]:
                    raise MlabError("indexed assignment beyond the array (growth) is not supported")
                idx.append(sp.reshape(-1, order="F"))
        tshape = tuple(len(i) for i in idx)
        if v.size == 1:
            a[np.ix_(*idx)] = v.reshape(-1)[0]
        else:
            # MATLAB: sizes must agree after removing singleton dimensions
            if [d for d in tshape if d != 1] != [d for d in v.shape if d != 1]:
                raise MlabError(f"Unable to perform assignment because the size of the left side is {tshape} and the size of the right side is {v.shape}.")
            a[np.ix_(*idx)] = v.reshape(tshape, order="F")
        return out

    # -- expressions
    def eval(self, e, ws, unit, end_ctx=None):
        vals = self.eval_multi(e, ws, unit, 1, end_ctx)
        if not vals:
            raise MlabError("expression produced no value")
        return vals[0]

    def eval_multi(self, e, ws, unit, nargout, end_ctx=None):
        k = e[0]
        ev = lambda x: self.eval(x, ws, unit, end_ctx)   # noqa: E731
        if k == "num":
            return [mat(e[1])]
        if k == "str":
            return [e[1]]
        if k == "paren":
            return [ev(e[1])]
        if k == "end":
            if end_ctx is None:
                raise MlabError("'end' outside an index expression")
            return [mat(float(end_ctx[0]))]
        if k == "name":
            nme = e[1]
            if nme in ws:
                return [ws[nme]]
            return self.call_named(nme, [], nargout, unit)
        if k == "call":
            tgt, args = e[1], e[2]
            if tgt[0] == "name" and tgt[1] not in ws:
                argv = [self.eval(a, ws, unit, end_ctx) for a in args]
                return self.call_named(tgt[1], argv, nargout, unit)
            base = self.eval(tgt, ws, unit, end_ctx)
            if isinstance(base, (dict, str)):
                raise MlabError("indexing into structs / strings is not supported")
            return [self.index_get(base, args, ws, unit)]
        if k == "getfield":
            base = ev(e[1])
            if not isinstance(base, dict):
                raise MlabError("Dot indexing is not supported for variables of this type.")
            if e[2] not in base:
                raise MlabError(f'Unrecognized field name "{e[2]}".')
            return [base[e[2]]]
        if k == "transpose":
            v = ev(e[1])
            if v.ndim != 2:
                raise MlabError("TRANSPOSE does not support N-D arrays.")
            return [mat(v.T)]
        if k == "un":
            v = ev(e[2])
            if e[1] == "-":
                return [mat(-num(v))]
            if e[1] == "+":
                return [mat(num(v))]
            return [mat(num(v) == 0)]
        if k == "andand":
            left = ev(e[1])
            if not self.truth(left):
                return [mat(False)]
            return [mat(self.truth(ev(e[2])))]
        if k == "oror":
            left = ev(e[1])
            if self.truth(left):
                return [mat(True)]
            return [mat(self.truth(ev(e[2])))]
        if k == "range":
            a = scalar(ev(e[1]), "range start")
            b = scalar(ev(e[3]), "range end")
            st = scalar(ev(e[2]), "range step") if e[2] is not None else 1.0
            if st == 0 or (st > 0 and a > b) or (st < 0 and a < b):
                return [mat(np.zeros((1, 0)))]
            nsteps = int(np.floor((b - a) / st * (1 + 4 * np.finfo(float).eps)))
            return [mat(a + st * np.arange(nsteps + 1, dtype=np.float64))]
        if k == "matrix":
            rows = []
            for row in e[1]:
                vals = [self.eval(x, ws, unit, end_ctx) for x in row]
                if any(isinstance(v, str) for v in vals):
                    if all(isinstance(v, str) for v in vals):
                        rows.append("".join(vals))
                        continue
                    raise MlabError("mixed string / numeric matrices are not supported")
                vals = [num(v) for v in vals if v.size > 0 or len(vals) == 1]
                if vals:
                    rows.append(np.concatenate([v if v.ndim > 1 else v.reshape(1, -1) for v in vals], axis=1))
            if rows and isinstance(rows[0], str):
                if len(rows) != 1:
                    raise MlabError("multi-row char arrays are not supported")
                return [rows[0]]
            if not rows:
                return [mat(np.zeros((0, 0)))]
            return [mat(np.concatenate(rows, axis=0))]
        if k == "bin":
            return [self.binop(e[1], ev(e[2]), ev(e[3]))]
        if k == "colon_all":
            raise MlabError("':' outside an index expression")
        raise MlabError(f"unknown expression {k}")

    def binop(self, op, a, b):
        if isinstance(a, str) or isinstance(b, str):
            if op == "==" and isinstance(a, str) and isinstance(b, str):
                return mat(a == b)
            raise MlabError(f"operator {op} on strings is not supported")
        if op == "*":
            if a.size == 1 or b.size == 1:
                op = ".*"
            else:
                if a.ndim != 2 or b.ndim != 2:
                    raise MlabError("Arguments must be 2-D, or at least one argument must be scalar.")
                if a.shape[1] != b.shape[0]:
                    raise MlabError(f"Incorrect dimensions for matrix multiplication: {a.shape} * {b.shape}")
                return mat(num(a) @ num(b))
        if op == "/":
            if b.size == 1:
                op = "./"
            else:
                raise MlabError("matrix right division is not supported")
        if op == "^":
            if a.size == 1 and b.size == 1:
                op = ".^"
            else:
                raise MlabError("matrix power is not supported")
        x, y = broadcast2(a, b)
        x, y = num(x), num(y)
        if op == "+":
            return mat(x + y)
        if op == "-":
            return mat(x - y)
        if op == ".*":
            return mat(x * y)
        if op == "./":
            with np.errstate(divide="ignore", invalid="ignore"):
                return mat(x / y)
        if op == ".^":
            return mat(np.power(x, y))
        if op == "==":
            return mat(x == y)
        if op == "~=":
            return mat(x != y)
        if op == "<":
            return mat(x < y)
        if op == "<=":
            return mat(x <= y)
        if op == ">":
            return mat(x > y)
        if op == ">=":
            return mat(x >= y)
        if op == "&":
            return mat((x != 0) & (y != 0))
        if op == "|":
            return mat((x != 0) | (y != 0))
        raise MlabError(f"operator {op} is not supported")


# ---------------------------------------------------------------------------------------------------------------
# built-ins (MATLAB semantics for exactly the calls the reference makes)
# ---------------------------------------------------------------------------------------------------------------
def _dims_from_args(args):
    if len(args) == 1:
        v = args[0]
        if v.size == 1:
            n = int(scalar(v))
            return (n, n)
        return tuple(int(x) for x in v.reshape(-1, order="F"))
    return tuple(int(scalar(a)) for a in args)


def _format(fmt, args):
    """fprintf / sprintf: C-style format cycled over the flattened arguments (MATLAB semantics)"""
    flat = []
    for a in args:
        if isinstance(a, str):
            flat.append(a)
        else:
            flat.extend(float(x) for x in num(a).reshape(-1, order="F"))
    spec = re.compile(r"%(?:%|[-+ 0#]*\d*(?:\.\d+)?[diufeEgGsc])")
    pieces = spec.findall(fmt)
    nconv = sum(1 for p in pieces if p != "%%")

    def esc(s):
        return (s.replace("\\n", "\n").replace("\\t", "\t").replace("\\\\", "\\"))

    def once(vals):
        it = iter(vals)

        def rep(m):
            p = m.group(0)
            if p == "%%":
                return "%"
            try:
                v = next(it)
            except StopIteration:
                return ""
            conv = p[-1]
            if conv in "di":
                if isinstance(v, float) and v != int(v):
                    return (p[:-1] + "e") % v
                return (p[:-1] + "d") % int(v)
            if conv == "u":
                return (p[:-1] + "d") % int(v)
            if conv in "sc":
                return (p[:-1] + "s") % (v if isinstance(v, str) else ("%g" % v))
            return p % float(v)
        return spec.sub(rep, fmt)

    if nconv == 0 or not flat:
        return esc(once(flat))
    out = []
    for i in range(0, len(flat), nconv):
        out.append(once(flat[i:i + nconv]))
    return esc("".join(out))


def _minmax(fn, npfn):
    def f(interp, args, nargout):
        if len(args) == 2:
            x, y = broadcast2(num(args[0]), num(args[1]))
            return [mat(npfn(x, y))]      # np.maximum / np.minimum: NaN-propagating differs from MATLAB only when NaNs occur
        a = num(args[0])
        if a.size == 0:
            return [mat(np.zeros((0, 0)))]
        ax = next((i for i, d in enumerate(a.shape) if d != 1), 0)
        return [mat(fn(a, axis=ax, keepdims=True))]
    return f


def _make_builtins():
    B = {}

    def reg(name):
        def deco(fn):
            B[name] = fn
            return fn
        return deco

    @reg("size")
    def _size(interp, args, nargout):
        a = args[0]
        shp = (1, len(a)) if isinstance(a, str) else a.shape
        if isinstance(a, dict):
            shp = (1, 1)
        arr = np.empty(shp, dtype=np.bool_) if not isinstance(a, np.ndarray) else a
        if len(args) == 2:
            d = int(scalar(args[1]))
            return [mat(float(shp[d - 1]) if d <= len(shp) else 1.0)]
        if nargout <= 1:
            return [mat(np.array(shp, dtype=np.float64).reshape(1, -1))]
        return [mat(float(x)) for x in dims_of(arr, nargout)]

    @reg("numel")
    def _numel(interp, args, nargout):
        return [mat(float(args[0].size if isinstance(args[0], np.ndarray) else len(args[0])))]

    @reg("length")
    def _length(interp, args, nargout):
        a = args[0]
        return [mat(float(0 if a.size == 0 else max(a.shape)))]

    @reg("ndims")
    def _ndims(interp, args, nargout):
        return [mat(float(args[0].ndim))]

    @reg("zeros")
    def _zeros(interp, args, nargout):
        return [mat(np.zeros(_dims_from_args(args) if args else (1, 1), order="F"))]

    @reg("ones")
    def _ones(interp, args, nargout):
        return [mat(np.ones(_dims_from_args(args) if args else (1, 1), order="F"))]

    @reg("true")
    def _true(interp, args, nargout):
        return [mat(np.ones(_dims_from_args(args) if args else (1, 1), dtype=np.bool_, order="F"))]

    @reg("false")
    def _false(interp, args, nargout):
        return [mat(np.zeros(_dims_from_args(args) if args else (1, 1), dtype=np.bool_, order="F"))]

    @reg("eye")
    def _eye(interp, args, nargout):
        d = _dims_from_args(args) if args else (1, 1)
        return [mat(np.eye(d[0], d[1] if len(d) > 1 else d[0]))]

    @reg("randn")
    def _randn(interp, args, nargout):
        raise MlabError("randn: MATLAB's generator is not reproduced; shadow it with Interp(overrides={'randn': ...})")

    B["rand"] = B["randn"]
    B["randperm"] = B["randn"]

    @reg("reshape")
    def _reshape(interp, args, nargout):
        a = args[0]
        d = _dims_from_args(args[1:])
        if int(np.prod(d)) != a.size:
            raise MlabError(f"Number of elements must not change: reshape {a.shape} -> {d}")
        return [mat(a.reshape(d, order="F"))]

    @reg("permute")
    def _permute(interp, args, nargout):
        a = args[0]
        order = [int(x) - 1 for x in args[1].reshape(-1, order="F")]
        if sorted(order) != list(range(len(order))) or len(order) < a.ndim:
            raise MlabError("permute: ORDER must be a permutation of 1:n with n >= ndims")
        a2 = a.reshape(tuple(a.shape) + (1,) * (len(order) - a.ndim), order="F")
        return [mat(np.transpose(a2, order))]

    @reg("squeeze")
    def _squeeze(interp, args, nargout):
        a = args[0]
        if a.ndim <= 2:
            return [a]
        shp = [d for d in a.shape if d != 1]
        while len(shp) < 2:
            shp.append(1)
        return [mat(a.reshape(shp, order="F"))]

    @reg("norm")
    def _norm(interp, args, nargout):
        a = num(args[0])
        if len(args) > 1:
            p = args[1]
            if isinstance(p, str):
                if p == "fro":
                    return [mat(np.sqrt(np.sum(a * a)))]
                raise MlabError("norm: unsupported norm type")
            pv = scalar(p)
            if 1 in a.shape and a.ndim == 2:
                v = a.reshape(-1)
                if np.isinf(pv):
                    return [mat(np.max(np.abs(v)))]
                return [mat(np.linalg.norm(v, pv))]
            raise MlabError("norm: matrix p-norm not supported")
        if a.ndim != 2:
            raise MlabError("norm: input must be 2-D")
        if 1 in a.shape or a.size == 0:
            return [mat(np.linalg.norm(a.reshape(-1)))]
        return [mat(np.linalg.norm(a, 2))]

    @reg("pinv")
    def _pinv(interp, args, nargout):
        return [matlab_pinv(args[0])]

    @reg("kron")
    def _kron(interp, args, nargout):
        return [mat(np.kron(num(args[0]), num(args[1])))]

    @reg("abs")
    def _abs(interp, args, nargout):
        return [mat(np.abs(num(args[0])))]

    @reg("sign")
    def _sign(interp, args, nargout):
        return [mat(np.sign(num(args[0])))]

    @reg("sqrt")
    def _sqrt(interp, args, nargout):
        return [mat(np.sqrt(num(args[0])))]

    @reg("round")
    def _round(interp, args, nargout):
        a = num(args[0])
        return [mat(np.sign(a) * np.floor(np.abs(a) + 0.5))]

    @reg("floor")
    def _floor(interp, args, nargout):
        return [mat(np.floor(num(args[0])))]

    @reg("ceil")
    def _ceil(interp, args, nargout):
        return [mat(np.ceil(num(args[0])))]

    @reg("double")
    def _double(interp, args, nargout):
        return [mat(num(args[0]))]

    @reg("mod")
    def _mod(interp, args, nargout):
        x, y = broadcast2(num(args[0]), num(args[1]))
        with np.errstate(divide="ignore", invalid="ignore"):
            r = np.where(y == 0, x, x - np.floor(x / y) * y)
        return [mat(r)]

    B["max"] = _minmax(np.max, np.maximum)
    B["min"] = _minmax(np.min, np.minimum)

    @reg("sum")
    def _sum(interp, args, nargout):
        a = num(args[0])
        if len(args) > 1:
            if isinstance(args[1], str):
                if args[1] == "all":
                    return [mat(np.sum(a))]
                raise MlabError("sum: unsupported option")
            ax = int(scalar(args[1])) - 1
            if ax >= a.ndim:
                return [a]
        else:
            ax = next((i for i, d in enumerate(a.shape) if d != 1), 0)
        return [mat(np.sum(a, axis=ax, keepdims=True))]

    @reg("isempty")
    def _isempty(interp, args, nargout):
        a = args[0]
        return [mat((len(a) if isinstance(a, (str, dict)) else a.size) == 0)]

    @reg("fprintf")
    def _fprintf(interp, args, nargout):
        if args and not isinstance(args[0], str):
            args = args[1:]              # file id
        s = _format(args[0], args[1:])
        interp.out.append(s)
        if interp.echo:
            print(s, end="")
        return []

    @reg("sprintf")
    def _sprintf(interp, args, nargout):
        return [_format(args[0], args[1:])]

    @reg("disp")
    def _disp(interp, args, nargout):
        s = (args[0] if isinstance(args[0], str) else str(args[0])) + "\n"
        interp.out.append(s)
        if interp.echo:
            print(s, end="")
        return []

    @reg("error")
    def _error(interp, args, nargout):
        raise MlabError(_format(args[0], args[1:]) if args else "error")

    @reg("tic")
    def _tic(interp, args, nargout):
        return []

    @reg("toc")
    def _toc(interp, args, nargout):
        return [mat(0.0)]

    return B


def run_function(path_dirs, name, args, nargout=1, overrides=None, entry_file=None):
    """Execute function `name` (found on `path_dirs`, or the main / a local function of `entry_file`) of the reference.
    -> (outputs, printed text, interpreter)."""
    it = Interp(path=path_dirs, overrides=overrides)
    unit = it.load(entry_file) if entry_file else None
    outs = it.call(name, args, nargout=nargout, unit=unit)
    return outs, "".join(it.out), it
