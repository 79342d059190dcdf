"""minilua -- a small tree-walking interpreter for the subset of Lua 5.1 / LuaJIT that the reference's
`cpu-raw.lua` is written in. TEST INFRASTRUCTURE ONLY (like everything under oracle/).

Why it exists: the reference is LuaJIT source and this image has no Lua runtime at all (probed: luajit, lua5.x,
liblua*, lupa), so the reference cannot be executed here by its own interpreter. This module executes the
reference's UNMODIFIED source text (`/root/reference/cpu-raw.lua`) so that the C oracle (`mg_oracle.c`) can be
pinned against what the reference's own statements compute, instead of only against hand-derived known answers.
`oracle/run_reference.py` is the driver that produces the committed fixtures `tests/golden/ref_*.npz`.

Fidelity argument (what makes the numbers equal to LuaJIT's on x86-64):
  * Lua numbers are IEEE doubles; every arithmetic operator the reference uses (+ - * /, unary minus, comparisons,
    math.sqrt, math.floor) is a single correctly rounded IEEE operation in LuaJIT (SSE2, no x87 excess precision,
    no FMA contraction on x86-64) and in CPython alike, evaluated in the same left-to-right order by the same
    operator precedences (implemented below from the Lua 5.1 manual, section 2.5.6).
  * FFI `float[?]` / `double[?]` buffers (the `image` library's `.buffer`) convert on store with round-to-nearest
    and widen exactly on load: numpy arrays of the same element type do the same.
  * `ffi.copy` is a byte copy; `bit.lshift` on the small non-negative integers used is a plain shift.
What is NOT reproduced: LuaJIT's tracing JIT (semantically transparent), `print` formatting (values are captured
as numbers, not as text), and the un-vendored libraries `ext`, `image`, whose few used entry points are shimmed in
`run_reference.py` from their call sites in cpu-raw.lua (`class()`, `image(w,h,ch,real).buffer`, `math.round`,
`math.isfinite`, `math.fabs`).

Supported: local/global variables, assignments (multiple), functions and closures, method definitions and calls
(`function T:m()`, `obj:m()`), varargs, numeric and generic `for`, `while`, `repeat`, `if/elseif/else`, `do`,
`return`, `break`, table constructors, all Lua 5.1 operators, the metamethods `__index` (table or function), `__call`,
`__add __sub __mul __div __mod __pow __unm __concat` (as in Lua 5.1, `#` ignores `__len` on tables), strings with the
usual escapes, long strings/comments. Not supported: coroutines, goto, string
library patterns, integer division, bitwise operators (Lua 5.3).
"""
import math
import re

__all__ = ["LuaError", "LuaExit", "LuaTable", "LuaFunction", "Interpreter", "lua_call", "lua_index", "lua_setindex", "lua_truth"]


class LuaError(Exception):
    pass


class LuaExit(Exception):
    """os.exit(code)"""

    def __init__(self, code):
        super().__init__(code)
        self.code = code


class _Break(Exception):
    pass


class _Return(Exception):
    def __init__(self, values):
        self.values = values


def _key(k):
    """Table-key normalisation: a float with an integral value is the same key as that integer (Lua has one
    number type)."""
    if type(k) is float and k.is_integer():
        return int(k)
    return k


class LuaTable:
    __slots__ = ("hash", "meta")

    def __init__(self):
        self.hash = {}
        self.meta = None

    def get(self, k):
        return self.hash.get(_key(k))

    def set(self, k, v):
        k = _key(k)
        if v is None:
            self.hash.pop(k, None)
        else:
            self.hash[k] = v

    def length(self):
        n = 0
        while (n + 1) in self.hash:
            n += 1
        return n


class LuaFunction:
    __slots__ = ("params", "vararg", "body", "scope", "name")

    def __init__(self, params, vararg, body, scope, name="?"):
        self.params, self.vararg, self.body, self.scope, self.name = params, vararg, body, scope, name


class Scope:
    __slots__ = ("vars", "parent", "varargs")

    def __init__(self, parent=None, varargs=None):
        self.vars = {}
        self.parent = parent
        self.varargs = varargs if varargs is not None else (parent.varargs if parent is not None else [])

    def lookup(self, name):
        s = self
        while s is not None:
            if name in s.vars:
                return s
            s = s.parent
        return None


def lua_truth(v):
    return v is not None and v is not False


def lua_index(obj, k):
    if isinstance(obj, LuaTable):
        v = obj.get(k)
        if v is None and obj.meta is not None:
            h = obj.meta.get("__index")
            if h is not None:
                if isinstance(h, (LuaFunction,)) or callable(h):
                    r = lua_call(h, [obj, k])
                    return r[0] if r else None
                return lua_index(h, k)
        return v
    if hasattr(obj, "lua_index"):
        return obj.lua_index(k)
    if isinstance(obj, str):
        return STRING_LIB.get(k)     # s:method(...) resolves through the string library, as in Lua
    raise LuaError(f"attempt to index a {type(obj).__name__} value with key {k!r}")


def lua_setindex(obj, k, v):
    if isinstance(obj, LuaTable):
        obj.set(k, v)
    elif hasattr(obj, "lua_setindex"):
        obj.lua_setindex(k, v)
    else:
        raise LuaError(f"attempt to index a {type(obj).__name__} value (assignment to key {k!r})")


def lua_call(f, args):
    """Calls a Lua function, a Python callable (which receives the arguments positionally and may return None, a
    single value or a tuple/list of values), or a table with a `__call` metamethod. Returns a list."""
    if isinstance(f, LuaFunction):
        sc = Scope(f.scope, varargs=args[len(f.params):] if f.vararg else [])
        v = sc.vars
        n = len(args)
        for i, p in enumerate(f.params):
            v[p] = args[i] if i < n else None
        try:
            f.body.execute(sc)
        except _Return as r:
            return r.values
        return []
    if isinstance(f, LuaTable):
        h = f.meta.get("__call") if f.meta is not None else None
        if h is None:
            raise LuaError("attempt to call a table value")
        return lua_call(h, [f] + list(args))
    if callable(f):
        r = f(*args)
        if r is None:
            return []
        if isinstance(r, (tuple, list)):
            return list(r)
        return [r]
    raise LuaError(f"attempt to call a {type(f).__name__} value")


# ----------------------------------------------------------------------------- string library (the few functions used)
_LUA_CLASS = {"a": "A-Za-z", "d": "0-9", "l": "a-z", "u": "A-Z", "w": "A-Za-z0-9", "s": r" \t\n\r\f\v", "x": "0-9A-Fa-f",
              "p": r"!-/:-@\[-`{-~"}


def lua_pattern_to_regex(pat):
    """Translates the common part of Lua patterns (%a %d %l %u %w %s %x %p and their complements, sets, anchors,
    * + - ?, captures, %-escapes) to a Python regular expression. %b and %f are not supported."""
    out, i, n = [], 0, len(pat)

    def cls(c, in_set):
        lo = c.lower()
        if lo in _LUA_CLASS:
            body = _LUA_CLASS[lo]
            if c.isupper():
                if in_set:
                    raise LuaError("complemented class inside a set is not supported by minilua")
                return "[^" + body + "]"
            return body if in_set else "[" + body + "]"
        return re.escape(c)

    while i < n:
        c = pat[i]
        if c == "%":
            i += 1
            if i >= n:
                raise LuaError("malformed pattern (ends with '%')")
            if pat[i] in "bf":
                raise LuaError("%b / %f patterns are not supported by minilua")
            out.append(cls(pat[i], False))
        elif c == "[":
            j = i + 1
            body = "["
            if j < n and pat[j] == "^":
                body += "^"
                j += 1
            first = True
            while j < n and (pat[j] != "]" or first):
                first = False
                if pat[j] == "%":
                    j += 1
                    body += cls(pat[j], True)
                else:
                    body += "\\" + pat[j] if pat[j] in "\\[]" else pat[j]
                j += 1
            out.append(body + "]")
            i = j
        elif c == "-":
            out.append("*?")
        elif c in "*+?()":
            out.append(c)
        elif c == "^" and i == 0:
            out.append("^")
        elif c == "$" and i == n - 1:
            out.append("$")
        elif c == ".":
            out.append("(?s:.)")
        else:
            out.append(re.escape(c))
        i += 1
    return "".join(out)


def _str_match(s, pat, init=1):
    m = re.compile(lua_pattern_to_regex(pat)).search(s, int(init) - 1)
    if m is None:
        return (None,)
    return tuple(m.groups()) if m.groups() else (m.group(0),)


def _str_gmatch(s, pat):
    """string.gmatch for patterns without anchors: an iterator over the matches (captures, or the whole match)"""
    rx = re.compile(lua_pattern_to_regex(pat))
    it = rx.finditer(s)

    def step(*_):
        for m in it:
            return tuple(m.groups()) if m.groups() else (m.group(0),)
        return None
    return (step, None, None)


def _str_find(s, pat, init=1, plain=None):
    if lua_truth(plain):
        i = s.find(pat, int(init) - 1)
        return (None,) if i < 0 else (float(i + 1), float(i + len(pat)))
    m = re.compile(lua_pattern_to_regex(pat)).search(s, int(init) - 1)
    if m is None:
        return (None,)
    return (float(m.start() + 1), float(m.end())) + tuple(m.groups())


def _str_format(fmt, *args):
    return fmt % tuple(int(a) if isinstance(a, float) and a.is_integer() and re.search(r"%[-+ #0]*\d*d", fmt) else a for a in args)


class _StringLib:
    def __init__(self):
        self.fns = {"lower": lambda s, *_: s.lower(), "upper": lambda s, *_: s.upper(), "len": lambda s: float(len(s)),
                    "sub": lambda s, i=1, j=-1: s[(int(i) - 1 if i > 0 else max(len(s) + int(i), 0)):(int(j) if j >= 0 else len(s) + int(j) + 1)],
                    "rep": lambda s, n: s * int(n), "match": _str_match, "find": _str_find, "format": _str_format,
                    "byte": lambda s, i=1: float(ord(s[int(i) - 1])), "char": lambda *a: "".join(chr(int(x)) for x in a),
                    "gmatch": _str_gmatch}

    def get(self, k):
        return self.fns.get(k)

    def as_table(self):
        t = LuaTable()
        for k, v in self.fns.items():
            t.set(k, v)
        return t


STRING_LIB = _StringLib()


# ----------------------------------------------------------------------------- lexer
_TOKEN = re.compile(r"""
    (?P<ws>\s+)
  | (?P<lcomment>--\[(?P<lceq>=*)\[)
  | (?P<comment>--[^\n]*)
  | (?P<lstring>\[(?P<lseq>=*)\[)
  | (?P<number>0[xX][0-9a-fA-F]+ | (?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?)
  | (?P<name>[A-Za-z_][A-Za-z_0-9]*)
  | (?P<string>"(?:\\.|[^"\\\n])*"|'(?:\\.|[^'\\\n])*')
  | (?P<op>\.\.\.|\.\.|==|~=|<=|>=|[-+*/%^\#<>=(){}\[\];:,.])
""", re.X)
_KEYWORDS = {"and", "break", "do", "else", "elseif", "end", "false", "for", "function", "if", "in", "local", "nil", "not",
             "or", "repeat", "return", "then", "true", "until", "while"}
_ESC = {"n": "\n", "t": "\t", "r": "\r", "\\": "\\", '"': '"', "'": "'", "a": "\a", "b": "\b", "f": "\f", "v": "\v", "\n": "\n"}


def tokenize(src):
    toks, pos, line = [], 0, 1
    if src.startswith("#"):  # shebang
        pos = src.index("\n") if "\n" in src else len(src)
    while pos < len(src):
        m = _TOKEN.match(src, pos)
        if not m:
            raise LuaError(f"line {line}: unexpected character {src[pos]!r}")
        kind = m.lastgroup
        text = m.group(kind)
        if kind in ("lcomment", "lstring"):
            eq = m.group("lceq") if kind == "lcomment" else m.group("lseq")
            close = "]" + eq + "]"
            end = src.find(close, m.end())
            if end < 0:
                raise LuaError(f"line {line}: unfinished long {'comment' if kind == 'lcomment' else 'string'}")
            body = src[m.end():end]
            if kind == "lstring":
                toks.append(("string", body[1:] if body.startswith("\n") else body, line))
            line += src.count("\n", pos, end + len(close))
            pos = end + len(close)
            continue
        line += text.count("\n")
        pos = m.end()
        if kind in ("ws", "comment"):
            continue
        if kind == "number":
            toks.append(("number", float(int(text, 16)) if text[:2].lower() == "0x" else float(text), line))
        elif kind == "name":
            toks.append(("kw" if text in _KEYWORDS else "name", text, line))
        elif kind == "string":
            toks.append(("string", re.sub(r"\\(\d{1,3}|.|\n)", lambda e: chr(int(e.group(1))) if e.group(1).isdigit()
                                          else _ESC.get(e.group(1), e.group(1)), text[1:-1]), line))
        else:
            toks.append(("op", text, line))
    toks.append(("eof", None, line))
    return toks


# ----------------------------------------------------------------------------- AST + evaluation
class Node:
    __slots__ = ("line",)


class Block(Node):
    __slots__ = ("stats",)

    def __init__(self, stats):
        self.stats = stats

    def execute(self, scope):
        for s in self.stats:
            s.execute(scope)


class Num(Node):
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = v

    def eval(self, scope):
        return self.v


Str = Num  # a literal is a literal


class Const(Node):
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = v

    def eval(self, scope):
        return self.v


class Vararg(Node):
    __slots__ = ()

    def eval(self, scope):
        va = scope.varargs
        return va[0] if va else None

    def eval_multi(self, scope):
        return list(scope.varargs)


class Name(Node):
    __slots__ = ("name",)

    def __init__(self, name):
        self.name = name

    def eval(self, scope):
        s = scope
        n = self.name
        while s is not None:
            v = s.vars
            if n in v:
                return v[n]
            s = s.parent
        return None

    def assign(self, scope, value):
        s = scope.lookup(self.name)
        if s is None:  # global
            s = scope
            while s.parent is not None:
                s = s.parent
        s.vars[self.name] = value


class Index(Node):
    __slots__ = ("obj", "key")

    def __init__(self, obj, key):
        self.obj, self.key = obj, key

    def eval(self, scope):
        o = self.obj.eval(scope)
        if o is None:
            raise LuaError(f"line {self.line}: attempt to index a nil value")
        return lua_index(o, self.key.eval(scope))

    def assign(self, scope, value):
        o = self.obj.eval(scope)
        if o is None:
            raise LuaError(f"line {self.line}: attempt to index a nil value")
        lua_setindex(o, self.key.eval(scope), value)


def _eval_list(exprs, scope):
    """explist semantics: every expression yields one value, except that a call or `...` in LAST position expands."""
    out = []
    last = len(exprs) - 1
    for i, e in enumerate(exprs):
        if i == last and hasattr(e, "eval_multi"):
            out.extend(e.eval_multi(scope))
        else:
            out.append(e.eval(scope))
    return out


class Call(Node):
    __slots__ = ("fn", "args")

    def __init__(self, fn, args):
        self.fn, self.args = fn, args

    def eval_multi(self, scope):
        f = self.fn.eval(scope)
        if f is None:
            raise LuaError(f"line {self.line}: attempt to call a nil value")
        return lua_call(f, _eval_list(self.args, scope))

    def eval(self, scope):
        r = self.eval_multi(scope)
        return r[0] if r else None

    execute = eval_multi


class MethodCall(Node):
    __slots__ = ("obj", "name", "args")

    def __init__(self, obj, name, args):
        self.obj, self.name, self.args = obj, name, args

    def eval_multi(self, scope):
        o = self.obj.eval(scope)
        if o is None:
            raise LuaError(f"line {self.line}: attempt to index a nil value (method {self.name})")
        f = lua_index(o, self.name)
        if f is None:
            raise LuaError(f"line {self.line}: attempt to call method '{self.name}' (a nil value)")
        return lua_call(f, [o] + _eval_list(self.args, scope))

    def eval(self, scope):
        r = self.eval_multi(scope)
        return r[0] if r else None

    execute = eval_multi


class Paren(Node):  # (f()) truncates to one value
    __slots__ = ("e",)

    def __init__(self, e):
        self.e = e

    def eval(self, scope):
        return self.e.eval(scope)


class Function(Node):
    __slots__ = ("params", "vararg", "body", "name")

    def __init__(self, params, vararg, body, name="?"):
        self.params, self.vararg, self.body, self.name = params, vararg, body, name

    def eval(self, scope):
        return LuaFunction(self.params, self.vararg, self.body, scope, self.name)


class TableCons(Node):
    __slots__ = ("array", "fields")

    def __init__(self, array, fields):
        self.array, self.fields = array, fields

    def eval(self, scope):
        t = LuaTable()
        for i, v in enumerate(_eval_list(self.array, scope)):
            t.set(i + 1, v)
        for k, v in self.fields:
            t.set(k.eval(scope), v.eval(scope))
        return t


_ARITH_EVENT = {"+": "__add", "-": "__sub", "*": "__mul", "/": "__div", "%": "__mod", "^": "__pow", "..": "__concat"}


def _metamethod(a, b, event):
    """Lua 5.1 manual 2.8: the handler of a binary event is looked up in the first operand, then in the second."""
    for v in (a, b):
        if isinstance(v, LuaTable) and v.meta is not None:
            h = v.meta.get(event)
            if h is not None:
                return h
    return None


def _arith_operand(v, line):
    if isinstance(v, bool) or not isinstance(v, (int, float)):
        if isinstance(v, str):
            try:
                return float(v)
            except ValueError:
                pass
        raise LuaError(f"line {line}: attempt to perform arithmetic on a {'nil' if v is None else type(v).__name__} value")
    return v


def _lua_div(a, b):
    a, b = float(a), float(b)
    if b == 0.0:
        if a == 0.0 or a != a:
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)
    return a / b


def _lua_mod(a, b):
    a, b = float(a), float(b)
    if b == 0.0:
        return math.nan
    return a - math.floor(a / b) * b


def _lua_pow(a, b):
    try:
        return math.pow(a, b)
    except (OverflowError, ValueError):
        return math.inf if a > 0 else math.nan


_ARITH = {"+": lambda a, b: a + b, "-": lambda a, b: a - b, "*": lambda a, b: a * b, "/": _lua_div, "%": _lua_mod, "^": _lua_pow}


def lua_tostring(v):
    if v is None:
        return "nil"
    if v is True:
        return "true"
    if v is False:
        return "false"
    if isinstance(v, (int, float)):
        return "%.14g" % v
    return str(v)


class BinOp(Node):
    __slots__ = ("op", "a", "b", "fn")

    def __init__(self, op, a, b):
        self.op, self.a, self.b = op, a, b
        self.fn = _ARITH.get(op)

    def eval(self, scope):
        op = self.op
        if op == "and":
            a = self.a.eval(scope)
            return self.b.eval(scope) if (a is not None and a is not False) else a
        if op == "or":
            a = self.a.eval(scope)
            return a if (a is not None and a is not False) else self.b.eval(scope)
        a = self.a.eval(scope)
        b = self.b.eval(scope)
        fn = self.fn
        if fn is not None:
            ta, tb = type(a), type(b)
            if (ta is float or ta is int) and (tb is float or tb is int):
                return fn(a, b)
            h = _metamethod(a, b, _ARITH_EVENT[op])
            if h is not None:
                r = lua_call(h, [a, b])
                return r[0] if r else None
            return fn(_arith_operand(a, self.line), _arith_operand(b, self.line))
        if op == "==":
            return a == b if not (isinstance(a, bool) ^ isinstance(b, bool)) else False
        if op == "~=":
            return not (a == b if not (isinstance(a, bool) ^ isinstance(b, bool)) else False)
        if op == "..":
            if not isinstance(a, (str, int, float)) or not isinstance(b, (str, int, float)):
                h = _metamethod(a, b, "__concat")
                if h is not None:
                    r = lua_call(h, [a, b])
                    return r[0] if r else None
            return lua_tostring(a) + lua_tostring(b)
        if (isinstance(a, (int, float)) and isinstance(b, (int, float))) or (isinstance(a, str) and isinstance(b, str)):
            if op == "<":
                return a < b
            if op == "<=":
                return a <= b
            if op == ">":
                return a > b
            return a >= b
        raise LuaError(f"line {self.line}: attempt to compare {type(a).__name__} with {type(b).__name__}")


class UnOp(Node):
    __slots__ = ("op", "e")

    def __init__(self, op, e):
        self.op, self.e = op, e

    def eval(self, scope):
        v = self.e.eval(scope)
        if self.op == "-":
            if isinstance(v, LuaTable) and v.meta is not None and v.meta.get("__unm") is not None:
                r = lua_call(v.meta.get("__unm"), [v, v])
                return r[0] if r else None
            return -_arith_operand(v, self.line)
        if self.op == "not":
            return not lua_truth(v)
        if isinstance(v, LuaTable):
            return v.length()
        if isinstance(v, str):
            return len(v)
        raise LuaError(f"line {self.line}: attempt to get length of a {type(v).__name__} value")


class Local(Node):
    __slots__ = ("names", "exprs")

    def __init__(self, names, exprs):
        self.names, self.exprs = names, exprs

    def execute(self, scope):
        vals = _eval_list(self.exprs, scope)
        n = len(vals)
        for i, name in enumerate(self.names):
            scope.vars[name] = vals[i] if i < n else None


class LocalFunction(Node):
    __slots__ = ("name", "func")

    def __init__(self, name, func):
        self.name, self.func = name, func

    def execute(self, scope):
        scope.vars[self.name] = None
        scope.vars[self.name] = self.func.eval(scope)


class Assign(Node):
    __slots__ = ("targets", "exprs")

    def __init__(self, targets, exprs):
        self.targets, self.exprs = targets, exprs

    def execute(self, scope):
        vals = _eval_list(self.exprs, scope)
        n = len(vals)
        for i, t in enumerate(self.targets):
            t.assign(scope, vals[i] if i < n else None)


class Do(Node):
    __slots__ = ("body",)

    def __init__(self, body):
        self.body = body

    def execute(self, scope):
        self.body.execute(Scope(scope))


class While(Node):
    __slots__ = ("cond", "body")

    def __init__(self, cond, body):
        self.cond, self.body = cond, body

    def execute(self, scope):
        try:
            while lua_truth(self.cond.eval(scope)):
                self.body.execute(Scope(scope))
        except _Break:
            pass


class Repeat(Node):
    __slots__ = ("body", "cond")

    def __init__(self, body, cond):
        self.body, self.cond = body, cond

    def execute(self, scope):
        try:
            while True:
                inner = Scope(scope)
                self.body.execute(inner)
                if lua_truth(self.cond.eval(inner)):
                    break
        except _Break:
            pass


class If(Node):
    __slots__ = ("clauses", "orelse")

    def __init__(self, clauses, orelse):
        self.clauses, self.orelse = clauses, orelse

    def execute(self, scope):
        for cond, body in self.clauses:
            c = cond.eval(scope)
            if c is not None and c is not False:
                body.execute(Scope(scope))
                return
        if self.orelse is not None:
            self.orelse.execute(Scope(scope))


class NumFor(Node):
    __slots__ = ("var", "start", "stop", "step", "body")

    def __init__(self, var, start, stop, step, body):
        self.var, self.start, self.stop, self.step, self.body = var, start, stop, step, body

    def execute(self, scope):
        i = _arith_operand(self.start.eval(scope), self.line)
        stop = _arith_operand(self.stop.eval(scope), self.line)
        step = _arith_operand(self.step.eval(scope), self.line) if self.step is not None else 1
        if step == 0:
            raise LuaError(f"line {self.line}: 'for' step is zero")
        body, var = self.body, self.var
        try:
            while (i <= stop) if step > 0 else (i >= stop):
                inner = Scope(scope)
                inner.vars[var] = i
                body.execute(inner)
                i = i + step
        except _Break:
            pass


class GenFor(Node):
    __slots__ = ("names", "exprs", "body")

    def __init__(self, names, exprs, body):
        self.names, self.exprs, self.body = names, exprs, body

    def execute(self, scope):
        vals = _eval_list(self.exprs, scope) + [None, None, None]
        f, s, ctl = vals[0], vals[1], vals[2]
        try:
            while True:
                rets = lua_call(f, [s, ctl])
                if not rets or rets[0] is None:
                    break
                ctl = rets[0]
                inner = Scope(scope)
                for i, n in enumerate(self.names):
                    inner.vars[n] = rets[i] if i < len(rets) else None
                self.body.execute(inner)
        except _Break:
            pass


class Return(Node):
    __slots__ = ("exprs",)

    def __init__(self, exprs):
        self.exprs = exprs

    def execute(self, scope):
        raise _Return(_eval_list(self.exprs, scope))


class Break(Node):
    __slots__ = ()

    def execute(self, scope):
        raise _Break()


# ----------------------------------------------------------------------------- parser
# (left, right) binding powers, Lua 5.1 manual 2.5.6; `..` and `^` are right associative
_BINPRI = {"or": (1, 1), "and": (2, 2), "<": (3, 3), ">": (3, 3), "<=": (3, 3), ">=": (3, 3), "~=": (3, 3), "==": (3, 3),
           "..": (5, 4), "+": (6, 6), "-": (6, 6), "*": (7, 7), "/": (7, 7), "%": (7, 7), "^": (10, 9)}
_UNARY_PRI = 8


class Parser:
    def __init__(self, src, chunkname="?"):
        self.toks = tokenize(src)
        self.i = 0
        self.chunk = chunkname

    # token helpers
    def peek(self):
        return self.toks[self.i]

    def next(self):
        t = self.toks[self.i]
        self.i += 1
        return t

    def check(self, kind, val=None):
        t = self.toks[self.i]
        return t[0] == kind and (val is None or t[1] == val)

    def accept(self, kind, val=None):
        if self.check(kind, val):
            return self.next()
        return None

    def expect(self, kind, val=None):
        t = self.next()
        if t[0] != kind or (val is not None and t[1] != val):
            raise LuaError(f"{self.chunk}:{t[2]}: expected {val or kind}, got {t[1]!r}")
        return t

    def mk(self, node, line):
        node.line = line
        return node

    # grammar
    def parse_chunk(self):
        b = self.block()
        self.expect("eof")
        return b

    def block(self):
        stats = []
        while True:
            t = self.peek()
            if t[0] == "eof" or (t[0] == "kw" and t[1] in ("end", "else", "elseif", "until")):
                break
            if t[0] == "kw" and t[1] == "return":
                self.next()
                exprs = []
                t2 = self.peek()
                if not (t2[0] == "eof" or (t2[0] == "kw" and t2[1] in ("end", "else", "elseif", "until")) or (t2[0] == "op" and t2[1] == ";")):
                    exprs = self.explist()
                self.accept("op", ";")
                stats.append(self.mk(Return(exprs), t[2]))
                break
            if t[0] == "kw" and t[1] == "break":
                self.next()
                self.accept("op", ";")
                stats.append(self.mk(Break(), t[2]))
                break
            stats.append(self.statement())
            self.accept("op", ";")
        return self.mk(Block(stats), 0)

    def statement(self):
        t = self.peek()
        line = t[2]
        if t[0] == "kw":
            kw = t[1]
            if kw == "if":
                self.next()
                clauses = []
                cond = self.expr()
                self.expect("kw", "then")
                clauses.append((cond, self.block()))
                orelse = None
                while True:
                    if self.accept("kw", "elseif"):
                        c = self.expr()
                        self.expect("kw", "then")
                        clauses.append((c, self.block()))
                    elif self.accept("kw", "else"):
                        orelse = self.block()
                        self.expect("kw", "end")
                        break
                    else:
                        self.expect("kw", "end")
                        break
                return self.mk(If(clauses, orelse), line)
            if kw == "while":
                self.next()
                c = self.expr()
                self.expect("kw", "do")
                b = self.block()
                self.expect("kw", "end")
                return self.mk(While(c, b), line)
            if kw == "do":
                self.next()
                b = self.block()
                self.expect("kw", "end")
                return self.mk(Do(b), line)
            if kw == "for":
                self.next()
                n1 = self.expect("name")[1]
                if self.accept("op", "="):
                    a = self.expr()
                    self.expect("op", ",")
                    b = self.expr()
                    c = self.expr() if self.accept("op", ",") else None
                    self.expect("kw", "do")
                    body = self.block()
                    self.expect("kw", "end")
                    return self.mk(NumFor(n1, a, b, c, body), line)
                names = [n1]
                while self.accept("op", ","):
                    names.append(self.expect("name")[1])
                self.expect("kw", "in")
                exprs = self.explist()
                self.expect("kw", "do")
                body = self.block()
                self.expect("kw", "end")
                return self.mk(GenFor(names, exprs, body), line)
            if kw == "repeat":
                self.next()
                b = self.block()
                self.expect("kw", "until")
                return self.mk(Repeat(b, self.expr()), line)
            if kw == "function":
                self.next()
                n = self.expect("name")
                target = self.mk(Name(n[1]), n[2])
                fname = n[1]
                is_method = False
                while True:
                    if self.accept("op", "."):
                        k = self.expect("name")
                        target = self.mk(Index(target, self.mk(Str(k[1]), k[2])), k[2])
                        fname += "." + k[1]
                    elif self.accept("op", ":"):
                        k = self.expect("name")
                        target = self.mk(Index(target, self.mk(Str(k[1]), k[2])), k[2])
                        fname += ":" + k[1]
                        is_method = True
                        break
                    else:
                        break
                f = self.funcbody(line, fname, is_method)
                return self.mk(Assign([target], [f]), line)
            if kw == "local":
                self.next()
                if self.accept("kw", "function"):
                    n = self.expect("name")[1]
                    return self.mk(LocalFunction(n, self.funcbody(line, n, False)), line)
                names = [self.expect("name")[1]]
                while self.accept("op", ","):
                    names.append(self.expect("name")[1])
                exprs = self.explist() if self.accept("op", "=") else []
                return self.mk(Local(names, exprs), line)
            raise LuaError(f"{self.chunk}:{line}: unexpected keyword {kw!r}")
        # exprstat: call or assignment
        e = self.suffixedexp()
        if self.check("op", "=") or self.check("op", ","):
            targets = [e]
            while self.accept("op", ","):
                targets.append(self.suffixedexp())
            self.expect("op", "=")
            exprs = self.explist()
            for tg in targets:
                if not isinstance(tg, (Name, Index)):
                    raise LuaError(f"{self.chunk}:{line}: cannot assign to this expression")
            return self.mk(Assign(targets, exprs), line)
        if not isinstance(e, (Call, MethodCall)):
            raise LuaError(f"{self.chunk}:{line}: syntax error (expression is not a statement)")
        return e

    def funcbody(self, line, name, is_method):
        self.expect("op", "(")
        params, vararg = (["self"] if is_method else []), False
        if not self.check("op", ")"):
            while True:
                if self.accept("op", "..."):
                    vararg = True
                    break
                params.append(self.expect("name")[1])
                if not self.accept("op", ","):
                    break
        self.expect("op", ")")
        body = self.block()
        self.expect("kw", "end")
        return self.mk(Function(params, vararg, body, name), line)

    def explist(self):
        out = [self.expr()]
        while self.accept("op", ","):
            out.append(self.expr())
        return out

    def primaryexp(self):
        t = self.next()
        if t[0] == "name":
            return self.mk(Name(t[1]), t[2])
        if t[0] == "op" and t[1] == "(":
            e = self.expr()
            self.expect("op", ")")
            return self.mk(Paren(e), t[2])
        raise LuaError(f"{self.chunk}:{t[2]}: unexpected symbol {t[1]!r}")

    def suffixedexp(self):
        e = self.primaryexp()
        while True:
            t = self.peek()
            if t[0] == "op" and t[1] == ".":
                self.next()
                k = self.expect("name")
                e = self.mk(Index(e, self.mk(Str(k[1]), k[2])), t[2])
            elif t[0] == "op" and t[1] == "[":
                self.next()
                k = self.expr()
                self.expect("op", "]")
                e = self.mk(Index(e, k), t[2])
            elif t[0] == "op" and t[1] == ":":
                self.next()
                n = self.expect("name")[1]
                e = self.mk(MethodCall(e, n, self.callargs()), t[2])
            elif (t[0] == "op" and t[1] in ("(", "{")) or t[0] == "string":
                e = self.mk(Call(e, self.callargs()), t[2])
            else:
                return e

    def callargs(self):
        t = self.peek()
        if t[0] == "string":
            self.next()
            return [self.mk(Str(t[1]), t[2])]
        if t[0] == "op" and t[1] == "{":
            return [self.tablecons()]
        self.expect("op", "(")
        args = []
        if not self.check("op", ")"):
            args = self.explist()
        self.expect("op", ")")
        return args

    def tablecons(self):
        line = self.expect("op", "{")[2]
        array, fields = [], []
        while not self.check("op", "}"):
            if self.check("op", "["):
                self.next()
                k = self.expr()
                self.expect("op", "]")
                self.expect("op", "=")
                fields.append((k, self.expr()))
            elif self.check("name") and self.toks[self.i + 1][0] == "op" and self.toks[self.i + 1][1] == "=":
                k = self.next()
                self.next()
                fields.append((self.mk(Str(k[1]), k[2]), self.expr()))
            else:
                array.append(self.expr())
            if not (self.accept("op", ",") or self.accept("op", ";")):
                break
        self.expect("op", "}")
        return self.mk(TableCons(array, fields), line)

    def simpleexp(self):
        t = self.peek()
        if t[0] == "number":
            self.next()
            return self.mk(Num(t[1]), t[2])
        if t[0] == "string":
            self.next()
            return self.mk(Str(t[1]), t[2])
        if t[0] == "kw":
            if t[1] == "nil":
                self.next()
                return self.mk(Const(None), t[2])
            if t[1] == "true":
                self.next()
                return self.mk(Const(True), t[2])
            if t[1] == "false":
                self.next()
                return self.mk(Const(False), t[2])
            if t[1] == "function":
                self.next()
                return self.funcbody(t[2], "anonymous", False)
        if t[0] == "op" and t[1] == "...":
            self.next()
            return self.mk(Vararg(), t[2])
        if t[0] == "op" and t[1] == "{":
            return self.tablecons()
        return self.suffixedexp()

    def expr(self, limit=0):
        t = self.peek()
        if (t[0] == "kw" and t[1] == "not") or (t[0] == "op" and t[1] in ("-", "#")):
            self.next()
            e = self.mk(UnOp(t[1], self.expr(_UNARY_PRI)), t[2])
        else:
            e = self.simpleexp()
        while True:
            t = self.peek()
            op = t[1] if t[0] in ("op", "kw") else None
            pri = _BINPRI.get(op)
            if pri is None or pri[0] <= limit:
                return e
            self.next()
            rhs = self.expr(pri[1])
            e = self.mk(BinOp(op, e, rhs), t[2])


# ----------------------------------------------------------------------------- interpreter + base library
class Interpreter:
    """Global environment with the part of the base library the reference touches. `modules` maps `require` names to
    values (LuaTable, Python callable, ...); a missing module raises, like Lua's require."""

    def __init__(self, modules=None, stdout=None, search_dirs=()):
        self.globals = Scope()
        self.modules = dict(modules or {})
        self.search_dirs = list(search_dirs)   # `require 'a.b'` also tries <dir>/a/b.lua (after package.path's ?.lua entries)
        self.loaded = {}
        self.printed = []      # every print(...) call as a tuple of raw values
        self.written = []      # io.write arguments
        self._stdout = stdout
        g = self.globals.vars
        g["print"] = self._print
        g["error"] = self._error
        g["assert"] = self._assert
        g["require"] = self._require
        g["pcall"] = self._pcall
        self.package = self.table_from({"path": "./?.lua"})
        g["package"] = self.package
        g["type"] = self._type
        g["tostring"] = lua_tostring
        g["tonumber"] = self._tonumber
        g["ipairs"] = self._ipairs
        g["pairs"] = self._pairs
        g["next"] = self._next
        g["select"] = self._select
        g["unpack"] = self._unpack
        g["rawget"] = lambda t, k: (t.get(k),)
        g["rawset"] = lambda t, k, v: (t.set(k, v), t)[1]
        g["setmetatable"] = self._setmetatable
        g["getmetatable"] = lambda t: t.meta if isinstance(t, LuaTable) else None
        g["math"] = self.table_from(self.math_functions())
        io = LuaTable()
        io.set("write", self._io_write)
        io.set("open", self._io_open)
        g["io"] = io
        # the few `table` functions the shipped driver scripts (lua/test/*.lua) use
        def _t_insert(t, a, b=None):
            if b is None:
                t.set(t.length() + 1, a)
            else:
                n = t.length()
                for k in range(n, int(a) - 1, -1):
                    t.set(k + 1, t.get(k))
                t.set(int(a), b)
        def _t_concat(t, sep="", i=1, j=None):
            j = t.length() if j is None else int(j)
            return (sep.join(lua_tostring(t.get(k)) for k in range(int(i), j + 1)),)
        g["table"] = self.table_from({"insert": _t_insert, "concat": _t_concat,
                                      "unpack": lambda t, i=1, j=None: self._unpack(t, i, j)})
        g["string"] = STRING_LIB.as_table()
        import os as _os
        import time as _time
        def _exit(code=0):
            raise LuaExit(int(code or 0))
        def _execute(cmd=None):
            import subprocess as _sp
            return (float(_sp.call(cmd, shell=True)),) if cmd is not None else (True,)
        g["os"] = self.table_from({"getenv": lambda k: (_os.environ.get(k),), "clock": lambda: _time.process_time(),
                                   "execute": _execute,
                                   "time": lambda: float(int(_time.time())), "exit": _exit})
        # LuaJIT loads its `bit` library as a global as well (gpu.lua:257 uses it without a require)
        g["bit"] = self.table_from({"lshift": lambda a, n: float(int(a) << int(n)), "rshift": lambda a, n: float(int(a) >> int(n)),
                                    "band": lambda a, b: float(int(a) & int(b)), "bor": lambda a, b: float(int(a) | int(b))})
        g["_G"] = None

    # --- helpers for embedding
    @staticmethod
    def table_from(d):
        t = LuaTable()
        for k, v in d.items():
            t.set(k, v)
        return t

    @staticmethod
    def math_functions():
        def log(x, base=None):
            if x == 0:
                r = -math.inf
            elif x < 0 or x != x:
                return math.nan
            else:
                r = math.log(x)
            if base is None:
                return r
            return math.log2(x) if base == 2 else (math.log10(x) if base == 10 else r / math.log(base))
        return {"floor": lambda x: float(math.floor(x)), "ceil": lambda x: float(math.ceil(x)),
                "tointeger": lambda x: float(int(x)),
                "sqrt": lambda x: math.sqrt(x) if x >= 0 else math.nan, "abs": lambda x: abs(x), "huge": math.inf, "pi": math.pi,
                "log": log, "exp": math.exp, "sin": math.sin, "cos": math.cos,
                "max": lambda *a: max(a), "min": lambda *a: min(a), "pow": _lua_pow, "fmod": math.fmod}

    def run(self, src, chunkname="chunk", varargs=None):
        block = Parser(src, chunkname).parse_chunk()
        sc = Scope(self.globals, varargs=list(varargs or []))
        try:
            block.execute(sc)
        except _Return as r:
            return r.values
        return []

    def run_file(self, path):
        with open(path) as fh:
            return self.run(fh.read(), path)

    # --- base library
    def _print(self, *args):
        self.printed.append(tuple(args))
        if self._stdout is not None:
            self._stdout.write("\t".join(lua_tostring(a) for a in args) + "\n")

    def _io_write(self, *args):
        self.written.extend(args)
        if self._stdout is not None:
            self._stdout.write("".join(lua_tostring(a) for a in args))

    def _io_open(self, name, mode="r"):
        """io.open for the driver scripts: a file object with :write, :read('*a'), :lines-free, :flush, :close"""
        try:
            fh = open(name, mode.replace("b", ""))
        except OSError as e:
            return (None, str(e))
        f = LuaTable()
        f.set("write", lambda self_, *a: (fh.write("".join(lua_tostring(x) for x in a)), self_)[1])
        f.set("read", lambda self_, fmt="*l": (fh.read() if str(fmt).lstrip("*").startswith("a") else (fh.readline().rstrip("\n") or None),))
        f.set("flush", lambda self_: fh.flush())
        f.set("close", lambda self_: fh.close())
        return (f,)

    def _error(self, msg=None, level=None):
        raise LuaError(lua_tostring(msg))

    def _assert(self, v=None, msg="assertion failed!", *rest):
        if not lua_truth(v):
            raise LuaError(lua_tostring(msg))
        return (v, msg) + rest

    def _require(self, name):
        if name in self.modules:
            return self.modules[name]
        if name in self.loaded:
            return self.loaded[name]
        import os as _os
        rel = name.replace(".", _os.sep)
        cands = [t.replace("?", rel) for t in str(self.package.get("path") or "").split(";") if t]
        cands += [_os.path.join(d, rel + ".lua") for d in self.search_dirs]
        if "." in name:   # LuaJIT-style flat layouts: 'multigrid-poisson.cpu-raw' -> <dir>/cpu-raw.lua
            cands += [_os.path.join(d, name.split(".")[-1] + ".lua") for d in self.search_dirs]
        for c in cands:
            if _os.path.isfile(c):
                r = self.run_file(c)
                self.loaded[name] = r[0] if r else True
                return self.loaded[name]
        raise LuaError(f"module '{name}' not found")

    def _pcall(self, f=None, *args):
        try:
            return (True,) + tuple(lua_call(f, list(args)))
        except LuaError as e:
            return (False, str(e))

    def _type(self, v):
        if v is None:
            return "nil"
        if isinstance(v, bool):
            return "boolean"
        if isinstance(v, (int, float)):
            return "number"
        if isinstance(v, str):
            return "string"
        if isinstance(v, LuaTable):
            return "table"
        if isinstance(v, LuaFunction) or callable(v):
            return "function"
        return "cdata"

    def _tonumber(self, v, base=None):
        if isinstance(v, (int, float)) and not isinstance(v, bool):
            return (v,)
        try:
            return (float(int(v, int(base))) if base is not None else float(v),)
        except (TypeError, ValueError):
            return (None,)

    def _ipairs(self, t):
        def it(tt, i):
            i = int(i) + 1
            v = lua_index(tt, i)
            if v is None:
                return None
            return (float(i), v)
        return (it, t, 0.0)

    def _next(self, t, k=None):
        keys = list(t.hash.keys())
        if k is None:
            idx = 0
        else:
            idx = keys.index(_key(k)) + 1
        if idx >= len(keys):
            return None
        kk = keys[idx]
        return (float(kk) if isinstance(kk, int) else kk, t.hash[kk])

    def _pairs(self, t):
        return (self._next, t, None)

    def _select(self, n, *args):
        if n == "#":
            return float(len(args))
        return tuple(args[int(n) - 1:])

    def _unpack(self, t, i=1, j=None):
        j = t.length() if j is None else int(j)
        return tuple(t.get(k) for k in range(int(i), j + 1))

    def _setmetatable(self, t, m):
        t.meta = m
        return t
