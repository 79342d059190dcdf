"""Executes the reference's `cpu.lua` (UNMODIFIED; the table-of-tables solver that test/converge-multigrid-vs-krylov.lua
drives) under oracle/minilua.py and writes what it computes to tests/golden/refcpu_*.npz. TEST INFRASTRUCTURE ONLY.

    python oracle/run_reference_cpu.py            # regenerates every fixture (needs /root/reference)
    python oracle/run_reference_cpu.py 8 3        # one case: size, number of step() calls

cpu.lua is the variant whose coarse corrections start from ZERO in every cycle (`local V = matrix.zeros(...)`,
cpu.lua:138), whereas cpu-raw.lua / gpu.lua keep Vs[L] from the previous cycle: the library offers it as
mg_zero_corrections(), and BASELINE config 5 (multigrid vs Krylov) runs it. Everything else -- neighbour order, the
rounding sequence, restriction order, injection -- is the same arithmetic in another container, so the C oracle with its
V buffers zeroed before each step must reproduce these fields bit for bit.

The un-vendored `matrix` library (thenumbernine/lua-matrix) is shimmed for exactly what cpu.lua touches:
  matrix(t)             deep copy of a (nested) table / matrix                        cpu.lua:42,180,199,200
  matrix{a, b}          a vector                                                      cpu.lua:178,185
  matrix.zeros(w, h)    w x h zeros                                                   cpu.lua:111,127,138,142
  matrix.lambda(sz, f)  sz[1] x sz[2] matrix of f(i, j)                               cpu.lua:184-194
  -m, m - n, m / s      element-wise (n may be a plain table, s a number)             cpu.lua:182,192,197,203
  m:normSq(), m:normLInf(), m:size(), v:prod()                                        cpu.lua:192,203
`m:normSq()` adds the squares row by row (first index outermost); the real library's order is not known here, so the
`err` values of these fixtures are compared with a tolerance, the fields exactly.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import minilua as ml  # noqa: E402
import run_reference as rr  # noqa: E402

REFERENCE = rr.REFERENCE
GOLDEN = rr.GOLDEN


def make_matrix_module():
    meta = ml.LuaTable()
    methods = ml.LuaTable()

    def is_tab(v):
        return isinstance(v, ml.LuaTable)

    def new(values):
        """values: nested python lists of numbers -> nested matrices"""
        t = ml.LuaTable()
        t.meta = meta
        for i, v in enumerate(values):
            t.set(i + 1, new(v) if isinstance(v, list) else v)
        return t

    def to_list(t):
        return [to_list(t.get(i)) if is_tab(t.get(i)) else t.get(i) for i in range(1, t.length() + 1)]

    def ew(a, b, op):
        if isinstance(a, list) and isinstance(b, list):
            if len(a) != len(b):
                raise ml.LuaError("matrix: size mismatch")
            return [ew(x, y, op) for x, y in zip(a, b)]
        if isinstance(a, list):
            return [ew(x, b, op) for x in a]
        if isinstance(b, list):
            return [ew(a, y, op) for y in b]
        return op(a, b)

    def val(v):
        return to_list(v) if is_tab(v) else v

    def flat(l):
        for x in l:
            if isinstance(x, list):
                yield from flat(x)
            else:
                yield x

    def construct(cls, t=None):
        return new(to_list(t)) if is_tab(t) else new([])

    def zeros(*dims):
        def z(d):
            return [z(d[1:]) if len(d) > 1 else 0.0 for _ in range(int(d[0]))]
        return new(z(dims))

    def lam(size, f):
        dims = [int(x) for x in to_list(size)]

        def build(prefix, d):
            if not d:
                return (ml.lua_call(f, [float(i) for i in prefix]) or [None])[0]
            return [build(prefix + [i + 1], d[1:]) for i in range(d[0])]
        return new(build([], dims))

    def size(m):
        dims, t = [], m
        while is_tab(t):
            dims.append(float(t.length()))
            t = t.get(1)
        return new(dims)

    def norm_sq(m):
        s = 0.0
        for x in flat(to_list(m)):     # row by row, first index outermost
            s = s + x * x
        return s

    def norm_linf(m):
        return max((abs(x) for x in flat(to_list(m))), default=0.0)

    def prod(m):
        p = 1.0
        for x in flat(to_list(m)):
            p = p * x
        return p

    for k, v in {"normSq": norm_sq, "normLInf": norm_linf, "size": size, "prod": prod}.items():
        methods.set(k, v)
    meta.set("__index", methods)
    meta.set("__unm", lambda a, *_: new(ew(val(a), None, lambda x, _y: -x)))
    meta.set("__add", lambda a, b: new(ew(val(a), val(b), lambda x, y: x + y)))
    meta.set("__sub", lambda a, b: new(ew(val(a), val(b), lambda x, y: x - y)))
    meta.set("__mul", lambda a, b: new(ew(val(a), val(b), lambda x, y: x * y)))
    meta.set("__div", lambda a, b: new(ew(val(a), val(b), lambda x, y: x / y)))
    mod = ml.LuaTable()
    mod.set("zeros", zeros)
    mod.set("lambda", lam)
    mm = ml.LuaTable()
    mm.set("__call", construct)
    mod.meta = mm
    return mod, to_list


def run_reference_cpu(size, steps):
    """`MultigridCPU{size=..., maxiter=..., epsilon=...}` then `steps` x `:step()` (cpu.lua:196-206). Arrays are returned
    flattened as index = (i-1) + size*(j-1) for cpu.lua's u[i][j] (its first index plays the role of cpu-raw.lua's i)."""
    matrix, to_list = make_matrix_module()
    it = ml.Interpreter(modules={"ext.class": rr._class, "ext.math": rr._ext_math(), "matrix": matrix})
    (cls,) = it.run_file(os.path.join(REFERENCE, "cpu.lua"))
    trace = []

    def arr(m):
        return np.array(to_list(m), dtype=np.float64).T.ravel()     # [i][j] -> i + L*j

    def show(self, name, m, width, *_):
        trace.append((name, int(width), arr(m)))

    cls.set("show", show)
    args = ml.Interpreter.table_from({"size": float(size), "maxiter": float(steps), "epsilon": 1e-10, "debug": True})
    obj = ml.lua_call(cls, [args])[0]
    out = {"f0": arr(obj.get("f")), "psi0": arr(obj.get("psi")), "errs": [], "psis": [], "trace_len": []}
    for _ in range(steps):
        e = ml.lua_call(ml.lua_index(obj, "step"), [obj])[0]
        out["errs"].append(e)
        out["psis"].append(arr(obj.get("psi")))
        out["trace_len"].append(len(trace))
    out["trace"] = trace
    return out


CASES = [(2, 3, True), (4, 3, True), (8, 3, True), (16, 4, True), (32, 4, False), (64, 3, False)]


def save_case(size, steps, keep_trace):
    r = run_reference_cpu(size, steps)
    d = {"errs": np.array(r["errs"]), "f0": r["f0"], "psi0": r["psi0"], "trace_len": np.array(r["trace_len"], dtype=np.int32)}
    for i, p in enumerate(r["psis"]):
        d[f"psi{i + 1}"] = p
    tr = r["trace"]
    d["trace_names"] = np.array([n for (n, L, a) in tr])
    d["trace_L"] = np.array([L for (n, L, a) in tr], dtype=np.int32)
    if keep_trace:
        for i, (n, L, a) in enumerate(tr[:r["trace_len"][0]]):      # the first cycle's dumps
            d[f"t{i:05d}"] = a
    path = os.path.join(GOLDEN, f"refcpu_2d_{size}_f64.npz")
    np.savez_compressed(path, meta=np.array([2, size, 0, steps]), **d)
    print(f"{os.path.basename(path)}: {len(tr)} dumps, err = {[float(e) for e in r['errs']]}")


def main():
    if len(sys.argv) >= 2:
        r = run_reference_cpu(int(sys.argv[1]), int(sys.argv[2]) if len(sys.argv) > 2 else 2)
        for i, e in enumerate(r["errs"]):
            print(f"{i + 1}\t{e:.14g}")
        return
    for c in CASES:
        save_case(*c)


if __name__ == "__main__":
    main()
