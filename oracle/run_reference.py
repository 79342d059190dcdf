"""Executes the reference's own `cpu-raw.lua` (UNMODIFIED, read from /root/reference) under oracle/minilua.py and
writes what it computes to tests/golden/ref_*.npz. TEST INFRASTRUCTURE ONLY.

    python oracle/run_reference.py            # regenerates every fixture (needs /root/reference; ~1 min)
    python oracle/run_reference.py 8 double   # one case, prints the err lines like the reference's print(iter, err)

This is the strongest pin of the C oracle available in this image: there is no Lua runtime here, so the reference
source is run by a purpose-built interpreter (see the fidelity argument at the top of minilua.py). The libraries the
reference `require`s are not vendored in its repository; the handful of entry points cpu-raw.lua uses are shimmed
below, each from its call site:

  require 'ffi'        ffi.copy(dst, src, nbytes), ffi.sizeof(ctype)              cpu-raw.lua:182,246
  require 'bit'        bit.lshift / bit.rshift                                    cpu-raw.lua:60-61,66-67,160
  require 'image'      image(w, h, channels, ctype) -> object with .buffer = zero-initialised `ctype[w*h*channels]`
                       (LuaJIT's ffi.new zero-fills)                              cpu-raw.lua:148-164
  require 'ext.class'  class() -> table; calling it builds an instance whose missing keys fall back to the class and
                       runs :init(...)                                            cpu-raw.lua:118,142
  require 'ext.math'   Lua's math plus round (floor(x + .5)), isfinite, fabs      cpu-raw.lua:10,90,136,159,254,256

The `show` method (cpu-raw.lua:126-140, a debug dump that prints only when `debugging` is set) is replaced by a
recorder and `debugging` is switched on, so that every buffer the reference would dump -- at every stage of twoGrid, in
the reference's own call order -- is captured as numbers instead of text. Nothing else of the class is touched.
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import minilua as ml  # noqa: E402

REFERENCE = os.environ.get("MG_REFERENCE_DIR", "/root/reference")
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
_CTYPES = {"double": np.float64, "float": np.float32}


class CArray:
    """An FFI array `ctype[n]`: loads widen to a Lua number (double), stores convert with round-to-nearest."""

    def __init__(self, n, ctype):
        self.a = np.zeros(int(n), dtype=_CTYPES[ctype])
        self.ctype = ctype

    def lua_index(self, k):
        i = int(k)
        if i != k or not 0 <= i < self.a.size:
            raise ml.LuaError(f"cdata index {k!r} outside [0, {self.a.size}) (LuaJIT would read out of bounds)")
        return float(self.a[i])

    def lua_setindex(self, k, v):
        i = int(k)
        if i != k or not 0 <= i < self.a.size:
            raise ml.LuaError(f"cdata index {k!r} outside [0, {self.a.size}) (LuaJIT would write out of bounds)")
        self.a[i] = v


def _ffi():
    def copy(dst, src, nbytes):
        n = int(nbytes)
        dst.a.view(np.uint8)[:n] = src.a.view(np.uint8)[:n]

    return ml.Interpreter.table_from({"copy": copy, "sizeof": lambda ct: float(np.dtype(_CTYPES[ct]).itemsize),
                                      "new": lambda ct, n=1: CArray(n, ct.split("[")[0])})


def _bit():
    return ml.Interpreter.table_from({"lshift": lambda a, n: float(int(a) << int(n)), "rshift": lambda a, n: float(int(a) >> int(n)),
                                      "band": lambda a, b: float(int(a) & int(b)), "bor": lambda a, b: float(int(a) | int(b))})


def _image(w, h, ch, fmt="double"):
    t = ml.LuaTable()
    t.set("width", w)
    t.set("height", h)
    t.set("channels", ch)
    t.set("format", fmt)
    t.set("buffer", CArray(int(w) * int(h) * int(ch), fmt))
    return t


def _class(*parents):
    cls = ml.LuaTable()
    meta = ml.LuaTable()

    def construct(c, *args):
        obj = ml.LuaTable()
        om = ml.LuaTable()
        om.set("__index", c)
        obj.meta = om
        init = ml.lua_index(c, "init")
        if init is not None:
            ml.lua_call(init, [obj] + list(args))
        return obj

    meta.set("__call", construct)
    if parents:
        meta.set("__index", parents[0])
    cls.meta = meta
    return cls


def _ext_math():
    d = ml.Interpreter.math_functions()
    d.update({"round": lambda x: float(math.floor(x + .5)), "isfinite": lambda x: math.isfinite(x), "fabs": lambda x: abs(x)})
    return ml.Interpreter.table_from(d)


def load_reference_class(interp=None):
    """Runs cpu-raw.lua and returns (interpreter, the MultigridCPURaw class table it returns)."""
    it = interp or ml.Interpreter(modules={"ffi": _ffi(), "bit": _bit(), "image": _image, "ext.class": _class, "ext.math": _ext_math()})
    (cls,) = it.run_file(os.path.join(REFERENCE, "cpu-raw.lua"))
    return it, cls


def run_reference(size, real="double"):
    """`MultigridCPURaw(size, real):run()` as the reference defines it (2 V-cycles, cpu-raw.lua:245). Returns a dict:
    errs (what it prints per iteration), f, psi, psiOld, per-level rs/Rs/vs/Vs, and `trace` = every show(name, buffer, L)
    call in order as (name, L, copy of the first L*L elements)."""
    it, cls = load_reference_class()
    trace = []

    def show(self, name, im, L, *_):
        n = int(L) * int(L)
        trace.append((name, int(L), im.a[:n].copy()))

    cls.set("show", show)
    cls.set("debugging", True)   # cpu-raw.lua:121,199-203: also dump f before every top-level pre-smoothing sweep
    obj = ml.lua_call(cls, [float(size), real])[0]
    f0 = obj.get("f").get("buffer").a.copy()
    psi0 = obj.get("psi").get("buffer").a.copy()
    ml.lua_call(ml.lua_index(obj, "run"), [obj])
    errs = [p[1] for p in it.printed if len(p) == 2 and isinstance(p[0], (int, float)) and not isinstance(p[0], bool)]
    out = {"errs": np.array(errs, dtype=np.float64), "f": obj.get("f").get("buffer").a.copy(), "f0": f0, "psi0": psi0,
           "psi": obj.get("psi").get("buffer").a.copy(), "psiOld": obj.get("psiOld").get("buffer").a.copy(), "trace": trace}
    L = 1
    while L <= size:
        for nm in ("rs", "Rs", "vs", "Vs"):
            out[f"{nm}{L}"] = obj.get(nm).get(L).get("buffer").a.copy()
        L *= 2
    return out


CASES = [  # (size, real, keep the full stage-by-stage trace?)
    (1, "double", True), (2, "double", True), (4, "double", True), (8, "double", True), (16, "double", True), (8, "float", True), (16, "float", True),
    (32, "double", False), (32, "float", False),   # test/test.lua:45 runs log2size = 5
    (64, "double", False), (64, "float", False),   # BASELINE config 1
]


def save_case(size, real, keep_trace):
    r = run_reference(size, real)
    d = {k: v for k, v in r.items() if k != "trace"}
    tr = r["trace"]
    # the top-level `u` dumps: after each smoothing sweep and after the correction (cpu-raw.lua:205,229,233)
    tops = [a for (n, L, a) in tr if n == "u" and L == size]
    d["psi_after_cycle1"] = tops[len(tops) // 2 - 1] if size > 1 else tops[0]
    d["trace_names"] = np.array([n for (n, L, a) in tr])
    d["trace_L"] = np.array([L for (n, L, a) in tr], dtype=np.int32)
    if keep_trace:
        for i, (n, L, a) in enumerate(tr):
            d[f"t{i:05d}"] = a
    path = os.path.join(GOLDEN, f"ref_2d_{size}_{'f64' if real == 'double' else 'f32'}.npz")
    np.savez_compressed(path, meta=np.array([2, size, 0 if real == "double" else 2, 2]), **d)
    print(f"{os.path.basename(path)}: {len(tr)} dumps, err = {[float(e) for e in r['errs']]}")


def main():
    if len(sys.argv) >= 2:
        r = run_reference(int(sys.argv[1]), sys.argv[2] if len(sys.argv) > 2 else "double")
        print("#iter\terr")
        for i, e in enumerate(r["errs"]):
            print(f"{i + 1}\t{e:.14g}")
        return
    for c in CASES:
        save_case(*c)


if __name__ == "__main__":
    main()
