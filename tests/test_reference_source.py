"""Pins the C oracle to the REFERENCE'S OWN SOURCE.

`tests/golden/ref_2d_*.npz` hold what `/root/reference/cpu-raw.lua` -- unmodified -- computes when it is executed by
oracle/minilua.py (generator: oracle/run_reference.py; there is no Lua runtime in this image). This file checks

  1. the C oracle (oracle/mg_oracle.c) against those fixtures, BIT FOR BIT: the err value of both iterations of
     run() (cpu-raw.lua:245-256), f, psi, psiOld, every level's rs / Rs / vs / Vs, and the complete stage-by-stage
     sequence of `show` dumps of twoGrid (cpu-raw.lua:186-237) -- same names, same level order, same numbers;
  2. the interpreter itself on small programs whose results follow from the Lua 5.1 manual (precedence,
     short-circuit values, varargs, closures, numeric for, method calls, metatables, float stores);
  3. when the reference tree is present (this container, not the GPU box): that re-running the reference source
     reproduces the committed fixture.

cpu-raw.lua's real = 'float' stores fp32 and computes in Lua numbers (doubles): the oracle's "float_acc64".

`tests/golden/refgpu_2d_*.npz` are the same for `gpu.lua` (generator: oracle/run_reference_gpu.py): its Lua host code
runs under minilua on a fake in-memory OpenCL device whose kernels are the reference's own OpenCL C source compiled by
gcc with -ffp-contract=off. real = 'float' there means fp32 storage AND fp32 arithmetic: the oracle's "float", the
arithmetic of the headline benchmark. gpu.lua never dumps the top-level f before the pre-smoothing sweeps (its test
`L==size` reads an undefined global, gpu.lua:268), so those seven dumps per cycle are dropped from the oracle's trace.
"""
import glob
import io
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402
import minilua as ml  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "ref_2d_*.npz"))) + sorted(glob.glob(os.path.join(GOLDEN, "refgpu_2d_*.npz")))


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint64 if a.dtype == np.float64 else np.uint32)


def _same(got, want, what):
    got, want = np.asarray(got).ravel(), np.asarray(want).ravel()
    assert got.dtype == want.dtype and got.shape == want.shape, (what, got.dtype, want.dtype, got.shape, want.shape)
    neq = _bits(got) != _bits(want)
    assert not neq.any(), f"{what}: {int(neq.sum())} of {neq.size} values differ bitwise, first at {int(np.argmax(neq))}: " \
                          f"oracle {got[np.argmax(neq)]!r} reference {want[np.argmax(neq)]!r}"


def test_fixtures_present():
    names = {os.path.basename(f) for f in FIXTURES}
    assert {"ref_2d_64_f64.npz", "ref_2d_64_f32.npz", "ref_2d_32_f64.npz", "ref_2d_8_f64.npz", "ref_2d_2_f64.npz",
            "refgpu_2d_64_f32.npz", "refgpu_2d_128_f32.npz", "refgpu_2d_8_f32.npz", "refgpu_2d_32_f64.npz"} <= names


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(f)[:-4] for f in FIXTURES])
def test_oracle_equals_reference_source(path):
    g = np.load(path)
    dim, size, real_kind, cycles = (int(x) for x in g["meta"])
    assert dim == 2 and cycles == 2
    real = {0: "double", 1: "float", 2: "float_acc64"}[real_kind]
    from_gpu_lua = os.path.basename(path).startswith("refgpu_")
    o = O.Oracle(size, real, dim)
    _same(o.f, g["f0"], "f after init (initCells, cpu-raw.lua:8-20)")
    _same(o.psi, g["psi0"], "psi after init")
    o.trace_enable(True)
    errs, psi1 = [], None
    for it in range(cycles):
        errs.append(o.step())
        if it == 0:
            psi1 = o.psi.copy()
    # err: both sides add size^2 squared differences sequentially in double -> exactly equal
    assert [float(e) for e in g["errs"]] == errs, (list(g["errs"]), errs)
    _same(psi1, g["psi_after_cycle1"], "psi after the first V-cycle")
    _same(o.psi, g["psi"], "psi after run()")
    _same(o.psiOld, g["psiOld"], "psiOld after run()")
    _same(o.f, g["f"], "f after run()")
    L = 1
    while L <= size:
        for nm, which in (("rs", O.BUF_r), ("Rs", O.BUF_R), ("vs", O.BUF_v), ("Vs", O.BUF_V)):
            _same(o.buffer(which, L), g[f"{nm}{L}"], f"{nm}[{L}] after run()")
        L *= 2
    # the stage-by-stage dumps, in the reference's own order
    tr = o.trace()
    if from_gpu_lua:   # drop the debugging-only dump of f before each of the 7 top-level pre-smoothing sweeps, per cycle
        per = len(tr) // cycles
        tr = [t for i, t in enumerate(tr) if not (i % per < 14 and i % per % 2 == 0)]
    assert [n for n, _, _ in tr] == [str(n) for n in g["trace_names"]], "sequence of dumped buffer names"
    assert [l for _, l, _ in tr] == [int(l) for l in g["trace_L"]], "sequence of dumped levels"
    if "t00000" in g.files:
        for i, (n, l, a) in enumerate(tr):
            _same(a, g[f"t{i:05d}"], f"dump #{i} ({n}, L = {l})")


@pytest.mark.skipif(not os.path.exists("/root/reference/gpu.lua"), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("size,fp64", [(8, False), (8, True)])
def test_rerunning_gpu_lua_reproduces_the_fixture(size, fp64):
    import run_reference_gpu as rg
    r = rg.run_reference_gpu(size, fp64)
    assert r["real"] == ("double" if fp64 else "float")     # gpu.lua:32 picks real from the device's fp64 extension
    g = np.load(os.path.join(GOLDEN, f"refgpu_2d_{size}_{'f64' if fp64 else 'f32'}.npz"))
    assert [float(e) for e in g["errs"]] == [float(e) for e in r["errs"]]
    _same(r["psi"], g["psi"], "psi")
    # the kernel sequence of one twoGrid level visit (gpu.lua:296-346): 7 x Jacobi, calcResidual, reduceResidual, ...
    names = [n for n, _ in r["launches"]]
    assert names[0] == "init" and names[1:8] == ["Jacobi"] * 7 and names[8:10] == ["calcResidual", "reduceResidual"]
    assert names.count("calcFrobErr") == 2 and names.count("expandResidual") == names.count("addTo") == 2 * 3
    for i, (n, l, a) in enumerate(r["trace"]):
        _same(a, g[f"t{i:05d}"], f"dump #{i}")


@pytest.mark.skipif(not os.path.exists("/root/reference/cpu-raw.lua"), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("size,real", [(8, "double"), (16, "float")])
def test_rerunning_the_reference_source_reproduces_the_fixture(size, real):
    import run_reference as rr
    r = rr.run_reference(size, real)
    g = np.load(os.path.join(GOLDEN, f"ref_2d_{size}_{'f64' if real == 'double' else 'f32'}.npz"))
    assert [float(e) for e in g["errs"]] == [float(e) for e in r["errs"]]
    _same(r["psi"], g["psi"], "psi")
    assert len(r["trace"]) == len(g["trace_names"])
    for i, (n, l, a) in enumerate(r["trace"]):
        _same(a, g[f"t{i:05d}"], f"dump #{i}")


CPU_FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "refcpu_2d_*.npz")))


@pytest.mark.parametrize("path", CPU_FIXTURES, ids=[os.path.basename(f)[:-4] for f in CPU_FIXTURES])
def test_oracle_with_zeroed_corrections_equals_cpu_lua(path):
    """cpu.lua (the solver test/converge-multigrid-vs-krylov.lua drives; generator oracle/run_reference_cpu.py) starts
    every cycle's coarse corrections from zero (cpu.lua:138); otherwise it is cpu-raw.lua's arithmetic in tables of
    tables. Oracle + Vs[*] = 0 before each step (what mg_zero_corrections does) must give the same fields, bit for
    bit, cycle after cycle, and the same sequence of stage dumps; err within 1e-12 (the un-vendored matrix library's
    summation order in normSq is not known, the shim adds row by row)."""
    g = np.load(path)
    dim, size, real_kind, steps = (int(x) for x in g["meta"])
    assert len(CPU_FIXTURES) >= 5 and dim == 2 and real_kind == 0
    o = O.Oracle(size, "double", dim)
    _same(o.f, g["f0"], "f after init (matrix.lambda, cpu.lua:184-194)")
    _same(o.psi, g["psi0"], "psi after init (-f, cpu.lua:197)")
    o.trace_enable(True)
    for c in range(steps):
        L = size // 2
        while L >= 1:
            o.buffer(O.BUF_V, L)[...] = 0
            L //= 2
        e = o.step()
        assert abs(e - float(g["errs"][c])) <= 1e-12 * float(g["errs"][c]), (c, e, float(g["errs"][c]))
        _same(o.psi, g[f"psi{c + 1}"], f"psi after step {c + 1}")
        if c == 0:
            tr = o.trace()
            n1 = int(g["trace_len"][0])
            assert [n for n, _, _ in tr] == [str(n) for n in g["trace_names"][:n1]], "sequence of dumped buffer names"
            assert [l for _, l, _ in tr] == [int(l) for l in g["trace_L"][:n1]], "sequence of dumped levels"
            if "t00000" in g.files:
                for i, (n, l, a) in enumerate(tr):
                    _same(a, g[f"t{i:05d}"], f"dump #{i} ({n}, L = {l})")


@pytest.mark.skipif(not os.path.exists("/root/reference/cpu.lua"), reason="reference tree not present (GPU box)")
def test_rerunning_cpu_lua_reproduces_the_fixture():
    import run_reference_cpu as rc
    r = rc.run_reference_cpu(8, 3)
    g = np.load(os.path.join(GOLDEN, "refcpu_2d_8_f64.npz"))
    assert [float(e) for e in g["errs"]] == [float(e) for e in r["errs"]]
    for c in range(3):
        _same(r["psis"][c], g[f"psi{c + 1}"], f"psi after step {c + 1}")


def test_minilua_arithmetic_metamethods():
    src = """
    local mt = {}
    local function V(x) return setmetatable({v = x}, mt) end
    local function val(a) return type(a) == 'table' and a.v or a end
    mt.__add = function(a, b) return V(val(a) + val(b)) end
    mt.__sub = function(a, b) return V(val(a) - val(b)) end
    mt.__div = function(a, b) return V(val(a) / val(b)) end
    mt.__unm = function(a) return V(-a.v) end
    local x = V(3)
    return (x + 4).v, (5 + x).v, (x - {v = 1}).v, (-x).v, (x / 2).v, #setmetatable({1, 2, 3}, {__len = function() return 99 end})
    """
    r, _, _ = _run(src)
    assert r == [7.0, 8.0, 2.0, -3.0, 1.5, 3]       # Lua 5.1: # ignores __len on tables


# ------------------------------------------------------------------ the interpreter on programs with known results
def _run(src, **mods):
    out = io.StringIO()
    it = ml.Interpreter(modules=mods, stdout=out)
    return it.run(src), it, out.getvalue()


def test_minilua_operator_precedence_and_associativity():
    r, _, _ = _run("return 1 + 2 * 3 - 4 / 2, 2 ^ 3 ^ 2, -2 ^ 2, 7 % 3, -7 % 3, 1 .. 2 .. 3, not nil == true, 1 < 2 == true, 10 - 4 - 3")
    assert r == [5.0, 512.0, -4.0, 1.0, 2.0, "123", True, True, 3.0]


def test_minilua_and_or_return_operands_not_booleans():
    # the idiom of cpu-raw.lua:23-26:  i > 0 and u[...] or 0
    r, _, _ = _run("local u = {[0]=5} return 1 > 0 and u[0] or 0, 0 > 0 and u[0] or 0, nil or false, false and nil, 0 and 'zero is true'")
    assert r == [5.0, 0.0, False, False, "zero is true"]
    # the right operand must not be evaluated when the left decides (an out-of-range index would raise)
    r, _, _ = _run("local t = nil return false and t.x or 7")
    assert r == [7.0]


def test_minilua_varargs_and_multiple_results():
    src = """
    local function call2D(w, h, kernel, ...)      -- the shape of cpu-raw.lua:108-114
        local acc = 0
        for j = 0, h - 1 do for i = 0, w - 1 do acc = acc + kernel(w, h, i, j, ...) end end
        return acc
    end
    local function k(w, h, i, j, a, b) return (i + w * j) * a + b end
    local function mr() return 1, 2, 3 end
    local t = {mr(), mr()}
    return call2D(3, 2, k, 10, 1), #t, (mr()), select('#', mr()), select(2, mr())
    """
    r, _, _ = _run(src)
    assert r == [156.0, 4, 1.0, 3.0, 2.0, 3.0]


def test_minilua_numeric_for_and_closures():
    src = """
    local s, n = 0, 0
    for i = 10, 1, -3 do s = s + i n = n + 1 end        -- 10, 7, 4, 1
    for i = 1, 0 do s = s + 1000 end                      -- never runs
    local fs = {}
    for i = 1, 3 do fs[i] = function() return i end end   -- a fresh variable per iteration
    local function counter() local c = 0 return function() c = c + 1 return c end end
    local c1, c2 = counter(), counter()
    c1() c1()
    return s, n, fs[1]() + fs[2]() + fs[3](), c1(), c2()
    """
    r, _, _ = _run(src)
    assert r == [22.0, 4.0, 6.0, 3.0, 1.0]


def test_minilua_methods_metatables_and_class_shim():
    import run_reference as rr
    src = """
    local class = require 'ext.class'
    local A = class()
    A.smooth = 7
    function A:init(x) self.x = x end
    function A:get() return self.x + self.smooth end
    local a, b = A(1), A(10)
    b.smooth = 100                                         -- shadows the class field for b only
    local t = setmetatable({}, {__index = function(t, k) return k .. '!' end})
    return a:get(), b:get(), A.smooth, t.hello, rawget(t, 'hello')
    """
    r, _, _ = _run(src, **{"ext.class": rr._class})
    assert r == [8.0, 110.0, 7.0, "hello!", None]


def test_minilua_comments_and_number_literals():
    src = """
    --[[ a long comment
    with function end end
    --]]
    -- [[ a line comment that looks like the one at cpu-raw.lua:180
    local x = .25 + 4. + 1e+6 + 0x10     --]]
    return x, 1e-10, 'a\\tb', [[long
string]]
    """
    r, _, _ = _run(src)
    assert r == [1000020.25, 1e-10, "a\tb", "long\nstring"]


def test_minilua_ffi_float_store_rounds_and_copy_is_bytewise():
    import run_reference as rr
    src = """
    local ffi = require 'ffi'
    local image = require 'image'
    local a, b = image(2, 2, 1, 'float'), image(2, 2, 1, 'float')
    a.buffer[0] = 0.1                      -- stored as fp32, read back widened
    a.buffer[3] = 1/3
    ffi.copy(b.buffer, a.buffer, 4 * ffi.sizeof('float'))
    return a.buffer[0], b.buffer[3], b.buffer[1], ffi.sizeof('double')
    """
    r, _, _ = _run(src, ffi=rr._ffi(), image=rr._image)
    assert r == [float(np.float32(0.1)), float(np.float32(1 / 3)), 0.0, 8.0]


def test_minilua_errors_are_reported_not_swallowed():
    with pytest.raises(ml.LuaError):
        _run("local t = nil return t.x")
    with pytest.raises(ml.LuaError):
        _run("error('found a nan')")
    with pytest.raises(ml.LuaError):
        _run("local x = 1 +")
    with pytest.raises(ml.LuaError):
        _run("return require 'no.such.module'")


@pytest.mark.skipif(not os.path.exists("/root/reference/cpu-raw.lua"), reason="reference tree not present (GPU box)")
def test_luajit_timing_hook_runs_and_skips_cleanly():
    """baseline/run_cpu_raw.lua is meant for a machine that HAS LuaJIT; here it is at least executed by minilua: with
    the dependency shims it runs the reference and prints its timing line, without them it prints SKIP and exits 0."""
    import run_reference as rr
    hook = os.path.join(ROOT, "baseline", "run_cpu_raw.lua")
    out = io.StringIO()
    it = ml.Interpreter(modules={"ffi": rr._ffi(), "bit": rr._bit(), "image": rr._image, "ext.class": rr._class,
                                 "ext.math": rr._ext_math()}, stdout=out)
    it.globals.vars["arg"] = ml.Interpreter.table_from({1: "8", 2: "double", 3: "/root/reference"})
    it.run_file(hook)
    lines = out.getvalue().splitlines()
    assert lines[-1].startswith("cpu-raw.lua 8 double seconds_for_run ")
    assert "6828.1698388918" in out.getvalue()                 # the reference's own print(iter, err) of cycle 2
    out2 = io.StringIO()
    it2 = ml.Interpreter(modules={"ffi": rr._ffi()}, stdout=out2)
    with pytest.raises(ml.LuaExit) as e:
        it2.run_file(hook)
    assert e.value.code == 0 and out2.getvalue().startswith("SKIP: ")
