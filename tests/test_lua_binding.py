"""Executes the LuaJIT binding this project ships -- lua/multigrid-poisson/cuda.lua, the class a maintainer of the
reference would `require` (test/test.lua:53-56: `cl(size, real, cpuDepth)`, `multigrid:run()`) -- from its Lua source.

There is no Lua runtime in the image, so the Lua is interpreted by oracle/minilua.py and `require 'ffi'` is served by
oracle/minilua_ffi.py (cdef / load / new / gc / string on top of ctypes): every call below goes Lua -> ffi -> C ABI ->
CUDA, exactly the path LuaJIT would take.

CPU tier: the binding parses, its ffi.cdef block declares nothing the library does not export, and constructing a solver
without a GPU fails loudly with the library's own message. GPU tier: run() / twoGrid() / getbuffer() through the Lua
class reproduce, bit for bit, what the reference's own cpu-raw.lua and gpu.lua compute (tests/golden/ref*_2d_*.npz).
"""
import io
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import minilua as ml  # noqa: E402
import minilua_ffi  # noqa: E402
import run_reference as rr  # noqa: E402

PKG = os.path.join(ROOT, "lua-multigrid-poisson_b200")
CUDA_LUA = os.path.join(PKG, "lua", "multigrid-poisson", "cuda.lua")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_binding():
    ffi = minilua_ffi.FFI(search_dirs=[PKG])
    out = io.StringIO()
    it = ml.Interpreter(modules={"ffi": ffi.module(), "ext.class": rr._class, "ext.math": rr._ext_math()}, stdout=out)
    os.environ.setdefault("MGPOISSON_LIB", os.path.join(PKG, "libmgpoisson.so"))
    (cls,) = it.run_file(CUDA_LUA)
    return it, ffi, cls


def test_binding_loads_and_every_declared_function_resolves():
    it, ffi, cls = load_binding()
    assert isinstance(cls, ml.LuaTable) and ml.lua_index(cls, "run") is not None
    assert ml.lua_index(cls, "smooth") == 7 and ml.lua_index(cls, "accuracy") == 1e-10      # cpu-raw.lua:123-124
    assert {"MG_REAL_F64", "MG_REAL_F32", "MG_REAL_F32_ACC64", "MG_BUF_PSI", "MG_MODE_REFSEQ", "MG_OK"} <= set(ffi.enums)
    assert len(ffi.protos) >= 30
    lib = ffi.libs[0]
    for name in ffi.protos:
        assert callable(lib.lua_index(name)), name          # dlsym + prototype binding
    v = ml.lua_call(lib.lua_index("mg_version"), [])[0]
    assert "sm_100a" in ffi.string(v)


def test_cdef_block_is_the_header_block():
    hdr = open(os.path.join(ROOT, "include", "mgpoisson.h")).read()
    lua = open(CUDA_LUA).read()
    a = hdr[hdr.index("/* MGPOISSON_CDEF_BEGIN */") + len("/* MGPOISSON_CDEF_BEGIN */"):hdr.index("/* MGPOISSON_CDEF_END */")]
    b = lua[lua.index("ffi.cdef[[") + len("ffi.cdef[["):lua.index("]]", lua.index("ffi.cdef[["))]
    assert a.strip() == b.strip()


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present: the constructor succeeds")
def test_without_a_gpu_the_constructor_fails_loudly_in_lua():
    it, ffi, cls = load_binding()
    with pytest.raises(ml.LuaError, match="libmgpoisson: .*(CUDA|device)"):
        ml.lua_call(cls, [64.0, "double"])


CASES = [("ref_2d_64_f64", "double"), ("ref_2d_64_f32", "float_acc64"), ("refgpu_2d_64_f32", "float"), ("ref_2d_16_f64", "double")]


def _bits_equal(got, want, what):
    got, want = np.ascontiguousarray(got).ravel(), np.ascontiguousarray(want).ravel()
    assert got.dtype == want.dtype and got.shape == want.shape, what
    u = np.uint64 if got.dtype == np.float64 else np.uint32
    assert np.array_equal(got.view(u), want.view(u)), what


@pytest.mark.gpu
@pytest.mark.parametrize("fixture,real", CASES)
@pytest.mark.parametrize("how", ["plain", "cpuDepth3", "debugging"])
def test_lua_class_run_matches_the_reference_source_run(fixture, real, how):
    g = np.load(os.path.join(GOLDEN, fixture + ".npz"))
    size = int(g["meta"][1])
    dt = np.float64 if real == "double" else np.float32
    it, ffi, cls = load_binding()
    if how == "debugging":
        cls.set("debugging", True)          # -> MG_MODE_REFSEQ: one kernel per reference operator (cuda.lua, init)
    args = [float(size), real] + ([3.0] if how == "cpuDepth3" else [])
    obj = ml.lua_call(cls, args)[0]          # cl(size, real, cpuDepth)            test/test.lua:54
    n0 = len(it.printed)
    ml.lua_call(ml.lua_index(obj, "run"), [obj])          # multigrid:run()         test/test.lua:56
    errs = [p[1] for p in it.printed[n0:] if len(p) == 2 and isinstance(p[0], float)]
    assert len(errs) == 2
    for e, w in zip(errs, g["errs"]):
        assert abs(e - w) <= (1e-12 + size * size * 2.0 ** -55) * w
    lib = ffi.libs[0]
    psi = ml.lua_call(ml.lua_index(obj, "getbuffer"), [obj, lib.lua_index("MG_BUF_PSI"), float(size)])[0]
    _bits_equal(psi.numpy(dt), g["psi"], "psi after run()")
    old = ml.lua_call(ml.lua_index(obj, "getbuffer"), [obj, lib.lua_index("MG_BUF_PSIOLD"), float(size)])[0]
    _bits_equal(old.numpy(dt), g["psiOld"], "psiOld after run()")
    L = size // 2
    while L >= 1:
        V = ml.lua_call(ml.lua_index(obj, "getbuffer"), [obj, lib.lua_index("MG_BUF_V"), float(L)])[0]
        _bits_equal(V.numpy(dt), g[f"Vs{L}"], f"Vs[{L}]")
        L //= 2
    ffi.close()


@pytest.mark.gpu
def test_lua_class_twogrid_and_smoother_on_device_pointers():
    """obj:twoGrid(h, obj.psi.buffer, obj.f.buffer, size) and obj:inPlaceIterativeSolver(L, u, f, h) take the device
    pointers the class exposes, like gpu.lua's methods take cl buffers (gpu.lua:252-346)."""
    g = np.load(os.path.join(GOLDEN, "ref_2d_16_f64.npz"))
    size = 16
    it, ffi, cls = load_binding()
    src = """
    local cls, size = ...
    local m = cls(size, 'double')
    m:twoGrid(1/size, m.psi.buffer, m.f.buffer, size)
    return m
    """
    (obj,) = it.run(src, "driver", [cls, float(size)])
    lib = ffi.libs[0]
    psi = ml.lua_call(ml.lua_index(obj, "getbuffer"), [obj, lib.lua_index("MG_BUF_PSI"), float(size)])[0]
    _bits_equal(psi.numpy(np.float64), g["psi_after_cycle1"], "psi after one twoGrid")
    # seven sweeps of the smoother on a fresh solver = the reference's 7th top-level `u` dump: with `debugging` the
    # reference dumps f, u for each pre-smoothing sweep (cpu-raw.lua:198-206), so that is dump #13
    (obj2,) = it.run("local cls, size = ... local m = cls(size, 'double') for i = 1, 7 do "
                     "m:inPlaceIterativeSolver(size, m.psi.buffer, m.f.buffer, 1/size) end return m", "driver2", [cls, float(size)])
    psi2 = ml.lua_call(ml.lua_index(obj2, "getbuffer"), [obj2, lib.lua_index("MG_BUF_PSI"), float(size)])[0]
    names, Ls = [str(n) for n in g["trace_names"]], [int(l) for l in g["trace_L"]]
    assert names[12:14] == ["f", "u"] and Ls[13] == size
    _bits_equal(psi2.numpy(np.float64), g["t00013"], "psi after 7 Jacobi sweeps")
    ffi.close()


@pytest.mark.gpu
@pytest.mark.parametrize("size", [16, 32, 64])
def test_lua_class_in_cpu_lua_style_matches_cpu_lua(size):
    """The cpu.lua form of the API, as test/converge-multigrid-vs-krylov.lua:20-30 uses it: a table constructor with an
    errorCallback that records mg.psi:normLInf() per cycle, then :solve(). cpu.lua re-zeroes the coarse corrections
    every cycle (cpu.lua:138); fixtures tests/golden/refcpu_2d_*.npz are cpu.lua itself executed by minilua."""
    g = np.load(os.path.join(GOLDEN, f"refcpu_2d_{size}_f64.npz"))
    steps = int(g["meta"][3])
    it, ffi, cls = load_binding()
    src = """
    local cls, size, steps = ...
    local data = {}
    local mg
    mg = cls{
        size = size,
        maxiter = steps,
        epsilon = 1e-20,
        errorCallback = function(iter, err)
            data[iter] = {err, mg.psi:normLInf()}
        end,
    }
    mg:solve()
    return mg, data
    """
    obj, data = it.run(src, "driver", [cls, float(size), float(steps)])
    lib = ffi.libs[0]
    assert data.length() == steps
    for c in range(steps):
        err, linf = data.get(c + 1).get(1), data.get(c + 1).get(2)
        w = float(g["errs"][c])
        assert abs(err - w) <= (1e-12 + size * size * 2.0 ** -55) * w
        assert linf == float(np.max(np.abs(g[f"psi{c + 1}"])))
    psi = ml.lua_call(ml.lua_index(obj, "getbuffer"), [obj, lib.lua_index("MG_BUF_PSI"), float(size)])[0]
    _bits_equal(psi.numpy(np.float64), g[f"psi{steps}"], f"psi after {steps} cycles of solve()")
    ffi.close()


def _show_text(name, a, L):
    """cpu-raw.lua:126-134: the name, then L rows of L values, each preceded by a blank"""
    a = np.asarray(a).ravel()
    rows = ["".join(" " + ml.lua_tostring(float(a[j + L * i])) for j in range(L)) for i in range(L)]
    return name + "\n" + "".join(r + "\n" for r in rows)


@pytest.mark.gpu
@pytest.mark.parametrize("fixture,real", [("ref_2d_16_f64", "double"), ("ref_2d_8_f32", "float_acc64")])
def test_debugging_show_dumps_have_the_reference_text(fixture, real):
    """`debugging = true`: run() prints, for every field the reference's twoGrid shows (cpu-raw.lua:187-236), the text
    its show() prints (:126-134), in the reference's order. Expected text = that layout applied to the stage dumps of the
    reference-source run (tests/golden/ref_2d_*.npz). (cpu-raw.lua's stray `print('L', L)` / `print('smooth', i)` lines
    are not part of show() and are not reproduced.)"""
    g = np.load(os.path.join(GOLDEN, fixture + ".npz"))
    size = int(g["meta"][1])
    names, Ls = [str(n) for n in g["trace_names"]], [int(l) for l in g["trace_L"]]
    per_cycle = len(names) // 2
    ffi = minilua_ffi.FFI(search_dirs=[PKG])
    out = io.StringIO()
    it = ml.Interpreter(modules={"ffi": ffi.module(), "ext.class": rr._class, "ext.math": rr._ext_math()}, stdout=out)
    os.environ.setdefault("MGPOISSON_LIB", os.path.join(PKG, "libmgpoisson.so"))
    (cls,) = it.run_file(CUDA_LUA)
    cls.set("debugging", True)
    obj = ml.lua_call(cls, [float(size), real])[0]
    ml.lua_call(ml.lua_index(obj, "run"), [obj])
    text = out.getvalue()
    want = "#iter\terr\n"
    for cyc in range(2):
        for k in range(per_cycle * cyc, per_cycle * (cyc + 1)):
            want += _show_text(names[k], g[f"t{k:05d}"], Ls[k])
        want += f"{cyc + 1}\t"
        assert text.startswith(want), f"cycle {cyc + 1}: dump text differs at char {next(i for i, (a, b) in enumerate(zip(text, want)) if a != b) if not text.startswith(want) else -1}"
        rest = text[len(want):]
        line = rest[:rest.index("\n")]
        assert abs(float(line) - float(g["errs"][cyc])) <= 1e-12 * float(g["errs"][cyc])
        want += line + "\n"
    assert text == want
    ffi.close()


def _run_script(path, args, env):
    ffi = minilua_ffi.FFI(search_dirs=[PKG])
    out = io.StringIO()
    it = ml.Interpreter(modules={"ffi": ffi.module(), "ext.class": rr._class, "ext.math": rr._ext_math()}, stdout=out,
                        search_dirs=[os.path.join(PKG, "lua")])
    it.modules["bit"] = it.globals.vars["bit"]
    it.globals.vars["arg"] = ml.Interpreter.table_from({i + 1: a for i, a in enumerate(args)})
    os.environ.setdefault("MGPOISSON_LIB", os.path.join(PKG, "libmgpoisson.so"))
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        it.run_file(path)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        ffi.close()
    return out.getvalue()


@pytest.mark.gpu
def test_timing_driver_writes_the_reference_table(tmp_path):
    """lua/test/test-cuda.lua: the `cpu-vs-gpu.txt` artefact of test/test.lua:16-33,45-65 with a `cuda` column."""
    tsv = tmp_path / "cpu-vs-gpu.txt"
    text = _run_script(os.path.join(PKG, "lua", "test", "test-cuda.lua"), ["4", "5"], {"MGPOISSON_TSV": str(tsv)})
    lines = tsv.read_text().splitlines()
    assert lines[0] == "#size\tcuda" and len(lines) == 3
    for ln, size in zip(lines[1:], (16, 32)):
        a, b = ln.split("\t")
        assert int(a) == size and float(b) >= 0.0
    assert "#size\tcuda" in text and "#iter\terr" in text          # echoed to the terminal, and run() printed its lines


@pytest.mark.gpu
def test_convergence_driver_writes_the_reference_tables(tmp_path, mgp):
    """lua/test/converge-cuda.lua: converge/<size>.txt of test/converge-multigrid-vs-krylov.lua:79-89 -- per iteration
    |psi|_inf of the multigrid (cpu.lua form) and |x|_inf of conjugate gradient, minus the smallest recorded value."""
    _run_script(os.path.join(PKG, "lua", "test", "converge-cuda.lua"), ["40", "8", "16"], {"MGPOISSON_CONVERGE_DIR": str(tmp_path)})
    for size in (8, 16):
        rows = [ln.split("\t") for ln in (tmp_path / f"{size}.txt").read_text().splitlines()]
        s = mgp.MultigridCUDA(size, "double", dim=2, out=False)
        mgv = []
        for _ in range(40):
            s.zero_corrections()
            e = s.step()
            mgv.append(s.linf_norm())
            if e < 1e-20 or not np.isfinite(e):
                break
        s.init_cells()
        _, cgv = s.conjgrad(max_iter=size * size * 4, epsilon=1e-20)
        s.close()
        n = max(len(mgv), len(cgv))
        assert len(rows) == n and all(len(r) == 2 for r in rows)
        lo = min(mgv + cgv)
        for k in range(n):
            for col, vals in ((0, mgv), (1, cgv)):
                cell = rows[k][col]
                if k < len(vals):
                    assert cell == ml.lua_tostring(vals[k] - lo), (size, k, col, cell, vals[k] - lo)
                else:
                    assert cell in ("nan", "-nan")
        assert any(float(c) == 0.0 for r in rows for c in r if "nan" not in c)
