"""Executes the reference's `gpu.lua` (UNMODIFIED host code AND its own OpenCL kernel source) without an OpenCL
device, and writes what it computes to tests/golden/refgpu_*.npz. TEST INFRASTRUCTURE ONLY.

    python oracle/run_reference_gpu.py             # regenerates every fixture (needs /root/reference, gcc)
    python oracle/run_reference_gpu.py 8 float     # one case, prints the err lines

What runs what:
  * gpu.lua's Lua host code (MultigridGPU:init, :twoGrid, :run, clcall1D/2D ...) is interpreted by oracle/minilua.py.
  * The OpenCL C kernel source that gpu.lua:36-199 hands to `cl.program` -- after gpu.lua's own template substitution of
    `size` and `real` -- is piped to GCC and compiled AS C into oracle/_ref/*.so (git-ignored; the reference text itself
    is never written into the repository tree) behind a four-line prelude: `kernel` and `global` are empty qualifiers and get_global_id /
    get_global_size read the work-item index the launcher sets. Flags: -O2 -ffp-contract=off, i.e. every operator
    correctly rounded and nothing contracted -- the strict reading of OpenCL C (a real device may contract a*b+c and
    may divide with <= 2.5 ulp error, so a GPU run of the reference is only defined up to that; this is the one
    deterministic member of that family, and it is what the product's MG_REAL_F32 mode computes).
  * The un-vendored libraries are shimmed from their call sites: `cl.platform/context/commandqueue/program` (a fake
    in-memory device: buffers are byte arrays, enqueueNDRangeKernel loops over the global range and calls the
    compiled kernel once per work item, Fill/Copy/Read are memset/memcpy), `template` (`<?=name?>` looked up in the
    environment table, gpu.lua:199), `ext.class`, `ext.math`, `ext.string`, and the list helpers map / sort / find that
    get64bit (gpu.lua:7-16) applies to the platform and device lists. The fake device advertises cl_khr_fp64 or not,
    which is how gpu.lua itself chooses `real` (gpu.lua:32): both are run.

`showAndCheck` (gpu.lua:234-250) is replaced by a recorder and `debugging` is switched on, as in run_reference.py.
"""
import ctypes as C
import functools
import os
import re
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import minilua as ml  # noqa: E402
import run_reference as rr  # noqa: E402

REFERENCE = rr.REFERENCE
GOLDEN = rr.GOLDEN
REFDIR = os.path.join(HERE, "_ref")
_NP = {"double": np.float64, "float": np.float32, "unsigned int": np.uint32, "int": np.int32}
_CT = {"double": C.c_double, "float": C.c_float, "unsigned int": C.c_uint, "int": C.c_int}

PRELUDE = """/* prelude written by oracle/run_reference_gpu.py; everything after it is the reference's kernel source */
#include <math.h>
#define kernel
#define global
int mg_gid[3], mg_gsz[3];
#define get_global_id(d) (mg_gid[d])
#define get_global_size(d) (mg_gsz[d])
"""


class HostArray(rr.CArray):
    """ffi.new(ctype .. '[n]' [, init]) : also used for scalar kernel arguments (`real[1]`, `unsigned int[1]`)."""

    def __init__(self, n, ctype, init=None):
        self.a = np.zeros(int(n), dtype=_NP[ctype])
        self.ctype = ctype
        if init is not None:
            self.a[0] = init


class DeviceBuffer:
    def __init__(self, nbytes):
        self.raw = np.zeros(int(nbytes), dtype=np.uint8)

    def view(self, real):
        return self.raw.view(_NP[real])


def make_list(items):
    """The list type the cl binding returns (ext.table): 1-based, with map / sort / find."""
    t = ml.LuaTable()
    for i, v in enumerate(items):
        t.set(i + 1, v)

    def values(tt):
        return [tt.get(i) for i in range(1, tt.length() + 1)]

    def tmap(tt, f):
        return make_list([(ml.lua_call(f, [v, float(i + 1), tt]) or [None])[0] for i, v in enumerate(values(tt))])

    def tsort(tt, cmp=None):
        vs = values(tt)
        if cmp is None:
            vs.sort()
        else:
            def c(a, b):
                if ml.lua_truth((ml.lua_call(cmp, [a, b]) or [None])[0]):
                    return -1
                if ml.lua_truth((ml.lua_call(cmp, [b, a]) or [None])[0]):
                    return 1
                return 0
            vs.sort(key=functools.cmp_to_key(c))
        for i, v in enumerate(vs):
            tt.set(i + 1, v)
        return tt

    def tfind(tt, value=None, eq=None):
        # ext.table.find(t, value, eq); gpu.lua:10 passes a predicate as `value`: treated as the test
        for i, v in enumerate(values(tt)):
            if eq is not None:
                hit = ml.lua_truth((ml.lua_call(eq, [v, value]) or [None])[0])
            elif isinstance(value, ml.LuaFunction) or callable(value):
                hit = ml.lua_truth((ml.lua_call(value, [v]) or [None])[0])
            else:
                hit = v == value
            if hit:
                return (float(i + 1), v)
        return (None,)

    meta = ml.LuaTable()
    meta.set("__index", ml.Interpreter.table_from({"map": tmap, "sort": tsort, "find": tfind}))
    t.meta = meta
    return t


def obj(**methods):
    return ml.Interpreter.table_from(methods)


class FakeCL:
    """An OpenCL platform with one GPU device that runs kernels on the host, one work item at a time."""

    def __init__(self, fp64):
        self.fp64 = fp64
        self.launches = []     # (kernel name, global size) in enqueue order
        exts = ["cl_khr_global_int32_base_atomics", "cl_khr_byte_addressable_store"] + (["cl_khr_fp64"] if fp64 else [])
        self.device = obj(getExtensions=lambda d: make_list(exts),
                          getInfo=lambda d, what: "256" if what == "CL_DEVICE_MAX_WORK_GROUP_SIZE" else None)
        self.platform = obj(getExtensions=lambda p: make_list(exts), getDevices=lambda p, flt=None: make_list([self.device]))
        self.real = None
        self.lib = None

    # require 'cl.platform'
    def platform_module(self):
        return obj(getAll=lambda: make_list([self.platform]))

    # require 'cl.context'{platform=, device=}
    def context(self, args):
        return obj(buffer=lambda ctx, a: DeviceBuffer(a.get("size")))

    # require 'cl.program'{context=, devices=, code=}
    def program(self, args):
        code = args.get("code")
        m = re.search(r"typedef\s+(\w+)\s+real\s*;", code)
        self.real = m.group(1)
        size = int(re.search(r"#define\s+size\s+(\d+)", code).group(1))
        os.makedirs(REFDIR, exist_ok=True)
        base = os.path.join(REFDIR, f"gpu_kernels_{self.real}_{size}")
        # the kernel text goes to gcc through a pipe: only the compiled object lands in oracle/_ref/, the reference's
        # source is never written into the repository tree
        subprocess.run(["gcc", "-x", "c", "-std=gnu11", "-O2", "-ffp-contract=off", "-fno-fast-math", "-Wno-unknown-pragmas",
                        "-shared", "-fPIC", "-o", base + ".so", "-", "-lm"], input=(PRELUDE + code).encode(), check=True)
        self.lib = C.CDLL(base + ".so")
        self.gid = (C.c_int * 3).in_dll(self.lib, "mg_gid")
        self.gsz = (C.c_int * 3).in_dll(self.lib, "mg_gsz")

        def kernel(prog, name):
            fn = getattr(self.lib, name)
            fn.restype = None
            state = {"args": []}

            def set_args(k, *a):
                state["args"] = a

            k = obj(setArgs=set_args)
            k.set("_name", name)
            k.set("_fn", lambda: (fn, state["args"]))
            return k

        return obj(kernel=kernel)

    # require 'cl.commandqueue'{context=, device=}
    def queue(self, args):
        def nd_range(q, a):
            k = a.get("kernel")
            fn, kargs = k.get("_fn")()
            gs = a.get("globalSize")
            dims = [int(gs.get(i)) for i in range(1, gs.length() + 1)] if isinstance(gs, ml.LuaTable) else [int(gs)]
            self.launches.append((k.get("_name"), tuple(dims)))
            cargs = []
            for x in kargs:
                if isinstance(x, DeviceBuffer):
                    cargs.append(x.raw.ctypes.data_as(C.c_void_p))
                elif isinstance(x, HostArray):          # `type[1]` = a scalar argument of that type
                    cargs.append(_CT[x.ctype](x.a[0].item()))
                else:
                    raise ml.LuaError(f"kernel argument of unsupported type {type(x).__name__}")
            dims3 = dims + [1] * (3 - len(dims))
            for d in range(3):
                self.gsz[d] = dims3[d]
            gid = self.gid
            for kk in range(dims3[2]):
                gid[2] = kk
                for j in range(dims3[1]):
                    gid[1] = j
                    for i in range(dims3[0]):
                        gid[0] = i
                        fn(*cargs)

        def fill(q, a):
            a.get("buffer").raw[:int(a.get("size"))] = 0      # no pattern given: zero fill

        def copy(q, a):
            n = int(a.get("size"))
            a.get("dst").raw[:n] = a.get("src").raw[:n]

        def read(q, a):
            n = int(a.get("size"))
            a.get("ptr").a.view(np.uint8)[:n] = a.get("buffer").raw[:n]

        return obj(enqueueNDRangeKernel=nd_range, enqueueFillBuffer=fill, enqueueCopyBuffer=copy, enqueueReadBuffer=read)


def _template(code, env=None):
    def sub(m):
        v = ml.lua_index(env, m.group(1).strip())
        return ml.lua_tostring(v)
    return re.sub(r"<\?=(.*?)\?>", sub, code)


def _ffi_gpu():
    def new(ct, a=None, b=None):
        m = re.match(r"\s*([\w ]+?)\s*\[\s*(\?|\d+)\s*\]\s*$", ct)
        if not m:
            raise ml.LuaError(f"ffi.new: unsupported ctype {ct!r}")
        if m.group(2) == "?":
            return HostArray(a, m.group(1))
        return HostArray(int(m.group(2)), m.group(1), a)

    def copy(dst, src, nbytes):
        n = int(nbytes)
        dst.a.view(np.uint8)[:n] = src.a.view(np.uint8)[:n]

    return ml.Interpreter.table_from({"new": new, "copy": copy, "cdef": lambda s: None,
                                      "sizeof": lambda ct: float(np.dtype(_NP[ct]).itemsize)})


def run_reference_gpu(size, fp64):
    """`MultigridGPU(size):run()` (gpu.lua:26-373) on the fake device. Returns errs, real, f, psi, psiOld, per-level
    rs/Rs/vs/Vs, the list of kernel launches and the showAndCheck trace."""
    cl = FakeCL(fp64)
    it = ml.Interpreter(modules={
        "ffi": _ffi_gpu(), "ext.class": rr._class, "ext.math": rr._ext_math(), "ext.string": ml.STRING_LIB.as_table(),
        "cl.platform": cl.platform_module(), "cl.context": lambda a: cl.context(a), "cl.commandqueue": lambda a: cl.queue(a),
        "cl.program": lambda a: cl.program(a), "template": _template})
    (cls,) = it.run_file(os.path.join(REFERENCE, "gpu.lua"))
    trace = []

    def show(self, name, mem, L, *_):
        n = int(L) * int(L)
        trace.append((name, int(L), mem.view(cl.real)[:n].copy()))

    cls.set("showAndCheck", show)
    cls.set("debugging", True)
    o = ml.lua_call(cls, [float(size)])[0]
    real = o.get("real")
    f0, psi0 = o.get("f").view(real).copy(), o.get("psi").view(real).copy()
    ml.lua_call(ml.lua_index(o, "run"), [o])
    errs = [p[1] for p in it.printed if len(p) == 2 and isinstance(p[0], (int, float)) and not isinstance(p[0], bool)]
    out = {"errs": np.array(errs, dtype=np.float64), "real": real, "f0": f0, "psi0": psi0, "f": o.get("f").view(real).copy(),
           "psi": o.get("psi").view(real).copy(), "psiOld": o.get("psiOld").view(real).copy(), "trace": trace,
           "launches": cl.launches}
    L = 1
    while L <= size:
        for nm in ("rs", "Rs", "vs", "Vs"):
            out[f"{nm}{L}"] = o.get(nm).get(L).view(real).copy()
        L *= 2
    return out


CASES = [(4, False, True), (8, False, True), (16, False, True), (32, False, False), (64, False, False), (128, False, False),   # float
         (8, True, True), (32, True, False)]                                                               # real = double


def save_case(size, fp64, keep_trace):
    r = run_reference_gpu(size, fp64)
    real = r["real"]
    d = {k: v for k, v in r.items() if k not in ("trace", "launches", "real")}
    tr = r["trace"]
    tops = [a for (n, L, a) in tr if n == "u" and L == size]
    d["psi_after_cycle1"] = tops[len(tops) // 2 - 1]
    d["trace_names"] = np.array([n for (n, L, a) in tr])
    d["trace_L"] = np.array([L for (n, L, a) in tr], dtype=np.int32)
    d["launch_names"] = np.array([n for (n, g) in r["launches"]])
    if keep_trace:
        for i, (n, L, a) in enumerate(tr):
            d[f"t{i:05d}"] = a
    path = os.path.join(GOLDEN, f"refgpu_2d_{size}_{'f64' if real == 'double' else 'f32'}.npz")
    # real kind in the oracle's numbering: 0 = double, 1 = float storage AND float arithmetic
    np.savez_compressed(path, meta=np.array([2, size, 0 if real == "double" else 1, 2]), **d)
    print(f"{os.path.basename(path)}: real = {real}, {len(r['launches'])} kernel launches, {len(tr)} dumps, "
          f"err = {[float(e) for e in r['errs']]}")


def main():
    if len(sys.argv) >= 2:
        r = run_reference_gpu(int(sys.argv[1]), (sys.argv[2] if len(sys.argv) > 2 else "float") == "double")
        print("#iter\terr    (real = %s, %d kernel launches)" % (r["real"], len(r["launches"])))
        for i, e in enumerate(r["errs"]):
            print(f"{i + 1}\t{e:.14g}")
        return
    for c in CASES:
        save_case(*c)


if __name__ == "__main__":
    main()
