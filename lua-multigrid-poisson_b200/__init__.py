"""Host-side mirror (Python, ctypes) of the reference's solver class for the V-cycle path.

The reference's host language is LuaJIT; no Lua runtime exists in this image, so the
executable host side lives here and the LuaJIT wrapper that a reference user would load
(`lua/multigrid-poisson/cuda.lua`) is shipped beside it, written against the same C ABI
(`include/mgpoisson.h`). Both are thin: every number is computed by libmgpoisson.so.

`MultigridCUDA` mirrors `MultigridCPURaw` / `MultigridGPU`:
    cl(size, real, cpuDepth)            test/test.lua:54, cpu-raw.lua:142, gpu.lua:26
    :run()                              cpu-raw.lua:239-258, gpu.lua:348-373
    :twoGrid(h, u, f, L)                cpu-raw.lua:186-237, gpu.lua:296-346
    :inPlaceIterativeSolver(L, u, f, h) cpu-raw.lua:176-184
    .size .real .smooth .accuracy .debugging, .f .psi .psiOld .errorBuf .tmpU .rs .Rs .vs .Vs

There is no CPU fallback: constructing a solver without a CUDA device raises.
The directory name contains '-' (it is the reference's name), so import it by path; see
`tests/conftest.py::load_package` or `__graft_entry__.load_package`.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import re
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("MGPOISSON_LIB") or os.path.join(_HERE, "libmgpoisson.so")  # same override as cuda.lua
HEADER_PATH = os.path.join(_ROOT, "include", "mgpoisson.h")

REAL_F64, REAL_F32, REAL_F32_ACC64 = 0, 1, 2
BUF_F, BUF_PSI, BUF_PSIOLD, BUF_ERRORBUF, BUF_TMPU, BUF_r, BUF_R, BUF_v, BUF_V = range(9)
MODE_FUSED, MODE_REFSEQ = 0, 1
REAL_NAMES = {"double": REAL_F64, "float": REAL_F32, "float_acc64": REAL_F32_ACC64}
REAL_KIND_NAMES = {v: k for k, v in REAL_NAMES.items()}


class MGError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile libmgpoisson.so for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise MGError("building libmgpoisson.so failed")
    return LIB_PATH


def header_cdef() -> str:
    """The plain-C block of include/mgpoisson.h that LuaJIT's ffi.cdef takes verbatim."""
    src = open(HEADER_PATH).read()
    m = re.search(r"/\* MGPOISSON_CDEF_BEGIN \*/(.*)/\* MGPOISSON_CDEF_END \*/", src, re.S)
    return m.group(1)


def header_functions():
    """Names of every function the header declares."""
    body = re.sub(r"/\*.*?\*/", "", header_cdef(), flags=re.S)
    return sorted(set(re.findall(r"\b(mg_\w+)\s*\(", body)))


_lib = None


def lib():
    """Load the C-ABI library. Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MGError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(the CUDA extension is mandatory; there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i, d, sz, u64 = C.c_void_p, C.c_int, C.c_double, C.c_size_t, C.c_uint64
    pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int)
    sig = {
        "mg_create": (i, [i, i, i, i, i, C.POINTER(vp)]),
        "mg_destroy": (i, [vp]),
        "mg_last_error": (C.c_char_p, [vp]),
        "mg_version": (C.c_char_p, []),
        "mg_set_mode": (i, [vp, i]),
        "mg_set_stream": (i, [vp, vp]),
        "mg_set_tuning": (i, [vp, i, i, i]),
        "mg_set_option": (i, [vp, C.c_char_p, i]),
        "mg_get_info": (i, [vp, pi, pi, pi, pi, pi, C.POINTER(u64)]),
        "mg_init_cells": (i, [vp]),
        "mg_zero_corrections": (i, [vp]),
        "mg_upload": (i, [vp, i, i, vp, sz]),
        "mg_download": (i, [vp, i, i, vp, sz]),
        "mg_device_ptr": (vp, [vp, i, i]),
        "mg_host_alloc": (vp, [sz]),
        "mg_host_free": (i, [vp]),
        "mg_vcycle": (i, [vp]),
        "mg_vcycle_async": (i, [vp]),
        "mg_synchronize": (i, [vp]),
        "mg_step": (i, [vp, pd]),
        "mg_run": (i, [vp, i, d, pd, pi]),
        "mg_step_host": (i, [vp, vp, vp, pd]),
        "mg_step_host_batch": (i, [vp, i, C.POINTER(vp), C.POINTER(vp), pd]),
        "mg_residual_norm": (i, [vp, pd]),
        "mg_twogrid": (i, [vp, d, vp, vp, i]),
        "mg_smooth": (i, [vp, i, vp, vp, d, i]),
        "mg_jacobi": (i, [vp, i, vp, vp, vp, d]),
        "mg_residual": (i, [vp, i, vp, vp, vp, d]),
        "mg_restrict": (i, [vp, i, vp, vp]),
        "mg_prolong": (i, [vp, i, vp, vp]),
        "mg_add_to": (i, [vp, sz, vp, vp]),
        "mg_frob_err": (i, [vp, pd]),
        "mg_smooth_residual_restrict": (i, [vp, i, vp, vp, d, i, vp]),
        "mg_prolong_add_smooth": (i, [vp, i, vp, vp, d, i, vp]),
        "mg_cg": (i, [vp, i, d, pd, pd, pi]),
        "mg_linf_norm": (i, [vp, i, i, pd]),
        "mg_trace_enable": (i, [vp, i]),
        "mg_trace_clear": (i, [vp]),
        "mg_trace_count": (sz, [vp]),
        "mg_trace_get": (i, [vp, sz, C.POINTER(C.c_char), pi, C.POINTER(vp), C.POINTER(sz)]),
        "mg_time_vcycles": (i, [vp, i, C.POINTER(C.c_float)]),
        "mg_launch_count": (u64, [vp]),
        "mg_profile_vcycle": (i, [vp, i, pi, pi, pi, C.POINTER(C.c_float), pi]),
        "mg_nccl_unique_id": (i, [vp, sz]),
        "mg_create_slab": (i, [i, i, i, i, i, i, i, vp, sz, C.POINTER(vp)]),
        "mg_create_slab_local": (i, [i, i, i, i, i, i, C.POINTER(vp)]),
        "mg_create_slab_multi": (i, [i, i, i, i, i, pi, C.POINTER(vp)]),
        "mg_set_global_option": (i, [C.c_char_p, i]),
        "mg_set_omega": (i, [vp, d]),
        "mg_slab_traffic": (i, [vp, C.POINTER(u64), C.POINTER(u64)]),
        "mg_slab_trace": (i, [vp, C.POINTER(u64), sz, C.POINTER(sz)]),
        "mg_slab_ipc_export": (i, [vp, vp, sz]),
        "mg_slab_ipc_attach": (i, [vp, vp, sz]),
        "mg_slab_info": (i, [vp, pi, pi, pi, pi, C.POINTER(u64), C.POINTER(u64)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def np_dtype(real_kind):
    return np.float64 if real_kind == REAL_F64 else np.float32


def _hptr(a: np.ndarray):
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("host array must be C-contiguous")
    return a.ctypes.data_as(C.c_void_p)


class DeviceBuffer:
    """A device field of the hierarchy (the reference's `image(...).buffer` / cl buffer)."""

    def __init__(self, owner, which, L):
        self.owner, self.which, self.L = owner, which, L

    @property
    def ptr(self) -> int:
        p = lib().mg_device_ptr(self.owner._h, self.which, self.L)
        if not p:
            raise MGError("buffer is not materialised in this mode (rs/vs/errorBuf/tmpU exist only in "
                          "the reference-sequence mode)")
        return p

    @property
    def shape(self):
        o = self.owner
        if o.slab_n > 1 and self.L >= 64 and self.L // o.slab_n >= max(8, int(os.environ.get("MGPOISSON_SLAB_MIN_PLANES", "32"))):
            return (self.L // o.slab_n,) + (self.L,) * (o.dim - 1)
        return (self.L,) * o.dim

    @property
    def nbytes(self):
        return self.L ** self.owner.dim * self.owner.dtype().itemsize

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": np.dtype(self.owner.dtype).str, "data": (self.ptr, False),
                "version": 3, "strides": None}

    def download(self) -> np.ndarray:
        out = np.empty(self.shape, self.owner.dtype)
        self.owner._ck(lib().mg_download(self.owner._h, self.which, self.L, _hptr(out), out.nbytes))
        return out

    def upload(self, a):
        a = np.ascontiguousarray(a, dtype=self.owner.dtype)
        if a.size != int(np.prod(self.shape)):
            raise ValueError("size mismatch")
        self.owner._ck(lib().mg_upload(self.owner._h, self.which, self.L, _hptr(a), a.nbytes))


class _LevelTable:
    """`self.rs[L]`-style access (cpu-raw.lua:155-164)."""

    def __init__(self, owner, which):
        self.owner, self.which = owner, which

    def __getitem__(self, L):
        return DeviceBuffer(self.owner, self.which, L)


class MultigridCUDA:
    """Drop-in column for test/test.lua's `cols`: `cl(size, real, cpuDepth)` then `:run()`.

    real: None/'double' (cpu-raw.lua:143 default), 'float' (gpu.lua fp32 device semantics) or
    'float_acc64' (cpu-raw.lua with real='float': fp32 storage, double arithmetic).
    cpuDepth: the reference's hybrid hands levels L <= 2^cpuDepth to another executor
    (cpu-gpu.lua:11-15,94); here they go to the persistent small-level kernel.
    dim=3 is this project's extension (SURVEY section 8(a')).
    """

    smooth = 7          # cpu-raw.lua:123
    accuracy = 1e-10    # cpu-raw.lua:124
    debugging = False   # cpu-raw.lua:121
    max_cycles = 2      # cpu-raw.lua:245 `for iter=1,2`

    def __init__(self, size, real=None, cpuDepth=None, dim=2, device=-1, smooth=None, out=None,
                 local_slabs=1, slab=None, devices=None):
        """local_slabs > 1: cut the 3-D grid into that many slabs inside this process on one GPU
        (exercises the multi-GPU schedule on a single device; the object still looks like one
        solver on the global grid). devices = [0, 1, ...]: one slab per listed GPU, all driven by
        this process (mg_create_slab_multi; what a single LuaJIT host would use). slab = (rank,
        nranks, nccl_id_bytes): this process's slab of a multi-process solver (see
        `create_distributed`); buffers then hold the owned planes."""
        self.real = real or "double"
        self.real_kind = REAL_NAMES[self.real] if isinstance(self.real, str) else int(self.real)
        self.dtype = np_dtype(self.real_kind)
        self.size, self.dim = int(size), int(dim)
        if smooth is not None:
            self.smooth = int(smooth)
        self.out = out  # where run() prints its `#iter err` lines (None = stdout, False = silent)
        h = C.c_void_p()
        self.slab_rank, self.slab_n = 0, 1
        if slab is not None:
            self.slab_rank, self.slab_n, nid = slab
            idbuf = C.create_string_buffer(bytes(nid), len(nid))
            rc = lib().mg_create_slab(self.dim, self.size, self.real_kind, self.smooth, device, self.slab_rank,
                                      self.slab_n, idbuf, len(nid), C.byref(h))
        elif devices is not None and len(devices) > 1:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = lib().mg_create_slab_multi(self.dim, self.size, self.real_kind, self.smooth, len(devices), arr, C.byref(h))
        elif local_slabs > 1:
            rc = lib().mg_create_slab_local(self.dim, self.size, self.real_kind, self.smooth, device,
                                            int(local_slabs), C.byref(h))
        else:
            rc = lib().mg_create(self.dim, self.size, self.real_kind, self.smooth, device, C.byref(h))
        if rc != 0:
            raise MGError(f"mg_create failed ({rc}): {lib().mg_last_error(None).decode()}")
        self._h = h
        self.f = DeviceBuffer(self, BUF_F, self.size)
        self.psi = DeviceBuffer(self, BUF_PSI, self.size)
        self.psiOld = DeviceBuffer(self, BUF_PSIOLD, self.size)
        self.errorBuf = DeviceBuffer(self, BUF_ERRORBUF, self.size)
        self.tmpU = DeviceBuffer(self, BUF_TMPU, self.size)
        self.rs, self.Rs = _LevelTable(self, BUF_r), _LevelTable(self, BUF_R)
        self.vs, self.Vs = _LevelTable(self, BUF_v), _LevelTable(self, BUF_V)
        if cpuDepth is not None:
            self.set_tuning(small_L=min(1 << int(cpuDepth), 256, self.size))

    # -------------------------------------------------------------- plumbing
    def _ck(self, rc):
        if rc != 0:
            raise MGError(f"libmgpoisson error {rc}: {lib().mg_last_error(self._h).decode()}")

    def close(self):
        if getattr(self, "_h", None):
            lib().mg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_mode(self, mode):
        self._ck(lib().mg_set_mode(self._h, mode))

    def set_debugging(self, on=True):
        """`debugging = true` (cpu-raw.lua:121): reference operator sequence + stage trace."""
        self.debugging = bool(on)
        self.set_mode(MODE_REFSEQ if on else MODE_FUSED)
        self._ck(lib().mg_trace_enable(self._h, int(on)))

    def set_stream(self, cuda_stream: int | None):
        self._ck(lib().mg_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def set_tuning(self, tb=-1, small_L=-1, use_graph=-1):
        self._ck(lib().mg_set_tuning(self._h, tb, small_L, use_graph))

    def set_option(self, name, value):
        self._ck(lib().mg_set_option(self._h, name.encode(), int(value)))

    def set_omega(self, omega: float):
        """Weighted Jacobi (an extension: the reference is omega = 1, the default)."""
        self._ck(lib().mg_set_omega(self._h, float(omega)))

    def slab_info(self):
        v = [C.c_int() for _ in range(4)]
        ex, eb = C.c_uint64(), C.c_uint64()
        self._ck(lib().mg_slab_info(self._h, *[C.byref(x) for x in v], C.byref(ex), C.byref(eb)))
        return dict(rank=v[0].value, nranks=v[1].value, own_planes=v[2].value, ghost=v[3].value,
                    exchanges=ex.value, exchanged_bytes=eb.value)

    def slab_traffic(self):
        """NVLink bytes so far: stored by this handle's kernels into other GPUs' memory, and moved by explicit exchanges."""
        a, b = C.c_uint64(), C.c_uint64()
        self._ck(lib().mg_slab_traffic(self._h, C.byref(a), C.byref(b)))
        return dict(peer_store_bytes=a.value, exchange_bytes=b.value)

    def slab_trace(self, cap=4096):
        """Timeline of this rank's distributed smoother passes since set_option("slab_trace", 1): an (n, 4) uint64 array
        {entry, after the lower-neighbour wait, ns waited for the upper neighbour, end} in this GPU's nanoseconds."""
        buf = (C.c_uint64 * (4 * cap))()
        n = C.c_size_t()
        self._ck(lib().mg_slab_trace(self._h, buf, cap, C.byref(n)))
        return np.ctypeslib.as_array(buf)[: 4 * n.value].reshape(-1, 4).copy()

    def info(self):
        v = [C.c_int() for _ in range(5)]
        ab = C.c_uint64()
        self._ck(lib().mg_get_info(self._h, *[C.byref(x) for x in v], C.byref(ab)))
        return dict(dim=v[0].value, size=v[1].value, real_kind=v[2].value, smooth=v[3].value,
                    nlevels=v[4].value, arena_bytes=ab.value)

    # -------------------------------------------------------------- reference methods
    def init_cells(self):
        self._ck(lib().mg_init_cells(self._h))

    def zero_corrections(self):
        self._ck(lib().mg_zero_corrections(self._h))

    def twoGrid(self, h, u, f, L):
        """cpu-raw.lua:186: V-cycle on level L; u, f are DeviceBuffers or raw device pointers."""
        self._ck(lib().mg_twogrid(self._h, float(h), _dptr(u), _dptr(f), int(L)))

    def inPlaceIterativeSolver(self, L, u, f, h, n=1):
        """cpu-raw.lua:176: one Jacobi sweep on u (n of them with n > 1)."""
        self._ck(lib().mg_smooth(self._h, int(L), _dptr(u), _dptr(f), float(h), int(n)))

    def vcycle(self):
        self._ck(lib().mg_vcycle(self._h))

    def step(self) -> float:
        e = C.c_double()
        self._ck(lib().mg_step(self._h, C.byref(e)))
        return e.value

    def step_host(self, f_host: np.ndarray, psi_host: np.ndarray) -> float:
        e = C.c_double()
        self._ck(lib().mg_step_host(self._h, _hptr(f_host), _hptr(psi_host), C.byref(e)))
        return e.value

    def step_host_batch(self, f_hosts, psi_hosts):
        """mg_step_host for a list of independent host problems, transfers overlapped with the cycles
        (psi_hosts are updated in place and must be distinct arrays); returns the list of errs."""
        n = len(psi_hosts)
        assert len(f_hosts) == n
        fp = (C.c_void_p * max(n, 1))(*[_hptr(a) for a in f_hosts])
        pp = (C.c_void_p * max(n, 1))(*[_hptr(a) for a in psi_hosts])
        errs = (C.c_double * max(n, 1))()
        self._ck(lib().mg_step_host_batch(self._h, n, fp, pp, errs))
        return [errs[k] for k in range(n)]

    def run(self, max_cycles=None, accuracy=None):
        """cpu-raw.lua:239-258. Prints the reference's `#iter err` table; returns the errs."""
        n = self.max_cycles if max_cycles is None else int(max_cycles)
        acc = self.accuracy if accuracy is None else float(accuracy)
        errs = (C.c_double * max(n, 1))()
        done = C.c_int()
        self._ck(lib().mg_run(self._h, n, acc, errs, C.byref(done)))
        res = [errs[k] for k in range(done.value)]
        if self.out is not False:
            o = self.out or sys.stdout
            print("#iter\terr", file=o)
            for k, e in enumerate(res):
                print(f"{k + 1}\t{_lua_number(e)}", file=o)
        return res

    def residual_norm(self) -> float:
        r = C.c_double()
        self._ck(lib().mg_residual_norm(self._h, C.byref(r)))
        return r.value

    def frob_err(self) -> float:
        e = C.c_double()
        self._ck(lib().mg_frob_err(self._h, C.byref(e)))
        return e.value

    # -------------------------------------------------------------- Krylov comparator
    def conjgrad(self, max_iter=1000, epsilon=1e-20):
        """test/converge-multigrid-vs-krylov.lua:38-69: CG on the same operator, b = f, x = psi as found.
        Returns (errs, linf_of_x) per iteration."""
        errs, linf = (C.c_double * max(max_iter, 1))(), (C.c_double * max(max_iter, 1))()
        n = C.c_int()
        self._ck(lib().mg_cg(self._h, int(max_iter), float(epsilon), errs, linf, C.byref(n)))
        return [errs[k] for k in range(n.value)], [linf[k] for k in range(n.value)]

    def linf_norm(self, which=BUF_PSI, L=None) -> float:
        v = C.c_double()
        self._ck(lib().mg_linf_norm(self._h, which, L or self.size, C.byref(v)))
        return v.value

    # -------------------------------------------------------------- per-operator (device ptrs)
    def jacobi(self, L, dest, u, f, h):
        self._ck(lib().mg_jacobi(self._h, L, _dptr(dest), _dptr(u), _dptr(f), float(h)))

    def residual(self, L, r, f, u, h):
        self._ck(lib().mg_residual(self._h, L, _dptr(r), _dptr(f), _dptr(u), float(h)))

    def restrict(self, L2, R, r):
        self._ck(lib().mg_restrict(self._h, L2, _dptr(R), _dptr(r)))

    def prolong(self, L2, v, V):
        self._ck(lib().mg_prolong(self._h, L2, _dptr(v), _dptr(V)))

    def add_to(self, n, u, v):
        self._ck(lib().mg_add_to(self._h, n, _dptr(u), _dptr(v)))

    def smooth_residual_restrict(self, L, u, f, h, n, R):
        self._ck(lib().mg_smooth_residual_restrict(self._h, L, _dptr(u), _dptr(f), float(h), n, _dptr(R)))

    def prolong_add_smooth(self, L, u, f, h, n, V):
        self._ck(lib().mg_prolong_add_smooth(self._h, L, _dptr(u), _dptr(f), float(h), n, _dptr(V)))

    # -------------------------------------------------------------- trace / measurement
    def trace_clear(self):
        self._ck(lib().mg_trace_clear(self._h))

    def trace(self):
        out = []
        for k in range(lib().mg_trace_count(self._h)):
            name, L, data, nb = C.c_char(), C.c_int(), C.c_void_p(), C.c_size_t()
            self._ck(lib().mg_trace_get(self._h, k, C.byref(name), C.byref(L), C.byref(data), C.byref(nb)))
            buf = (C.c_char * nb.value).from_address(data.value)
            out.append((name.value.decode(), L.value,
                        np.frombuffer(buf, dtype=self.dtype).copy().reshape((L.value,) * self.dim)))
        return out

    def show_text(self, name, field, L) -> str:
        """The reference's debug dump layout (cpu-raw.lua:126-134), 2-D only."""
        lines = [name]
        for i in range(L):
            lines.append("".join(" " + _lua_number(float(field[i, j])) for j in range(L)))
        return "\n".join(lines) + "\n"

    def time_vcycles(self, n) -> float:
        ms = C.c_float()
        self._ck(lib().mg_time_vcycles(self._h, int(n), C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        return int(lib().mg_launch_count(self._h))

    KERNEL_KINDS = {0: "sweep", 1: "residual_restrict", 2: "small_levels", 3: "prolong_add", 4: "copy",
                    5: "sweep+prolong_add", 6: "sweep+residual_restrict"}

    def profile_vcycle(self):
        """One ungraphed fused V-cycle with CUDA events around every launch.
        Returns a list of dicts {kind, L, sweeps, ms}."""
        cap = 4096
        kind, L, sw = (C.c_int * cap)(), (C.c_int * cap)(), (C.c_int * cap)()
        ms, n = (C.c_float * cap)(), C.c_int()
        self._ck(lib().mg_profile_vcycle(self._h, cap, kind, L, sw, ms, C.byref(n)))
        return [dict(kind=self.KERNEL_KINDS[kind[k]], L=L[k], sweeps=sw[k], ms=ms[k])
                for k in range(min(n.value, cap))]


def _dptr(x):
    if isinstance(x, DeviceBuffer):
        return C.c_void_p(x.ptr)
    if hasattr(x, "data_ptr"):  # torch tensor
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(int(x))


def _lua_number(x: float) -> str:
    """Lua's tostring(number): '%.14g'."""
    if math.isnan(x):
        return "nan"
    if math.isinf(x):
        return "inf" if x > 0 else "-inf"
    return "%.14g" % x


class PinnedArray:
    """numpy view over cudaMallocHost memory (for the host-buffer entry point)."""

    def __init__(self, shape, dtype):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._p = lib().mg_host_alloc(self.nbytes)
        if not self._p:
            raise MGError("mg_host_alloc failed")
        buf = (C.c_char * self.nbytes).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self._p:
            self.array = None
            lib().mg_host_free(self._p)
            self._p = None


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = lib().mg_nccl_unique_id(buf, 128)
    if rc != 0:
        raise MGError(f"mg_nccl_unique_id failed ({rc}): {lib().mg_last_error(None).decode()}")
    return buf.raw


def create_distributed(size, real="float", dim=3, smooth=None, out=False, p2p=True):
    """One slab per process (torchrun: one process per GPU). torch.distributed is only the
    plumbing that ships rank 0's NCCL unique id to the other ranks; every halo exchange,
    all-gather and reduction afterwards is issued by libmgpoisson on its own communicator."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return MultigridCUDA(size, real, dim=dim, smooth=smooth, out=out, device=torch.cuda.current_device())
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        t = torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8).to(dev)
    dist.broadcast(t, 0)
    nid = bytes(t.cpu().numpy().tobytes())
    s = MultigridCUDA(size, real, dim=dim, smooth=smooth, out=out, device=torch.cuda.current_device(),
                      slab=(rank, world, nid))
    if p2p:
        # fused halo exchange: all-gather the CUDA IPC handles of the arenas, map the neighbours
        hbuf = C.create_string_buffer(64)
        s._ck(lib().mg_slab_ipc_export(s._h, hbuf, 64))
        mine = torch.frombuffer(bytearray(hbuf.raw), dtype=torch.uint8).to(dev)
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        blob = b"".join(bytes(x.cpu().numpy().tobytes()) for x in allh)
        s._ck(lib().mg_slab_ipc_attach(s._h, C.create_string_buffer(blob, len(blob)), len(blob)))
        dist.barrier()
    return s


def slab_partition(size, nranks):
    """Host-side description of the decomposition libmgpoisson uses (mg_slab.cuh): which level
    widths are cut across the ranks, planes per rank, ghost depth, and the replicated levels."""
    levels, L = [], size
    while L >= 1:
        d = nranks > 1 and L >= 64 and L // nranks >= max(8, int(os.environ.get("MGPOISSON_SLAB_MIN_PLANES", "32")))
        levels.append(dict(L=L, distributed=d, planes_per_rank=L // nranks if d else L, ghost=4 if d else 0))
        L //= 2
    return levels


def plan_passes(n, tb, has_res, extra_pass=False):
    """Host mirror of EngineT::plan_passes (csrc/mg_engine.cuh): how n Jacobi sweeps are split
    into smoother passes of <= tb sweeps; with a fused residual stage the last pass has <= 3.
    extra_pass: one pass more than necessary (the slab schedule keeps the pass count of a level
    visit even, so that the ping-pong ends in u without a copy)."""
    plan, rem, last = [], n, 0
    if has_res:
        last = min(n, 3, tb)
        rem = n - last
    if rem > 0:
        k = (rem + tb - 1) // tb
        if extra_pass and k < rem:
            k += 1
        base, extra = divmod(rem, k)
        plan += [base + (1 if i < extra else 0) for i in range(k)]
    if has_res:
        plan.append(last)
    return plan


def slab_schedule(size, nranks, smooth=7, tb=4):
    """The communication schedule of one slab V-cycle (csrc/mg_engine.cuh::slab_twogrid), as a
    list of (op, level width, detail) tuples: what is exchanged, to which depth, and when."""
    ops = []

    def visit(L):
        lv = [x for x in slab_partition(size, nranks) if x["L"] == L][0]
        if not lv["distributed"]:
            ops.append(("replicated_vcycle", L, None))
            return
        coarse = [x for x in slab_partition(size, nranks) if x["L"] == L // 2][0]
        pre = plan_passes(smooth, tb, True)
        post = plan_passes(smooth, tb, False)
        if (len(pre) + len(post)) % 2:          # keep the ping-pong even: the result must land in u by itself
            post = plan_passes(smooth, tb, False, extra_pass=True)
        for i, s in enumerate(pre):
            res = i == len(pre) - 1
            ops.append(("exchange_u", L, s + (1 if res else 0)))
            ops.append(("pass", L, dict(sweeps=s, res=res, pro=False)))
        ops.append(("exchange_R" if coarse["distributed"] else "allgather_R", L // 2, 4 if coarse["distributed"] else None))
        visit(L // 2)
        if coarse["distributed"]:
            ops.append(("exchange_V", L // 2, 2))
        for i, s in enumerate(post):
            ops.append(("exchange_u", L, s))
            ops.append(("pass", L, dict(sweeps=s, res=False, pro=i == 0)))

    ops.append(("exchange_f", size, 4))
    visit(size)
    return ops
