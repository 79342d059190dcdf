"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE -- see oracle/mg_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may
import this module. The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmgoracle.so")

REAL_F64, REAL_F32, REAL_F32_ACC64 = 0, 1, 2
BUF_F, BUF_PSI, BUF_PSIOLD, BUF_ERRORBUF, BUF_TMPU, BUF_r, BUF_R, BUF_v, BUF_V = range(9)

REAL_NAMES = {"double": REAL_F64, "float": REAL_F32, "float_acc64": REAL_F32_ACC64}


def np_dtype(real_kind: int):
    return np.float64 if real_kind == REAL_F64 else np.float32


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, n) for n in ("mg_oracle.c", "mg_oracle_impl.h", "Makefile")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src if os.path.exists(s))
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i, d, sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t
        L.orc_create.restype = vp
        L.orc_create.argtypes = [i, i, i, i, i]
        L.orc_destroy.argtypes = [vp]
        L.orc_set_threads.argtypes = [vp, i]
        L.orc_buffer.restype = vp
        L.orc_buffer.argtypes = [vp, i, i]
        L.orc_two_grid.argtypes = [vp, d, vp, vp, i]
        L.orc_vcycle.argtypes = [vp]
        L.orc_step.argtypes = [vp, C.POINTER(d)]
        L.orc_run.argtypes = [vp, i, d, C.POINTER(d), C.POINTER(i)]
        L.orc_residual_rms.argtypes = [vp, C.POINTER(d)]
        L.orc_trace_enable.argtypes = [vp, i]
        L.orc_trace_clear.argtypes = [vp]
        L.orc_trace_count.restype = sz
        L.orc_trace_count.argtypes = [vp]
        L.orc_trace_get.argtypes = [vp, sz, C.POINTER(C.c_char), C.POINTER(i), C.POINTER(vp),
                                    C.POINTER(sz)]
        L.orc_op_init_cells.argtypes = [i, i, i, vp, vp]
        L.orc_op_jacobi.argtypes = [i, i, i, vp, vp, vp, d, i]
        L.orc_op_gauss_seidel.argtypes = [i, i, i, vp, vp, d]
        L.orc_op_residual.argtypes = [i, i, i, vp, vp, vp, d, i]
        L.orc_op_restrict.argtypes = [i, i, i, vp, vp, i]
        L.orc_op_prolong.argtypes = [i, i, i, vp, vp, i]
        L.orc_op_add_to.argtypes = [i, sz, vp, vp, i]
        L.orc_op_rel_err.argtypes = [i, sz, vp, vp, vp]
        L.orc_op_frob_err.argtypes = [i, i, i, vp, vp, vp, C.POINTER(d)]
        L.orc_op_apply_A.argtypes = [i, i, i, vp, vp]
        L.orc_op_cg.argtypes = [i, i, i, vp, vp, i, d, C.POINTER(d), C.POINTER(d), C.POINTER(i)]
        L.orc_max_threads.restype = i
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def _shape(dim, L):
    return (L,) * dim  # numpy index order [k][j][i]; i (x) fastest, as `i + L*j` (cpu-raw.lua:9)


# ------------------------------------------------------------------ stateless operators
def init_cells(dim, real_kind, L):
    f = np.zeros(_shape(dim, L), np_dtype(real_kind))
    psi = np.zeros_like(f)
    assert lib().orc_op_init_cells(dim, real_kind, L, _ptr(f), _ptr(psi)) == 0
    return f, psi


def jacobi(dim, real_kind, u, f, h, nthreads=1):
    dest = np.empty_like(u)
    assert lib().orc_op_jacobi(dim, real_kind, u.shape[-1], _ptr(dest), _ptr(u), _ptr(f), h,
                               nthreads) == 0
    return dest


def gauss_seidel(dim, real_kind, u, f, h):
    out = u.copy()
    assert lib().orc_op_gauss_seidel(dim, real_kind, u.shape[-1], _ptr(out), _ptr(f), h) == 0
    return out


def residual(dim, real_kind, f, u, h, nthreads=1):
    r = np.empty_like(u)
    assert lib().orc_op_residual(dim, real_kind, u.shape[-1], _ptr(r), _ptr(f), _ptr(u), h,
                                 nthreads) == 0
    return r


def restrict(dim, real_kind, r, nthreads=1):
    L2 = r.shape[-1] // 2
    R = np.empty(_shape(dim, L2), r.dtype)
    assert lib().orc_op_restrict(dim, real_kind, L2, _ptr(R), _ptr(r), nthreads) == 0
    return R


def prolong(dim, real_kind, V, nthreads=1):
    L2 = V.shape[-1]
    v = np.empty(_shape(dim, 2 * L2), V.dtype)
    assert lib().orc_op_prolong(dim, real_kind, L2, _ptr(v), _ptr(V), nthreads) == 0
    return v


def add_to(real_kind, u, v, nthreads=1):
    out = u.copy()
    assert lib().orc_op_add_to(real_kind, out.size, _ptr(out), _ptr(v), nthreads) == 0
    return out


def frob_err(dim, real_kind, psi, psiOld):
    eb = np.empty_like(psi)
    err = C.c_double()
    assert lib().orc_op_frob_err(dim, real_kind, psi.shape[-1], _ptr(eb), _ptr(psi), _ptr(psiOld),
                                 C.byref(err)) == 0
    return err.value, eb


def apply_A(dim, real_kind, u):
    out = np.empty_like(u)
    assert lib().orc_op_apply_A(dim, real_kind, u.shape[-1], _ptr(out), _ptr(u)) == 0
    return out


def conjgrad(dim, real_kind, x0, b, max_iter=1000, epsilon=1e-20):
    """test/converge-multigrid-vs-krylov.lua:38-69 (textbook CG; the reference's solver lib is un-vendored).
    Returns (x, errs, linf_of_x)."""
    x = np.ascontiguousarray(x0).copy()
    errs, linf = (C.c_double * max(max_iter, 1))(), (C.c_double * max(max_iter, 1))()
    n = C.c_int()
    assert lib().orc_op_cg(dim, real_kind, x.shape[-1], _ptr(x), _ptr(np.ascontiguousarray(b)), max_iter, epsilon,
                           errs, linf, C.byref(n)) == 0
    return x, [errs[k] for k in range(n.value)], [linf[k] for k in range(n.value)]


# ------------------------------------------------------------------ solver object
class Oracle:
    """Mirror of `MultigridCPURaw(size, real)` (cpu-raw.lua:118-258), plus `dim`."""

    smooth = 7          # cpu-raw.lua:123
    accuracy = 1e-10    # cpu-raw.lua:124

    def __init__(self, size, real="double", dim=2, smooth=7, nthreads=1):
        self.real_kind = REAL_NAMES[real] if isinstance(real, str) else int(real)
        self.dim, self.size, self.smooth = dim, size, smooth
        self._h = lib().orc_create(dim, size, self.real_kind, smooth, nthreads)
        if not self._h:
            raise ValueError("orc_create failed (size must be a power of two, dim 2 or 3)")

    def close(self):
        if self._h:
            lib().orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def buffer(self, which, L=None) -> np.ndarray:
        """numpy VIEW of an oracle buffer (no copy)."""
        L = self.size if (L is None or which <= BUF_TMPU) else L
        p = lib().orc_buffer(self._h, which, L)
        n = L ** self.dim
        ct = C.c_double if self.real_kind == REAL_F64 else C.c_float
        arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(n,))
        return arr.reshape(_shape(self.dim, L))

    f = property(lambda s: s.buffer(BUF_F))
    psi = property(lambda s: s.buffer(BUF_PSI))
    psiOld = property(lambda s: s.buffer(BUF_PSIOLD))

    def set_threads(self, n):
        lib().orc_set_threads(self._h, n)

    def vcycle(self):
        assert lib().orc_vcycle(self._h) == 0

    def two_grid(self, h, u: np.ndarray, f: np.ndarray, L):
        assert lib().orc_two_grid(self._h, h, _ptr(u), _ptr(f), L) == 0

    def step(self) -> float:
        err = C.c_double()
        assert lib().orc_step(self._h, C.byref(err)) == 0
        return err.value

    def run(self, max_cycles=2, accuracy=None):
        errs = (C.c_double * max(max_cycles, 1))()
        n = C.c_int()
        acc = self.accuracy if accuracy is None else accuracy
        assert lib().orc_run(self._h, max_cycles, acc, errs, C.byref(n)) == 0
        return [errs[i] for i in range(n.value)]

    def residual_rms(self) -> float:
        r = C.c_double()
        assert lib().orc_residual_rms(self._h, C.byref(r)) == 0
        return r.value

    # trace = the reference's `debugging` dumps (cpu-raw.lua:121,126-140)
    def trace_enable(self, on=True):
        lib().orc_trace_enable(self._h, int(on))

    def trace_clear(self):
        lib().orc_trace_clear(self._h)

    def trace(self):
        out = []
        n = lib().orc_trace_count(self._h)
        dt = np_dtype(self.real_kind)
        for i in range(n):
            name, L, data, nb = C.c_char(), C.c_int(), C.c_void_p(), C.c_size_t()
            assert lib().orc_trace_get(self._h, i, C.byref(name), C.byref(L), C.byref(data),
                                       C.byref(nb)) == 0
            buf = (C.c_char * nb.value).from_address(data.value)
            arr = np.frombuffer(buf, dtype=dt).copy().reshape(_shape(self.dim, L.value))
            out.append((name.value.decode(), L.value, arr))
        return out
