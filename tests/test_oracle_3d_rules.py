"""The 3-D rules do not exist in the reference (SURVEY.md 8(a')): six neighbours with out-of-range = 0, neighbour sum
((((xl+xr)+yl)+yr)+zl)+zr, adiag = -6/h^2, restriction .125 x (children summed i fastest, then j, then k), injection
prolongation, err over size^3. The C oracle is the pin for every 3-D parity test, so it is checked here against a SECOND,
independent statement of those rules: a few lines of numpy (whole-array IEEE operations in the same association order),
bit for bit in fp64 over whole V-cycles, and in the fp32 modes with the matching rounding points. The same numpy code
restricted to two dimensions is checked against the fixtures produced by the reference's own cpu-raw.lua, which ties
this second statement to the reference where the reference exists."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


class NumpyMultigrid:
    """cpu-raw.lua:142-258 in numpy, any dimension. store = dtype of the arrays, arith = dtype of the arithmetic."""

    def __init__(self, dim, size, store=np.float64, arith=np.float64, smooth=7):
        self.dim, self.size, self.store, self.arith, self.smooth = dim, size, store, arith, smooth
        c = size // 2
        self.f = np.zeros((size,) * dim, store)
        self.f[(c,) * dim] = -1e6                      # cpu-raw.lua:8-20 (index order is symmetric)
        self.psi = (-self.f).astype(store)
        self.V = {}
        L = size // 2
        while L >= 1:
            self.V[L] = np.zeros((L,) * dim, store)    # zeroed once, never again (cpu-raw.lua:165-170)
            L //= 2

    def nsum(self, u):
        a = u.astype(self.arith)
        p = np.pad(a, 1)
        core = (slice(1, -1),) * self.dim

        def sh(axis, d):       # neighbour at offset d along the axis; axis -1 is x (fastest)
            idx = list(core)
            idx[axis] = slice(1 + d, p.shape[axis] - 1 + d)
            return p[tuple(idx)]
        s = sh(-1, -1) + sh(-1, 1)                     # (xl + xr)
        for ax in range(2, self.dim + 1):              # + yl + yr [+ zl + zr], left to right
            s = s + sh(-ax, -1)
            s = s + sh(-ax, 1)
        return s

    def jacobi(self, u, f, h):
        A = self.arith
        h2 = A(h) * A(h)
        askew = self.nsum(u) / h2
        adiag = A(-2 * self.dim) / h2
        return ((f.astype(A) - askew) / adiag).astype(self.store)

    def residual(self, u, f, h):
        A = self.arith
        h2 = A(h) * A(h)
        askew = self.nsum(u) / h2
        adiag = A(-2 * self.dim) / h2
        a_u = askew + adiag * u.astype(A)
        return (f.astype(A) - a_u).astype(self.store)

    def restrict(self, r):
        a = r.astype(self.arith)
        s = None
        for off in np.ndindex(*(2,) * self.dim):       # last index (x) fastest, then y, then z
            child = a[tuple(slice(o, None, 2) for o in off)]
            s = child if s is None else s + child
        return (self.arith(0.5 ** self.dim) * s).astype(self.store)

    def prolong(self, V):
        v = V
        for ax in range(self.dim):
            v = np.repeat(v, 2, axis=ax)
        return v

    def two_grid(self, h, u, f, L):
        if L == 1:
            return self.jacobi(u, f, h)
        for _ in range(self.smooth):
            u = self.jacobi(u, f, h)
        R = self.restrict(self.residual(u, f, h))
        self.V[L // 2] = self.two_grid(2 * h, self.V[L // 2], R, L // 2)
        u = (u.astype(self.arith) + self.prolong(self.V[L // 2]).astype(self.arith)).astype(self.store)
        for _ in range(self.smooth):
            u = self.jacobi(u, f, h)
        return u

    def step(self):
        old = self.psi.copy()
        self.psi = self.two_grid(1.0 / self.size, self.psi, self.f, self.size)
        d = (self.psi.astype(np.float64) - old.astype(np.float64)).ravel()
        if self.arith == np.float32:
            d = (self.psi - old).astype(np.float32).ravel().astype(np.float64)
            sq = (d.astype(np.float32) * d.astype(np.float32)).astype(np.float64)
        else:
            sq = (d * d).astype(self.store).astype(np.float64)     # errorBuf is stored in `real` (cpu-raw.lua:99)
        err = 0.0
        for x in sq:                                               # sequential sum, cpu-raw.lua:250-253
            err += x
        return float(np.sqrt(err / self.size ** self.dim))


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint64 if a.dtype == np.float64 else np.uint32)


MODES = {"double": (np.float64, np.float64), "float_acc64": (np.float32, np.float64), "float": (np.float32, np.float32)}


@pytest.mark.parametrize("real", ["double", "float_acc64", "float"])
@pytest.mark.parametrize("dim,size", [(3, 8), (3, 16), (2, 16)])
def test_oracle_equals_the_numpy_statement_of_the_rules(dim, size, real):
    store, arith = MODES[real]
    n = NumpyMultigrid(dim, size, store, arith)
    o = O.Oracle(size, real, dim)
    assert np.array_equal(_bits(o.psi), _bits(n.psi))
    for cyc in range(3):
        en, eo = n.step(), o.step()
        assert np.array_equal(_bits(o.psi), _bits(n.psi)), (dim, size, real, cyc)
        assert abs(en - eo) <= 1e-12 * eo
        L = size // 2
        while L >= 1:
            assert np.array_equal(_bits(o.buffer(O.BUF_V, L)), _bits(n.V[L])), (dim, size, real, cyc, L)
            L //= 2


@pytest.mark.parametrize("name,real", [("ref_2d_32_f64", "double"), ("ref_2d_32_f32", "float_acc64"), ("refgpu_2d_32_f32", "float")])
def test_the_numpy_statement_in_2d_equals_the_reference_source_run(name, real):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    store, arith = MODES[real]
    n = NumpyMultigrid(2, 32, store, arith)
    errs = [n.step(), n.step()]
    assert np.array_equal(_bits(n.psi.ravel()), _bits(g["psi"]))
    for e, w in zip(errs, g["errs"]):
        assert abs(e - float(w)) <= 1e-12 * float(w)
