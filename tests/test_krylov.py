"""The Krylov comparator of test/converge-multigrid-vs-krylov.lua.
CPU: the oracle's CG against a dense solve and numpy's operator. GPU: mg_cg against the oracle's CG.
The reference's `solver.conjgrad` is un-vendored and unpinned, so this row is parity-unpinned."""
import numpy as np
import pytest

from gpu_util import err_rtol


def test_oracle_operator_and_cg_solve_the_poisson_problem(orc):
    L = 16
    rng = np.random.default_rng(2)
    u = rng.uniform(-1, 1, (L, L))
    p = np.pad(u, 1)
    S = ((p[1:-1, :-2] + p[1:-1, 2:]) + p[:-2, 1:-1]) + p[2:, 1:-1]
    assert np.array_equal(orc.apply_A(2, orc.REAL_F64, u), (S - 4 * u) / (1.0 / L) ** 2)
    # the experiment's call: x0 = -f, b = f, point source (converge-multigrid-vs-krylov.lua:45-46)
    f, psi = orc.init_cells(2, orc.REAL_F64, L)
    x, errs, linf = orc.conjgrad(2, orc.REAL_F64, psi, f, max_iter=400, epsilon=1e-12)
    assert errs[-1] < 1e-12 and len(errs) < 200 and len(linf) == len(errs)
    assert np.allclose(orc.apply_A(2, orc.REAL_F64, x), f, rtol=0, atol=1e-5 * 1e6)
    # dense check of the solution
    n = L * L
    Amat = np.zeros((n, n))
    for k in range(n):
        e = np.zeros(n); e[k] = 1
        Amat[:, k] = orc.apply_A(2, orc.REAL_F64, e.reshape(L, L)).ravel()
    assert np.allclose(x.ravel(), np.linalg.solve(Amat, f.ravel()), rtol=1e-8, atol=1e-9)
    assert linf[-1] == np.abs(x).max()


@pytest.mark.gpu
@pytest.mark.parametrize("dim,size,real", [(2, 64, "double"), (2, 256, "double"), (3, 32, "double"), (2, 128, "float")])
def test_cuda_cg_matches_oracle_cg(mgp, orc, dim, size, real):
    k = orc.REAL_NAMES[real]
    s = mgp.MultigridCUDA(size, real, dim=dim, out=False)
    f, psi = orc.init_cells(dim, k, size)
    eps = 1e-10 if real == "double" else 1e-5
    x, errs_o, linf_o = orc.conjgrad(dim, k, psi, f, max_iter=2000, epsilon=eps)
    errs, linf = s.conjgrad(max_iter=2000, epsilon=eps)
    assert errs[-1] < eps
    assert abs(len(errs) - len(errs_o)) <= max(2, len(errs_o) // 50)     # same iteration count (dot order differs)
    m = min(len(errs), len(errs_o), 40)
    tol = 1e-9 if real == "double" else 1e-3
    assert np.allclose(errs[:m], errs_o[:m], rtol=tol, atol=0)
    assert np.allclose(linf[:m], linf_o[:m], rtol=tol, atol=0)
    xs = s.psi.download()
    scale = np.abs(x).max()
    assert np.abs(xs.astype(np.float64) - x).max() <= (1e-8 if real == "double" else 2e-3) * scale
    assert abs(s.linf_norm() - np.abs(xs).max()) == 0
    s.close()


@pytest.mark.gpu
def test_multigrid_vs_krylov_records_like_the_experiment(mgp, orc):
    """converge-multigrid-vs-krylov.lua:20-29: MultigridCPU (cpu.lua, V re-zeroed each cycle, :138) run
    for a fixed number of cycles recording ||psi||_inf per cycle, then CG on the same problem."""
    size = 64
    s = mgp.MultigridCUDA(size, "double", out=False)
    rec = []
    for _ in range(5):
        s.zero_corrections()        # the cpu.lua variant
        s.step()
        rec.append(s.linf_norm())
    o = orc.Oracle(size, "double", 2)
    want = []
    for _ in range(5):
        L = size // 2
        while L >= 1:
            o.buffer(orc.BUF_V, L)[...] = 0
            L //= 2
        o.step()
        want.append(np.abs(o.psi).max())
    assert rec == want
    s.init_cells()
    errs, linf = s.conjgrad(max_iter=500, epsilon=1e-10)
    assert errs[-1] < 1e-10 and len(linf) == len(errs)
    s.close()
