"""-m gpu: every reference operator (rows a1-a7, a11, a12 of SURVEY section 8) through the
C ABI, bit-for-bit against the CPU oracle, for the three real kinds and both dimensions,
including the degenerate sizes (L = 1, 2) the hierarchy reaches."""
import numpy as np
import pytest

from gpu_util import KINDS, assert_bits_equal, rand_field, to_dev, to_host

pytestmark = pytest.mark.gpu

SIZES = {2: [1, 2, 4, 8, 32, 128, 512], 3: [1, 2, 4, 8, 32, 64]}


@pytest.fixture(scope="module")
def solvers(mgp):
    cache = {}

    def get(dim, real):
        key = (dim, real)
        if key not in cache:
            cache[key] = mgp.MultigridCUDA(max(SIZES[dim]), real, dim=dim, out=False)
        return cache[key]
    yield get
    for s in cache.values():
        s.close()


@pytest.mark.parametrize("real", KINDS)
@pytest.mark.parametrize("dim", [2, 3])
def test_init_cells(mgp, orc, dim, real):
    for size in (1, 2, 8, 64):
        s = mgp.MultigridCUDA(size, real, dim=dim, out=False)
        f, psi = orc.init_cells(dim, orc.REAL_NAMES[real], size)
        assert_bits_equal(s.f.download(), f, "f")
        assert_bits_equal(s.psi.download(), psi, "psi")
        s.close()


@pytest.mark.parametrize("real", KINDS)
@pytest.mark.parametrize("dim", [2, 3])
def test_jacobi_and_residual(solvers, orc, dim, real):
    s = solvers(dim, real)
    k = orc.REAL_NAMES[real]
    rng = np.random.default_rng(1234)
    for L in SIZES[dim]:
        h = 1.0 / L
        for case in ("random", "point", "big"):
            if case == "point":
                f, u = orc.init_cells(dim, k, L)
            else:
                scale = 1e6 if case == "big" else 1.0
                u = rand_field(rng, dim, L, s.dtype) * s.dtype(scale)
                f = rand_field(rng, dim, L, s.dtype) * s.dtype(scale * L * L)
            du, df = to_dev(u), to_dev(f)
            dd = to_dev(np.zeros_like(u))
            s.jacobi(L, dd, du, df, h)
            assert_bits_equal(to_host(dd), orc.jacobi(dim, k, u, f, h), f"jacobi {case} L={L}")
            s.residual(L, dd, df, du, h)
            assert_bits_equal(to_host(dd), orc.residual(dim, k, f, u, h), f"residual {case} L={L}")
            # a coarse level of a finer hierarchy: h is not 1/L (cpu-raw.lua:222 passes 2*h)
            s.jacobi(L, dd, du, df, 4 * h)
            assert_bits_equal(to_host(dd), orc.jacobi(dim, k, u, f, 4 * h), f"jacobi 4h L={L}")


@pytest.mark.parametrize("real", KINDS)
@pytest.mark.parametrize("dim", [2, 3])
def test_transfer_operators_and_add(solvers, orc, dim, real):
    s = solvers(dim, real)
    k = orc.REAL_NAMES[real]
    rng = np.random.default_rng(99)
    for L in SIZES[dim]:
        if L < 2:
            continue
        r = rand_field(rng, dim, L, s.dtype)
        V = rand_field(rng, dim, L // 2, s.dtype)
        dr, dV = to_dev(r), to_dev(V)
        dR, dv = to_dev(np.zeros_like(V)), to_dev(np.zeros_like(r))
        s.restrict(L // 2, dR, dr)
        assert_bits_equal(to_host(dR), orc.restrict(dim, k, r), f"restrict L={L}")
        s.prolong(L // 2, dv, dV)
        assert_bits_equal(to_host(dv), orc.prolong(dim, k, V), f"prolong L={L}")
        u = rand_field(rng, dim, L, s.dtype)
        du = to_dev(u)
        s.add_to(u.size, du, dv)
        assert_bits_equal(to_host(du), orc.add_to(k, u, orc.prolong(dim, k, V)), f"addTo L={L}")


@pytest.mark.parametrize("real", KINDS)
@pytest.mark.parametrize("dim", [2, 3])
def test_frob_err(mgp, orc, dim, real):
    size = 64 if dim == 2 else 16
    s = mgp.MultigridCUDA(size, real, dim=dim, out=False)
    k = orc.REAL_NAMES[real]
    rng = np.random.default_rng(5)
    psi, old = rand_field(rng, dim, size, s.dtype), rand_field(rng, dim, size, s.dtype)
    s.psi.upload(psi)
    s.psiOld.upload(old)
    want, eb = orc.frob_err(dim, k, psi, old)
    got = s.frob_err()
    assert abs(got - want) <= 1e-13 * want          # summation order differs (SURVEY 8(a) notes)
    s.set_mode(mgp.MODE_REFSEQ)                      # materialises errorBuf as the reference does
    assert abs(s.frob_err() - want) <= 1e-13 * want
    assert_bits_equal(s.errorBuf.download(), eb, "errorBuf")
    s.close()


@pytest.mark.parametrize("real", KINDS)
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("mode", ["fused", "refseq"])
def test_in_place_solver_n_sweeps(solvers, mgp, orc, dim, real, mode):
    # n x inPlaceIterativeSolver (cpu-raw.lua:176-184); the fused path ping-pongs instead
    s = solvers(dim, real)
    s.set_mode(mgp.MODE_REFSEQ if mode == "refseq" else mgp.MODE_FUSED)
    k = orc.REAL_NAMES[real]
    rng = np.random.default_rng(11)
    for L in SIZES[dim][:6]:
        h = 1.0 / L
        u, f = rand_field(rng, dim, L, s.dtype), rand_field(rng, dim, L, s.dtype) * s.dtype(L * L)
        for n in (1, 2, 7):
            du, df = to_dev(u), to_dev(f)
            s.inPlaceIterativeSolver(L, du, df, h, n)
            want = u
            for _ in range(n):
                want = orc.jacobi(dim, k, want, f, h)
            assert_bits_equal(to_host(du), want, f"{n} sweeps L={L}")
    s.set_mode(mgp.MODE_FUSED)


@pytest.mark.parametrize("real", KINDS)
@pytest.mark.parametrize("dim", [2, 3])
def test_fused_building_blocks(solvers, orc, dim, real):
    """mg_smooth_residual_restrict == n x Jacobi ; calcResidual ; reduceResidual
       mg_prolong_add_smooth       == expandResidual ; addTo ; n x Jacobi   (cpu-raw.lua:198-236)"""
    s = solvers(dim, real)
    k = orc.REAL_NAMES[real]
    rng = np.random.default_rng(21)
    for L in [x for x in SIZES[dim] if x >= 2]:
        h = 1.0 / L
        u, f = rand_field(rng, dim, L, s.dtype), rand_field(rng, dim, L, s.dtype) * s.dtype(L * L)
        V = rand_field(rng, dim, L // 2, s.dtype)
        for n in (0, 1, 3, 7):
            du, df, dR = to_dev(u), to_dev(f), to_dev(np.zeros_like(V))
            s.smooth_residual_restrict(L, du, df, h, n, dR)
            w = u
            for _ in range(n):
                w = orc.jacobi(dim, k, w, f, h)
            assert_bits_equal(to_host(du), w, f"pre u n={n} L={L}")
            assert_bits_equal(to_host(dR), orc.restrict(dim, k, orc.residual(dim, k, f, w, h)), f"pre R n={n} L={L}")
            du, dV = to_dev(u), to_dev(V)
            s.prolong_add_smooth(L, du, df, h, n, dV)
            w = orc.add_to(k, u, orc.prolong(dim, k, V))
            for _ in range(n):
                w = orc.jacobi(dim, k, w, f, h)
            assert_bits_equal(to_host(du), w, f"post u n={n} L={L}")
