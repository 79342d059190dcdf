"""-m gpu: the relaxation weight omega (mg_set_omega). The reference is omega = 1 (cpu-raw.lua:34-44,176-184) and that must
stay bit-identical; any other value is a labelled extension checked against a numpy composition of the oracle's operators
(u + omega (J(u) - u)) to floating-point tolerance, and it must do what it is for: converge where omega = 1 does not."""
import numpy as np
import pytest

from gpu_util import assert_bits_equal

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,size,real", [(2, 64, "double"), (3, 32, "float"), (3, 128, "float")])
def test_omega_one_is_the_reference_path_bit_for_bit(mgp, dim, size, real):
    a = mgp.MultigridCUDA(size, real, dim=dim, out=False)
    b = mgp.MultigridCUDA(size, real, dim=dim, out=False)
    b.set_omega(1.0)
    for _ in range(2):
        ea, eb = a.step(), b.step()
        assert ea == eb
    assert_bits_equal(a.psi.download(), b.psi.download(), "omega = 1.0 set explicitly")
    a.close(); b.close()


def weighted_vcycle(orc, k, dim, u, f, h, L, smooth, omega, Vs):
    """twoGrid (cpu-raw.lua:186-237) with the smoother replaced by u + omega (J(u) - u), in numpy double"""
    def sm(u):
        j = orc.jacobi(dim, k, u, f, h, 8)
        return (u + omega * (j - u)).astype(u.dtype)
    if L == 1:
        return sm(u)
    for _ in range(smooth):
        u = sm(u)
    r = orc.residual(dim, k, f, u, h, 8)
    R = orc.restrict(dim, k, r)
    V = weighted_vcycle(orc, k, dim, Vs[L // 2], R, 2 * h, L // 2, smooth, omega, Vs)
    Vs[L // 2] = V
    u = orc.add_to(k, u, orc.prolong(dim, k, V))
    for _ in range(smooth):
        u = sm(u)
    return u


@pytest.mark.parametrize("dim,size,omega", [(2, 32, 0.8), (3, 16, 6.0 / 7.0)])
def test_weighted_cycle_matches_a_numpy_restatement(mgp, orc, dim, size, omega):
    s = mgp.MultigridCUDA(size, "double", dim=dim, out=False)
    s.set_omega(omega)
    k = orc.REAL_NAMES["double"]
    f, u = s.f.download(), s.psi.download()
    Vs = {L: np.zeros((L,) * dim) for L in (2 ** i for i in range(0, 12)) if L < size}
    for cyc in range(3):
        s.vcycle()
        u = weighted_vcycle(orc, k, dim, u, f, 1.0 / size, size, 7, omega, Vs)
        got = s.psi.download()
        assert np.allclose(got, u, rtol=1e-11, atol=1e-11 * np.abs(u).max()), (cyc, np.abs(got - u).max())
    s.close()


def test_weighted_jacobi_converges_where_the_reference_does_not(mgp):
    """3-D 64^3 fp64, point source, coarse corrections re-zeroed every cycle (the cpu.lua form): with omega = 1 the top
    mode is almost undamped and 25 cycles barely move the residual (850 cycles to 1e-8); omega = 6/7
    damps it and the same 25 cycles gain orders of magnitude."""
    res = {}
    for omega in (1.0, 6.0 / 7.0):
        s = mgp.MultigridCUDA(64, "double", dim=3, out=False)
        s.set_omega(omega)
        r0 = s.residual_norm()
        for _ in range(25):
            s.zero_corrections()
            s.vcycle()
        res[omega] = s.residual_norm() / r0
        s.close()
    print("residual reduction after 25 cycles:", res)
    assert res[6.0 / 7.0] < 1e-3 and res[6.0 / 7.0] < 1e-2 * res[1.0], res


def test_omega_is_validated(mgp):
    s = mgp.MultigridCUDA(64, "float", dim=3, out=False)
    for bad in (0.0, 2.0, -1.0, float("nan")):
        with pytest.raises(mgp.MGError):
            s.set_omega(bad)
    s.close()
    g = mgp.MultigridCUDA(128, "float", dim=3, out=False, local_slabs=2)
    with pytest.raises(mgp.MGError):
        g.set_omega(0.9)
    g.set_omega(1.0)
    g.close()
