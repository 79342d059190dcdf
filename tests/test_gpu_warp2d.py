"""-m gpu: the 2-D warp-synchronous, temporally blocked smoother (mg_warp2d.cuh) against the
composition of reference operators it replaces (cpu-raw.lua:198-236), bit for bit: every sweep
count 1..7, with and without fused prolong+add / residual+restrict, every real kind."""
import numpy as np
import pytest

from gpu_util import KINDS, assert_bits_equal, rand_field, to_dev, to_host

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def solvers(mgp):
    cache = {}

    def get(real):
        if real not in cache:
            s = mgp.MultigridCUDA(1024, real, dim=2, out=False)
            s.set_option("warp2d_min_L", 32)
            cache[real] = s
        return cache[real]
    yield get
    for s in cache.values():
        s.close()


def ref_sweeps(orc, k, u, f, h, n):
    for _ in range(n):
        u = orc.jacobi(2, k, u, f, h, 8)
    return u


@pytest.mark.parametrize("real", KINDS)
@pytest.mark.parametrize("L,ty", [(32, 0), (128, 0), (128, 16), (512, 0), (1024, 64)])
def test_warp_streaming_passes(solvers, orc, real, L, ty):
    s = solvers(real)
    k = orc.REAL_NAMES[real]
    rng = np.random.default_rng(L + ty)
    h = 1.0 / L
    u = rand_field(rng, 2, L, s.dtype)
    f = rand_field(rng, 2, L, s.dtype) * s.dtype(L * L)
    V = rand_field(rng, 2, L // 2, s.dtype)
    s.set_option("ty", ty)
    for S in (1, 2, 3, 4, 5, 6, 7):
        s.set_option("tb2", S)
        du, df = to_dev(u), to_dev(f)
        s.inPlaceIterativeSolver(L, du, df, h, 7)
        assert_bits_equal(to_host(du), ref_sweeps(orc, k, u, f, h, 7), f"plain S={S} L={L} ty={ty}")
        du, dV = to_dev(u), to_dev(V)
        s.prolong_add_smooth(L, du, df, h, S, dV)
        w = ref_sweeps(orc, k, orc.add_to(k, u, orc.prolong(2, k, V)), f, h, S)
        assert_bits_equal(to_host(du), w, f"PRO S={S} L={L} ty={ty}")
        du, dR = to_dev(u), to_dev(np.zeros_like(V))
        s.smooth_residual_restrict(L, du, df, h, S, dR)
        w = ref_sweeps(orc, k, u, f, h, S)
        assert_bits_equal(to_host(du), w, f"RES u S={S} L={L} ty={ty}")
        assert_bits_equal(to_host(dR), orc.restrict(2, k, orc.residual(2, k, f, w, h, 8)), f"RES R S={S} L={L} ty={ty}")
    s.set_option("ty", 0)


@pytest.mark.parametrize("real", KINDS)
def test_sources_on_the_boundary(solvers, orc, real):
    s = solvers(real)
    k = orc.REAL_NAMES[real]
    L, h = 256, 1.0 / 256
    u = np.zeros((L, L), s.dtype)
    f = np.zeros((L, L), s.dtype)
    for idx in ((0, 0), (L - 1, L - 1), (0, 111), (112, 0), (255, 113), (120, 255)):
        u[idx] = 1e6
        f[idx] = -1e6
    for S in (1, 4, 7):
        s.set_option("tb2", S)
        du, df = to_dev(u), to_dev(f)
        s.inPlaceIterativeSolver(L, du, df, h, 14)
        assert_bits_equal(to_host(du), ref_sweeps(orc, k, u, f, h, 14), f"boundary S={S}")
