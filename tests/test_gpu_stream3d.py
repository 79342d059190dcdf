"""-m gpu: the temporally blocked, TMA-fed streaming smoother (mg_stream3d.cuh) against the
composition of reference operators it replaces, bit for bit, for every sweep count S, with and
without the fused prolong+add / residual+restrict, every real kind, several z-chunkings."""
import numpy as np
import pytest

from gpu_util import KINDS, assert_bits_equal, rand_field, to_dev, to_host

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def solvers(mgp):
    cache = {}

    def get(real):
        if real not in cache:
            s = mgp.MultigridCUDA(256, real, dim=3, out=False)
            s.set_option("stream_min_L", 64)
            cache[real] = s
        return cache[real]
    yield get
    for s in cache.values():
        s.close()


def ref_sweeps(orc, k, u, f, h, n):
    for _ in range(n):
        u = orc.jacobi(3, k, u, f, h, 8)
    return u


@pytest.mark.parametrize("real", KINDS)
@pytest.mark.parametrize("L,tz", [(64, 0), (64, 16), (128, 0), (128, 8), (128, 64), (256, 0)])
def test_streaming_passes(solvers, orc, real, L, tz):
    s = solvers(real)
    k = orc.REAL_NAMES[real]
    rng = np.random.default_rng(L + tz)
    h = 1.0 / L
    u = rand_field(rng, 3, L, s.dtype)
    f = rand_field(rng, 3, L, s.dtype) * s.dtype(L * L)
    V = rand_field(rng, 3, L // 2, s.dtype)
    s.set_option("tz", tz)
    for S in (1, 2, 3, 4):
        s.set_option("tb", S)
        # plain: n = S sweeps in one pass, and 7 sweeps split into passes of <= S
        for n in (S, 7):
            du, df = to_dev(u), to_dev(f)
            s.inPlaceIterativeSolver(L, du, df, h, n)
            assert_bits_equal(to_host(du), ref_sweeps(orc, k, u, f, h, n), f"plain S={S} n={n} L={L} tz={tz}")
        # prolong + add fused into the first pass
        du, df, dV = to_dev(u), to_dev(f), to_dev(V)
        s.prolong_add_smooth(L, du, df, h, S, dV)
        w = ref_sweeps(orc, k, orc.add_to(k, u, orc.prolong(3, k, V)), f, h, S)
        assert_bits_equal(to_host(du), w, f"PRO S={S} L={L} tz={tz}")
        # residual + restriction fused into the last pass
        for n in (S, 7):
            du, dR = to_dev(u), to_dev(np.zeros_like(V))
            s.smooth_residual_restrict(L, du, df, h, n, dR)
            w = ref_sweeps(orc, k, u, f, h, n)
            assert_bits_equal(to_host(du), w, f"RES u S={S} n={n} L={L} tz={tz}")
            assert_bits_equal(to_host(dR), orc.restrict(3, k, orc.residual(3, k, f, w, h, 8)),
                              f"RES R S={S} n={n} L={L} tz={tz}")
    s.set_option("tz", 0)


@pytest.mark.parametrize("real", KINDS)
def test_point_source_boundary_interaction(solvers, orc, real):
    """Dirichlet rule under temporal blocking: a source next to the boundary, many sweeps, so
    that the zero exterior is exercised at every pipeline stage."""
    s = solvers(real)
    k = orc.REAL_NAMES[real]
    L, h = 64, 1.0 / 64
    u = np.zeros((L, L, L), s.dtype)
    f = np.zeros((L, L, L), s.dtype)
    for idx in ((0, 0, 0), (L - 1, L - 1, L - 1), (0, L - 1, 31), (63, 0, 32), (31, 32, 0)):
        u[idx] = 1e6
        f[idx] = -1e6
    for S in (1, 2, 3, 4):
        s.set_option("tb", S)
        du, df = to_dev(u), to_dev(f)
        s.inPlaceIterativeSolver(L, du, df, h, 12)
        assert_bits_equal(to_host(du), ref_sweeps(orc, k, u, f, h, 12), f"boundary S={S}")


@pytest.mark.parametrize("flags", [0, 4])
def test_branch_free_division_and_guarded_rerun(solvers, orc, flags):
    """fp32: the pass is the branch-free kernel (Markstein division everywhere, sticky guard) followed by the
    guarded re-run kernel. flags = 4 forces the re-run of every pass; tiny but non-zero numerators (where the
    Markstein sequence is NOT exact) must trip the sticky guard by themselves. Bit for bit either way."""
    s = solvers("float")
    k = orc.REAL_NAMES["float"]
    L, h = 128, 1.0 / 128
    s.set_option("fast_min_L", 64)
    s.set_option("stream_flags", flags)
    rng = np.random.default_rng(7)
    try:
        for scale in (1.0, 1e-33, 1e-38):
            u = (rand_field(rng, 3, L, s.dtype).astype(np.float64) * scale).astype(s.dtype)
            f = (rand_field(rng, 3, L, s.dtype).astype(np.float64) * scale * L * L).astype(s.dtype)
            V = (rand_field(rng, 3, L // 2, s.dtype).astype(np.float64) * scale).astype(s.dtype)
            u[::3, ::5, ::7] = 0          # exact zeros stay on the fast path
            for S in (3, 4):
                s.set_option("tb", S)
                du, df = to_dev(u), to_dev(f)
                s.inPlaceIterativeSolver(L, du, df, h, S)
                assert_bits_equal(to_host(du), ref_sweeps(orc, k, u, f, h, S), f"plain S={S} scale={scale} flags={flags}")
                du, dV = to_dev(u), to_dev(V)
                s.prolong_add_smooth(L, du, df, h, S, dV)
                w = ref_sweeps(orc, k, orc.add_to(k, u, orc.prolong(3, k, V)), f, h, S)
                assert_bits_equal(to_host(du), w, f"PRO S={S} scale={scale} flags={flags}")
            du, dR = to_dev(u), to_dev(np.zeros_like(V))
            s.set_option("tb", 4)
            s.smooth_residual_restrict(L, du, df, h, 7, dR)
            w = ref_sweeps(orc, k, u, f, h, 7)
            assert_bits_equal(to_host(du), w, f"RES u scale={scale} flags={flags}")
            assert_bits_equal(to_host(dR), orc.restrict(3, k, orc.residual(3, k, f, w, h, 8)), f"RES R scale={scale} flags={flags}")
    finally:
        s.set_option("stream_flags", 0)
        s.set_option("fast_min_L", 256)
        s.set_option("tb", 4)


@pytest.mark.parametrize("ncta", [8, 40, 3])
def test_lock_step_partitions(solvers, orc, ncta):
    """The work partitions of the streaming smoother (Stream3DArgs::ncol ...): whole tile columns in waves + a balanced
    remainder (more tiles than CTAs), whole columns + helper CTAs on the top planes (fewer), forced at 256^3 (35 tiles)
    by overriding the CTA count. Any partition must give the same bits."""
    s = solvers("float")
    k = orc.REAL_NAMES["float"]
    L, h = 256, 1.0 / 256
    rng = np.random.default_rng(ncta)
    u = rand_field(rng, 3, L, s.dtype)
    f = rand_field(rng, 3, L, s.dtype) * s.dtype(L * L)
    V = rand_field(rng, 3, L // 2, s.dtype)
    s.set_option("ncta", ncta)
    s.set_option("tb", 4)
    try:
        du, df, dV, dR = to_dev(u), to_dev(f), to_dev(V), to_dev(np.zeros_like(V))
        s.prolong_add_smooth(L, du, df, h, 7, dV)
        w = ref_sweeps(orc, k, orc.add_to(k, u, orc.prolong(3, k, V)), f, h, 7)
        assert_bits_equal(to_host(du), w, f"PRO + 7 sweeps, ncta={ncta}")
        s.smooth_residual_restrict(L, du, df, h, 7, dR)
        w2 = ref_sweeps(orc, k, w, f, h, 7)
        assert_bits_equal(to_host(du), w2, f"7 sweeps + RES (u), ncta={ncta}")
        assert_bits_equal(to_host(dR), orc.restrict(3, k, orc.residual(3, k, f, w2, h, 8)), f"7 sweeps + RES (R), ncta={ncta}")
    finally:
        s.set_option("ncta", 0)
