"""Pins the CPU oracle (oracle/mg_oracle.c) to the hand-derived known answers of SURVEY.md
section 8(c) (KA1-KA5, derived from cpu-raw.lua:8-73,186-258 by hand) and to the committed
golden fixtures. CPU only."""
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
KINDS = ["double", "float", "float_acc64"]


@pytest.mark.parametrize("real", KINDS)
def test_ka1_coarsest_solve(orc, real):
    # L = 1 with h = 1: u = -f/4 (cpu-raw.lua:190-196 with :24-31); 3-D: -f/6
    k = orc.REAL_NAMES[real]
    f = np.array([[3.0]], orc.np_dtype(k))
    assert orc.jacobi(2, k, np.zeros_like(f), f, 1.0)[0, 0] == -0.75
    f3 = np.array([[[3.0]]], orc.np_dtype(k))
    assert orc.jacobi(3, k, np.zeros_like(f3), f3, 1.0)[0, 0, 0] == orc.np_dtype(k)(-0.5)
    # general h: u = -f h^2 / 4
    assert orc.jacobi(2, k, np.zeros_like(f), f, 0.25)[0, 0] == -3.0 * 0.0625 / 4


@pytest.mark.parametrize("real", KINDS)
def test_ka2_size2_one_sweep(orc, real):
    k = orc.REAL_NAMES[real]
    f, psi = orc.init_cells(2, k, 2)
    assert f.ravel().tolist() == [0, 0, 0, -1e6] and psi.ravel().tolist() == [0, 0, 0, 1e6]
    u = orc.jacobi(2, k, psi, f, 0.5)
    assert u.ravel().tolist() == [0, 250000, 250000, 62500]


@pytest.mark.parametrize("real", KINDS)
@pytest.mark.parametrize("N", [8, 64, 256])
def test_ka3_first_two_sweeps(orc, real, N):
    k = orc.REAL_NAMES[real]
    f, psi = orc.init_cells(2, k, N)
    c, h = N // 2, 1.0 / N
    assert f[c, c] == -1e6 and np.count_nonzero(f) == 1
    u1 = orc.jacobi(2, k, psi, f, h)
    assert u1[c, c] == 250000.0 / N**2
    for dj, di in ((0, 1), (0, -1), (1, 0), (-1, 0)):
        assert u1[c + dj, c + di] == 250000.0
    assert np.count_nonzero(u1) == 5
    u2 = orc.jacobi(2, k, u1, f, h)
    dt = orc.np_dtype(k)
    assert u2[c, c] == dt(250000.0 + 250000.0 / N**2)
    assert u2[c, c + 1] == dt(62500.0 / N**2)
    assert u2[c + 1, c + 1] == 125000.0 and u2[c - 1, c + 1] == 125000.0
    assert u2[c, c + 2] == 62500.0 and u2[c - 2, c] == 62500.0


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("real", KINDS)
def test_ka4_exact_scaling_and_transfer_identities(orc, real, dim):
    k = orc.REAL_NAMES[real]
    rng = np.random.default_rng(7)
    L = 16
    shp = (L,) * dim
    u = rng.uniform(-1, 1, shp).astype(orc.np_dtype(k))
    f = rng.uniform(-1, 1, shp).astype(orc.np_dtype(k))
    h = 1.0 / L
    # power-of-two scaling commutes with every operator exactly
    for s in (2.0, 0.125, 1024.0):
        assert np.array_equal(orc.jacobi(dim, k, u * s, f * s, h), orc.jacobi(dim, k, u, f, h) * s)
        assert np.array_equal(orc.residual(dim, k, f * s, u * s, h), orc.residual(dim, k, f, u, h) * s)
    # restriction of a constant is the constant; prolong then restrict is the identity
    const = np.full(shp, 3.25, orc.np_dtype(k))
    assert np.all(orc.restrict(dim, k, const) == 3.25)
    V = rng.uniform(-1, 1, (L // 2,) * dim).astype(orc.np_dtype(k))
    back = orc.restrict(dim, k, orc.prolong(dim, k, V))
    if dim == 2:
        assert np.array_equal(back, V)
    else:  # 3x, 5x, 6x, 7x round in the left-to-right sum of 8 equal children
        assert np.allclose(back, V, rtol=4 * np.finfo(orc.np_dtype(k)).eps, atol=0)
    # prolongation is injection to the 2^dim children
    v = orc.prolong(dim, k, V)
    idx = tuple(np.arange(L) // 2 for _ in range(dim))
    assert np.array_equal(v, V[np.ix_(*idx)])


def test_ka5_corrections_persist_between_cycles(orc):
    # cpu-raw.lua:221-222: Vs[L/2] is NOT re-zeroed; cycle 2 starts from cycle 1's V
    a = orc.Oracle(32, "double", 2)
    a.step()
    V16 = a.buffer(orc.BUF_V, 16).copy()
    assert np.count_nonzero(V16) > 0
    psi1 = a.psi.copy()
    a.step()
    psi2_persist = a.psi.copy()
    b = orc.Oracle(32, "double", 2)
    b.step()
    assert np.array_equal(b.psi, psi1)
    L = 16
    while L >= 1:
        b.buffer(orc.BUF_V, L)[...] = 0   # what cpu.lua:138 does
        L //= 2
    b.step()
    assert not np.array_equal(b.psi, psi2_persist)


def test_residual_definition_matches_operator(orc):
    # r = f - A u with A u = (sum nb - 4u)/h^2 (cpu-raw.lua:53-56), checked against numpy
    rng = np.random.default_rng(3)
    L, h = 32, 1.0 / 32
    u = rng.uniform(-1, 1, (L, L))
    f = rng.uniform(-1, 1, (L, L))
    p = np.pad(u, 1)
    S = ((p[1:-1, :-2] + p[1:-1, 2:]) + p[:-2, 1:-1]) + p[2:, 1:-1]
    ref = f - (S / h**2 + (-4.0 / h**2) * u)
    assert np.array_equal(orc.residual(2, orc.REAL_F64, f, u, h), ref)
    Sj = (f - S / h**2) / (-4.0 / h**2)
    assert np.array_equal(orc.jacobi(2, orc.REAL_F64, u, f, h), Sj)


def test_run_is_capped_at_two_cycles_and_stops_on_accuracy(orc):
    o = orc.Oracle(64, "double", 2)
    errs = o.run()  # cpu-raw.lua:245 `for iter=1,2`
    assert len(errs) == 2
    o = orc.Oracle(64, "double", 2)
    assert len(o.run(max_cycles=5, accuracy=1e3)) == 2  # err[1] = 800 < 1e3 stops the loop (cpu-raw.lua:256)


def test_threads_do_not_change_results(orc):
    a, b = orc.Oracle(32, "float", 3, nthreads=1), orc.Oracle(32, "float", 3, nthreads=4)
    for _ in range(2):
        ea, eb = a.step(), b.step()
        assert ea == eb
    assert np.array_equal(a.psi, b.psi)


def test_gauss_seidel_is_lexicographic(orc):
    # cpu-raw.lua:22-32 + call2D order (j outer, i inner): first cell sees old values only
    rng = np.random.default_rng(5)
    u = rng.uniform(-1, 1, (8, 8))
    f = rng.uniform(-1, 1, (8, 8))
    g = orc.gauss_seidel(2, orc.REAL_F64, u, f, 0.125)
    j = orc.jacobi(2, orc.REAL_F64, u, f, 0.125)
    assert g[0, 0] == j[0, 0] and not np.array_equal(g, j)


@pytest.mark.parametrize("path", sorted(p for p in glob.glob(os.path.join(HERE, "golden", "*.npz"))
                                        if not os.path.basename(p).startswith("ref")),   # ref*: test_reference_source.py
                         ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_matches_golden(orc, path):
    g = np.load(path)
    dim, size, kind, cycles = (int(x) for x in g["meta"])
    o = orc.Oracle(size, kind, dim)
    errs = []
    for c in range(cycles):
        errs.append(o.step())
        if c == 0:
            assert np.array_equal(o.psi, g["psi_first"])
    assert errs == g["errs"].tolist()
    assert np.array_equal(o.psi, g["psi_last"])
    L = size // 2
    while L >= 1:
        assert np.array_equal(o.buffer(orc.BUF_R, L), g[f"R{L}"])
        assert np.array_equal(o.buffer(orc.BUF_V, L), g[f"V{L}"])
        L //= 2
    assert o.residual_rms() == float(g["residual_rms"])


def test_trace_sites_follow_the_reference_show_sites(orc):
    # cpu-raw.lua:192-235: per level visit 7 u (+7 f at the top), f u r R, [recursion], V v u, 7 u
    o = orc.Oracle(4, "double", 2)
    o.trace_enable()
    o.vcycle()
    names = "".join(n for n, _, _ in o.trace())
    top = "fu" * 7 + "fur" + "R"
    mid = "u" * 7 + "fur" + "R"
    bottom = "fu"
    assert names == top + mid + bottom + "Vvu" + "u" * 7 + "Vvu" + "u" * 7
