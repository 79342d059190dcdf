"""CPU checks of the streaming smoother's pipeline logic (tools/emulate_stream3d.py, a numpy model of k_stream3d):
step/stage timing, active/emit predicates, masks and ring parity reproduce S plain Jacobi sweeps (+ residual and
restriction, + prolongation) exactly, and the TMA ring bookkeeping -- source plane t+2 and f plane t+1 issued on ONE
mbarrier per step -- never refills a slot in use nor reads an f plane whose mbarrier phase has not been awaited."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("emulate_stream3d", os.path.join(ROOT, "tools", "emulate_stream3d.py"))
em = importlib.util.module_from_spec(spec)
spec.loader.exec_module(em)


@pytest.mark.parametrize("NST", [1, 2, 3, 4, 5])
def test_ring_bookkeeping(NST):
    for TZ in (2, 4, 6, 16, 33, 64, 512):
        assert em.check_rings(NST, TZ)


@pytest.mark.parametrize("S,RES,PRO", [(2, False, False), (2, True, False), (3, False, True)])
def test_pipeline_equals_plain_sweeps(S, RES, PRO):
    rng = np.random.default_rng(1)
    L, TX, TY, TZ = 16, 8, 4, 8
    h = 1.0 / L
    src = rng.uniform(-1, 1, (L, L, L))
    f = rng.uniform(-1, 1, (L, L, L)) * L * L
    Vp = rng.uniform(-1, 1, (L // 2,) * 3)
    u = src + np.repeat(np.repeat(np.repeat(Vp, 2, 0), 2, 1), 2, 2) if PRO else src.copy()
    for _ in range(S):
        u = em.jacobi_ref(u, f, h)
    dst, Rout = em.emulate(src, f, h, S, RES, PRO, Vp, TX, TY, TZ)
    assert np.array_equal(dst, u)
    if RES:
        assert np.array_equal(Rout, em.restrict_ref(em.residual_ref(u, f, h)))
