"""The one-CTA kernel with the sub-hierarchy in shared memory (mg_small.cuh, K-d3) and programmatic dependent
launch (mg_math.cuh, pdl_enter) are schedule choices: every field they leave behind -- psi, the persistent
corrections Vs and the restricted right-hand sides Rs at every level -- must equal, bit for bit, what the
global-memory walker / plain launches leave, and the CPU oracle (cpu-raw.lua:186-237)."""
import numpy as np
import pytest

from gpu_util import KINDS, assert_bits_equal

pytestmark = pytest.mark.gpu


def _levels(s, size):
    out = {"psi": s.psi.download()}
    L = size // 2
    while L >= 1:
        out[f"V{L}"] = s.Vs[L].download()
        out[f"R{L}"] = s.Rs[L].download()
        L //= 2
    return out


@pytest.mark.parametrize("real", KINDS)
@pytest.mark.parametrize("dim,size", [(2, 64), (2, 32), (2, 8), (2, 2), (3, 16), (3, 8), (3, 2)])
def test_shared_memory_small_kernel_equals_the_global_memory_walker(mgp, dim, size, real):
    got = {}
    for smem in (1, 0):
        s = mgp.MultigridCUDA(size, real, dim=dim, out=False)
        s.set_tuning(small_L=size)          # the whole V-cycle is the small-level kernel
        s.set_option("small_smem", smem)
        errs = [s.step() for _ in range(3)]  # three cycles: the corrections persist in between (F4)
        got[smem] = (errs, _levels(s, size))
        s.close()
    assert got[1][0] == got[0][0]
    for k, v in got[0][1].items():
        assert_bits_equal(got[1][1][k], v, f"{k} ({dim}-D {size} {real})")


@pytest.mark.parametrize("real", KINDS)
@pytest.mark.parametrize("dim,size", [(2, 64), (3, 16)])
def test_shared_memory_small_kernel_equals_the_oracle(mgp, orc, dim, size, real):
    s = mgp.MultigridCUDA(size, real, dim=dim, out=False)
    s.set_tuning(small_L=size)
    o = orc.Oracle(size, real, dim)
    for _ in range(2):
        s.vcycle()
        o.vcycle()
    assert_bits_equal(s.psi.download(), o.psi, "psi")
    L = size // 2
    while L >= 1:
        assert_bits_equal(s.Vs[L].download(), o.buffer(orc.BUF_V, L), f"Vs[{L}]")
        assert_bits_equal(s.Rs[L].download(), o.buffer(orc.BUF_R, L), f"Rs[{L}]")
        L //= 2
    s.close()


def test_shared_memory_small_kernel_with_smooth_counts_and_omega(mgp):
    """even / odd / zero sweep counts (ping-pong parity) and the labelled omega extension."""
    for dim, size in ((2, 64), (3, 16)):
        for smooth, omega in ((7, 1.0), (4, 1.0), (1, 1.0), (7, 0.8)):
            got = {}
            for smem in (1, 0):
                s = mgp.MultigridCUDA(size, "double", dim=dim, smooth=smooth, out=False)
                s.set_tuning(small_L=size)
                s.set_option("small_smem", smem)
                if omega != 1.0:
                    s.set_omega(omega)
                s.vcycle(); s.vcycle()
                got[smem] = _levels(s, size)
                s.close()
            for k, v in got[0].items():
                assert_bits_equal(got[1][k], v, f"{k} ({dim}-D smooth={smooth} omega={omega})")


@pytest.mark.parametrize("dim,size,real", [(3, 128, "float"), (3, 64, "double"), (2, 512, "float"), (2, 256, "double")])
def test_programmatic_dependent_launch_does_not_change_a_single_bit(mgp, dim, size, real):
    got = {}
    for pdl in (1, 0):
        for graph in (1, 0):
            s = mgp.MultigridCUDA(size, real, dim=dim, out=False)
            s.set_tuning(use_graph=graph)
            s.set_option("pdl", pdl)
            errs = [s.step() for _ in range(4)]
            got[(pdl, graph)] = (errs, _levels(s, size))
            s.close()
    ref = got[(0, 0)]
    for key, (errs, lv) in got.items():
        assert errs == ref[0], key
        for k, v in ref[1].items():
            assert_bits_equal(lv[k], v, f"{k} pdl,graph={key}")


@pytest.mark.parametrize("real", KINDS)
def test_block_temporal_mid_level_kernel_does_not_change_a_single_bit(mgp, real):
    """mg_block3d.cuh: up to 4 sweeps per launch on the 32^3 / 64^3 (optionally 128^3) levels, against one sweep per
    launch (block_max_L = 0) -- every field of the hierarchy, three cycles."""
    got = {}
    for bmax in (0, 32, 64, 128):
        s = mgp.MultigridCUDA(128, real, dim=3, out=False)
        s.set_option("block_max_L", bmax)
        errs = [s.step() for _ in range(3)]
        got[bmax] = (errs, _levels(s, 128))
        s.close()
    for bmax in (32, 64, 128):
        assert got[bmax][0] == got[0][0], bmax
        for k, v in got[0][1].items():
            assert_bits_equal(got[bmax][1][k], v, f"{k} block_max_L={bmax} {real}")


@pytest.mark.parametrize("smooth", [1, 2, 3, 4, 5, 8])
def test_block_temporal_kernel_sweep_counts(mgp, smooth):
    """every split of the sweep count into passes of <= 4, with and without the fused prolongation"""
    got = {}
    for bmax in (0, 64):
        s = mgp.MultigridCUDA(64, "float", dim=3, smooth=smooth, out=False)
        s.set_option("block_max_L", bmax)
        s.vcycle(); s.vcycle()
        got[bmax] = _levels(s, 64)
        s.close()
    for k, v in got[0].items():
        assert_bits_equal(got[64][k], v, f"{k} smooth={smooth}")
