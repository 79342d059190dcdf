"""-m gpu: the multi-GPU slab schedule (mg_slab.cuh) exercised on ONE device: the grid is cut
into P slabs inside the process, halo planes are exchanged between the slabs exactly as the
NCCL transport does, coarse levels are replicated. Because Jacobi is order independent, the
result must be bit-identical to the single-solver run (and hence to the oracle)."""
import numpy as np
import pytest

from gpu_util import assert_bits_equal, err_rtol, rand_field

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size,P,real", [(128, 2, "float"), (128, 4, "float"), (256, 8, "float"),
                                         (128, 2, "double"), (128, 4, "float_acc64"), (256, 4, "float")])
@pytest.mark.parametrize("p2p,min_planes", [(1, 8), (0, 8), (1, 32)])
def test_local_slab_group_matches_single_solver(mgp, monkeypatch, size, P, real, p2p, min_planes):
    """p2p=1: the smoother kernel itself stores its boundary planes into the neighbour slab's ghost
    planes (the fused NVLink exchange, here through same-device pointers); p2p=0: separate copies
    (what ncclSend/ncclRecv do)."""
    monkeypatch.setenv("MGPOISSON_SLAB_MIN_PLANES", str(min_planes))   # replication threshold (default 32)
    if size // P < min_planes:
        pytest.skip("top level would not be distributed")
    one = mgp.MultigridCUDA(size, real, dim=3, out=False)
    one.set_tuning(tb=4)
    one.set_option("stream_min_L", 64)
    grp = mgp.MultigridCUDA(size, real, dim=3, out=False, local_slabs=P)
    grp.set_option("slab_p2p", p2p)
    info = grp.slab_info()
    assert info["nranks"] == P and info["own_planes"] == size // P and info["ghost"] == 4
    assert_bits_equal(grp.f.download(), one.f.download(), "f after initCells")
    assert_bits_equal(grp.psi.download(), one.psi.download(), "psi after initCells")
    for cyc in range(3):
        e1, eg = one.step(), grp.step()
        assert abs(e1 - eg) <= err_rtol(size ** 3) * e1, (cyc, e1, eg)
        assert_bits_equal(grp.psi.download(), one.psi.download(), f"psi after cycle {cyc + 1}")
    # replicated coarse levels are identical too (rank 0's copy is read back)
    for L in (32, 16, 1):
        assert_bits_equal(grp.Vs[L].download(), one.Vs[L].download(), f"Vs[{L}]")
        assert_bits_equal(grp.Rs[L].download(), one.Rs[L].download(), f"Rs[{L}]")
    assert (grp.slab_info()["exchanges"] > 6) == (p2p == 0)
    # the true residual norm and the cpu.lua variant (corrections re-zeroed) are available on slabs as well
    r1, rg = one.residual_norm(), grp.residual_norm()
    assert abs(r1 - rg) <= err_rtol(size ** 3) * r1, (r1, rg)
    one.zero_corrections(); grp.zero_corrections()
    e1, eg = one.step(), grp.step()
    assert abs(e1 - eg) <= err_rtol(size ** 3) * e1
    assert_bits_equal(grp.psi.download(), one.psi.download(), "psi after a cycle with re-zeroed corrections")
    one.close(); grp.close()


@pytest.mark.parametrize("smooth", [1, 2, 4, 5, 8])
def test_local_slab_group_other_sweep_counts(mgp, monkeypatch, smooth):
    """smooth = 4 and 8 give an odd number of ping-pong passes per level visit with tb = 4 unless the slab schedule
    pads them (one extra, shorter, post-smoothing pass); the result is the same field bit for bit."""
    monkeypatch.setenv("MGPOISSON_SLAB_MIN_PLANES", "8")
    one = mgp.MultigridCUDA(128, "float", dim=3, out=False, smooth=smooth)
    grp = mgp.MultigridCUDA(128, "float", dim=3, out=False, local_slabs=4, smooth=smooth)
    for cyc in range(2):
        e1, eg = one.step(), grp.step()
        assert abs(e1 - eg) <= err_rtol(128 ** 3) * e1, (cyc, e1, eg)
    assert_bits_equal(grp.psi.download(), one.psi.download(), f"smooth={smooth}")
    with pytest.raises(mgp.MGError):
        grp.set_option("tb", 0)          # the slab schedule needs the streaming smoother
    one.close(); grp.close()


def test_local_slab_group_random_rhs_vs_oracle(mgp, orc, monkeypatch):
    monkeypatch.setenv("MGPOISSON_SLAB_MIN_PLANES", "8")
    size, P = 128, 4
    rng = np.random.default_rng(1234)
    grp = mgp.MultigridCUDA(size, "float", dim=3, out=False, local_slabs=P)
    o = orc.Oracle(size, "float", 3, nthreads=8)
    f = rand_field(rng, 3, size, np.float32) * np.float32(size * size)
    psi = rand_field(rng, 3, size, np.float32)
    grp.f.upload(f); grp.psi.upload(psi)
    o.f[...] = f; o.psi[...] = psi
    for _ in range(2):
        eg, eo = grp.step(), o.step()
        assert abs(eg - eo) <= err_rtol(size ** 3) * eo
    assert_bits_equal(grp.psi.download(), o.psi, "slab group vs oracle")
    grp.close()


def test_slab_partition_description(mgp, monkeypatch):
    monkeypatch.delenv("MGPOISSON_SLAB_MIN_PLANES", raising=False)
    lv = mgp.slab_partition(1024, 8)            # default: a level stays cut while every rank keeps >= 32 planes
    assert [x["L"] for x in lv if x["distributed"]] == [1024, 512, 256]
    monkeypatch.setenv("MGPOISSON_SLAB_MIN_PLANES", "8")
    lv = mgp.slab_partition(1024, 8)
    assert [x["L"] for x in lv if x["distributed"]] == [1024, 512, 256, 128, 64]
    assert lv[0]["planes_per_rank"] == 128 and lv[4]["planes_per_rank"] == 8
    assert all(not x["distributed"] for x in mgp.slab_partition(512, 1))


def test_slab_rejects_unsupported(mgp):
    with pytest.raises(mgp.MGError):
        mgp.MultigridCUDA(32, "float", dim=3, out=False, local_slabs=2)      # too small to cut
    with pytest.raises(mgp.MGError):
        mgp.MultigridCUDA(256, "float", dim=2, out=False, local_slabs=2)     # 3-D only
