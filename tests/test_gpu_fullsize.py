"""-m gpu: BASELINE.json's full sizes. The oracle is run once where it finishes in tens of
seconds (all host threads; bit-identical to 1 thread); beyond that, size-independent
properties: exact power-of-two scaling (KA4), fused == reference sequence bit-for-bit."""
import numpy as np
import pytest

from gpu_util import err_rtol, assert_bits_equal

pytestmark = pytest.mark.gpu


def test_c3_3d_512_fp32_one_cycle_vs_oracle(mgp, orc):
    s = mgp.MultigridCUDA(512, "float", dim=3, out=False)
    o = orc.Oracle(512, "float", 3, nthreads=orc.lib().orc_max_threads())
    es, eo = s.step(), o.step()
    assert abs(es - eo) <= err_rtol(s.size ** s.dim) * eo
    assert_bits_equal(s.psi.download(), o.psi, "512^3 psi after 1 cycle")
    assert_bits_equal(s.Rs[256].download(), o.buffer(orc.BUF_R, 256), "Rs[256]")
    assert_bits_equal(s.Vs[256].download(), o.buffer(orc.BUF_V, 256), "Vs[256]")
    s.close()


def test_c2_2d_4096_fp32_two_cycles_vs_oracle(mgp, orc):
    s = mgp.MultigridCUDA(4096, "float", dim=2, out=False)
    o = orc.Oracle(4096, "float", 2, nthreads=orc.lib().orc_max_threads())
    for _ in range(2):
        es, eo = s.step(), o.step()
        assert abs(es - eo) <= err_rtol(s.size ** s.dim) * eo
    assert_bits_equal(s.psi.download(), o.psi, "4096^2 psi after 2 cycles")
    s.close()


def test_c5_2d_2048_fp64_two_cycles_vs_oracle(mgp, orc):
    s = mgp.MultigridCUDA(2048, "double", dim=2, out=False)
    o = orc.Oracle(2048, "double", 2, nthreads=orc.lib().orc_max_threads())
    for _ in range(2):
        es, eo = s.step(), o.step()
        assert abs(es - eo) <= err_rtol(s.size ** s.dim) * eo
    assert_bits_equal(s.psi.download(), o.psi, "2048^2 psi after 2 cycles")
    s.close()


@pytest.mark.parametrize("dim,size,real", [(3, 512, "float"), (2, 4096, "float")])
def test_fullsize_properties(mgp, dim, size, real):
    a = mgp.MultigridCUDA(size, real, dim=dim, out=False)
    b = mgp.MultigridCUDA(size, real, dim=dim, out=False)
    # KA4: scaling f and psi0 by a power of two scales every field exactly
    f, psi = a.f.download(), a.psi.download()
    b.f.upload(f * np.float32(8)); b.psi.upload(psi * np.float32(8))
    del f, psi
    ea = [a.step() for _ in range(2)]
    eb = [b.step() for _ in range(2)]
    pa = a.psi.download()
    assert_bits_equal(b.psi.download(), pa * np.float32(8), "scaled run")
    assert np.allclose(eb, [8 * e for e in ea], rtol=1e-12, atol=0)
    b.close()
    # fused path == the reference operator sequence, bit for bit, at full size
    c = mgp.MultigridCUDA(size, real, dim=dim, out=False)
    c.set_mode(mgp.MODE_REFSEQ)
    ec = [c.step() for _ in range(2)]
    assert np.allclose(ec, ea, rtol=1e-12, atol=0)
    assert_bits_equal(c.psi.download(), pa, "refseq vs fused")
    a.close(); c.close()
