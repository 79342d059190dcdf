"""-m gpu: the V-cycle (row a8), hierarchy (a9), driver + error metric (a11) through the C ABI
against the CPU oracle and the committed golden fixtures."""
import glob
import io
import os

import numpy as np
import pytest

from gpu_util import err_rtol, KINDS, assert_bits_equal, rand_field, to_dev, to_host

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def compare_hierarchy(s, o, orc, what):
    assert_bits_equal(s.psi.download(), o.psi, f"{what} psi")
    L = s.size // 2
    while L >= 1:
        assert_bits_equal(s.Rs[L].download(), o.buffer(orc.BUF_R, L), f"{what} Rs[{L}]")
        assert_bits_equal(s.Vs[L].download(), o.buffer(orc.BUF_V, L), f"{what} Vs[{L}]")
        L //= 2


@pytest.mark.parametrize("real", KINDS)
@pytest.mark.parametrize("dim,size", [(2, 1), (2, 2), (2, 16), (3, 2), (3, 8)])
def test_refseq_trace_matches_oracle_stage_by_stage(mgp, orc, dim, size, real):
    """The reference's own cross-variant check (debug dumps f,u,r,R,V,v per level,
    cpu-raw.lua:126-140,192-235), record by record, two cycles."""
    s = mgp.MultigridCUDA(size, real, dim=dim, out=False)
    s.set_debugging(True)
    o = orc.Oracle(size, real, dim)
    o.trace_enable()
    for cyc in range(2):
        es, eo = s.step(), o.step()
        assert abs(es - eo) <= err_rtol(size ** dim) * abs(eo) + 1e-300
    ts, to = s.trace(), o.trace()
    assert [(n, L) for n, L, _ in ts] == [(n, L) for n, L, _ in to]
    for i, ((n, L, a), (_, _, b)) in enumerate(zip(ts, to)):
        assert_bits_equal(a, b, f"trace[{i}] {n} L={L}")
    for nm in ("tmpU", "errorBuf", "psiOld"):
        assert_bits_equal(getattr(s, nm).download(), o.buffer(getattr(orc, "BUF_" + nm.upper()), size), nm)
    L = size
    while L >= 1:
        assert_bits_equal(s.rs[L].download(), o.buffer(orc.BUF_r, L), f"rs[{L}]")
        assert_bits_equal(s.vs[L].download(), o.buffer(orc.BUF_v, L), f"vs[{L}]")
        L //= 2
    s.close()


@pytest.mark.parametrize("real", KINDS)
@pytest.mark.parametrize("dim,size", [(2, 64), (2, 512), (3, 32), (3, 128)])
@pytest.mark.parametrize("mode", ["fused", "refseq"])
def test_vcycles_match_oracle(mgp, orc, dim, size, real, mode):
    s = mgp.MultigridCUDA(size, real, dim=dim, out=False)
    if mode == "refseq":
        s.set_mode(mgp.MODE_REFSEQ)
    o = orc.Oracle(size, real, dim, nthreads=8)
    for cyc in range(3):
        es, eo = s.step(), o.step()
        assert abs(es - eo) <= err_rtol(size ** dim) * abs(eo), (cyc, es, eo)
        compare_hierarchy(s, o, orc, f"cycle {cyc + 1}")
    assert abs(s.residual_norm() - o.residual_rms()) <= err_rtol(size ** dim) * o.residual_rms()
    s.close()


@pytest.mark.parametrize("dim,size,real", [(2, 256, "double"), (3, 64, "float"), (3, 64, "double")])
def test_random_rhs_and_guess(mgp, orc, dim, size, real):
    rng = np.random.default_rng(1234)          # seeded random RHS
    s = mgp.MultigridCUDA(size, real, dim=dim, out=False)
    o = orc.Oracle(size, real, dim, nthreads=8)
    f = rand_field(rng, dim, size, s.dtype) * s.dtype(size * size)
    psi = rand_field(rng, dim, size, s.dtype)
    s.f.upload(f); s.psi.upload(psi)
    o.f[...] = f; o.psi[...] = psi
    for cyc in range(2):
        es, eo = s.step(), o.step()
        assert abs(es - eo) <= err_rtol(size ** dim) * abs(eo)
    compare_hierarchy(s, o, orc, "random rhs")
    s.close()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(HERE, "golden", "ref_2d_*.npz")) +
                                        glob.glob(os.path.join(HERE, "golden", "refgpu_2d_*.npz"))),
                         ids=lambda p: os.path.basename(p)[:-4])
def test_cuda_matches_the_reference_source_run(mgp, path):
    """tests/golden/ref_2d_*.npz = what the reference's own cpu-raw.lua computes in run() (2 V-cycles) when executed by
    oracle/minilua.py (oracle/run_reference.py); refgpu_2d_*.npz = the same for gpu.lua with its own OpenCL kernel
    source compiled strictly by gcc (oracle/run_reference_gpu.py; real = float there is fp32 arithmetic, MG_REAL_F32).
    The CUDA path, through the C ABI, must give the same bits."""
    g = np.load(path)
    dim, size, kind, cycles = (int(x) for x in g["meta"])
    s = mgp.MultigridCUDA(size, kind, dim=dim, out=False)
    assert_bits_equal(s.f.download().ravel(), g["f0"], "f after init")
    assert_bits_equal(s.psi.download().ravel(), g["psi0"], "psi after init")
    for c in range(cycles):
        e = s.step()
        assert abs(e - g["errs"][c]) <= err_rtol(size ** dim) * g["errs"][c]
        if c == 0:
            assert_bits_equal(s.psi.download().ravel(), g["psi_after_cycle1"], "psi after cycle 1")
    assert_bits_equal(s.psi.download().ravel(), g["psi"], "psi after run()")
    assert_bits_equal(s.psiOld.download().ravel(), g["psiOld"], "psiOld after run()")
    L = size // 2
    while L >= 1:
        assert_bits_equal(s.Rs[L].download().ravel(), g[f"Rs{L}"], f"Rs[{L}]")
        assert_bits_equal(s.Vs[L].download().ravel(), g[f"Vs{L}"], f"Vs[{L}]")
        L //= 2
    s.close()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(HERE, "golden", "refcpu_2d_*.npz"))),
                         ids=lambda p: os.path.basename(p)[:-4])
def test_cuda_with_zeroed_corrections_matches_cpu_lua(mgp, path):
    """refcpu_2d_*.npz = the reference's cpu.lua (coarse corrections re-zeroed every cycle, cpu.lua:138; the solver
    of test/converge-multigrid-vs-krylov.lua) executed by oracle/minilua.py (oracle/run_reference_cpu.py).
    mg_zero_corrections() + mg_step(), cycle after cycle, must give the same bits."""
    g = np.load(path)
    dim, size, kind, steps = (int(x) for x in g["meta"])
    s = mgp.MultigridCUDA(size, kind, dim=dim, out=False)
    assert_bits_equal(s.psi.download().ravel(), g["psi0"], "psi after init")
    for c in range(steps):
        s.zero_corrections()
        e = s.step()
        assert abs(e - g["errs"][c]) <= err_rtol(size ** dim) * g["errs"][c]
        assert_bits_equal(s.psi.download().ravel(), g[f"psi{c + 1}"], f"psi after step {c + 1}")
    s.close()


@pytest.mark.parametrize("path", sorted(p for p in glob.glob(os.path.join(HERE, "golden", "*.npz"))
                                        if not os.path.basename(p).startswith("ref")),
                         ids=lambda p: os.path.basename(p)[:-4])
def test_cuda_matches_golden_fixtures(mgp, path):
    g = np.load(path)
    dim, size, kind, cycles = (int(x) for x in g["meta"])
    s = mgp.MultigridCUDA(size, kind, dim=dim, out=False)
    for c in range(cycles):
        e = s.step()
        assert abs(e - g["errs"][c]) <= 1e-12 * g["errs"][c]
        if c == 0:
            assert_bits_equal(s.psi.download(), g["psi_first"], "psi after cycle 1")
    assert_bits_equal(s.psi.download(), g["psi_last"], "psi after last cycle")
    L = size // 2
    while L >= 1:
        assert_bits_equal(s.Rs[L].download(), g[f"R{L}"], f"Rs[{L}]")
        assert_bits_equal(s.Vs[L].download(), g[f"V{L}"], f"Vs[{L}]")
        L //= 2
    assert abs(s.residual_norm() - float(g["residual_rms"])) <= 1e-12 * float(g["residual_rms"])
    s.close()


@pytest.mark.parametrize("dim,size", [(2, 512), (3, 128)])
def test_tuning_knobs_do_not_change_a_single_bit(mgp, dim, size):
    """temporal-blocking depth, tile chunking, small-level threshold and graph replay are schedule
    choices; per-point arithmetic is shared (mg_math.cuh), so every setting gives identical fields."""
    ref = None
    depths = (0, 1, 2, 3, 4) if dim == 3 else (0, 1, 2, 3, 5, 7)
    for tb in depths:
        for small_L in (1, 4, 16, 32):
            for graph in (0, 1):
                if tb > 1 and graph == 0 and small_L != 16:
                    continue
                s = mgp.MultigridCUDA(size, "float", dim=dim, out=False)
                s.set_tuning(small_L=small_L, use_graph=graph)
                s.set_option("tb" if dim == 3 else "tb2", tb)
                s.set_option("stream_min_L", 64)
                s.set_option("warp2d_min_L", 64)
                errs = [s.step() for _ in range(3)]
                psi = s.psi.download()
                if ref is None:
                    ref = (errs, psi)
                else:
                    assert errs == ref[0], (tb, small_L, graph)
                    assert_bits_equal(psi, ref[1], f"tb={tb} small_L={small_L} graph={graph}")
                s.close()


def test_run_contract(mgp, orc):
    """cpu-raw.lua:239-258: at most 2 cycles, `#iter err` table, stop on err < accuracy."""
    buf = io.StringIO()
    s = mgp.MultigridCUDA(64, None, 3, out=buf)          # cl(size, real=nil, cpuDepth) test/test.lua:54
    assert s.real == "double"
    errs = s.run()
    o = orc.Oracle(64, "double", 2)
    want = o.run()
    assert len(errs) == 2 and np.allclose(errs, want, rtol=1e-12, atol=0)
    lines = buf.getvalue().splitlines()
    assert lines[0] == "#iter\terr" and lines[1].startswith("1\t15402.468010") and len(lines) == 3
    assert_bits_equal(s.psi.download(), o.psi, "psi after run()")
    s2 = mgp.MultigridCUDA(64, out=False)
    assert len(s2.run(max_cycles=5, accuracy=1e3)) == 2   # 800.68 < 1e3 stops it
    s.close(); s2.close()


def test_corrections_persist_between_cycles(mgp, orc):
    # Vs[L] carries over; zeroing it (cpu.lua:138) gives a different cycle 2
    a, b = mgp.MultigridCUDA(32, out=False), mgp.MultigridCUDA(32, out=False)
    a.step(); b.step()
    b.zero_corrections()
    a.step(); b.step()
    assert not np.array_equal(a.psi.download(), b.psi.download())
    o = orc.Oracle(32)
    o.step(); o.step()
    assert_bits_equal(a.psi.download(), o.psi, "persisting V")


def test_twogrid_on_a_sub_level_and_user_buffers(mgp, orc):
    # obj:twoGrid(h, u, f, L) on caller buffers at L < size (what cpu-gpu.lua:35 does)
    rng = np.random.default_rng(8)
    s = mgp.MultigridCUDA(128, "double", out=False)
    o = orc.Oracle(128, "double")
    for mode in (mgp.MODE_FUSED, mgp.MODE_REFSEQ):
        s.set_mode(mode)
        for L in (1, 2, 8, 32):
            u, f = rand_field(rng, 2, L, np.float64), rand_field(rng, 2, L, np.float64) * L * L
            du, df = to_dev(u), to_dev(f)
            s.twoGrid(1.0 / L, du, df, L)
            uo = u.copy()
            o.two_grid(1.0 / L, uo, f.copy(), L)
            assert_bits_equal(to_host(du), uo, f"twoGrid L={L} mode={mode}")


def test_host_buffer_entry_point(mgp, orc):
    s = mgp.MultigridCUDA(64, "float", dim=3, out=False)
    o = orc.Oracle(64, "float", 3, nthreads=8)
    f = mgp.PinnedArray((64,) * 3, np.float32)
    psi = mgp.PinnedArray((64,) * 3, np.float32)
    f.array[...] = o.f; psi.array[...] = o.psi
    for _ in range(2):
        e = s.step_host(f.array, psi.array)
        eo = o.step()
        assert abs(e - eo) <= err_rtol(64 ** 3) * eo
        assert_bits_equal(psi.array, o.psi, "psi via host buffers")
    f.free(); psi.free(); s.close()


@pytest.mark.parametrize("dim,size,real", [(3, 64, "float"), (2, 256, "double"), (3, 128, "float")])
def test_pipelined_host_batch_equals_one_call_at_a_time(mgp, orc, dim, size, real):
    """mg_step_host_batch (uploads, cycles and downloads overlapped on three streams, two staging slots per direction):
    every problem's psi and err equal what mg_step_host gives for it alone, and the oracle's for the first problem.
    The problems differ (scaled right-hand sides, shifted sources), so a mixed-up staging slot cannot pass."""
    dt = np.float64 if real == "double" else np.float32
    shape = (size,) * dim
    n = 5
    o = orc.Oracle(size, real, dim, nthreads=8)
    f0, p0 = np.array(o.f, dtype=dt).reshape(shape), np.array(o.psi, dtype=dt).reshape(shape)
    fs = [mgp.PinnedArray(shape, dt) for _ in range(n)]
    ps = [mgp.PinnedArray(shape, dt) for _ in range(n)]
    want = []
    s = mgp.MultigridCUDA(size, real, dim=dim, out=False)
    for k in range(n):
        fs[k].array[...] = f0 * dt(1 + 0.25 * k)
        ps[k].array[...] = np.roll(p0, k, axis=0) * dt(1 - 0.125 * k)
    one = mgp.PinnedArray(shape, dt)
    for k in range(n):      # the serial entry point, one problem at a time, on the same handle (the coarse
        one.array[...] = ps[k].array     # corrections Vs persist from call to call: same sequence in both runs)
        want.append((s.step_host(fs[k].array, one.array), one.array.copy()))
    s.close()
    s = mgp.MultigridCUDA(size, real, dim=dim, out=False)
    errs = s.step_host_batch([a.array for a in fs], [a.array for a in ps])
    for k in range(n):
        assert errs[k] == want[k][0], (k, errs[k], want[k][0])
        assert_bits_equal(ps[k].array, want[k][1], f"problem {k} of the batch")
    eo = o.step()
    assert abs(errs[0] - eo) <= err_rtol(size ** dim) * eo
    assert_bits_equal(ps[0].array, o.psi, "first problem of the batch against the oracle")
    # second batch on the same handle (slots and events are reused), n = 2 and the degenerate sizes
    errs2 = s.step_host_batch([fs[0].array, fs[1].array], [ps[0].array, ps[1].array])
    assert len(errs2) == 2 and all(np.isfinite(errs2))
    assert s.step_host_batch([], []) == []
    assert len(s.step_host_batch([fs[2].array], [ps[2].array])) == 1
    with pytest.raises(mgp.MGError):
        s.step_host_batch([fs[0].array, fs[0].array], [ps[0].array, ps[0].array])   # outputs must be distinct
    for a in fs + ps + [one]:
        a.free()
    s.close()


def test_fp32_arithmetic_within_stated_tolerance_of_cpu_raw_float(mgp, orc):
    """north_star: fp32 ~1e-5 relative to the initial residual.
    MG_REAL_F32 (fp32 arithmetic, gpu.lua float semantics) against cpu-raw.lua's float mode
    (fp32 storage, double arithmetic -- which MG_REAL_F32_ACC64 reproduces bit for bit).
      * true residual RMS per cycle:  |r_cuda - r_ref| <= 1e-5 |r_0|
      * the reference's `err` for the cycles its run() performs (2, cpu-raw.lua:245), and the next:
                                      |err_cuda - err_ref| <= 1e-5 err_1
      * solution field:               rms(psi_cuda - psi_ref) <= tol * rms(psi_ref), tol = 1e-5 in 3-D.
        In 2-D the reference's own iteration amplifies rounding (its `err` grows from cycle 4 on, and
        its float and double modes already differ by 1.6e-5 after ONE cycle), so the 2-D field
        tolerance is 1e-4 over the reference's two cycles."""
    for dim, size, ftol in ((2, 256, 1e-4), (3, 64, 1e-5)):
        s = mgp.MultigridCUDA(size, "float", dim=dim, out=False)
        o = orc.Oracle(size, "float_acc64", dim, nthreads=8)
        r0 = o.residual_rms()
        e1 = None
        for cyc in range(3):
            es, eo = s.step(), o.step()
            e1 = eo if e1 is None else e1
            assert abs(es - eo) <= 1e-5 * e1, (dim, cyc, es, eo)
            assert abs(s.residual_norm() - o.residual_rms()) <= 1e-5 * r0
            if cyc < 2:
                dpsi = s.psi.download().astype(np.float64) - o.psi
                assert np.sqrt(np.mean(dpsi**2)) <= ftol * np.sqrt(np.mean(o.psi.astype(np.float64)**2)), (dim, cyc)
        s.close()


@pytest.mark.parametrize("dim,size,real,cluster_L", [(3, 128, "float", 64), (3, 64, "double", 64), (2, 512, "double", 256), (3, 64, "float", 32)])
def test_one_cluster_kernel_for_the_mid_levels(mgp, dim, size, real, cluster_L):
    """Option cluster_L: the levels between the streaming kernels and the one-CTA kernel in ONE launch of one thread-block
    cluster (hardware cluster barriers where the reference has kernel boundaries). Off by default (measured slower than
    the separate launches); whatever executes the levels, the bits are the same."""
    a = mgp.MultigridCUDA(size, real, dim=dim, out=False)
    b = mgp.MultigridCUDA(size, real, dim=dim, out=False)
    b.set_option("cluster_L", cluster_L)
    n0 = b.launch_count()
    for _ in range(3):
        ea, eb = a.step(), b.step()
        assert ea == eb
    assert b.psi.download().tobytes() == a.psi.download().tobytes()
    for L in (size // 2, 16, 2, 1):
        assert b.Vs[L].download().tobytes() == a.Vs[L].download().tobytes(), L
    assert b.launch_count() - n0 < a.launch_count()
    a.close(); b.close()
