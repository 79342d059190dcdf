#!/usr/bin/env python
"""bench.py -- V-cycles/s of the B200 multigrid Poisson V-cycle (BASELINE.json metric).

A "step" is one V-cycle (`twoGrid(1/size, psi, f, size)`, cpu-raw.lua:247) over the synthetic
point-source problem of the reference (cpu-raw.lua:8-20). N=1 workload: 3-D 512^3 fp32
(BASELINE.json configs[2], the configuration the metric is quoted on).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--config c1|c2|c3|c5] [--dim 3 --size 512 --real float] [--tb T --small-L S --no-graph]

--config picks a BASELINE.json configuration: c1 = configs[0] 2-D 64^2 fp64, c2 = configs[1] 2-D 4096^2 fp32,
c3 = configs[2] 3-D 512^3 fp32 (default, the headline), c5 = configs[4] 2-D 2048^2 fp64 with the multigrid-vs-Krylov leg
(`mg_vs_cg`); --gpus N > 1 is configs[3], the 1024^3 grid in z-slabs.

Prints ONE JSON line (rank 0). Keys beyond the base contract:
  roofline      dominant kernel. `achieved` / `frac` = DRAM bytes the launch really moves (ncu dram__bytes, from
                profiles/dram_traffic.json; the compulsory bytes of the fused launch when no capture exists) / CUDA-event
                duration, against the measured copy peak: a true roofline fraction (<= ~1). The A_op figure of SURVEY
                8(d) (bytes of the un-fused reference operators the launch replaces; temporal blocking pushes it above
                the peak) is kept beside it as `effective_*`.
  vcycle        whole V-cycle: A_op effective bandwidth, the fused lower bound A_min, DRAM bytes per cycle
  cpu_baseline  the CPU oracle (C restatement of cpu-raw.lua) timed on this host
  e2e           same metric through the C ABI with pinned HOST buffers, every step's copies inside the timed region:
                `value` through mg_step_host_batch (independent host problems, the copies of neighbouring steps overlap
                the cycle), `one_call_at_a_time` through mg_step_host (upload + cycle + download in sequence)
  parity        (--gpus N > 1) before timing: one V-cycle at 256^3 and at the timed 1024^3 slab shape, every rank's
                slab of psi / Rs[L/2] / Vs[L/2] compared (CRC-32 of the bytes) with a single-GPU solver on rank 0;
                a mismatch ends the run with a non-zero exit code
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "V-cycles/s"
NOMINAL_HBM_GBS = 8000.0
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


CONFIGS = {
    "c1": dict(dim=2, size=64, real="double", name="configs[0] (2D 64x64 Poisson, the reference's own cpu-raw.lua case)"),
    "c2": dict(dim=2, size=4096, real="float", name="configs[1] (2D 4096x4096 fp32 V-cycle on 1xB200)"),
    "c3": dict(dim=3, size=512, real="float", name="configs[2] (3D 512^3 fp32 V-cycle on 1xB200)"),
    "c5": dict(dim=2, size=2048, real="double", name="configs[4] (2D 2048x2048 fp64 multigrid-vs-Krylov convergence)"),
}


def level_sizes(dim, size):
    L, out = size, []
    while L >= 1:
        out.append(L)
        L //= 2
    return out


def a_op_bytes(dim, size, elem, smooth=7):
    """Algorithmic bytes of one V-cycle: every reference operator reads each input once and
    writes each output once; no copy-back, no fusion."""
    c = 2.0 ** -dim
    words = 0.0
    for L in level_sizes(dim, size):
        n = float(L) ** dim
        if L == 1:
            words += 3 * n
        else:
            words += (2 * smooth * 3 + 3 + (1 + c) + (1 + c) + 3) * n
    return words * elem


def a_min_bytes(dim, size, elem):
    """Fused lower bound of one V-cycle (SURVEY 8(d) A_min): per level visit the pre-smoothing leg reads u, f and
    writes u, R/2^dim; the post-smoothing leg reads u, f, V/2^dim and writes u: 6 + 2c words per point."""
    c = 2.0 ** -dim
    return sum((6 + 2 * c) * float(L) ** dim for L in level_sizes(dim, size) if L > 1) * elem + 3 * elem


def compulsory_bytes(kind, dim, L, sweeps, elem):
    """Bytes ONE fused launch cannot avoid: read u and f, write u (+ read V or write R, 2^-dim of a field)."""
    c = 2.0 ** -dim
    n = float(L) ** dim
    if kind == "copy":
        return 2 * n * elem
    if kind == "small_levels":
        return a_min_bytes(dim, L, elem)
    w = 3.0
    if kind in ("sweep+prolong_add", "sweep+residual_restrict", "prolong_add", "residual_restrict"):
        w += c
    return w * n * elem


def launch_bytes(kind, dim, L, sweeps, elem):
    """A_op bytes of ONE launch of a fused kernel = the reference operators it replaces."""
    c = 2.0 ** -dim
    n = float(L) ** dim
    w = 3.0 * sweeps
    if kind in ("sweep+prolong_add", "prolong_add"):
        w += (1 + c) + 3
    if kind in ("sweep+residual_restrict", "residual_restrict"):
        w += 3 + (1 + c)
    if kind == "copy":
        w = 2
    if kind == "small_levels":
        return a_op_bytes(dim, L, elem, sweeps // 2)
    return w * n * elem


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (pynvml)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def nvlink_kib(self):
        """NVLink data payload counters of this GPU, summed over its links, in KiB since the driver was loaded
        (NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX / _RX); None where NVML does not provide them."""
        if not self.nv:
            return None
        try:
            nv = self.nv
            ids = [getattr(nv, "NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX", 138), getattr(nv, "NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX", 139)]
            vals = nv.nvmlDeviceGetFieldValues(self.h, [(i, 0xFFFFFFFF) for i in ids])   # scope: all links
            out = []
            for v in vals:
                if v.nvmlReturn != 0:
                    return None
                out.append(int(v.value.ullVal))
            return out
        except Exception:
            return None

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        if self._t:
            self._stop.set()
            self._t.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def oracle_rate(dim, size, real, nthreads, budget_s, min_cycles=1):
    """V-cycles/s of the CPU oracle at the largest cube <= size whose cycle fits budget_s,
    scaled to `size` by the point count (the V-cycle is linear in N)."""
    import oracle
    sample = min(size, 64)
    rate_small = None
    while True:
        o = oracle.Oracle(sample, real, dim, nthreads=nthreads)
        o.vcycle()  # warm
        t0 = time.perf_counter()
        n = 0
        while True:
            o.vcycle()
            n += 1
            dt = time.perf_counter() - t0
            if n >= min_cycles and dt > 0.2:
                break
        o.close()
        rate_small = n / dt
        nxt = sample * 2
        est_next = (2 ** dim) / rate_small
        if nxt > size or est_next * (min_cycles + 1) > budget_s:
            break
        sample = nxt
    scale = (float(sample) / size) ** dim
    return rate_small * scale, sample, n


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path. The reference is
    LuaJIT Lua and no Lua runtime exists on this image, so this is the oracle port
    (oracle/mg_oracle.c, a C restatement of cpu-raw.lua) with every host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    # every host thread this process may use (torchrun exports OMP_NUM_THREADS=1; the oracle takes an
    # explicit thread count, which overrides it)
    try:
        nthreads = len(os.sched_getaffinity(0))
    except AttributeError:
        nthreads = os.cpu_count() or 1
    steps, warm = max(args.steps, 1), max(args.warmup, 0)
    # the whole run (warm-up + timed steps) gets ~150 s: the unit workload at its REAL size when that fits,
    # else the largest cube that does, with the rate scaled by the point count (the V-cycle is linear in N)
    per_step_budget = 150.0 / (steps + warm)
    _, sample, _ = oracle_rate(args.dim, args.size, args.real, nthreads, per_step_budget)
    o = oracle.Oracle(sample, args.real, args.dim, nthreads=nthreads)
    for _ in range(warm):
        o.vcycle()
    t0 = time.perf_counter()
    for _ in range(steps):
        o.vcycle()
    dt = time.perf_counter() - t0
    o.close()
    scale = (float(sample) / args.size) ** args.dim
    value = steps / dt * scale
    extrapolated = sample != args.size
    sample_txt = (f"{steps} V-cycles of the {args.dim}-D {sample}^{args.dim} {args.real} point-source problem on {nthreads} threads"
                  + (f", rate scaled by ({sample}/{args.size})^{args.dim} to the {args.size}^{args.dim} unit workload (EXTRAPOLATED: "
                     f"the full size does not fit the time budget)" if extrapolated else " (the unit workload at its real size)"))
    world = max(1, args.gpus)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "V-cycles/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": dtype_name(args.real),
        "data": "synthetic", "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": "V-cycles/s", "cores": nthreads, "kind": "port",
                         "sample": sample_txt, "extrapolated": extrapolated},
        "e2e": {"value": value, "unit": "V-cycles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": ("reference = C restatement of cpu-raw.lua (oracle port); LuaJIT is not available on this image. The CPU path runs "
                 "the unit workload (one V-cycle of the N = 1 grid) on this one host whatever --gpus is; the GPU arm at --gpus N > 1 "
                 "runs the 1024^3 grid and reports the same unit (V-cycles of the unit workload per second)"),
    }
    print(json.dumps(line))


def dtype_name(real):
    return {"float": "f32", "double": "f64", "float_acc64": "f32 storage / f64 arithmetic"}[real]


def l2_policy(args):
    mib = args.size ** args.dim * (8 if args.real == "double" else 4) / 2 ** 20
    if 3 * mib > 126:
        return ("inputs larger than L2: one smoother pass touches 3 top-level fields of %.0f MiB each (L2 is 126 MB); "
                "no flush needed" % mib)
    if 4 * mib > 126:
        return ("PARTLY L2 RESIDENT: the 4 top-level fields (%.0f MiB each) exceed the 126 MB L2 together, but one smoother pass "
                "touches only 3 of them (%.0f MiB), which fit: not a clean HBM measurement" % (mib, 3 * mib))
    return ("working set (4 fields x %.2f MiB) FITS in the 126 MB L2: not an HBM measurement; this size is a parity/"
            "latency case, not the roofline workload" % mib)


def baseline_config_name(args):
    for c in CONFIGS.values():
        if (c["dim"], c["size"], c["real"]) == (args.dim, args.size, args.real):
            return c["name"]
    return "custom"


def workload_config(args, world):
    """Pure function of the command line: both arms (--impl ours / reference) print the same dict."""
    size = args.size
    cfg = {
        "workload": f"{args.dim}D {size}^{args.dim} {dtype_name(args.real)} Poisson V-cycle, Dirichlet-0, point-source RHS, "
                    f"7+7 Jacobi(omega=1) sweeps per level, {len(level_sizes(args.dim, size))} levels",
        "baseline_config": baseline_config_name(args),
        "grid": [size] * args.dim, "parallelism": f"slab x{world}" if world > 1 else "single GPU",
        "l2_policy": l2_policy(args),
    }
    if world > 1:
        g = args.mgpu_size
        unit_scale = (float(g) / size) ** args.dim
        cfg.update({
            "workload": f"{args.dim}D {g}^{args.dim} {dtype_name(args.real)} Poisson V-cycle cut into {world} z-slabs "
                        f"({g // world} planes + 4 ghost planes per side per GPU); halo planes are stored into the neighbour's "
                        f"ghost planes by the smoother kernel itself over NVLink peer memory (CUDA IPC), handshake inside the "
                        f"kernel; levels with < 32 planes per GPU replicated, their right-hand side all-gathered by the "
                        f"restricting kernel's peer stores; no NCCL call inside a V-cycle; the cycle is one CUDA graph per GPU",
            "baseline_config": "configs[3] (3D 1024^3 fp32 slab-decomposed across 2/4/8 B200)",
            "grid": [g] * args.dim, "parallelism": f"z-slabs x{world}",
            "unit": f"value counts V-cycles of the N=1 workload ({size}^{args.dim}); one {g}^{args.dim} V-cycle = {unit_scale:g} units",
            "points_per_gpu": g ** args.dim // world, "points_per_gpu_at_n1": size ** args.dim,
            "scaling_note": (f"the grid is {g}^{args.dim} at every N > 1: N = {int(unit_scale)} holds the N = 1 points per GPU (weak), "
                             f"smaller N hold {int(unit_scale)}/N times as many (read the 2 -> 4 -> 8 steps as strong scaling)"),
            "l2_policy": "inputs larger than L2 (a slab's fields are >= 512 MiB each)",
        })
    return cfg


def mg_vs_cg(pkg, args, budget_s=8.0, tol=1e-10):
    """BASELINE config [4] (test/converge-multigrid-vs-krylov.lua): the multigrid the experiment drives (cpu.lua: coarse
    corrections re-zeroed every cycle, cpu.lua:138) against conjugate gradient on the same operator with b = f, x0 = -f
    (:38-69), both run towards a relative residual `tol`. The multigrid leg is bounded by `budget_s` (omega = 1 Jacobi leaves
    the top mode nearly undamped: the cycle count grows ~3.7x per grid doubling)."""
    s = pkg.MultigridCUDA(args.size, args.real, dim=args.dim, out=False)
    r0 = s.residual_norm()
    s.zero_corrections(); s.vcycle(); s.residual_norm()      # warm-up: graph capture
    s.init_cells()
    t0, c, r, check = time.perf_counter(), 0, r0, 200
    while time.perf_counter() - t0 < budget_s:
        for _ in range(check):
            s.zero_corrections()
            s.vcycle()
        c += check
        r = s.residual_norm()
        if not (r > tol * r0):
            break
    mg = {"variant": "cpu.lua (mg_zero_corrections + mg_vcycle per cycle)", "cycles": c, "seconds": time.perf_counter() - t0,
          "residual_rel": r / r0, "reached_tol": bool(r <= tol * r0), "linf_psi": s.linf_norm(), "budget_s": budget_s}
    s.init_cells()
    t0 = time.perf_counter()
    errs, linf = s.conjgrad(max_iter=50000, epsilon=tol)
    dt = time.perf_counter() - t0
    cg = {"iterations": len(errs), "seconds": dt, "err_r_over_b": errs[-1] if errs else None,
          "residual_rel": s.residual_norm() / r0, "reached_tol": bool(errs and errs[-1] < tol), "linf_x": linf[-1] if linf else None,
          "parity": "unpinned: the reference's solver.conjgrad is an un-vendored dependency; checked against the oracle's textbook CG"}
    s.close()
    return {"tol": tol, "quantity": "multigrid: ||f - A psi|| / ||f - A psi_0||; CG: ||r|| / ||b|| (what solver.conjgrad tests)",
            "multigrid": mg, "conjugate_gradient": cg}


def slab_parity(pkg, torch, dist, size, args, rank, world, local, solver):
    """One V-cycle (mg_step) on the slab solver and on a single-GPU solver of the same grid (rank 0); every rank's
    planes of psi, Rs[size/2] and Vs[size/2] must agree byte for byte (CRC-32 per slab), `err` to 1e-9 relative."""
    import zlib
    import numpy as np
    own = solver if solver is not None else pkg.create_distributed(size, args.real, dim=args.dim)
    own.init_cells()
    own.zero_corrections()
    err = own.step()
    half = size // 2
    crcs = [zlib.crc32(np.ascontiguousarray(b.download()).tobytes()) for b in (own.psi, own.Rs[half], own.Vs[half])]
    mine = torch.tensor(crcs + [0], dtype=torch.int64, device="cuda")
    mine[3] = int(np.float64(err).view(np.int64))
    allc = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine)
    ok = True
    if rank == 0:
        one = pkg.MultigridCUDA(size, args.real, dim=args.dim, device=local, out=False)
        one.set_option("stream_min_L", 64)
        ref_err = one.step()
        nz = size // world
        for name, buf, L in (("psi", one.psi, size), ("Rs", one.Rs[half], half), ("Vs", one.Vs[half], half)):
            full = np.ascontiguousarray(buf.download())
            idx = {"psi": 0, "Rs": 1, "Vs": 2}[name]
            dist_level = (L // world) >= 32 and L >= 64      # the library's rule (slab_partition): else replicated
            for r in range(world):
                part = full[r * (L // world):(r + 1) * (L // world)] if dist_level else full
                if zlib.crc32(part.tobytes()) != int(allc[r][idx].item()):
                    ok = False
                    print(f"[bench parity] {size}^3: {name} of rank {r} differs from the single-GPU solver", file=sys.stderr)
        errs = [float(np.int64(int(c[3].item())).view(np.float64)) for c in allc]
        if any(abs(e - ref_err) > 1e-9 * abs(ref_err) for e in errs):
            ok = False
            print(f"[bench parity] {size}^3: err {errs} vs single GPU {ref_err}", file=sys.stderr)
        one.close()
        del nz
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    if solver is None:
        own.close()
    return bool(int(flag.item()))


def run_ours(args):
    import numpy as np
    import torch

    from __graft_entry__ import load_package
    pkg = load_package()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    elem = 8 if args.real == "double" else 4
    # N > 1: BASELINE config [3], the 3-D 1024^3 grid cut into z-slabs with NCCL halo exchange.
    # One "unit" of work is a V-cycle of the N=1 workload (512^3 points); `value` is the
    # whole-job rate in those units, i.e. (global points / 512^3) x V-cycles/s.
    gsize = args.size
    if world > 1:
        gsize = args.mgpu_size
        s = pkg.create_distributed(gsize, args.real, dim=args.dim)
    else:
        s = pkg.MultigridCUDA(args.size, args.real, dim=args.dim, device=local, out=False)
    unit_scale = (float(gsize) / args.size) ** args.dim if world > 1 else 1.0
    if world == 1:
        s.set_tuning(tb=args.tb, small_L=args.small_L, use_graph=0 if args.no_graph else 1)
    else:
        s.set_tuning(tb=args.tb, use_graph=0 if args.no_graph else 1)
    for kv in args.opt:
        k, v = kv.split("=")
        s.set_option(k, int(v))
    stream = torch.cuda.Stream()
    s.set_stream(stream.cuda_stream)
    lib, h = pkg.lib(), s._h

    # ---- multi-GPU parity BEFORE timing: 256^3 and the timed slab shape against a single-GPU solver on rank 0
    parity = None
    if world > 1 and not args.no_parity:
        parity = {}
        for psize in sorted({256, gsize}):
            ok = slab_parity(pkg, torch, dist, psize, args, rank, world, local, s if psize == gsize else None)
            parity[str(psize)] = ok
        if not all(parity.values()):
            if rank == 0:
                print(json.dumps({"metric": METRIC, "n_gpus": world, "parity": parity,
                                  "error": "multi-GPU result differs from the single-GPU solver: no timing reported"}))
            dist.destroy_process_group()
            raise SystemExit(3)
        s.init_cells()
        s.zero_corrections()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: W warm-up + exactly K timed V-cycles
    for _ in range(max(args.warmup, 3)):
        lib.mg_vcycle_async(h)
    # The reference's iteration (omega = 1 Jacobi, injection prolongation) DIVERGES after a few
    # cycles (its `err` grows ~16x every 5 cycles from cycle ~5 on, oracle/ and tests/golden), so the
    # state is re-initialised after the warm-up to keep the timed fields finite. Kernel time does not depend
    # on the values as long as they are ordinary: only a pass that meets a tiny (< 1e-20) non-zero numerator is
    # run a second time by the guarded kernel (mg_stream3d.cuh), which does not happen on this problem.
    s.init_cells()
    s.zero_corrections()
    clocks = ClockSampler(local)     # NVML initialisation (tens of ms, different on every rank) stays in front of the barrier
    barrier()
    clocks.start()
    n0 = s.launch_count()
    tr0 = s.slab_traffic() if world > 1 else None
    nvl_hw0 = clocks.nvlink_kib() if world > 1 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        if world > 1:
            # The ranks leave the host barrier some 0.1-1 ms apart. One more UNTIMED cycle, enqueued in front of the first
            # event, absorbs that start skew in its neighbour handshakes, so that the events of every rank bracket exactly
            # K cycles in lock step instead of K cycles plus the wait for the last rank to arrive.
            rc = lib.mg_vcycle_async(h)
            assert rc == 0, lib.mg_last_error(h)
        e0.record()
        for _ in range(args.steps):
            rc = lib.mg_vcycle_async(h)
            assert rc == 0, lib.mg_last_error(h)
        e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    launches = s.launch_count() - n0
    if world > 1:
        launches = launches * args.steps // (args.steps + 1)    # the skew-absorbing cycle is not part of the timed region
    nvl = None
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tr1 = s.slab_traffic()
        nvl_hw1 = clocks.nvlink_kib()
        per_step = (tr1["peer_store_bytes"] - tr0["peer_store_bytes"]) / (args.steps + 1)    # + the skew-absorbing cycle
        tb = torch.tensor([per_step], device="cuda", dtype=torch.float64)
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        nvl = {"peer_store_bytes_per_step_this_rank": per_step, "peer_store_bytes_per_step_all_ranks": float(tb.item()),
               "gbs_per_gpu_out": per_step / (ms / args.steps * 1e-3) / 1e9,
               "explicit_exchange_bytes_in_timed_region": tr1["exchange_bytes"] - tr0["exchange_bytes"],
               "hw_counter_tx_rx_bytes_per_step_this_rank": ([1024.0 * (b - a) / (args.steps + 1) for a, b in zip(nvl_hw0, nvl_hw1)]
                                                              if nvl_hw0 and nvl_hw1 else None),
               "hw_counter_source": "NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX/_RX of this GPU, all links, read before and after the "
                                    "timed cycles (payload KiB; includes the handshake words). On this pool NVML answers NOT_SUPPORTED "
                                    "for these fields (probed on the B200 boxes), so the value is null here",
               "peak_per_direction_gbs": 770.0,
               "note": "bytes the smoother kernels store straight into other GPUs' memory (G = 4 boundary planes per neighbour "
                       "and pass, coarse boundary planes, the fused all-gather), counted per launch from the planes sent; "
                       "770 GB/s = measured peer-copy bandwidth per direction (B200_PROFILING.md)"}
    ms_per_step = ms / args.steps
    value = unit_scale * 1e3 / ms_per_step
    diverged = not bool(np.isfinite(s.step()))   # expected after enough cycles: the reference's iteration diverges
    # the trajectory a user of the reference sees: err of the first cycles from the reference's initial state
    s.init_cells()
    s.zero_corrections()
    first_errs = [s.step() for _ in range(4)]
    finite = bool(np.all(np.isfinite(first_errs)))

    # ---- per-launch CUDA-event timing of the same V-cycle (ungraphed), median of 5 cycles
    peak, peak_src = measured_peak()
    s.init_cells()
    s.zero_corrections()
    recs = [s.profile_vcycle() for _ in range(5)]
    med = []
    for i in range(len(recs[0])):
        r = dict(recs[0][i])
        r["ms"] = sorted(x[i]["ms"] for x in recs)[len(recs) // 2]
        med.append(r)
    groups = {}
    for r in med:
        g = groups.setdefault((r["kind"], r["L"], r["sweeps"]), {"ms": 0.0, "n": 0})
        g["ms"] += r["ms"]
        g["n"] += 1
    total_prof = sum(g["ms"] for g in groups.values())
    (dk, dL, dsw), dg = max(groups.items(), key=lambda kv: kv[1]["ms"])
    dom_ms = dg["ms"] / dg["n"]
    share = (world if dL >= 64 and world > 1 else 1)          # this rank's slab of a distributed level
    aop_launch = launch_bytes(dk, args.dim, dL, dsw, elem) / share
    comp_launch = compulsory_bytes(dk, args.dim, dL, dsw, elem) / share
    kname = f"{dk}[L={dL},sweeps={dsw}]"
    key = f"{args.dim}d_{gsize if world > 1 else args.size}_{args.real}" + (f"_n{world}" if world > 1 else "")
    traffic_tab = {}
    try:
        traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json"))).get(key, {})
    except Exception:
        pass
    traffic = traffic_tab.get(kname)
    pipes = None
    try:
        pipes = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json"))).get("_pipes", {}).get(key, {}).get(kname)
    except Exception:
        pass
    moved = traffic if traffic else comp_launch
    achieved = moved / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kname, "launches_per_step": dg["n"], "avg_launch_ms": dom_ms,
                "bytes_source": ("ncu dram__bytes_read.sum + dram__bytes_write.sum of this launch (profiles/dram_traffic.json)"
                                 if traffic else "compulsory bytes of the fused launch (no ncu capture of this kernel at this size)"),
                "compulsory_bytes_per_launch": comp_launch, "compulsory_gbs": comp_launch / (dom_ms * 1e-3) / 1e9,
                "traffic_over_compulsory": (traffic / comp_launch) if traffic else None,
                "effective_bytes_per_launch_A_op": aop_launch, "effective_gbs_A_op": aop_launch / (dom_ms * 1e-3) / 1e9,
                "effective_frac_A_op": aop_launch / (dom_ms * 1e-3) / 1e9 / peak,
                "share_of_step": dg["ms"] / total_prof, "peak_source": peak_src,
                "binding_resource": ({"from": "ncu capture of this launch (profiles/, cold cache, not this run)", **pipes,
                                      "reading": "the pass keeps the shared-memory data path busier than DRAM or the issue slots: "
                                                 "it is bound by on-chip traffic per point and stage, not by HBM"} if pipes else None),
                "frac_of_nominal_8000": achieved / NOMINAL_HBM_GBS,
                "note": "frac is DRAM bytes moved / time / measured copy peak; the A_op figure (SURVEY 8(d)) counts the bytes the "
                        "un-fused reference operators would move and exceeds the peak by design of temporal blocking"}
    aop = a_op_bytes(args.dim, gsize, elem)
    amin = a_min_bytes(args.dim, gsize, elem)
    v_gbs = aop / (ms_per_step * 1e-3) / 1e9 / world   # per GPU
    cyc_traffic = None
    if traffic_tab:   # DRAM bytes of a whole cycle where every stream/warp launch has a capture; the rest at compulsory bytes
        cyc_traffic = 0.0
        for (k, L, sw), g in groups.items():
            t = traffic_tab.get(f"{k}[L={L},sweeps={sw}]")
            cyc_traffic += g["n"] * (t if t else compulsory_bytes(k, args.dim, L, sw, elem) / (world if L >= 64 and world > 1 else 1))
    vcycle = {"a_op_bytes": aop, "effective_gbs_per_gpu": v_gbs, "effective_frac_of_measured": v_gbs / peak,
              "effective_frac_of_nominal_8000": v_gbs / NOMINAL_HBM_GBS,
              "a_min_bytes": amin, "a_min_gbs_per_gpu": amin / (ms_per_step * 1e-3) / 1e9 / world,
              "a_min_frac_of_measured": amin / (ms_per_step * 1e-3) / 1e9 / world / peak,
              "dram_bytes_per_cycle_per_gpu": cyc_traffic,
              "dram_gbs_per_gpu": (cyc_traffic / (ms_per_step * 1e-3) / 1e9) if cyc_traffic else None,
              "dram_frac_of_measured": (cyc_traffic / (ms_per_step * 1e-3) / 1e9 / peak) if cyc_traffic else None,
              "breakdown_ms": {f"{k}[L={L},sweeps={sw}]x{g['n']}": round(g["ms"], 4)
                               for (k, L, sw), g in sorted(groups.items(), key=lambda kv: -kv[1]["ms"])[:8]},
              "breakdown_all_ms": {f"{k}[L={L},sweeps={sw}]x{g['n']}": round(g["ms"], 4)
                                   for (k, L, sw), g in sorted(groups.items(), key=lambda kv: (-kv[0][1], kv[0][0], kv[0][2]))},
              "profiled_sum_ms": total_prof}

    # ---- end to end through the C ABI with HOST buffers (pinned), copies inside the timing
    N = gsize ** args.dim // world   # points whose host buffers this rank moves
    dt = np.float64 if args.real == "double" else np.float32
    e2e = None
    if not args.quick:
        fh, ph = pkg.PinnedArray((N,), dt), pkg.PinnedArray((N,), dt)
        fh.array[...] = s.f.download().ravel()
        ph.array[...] = 0
        ph.array[N // 2] = 1.0 if rank == world // 2 else 0.0
        s.step_host(fh.array, ph.array)  # warm
        barrier()
        ne2e = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(ne2e):
            s.step_host(fh.array, ph.array)
        torch.cuda.synchronize()
        serial_s = (time.perf_counter() - t0) / ne2e
        # the same steps as a batch of independent host problems (mg_step_host_batch): every step still uploads its f and
        # psi from pinned host memory and downloads its psi, but problem i+1's upload and problem i-1's download run on the
        # two copy engines while problem i's cycle runs (PCIe is full duplex). One input f serves every problem; every
        # problem has its own psi buffer (in: the point source, out: the result).
        nb = max(4, min(2 * args.steps, 16 if world == 1 else 6))
        pb, alloc_ok = [], 1
        try:
            pb = [pkg.PinnedArray((N,), dt) for _ in range(nb)]
        except Exception:
            alloc_ok = 0
        if dist is not None:   # every rank takes the batch leg or none does (its err reduction is a collective)
            t = torch.tensor([alloc_ok], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            alloc_ok = int(t.item())
        e2e_s = None
        if alloc_ok:
            def fill():
                for a in pb:
                    a.array[...] = 0
                    a.array[N // 2] = 1.0 if rank == world // 2 else 0.0
            fill()
            s.step_host_batch([fh.array] * nb, [a.array for a in pb])  # warm: staging slots, streams
            fill()
            barrier()
            t0 = time.perf_counter()
            s.step_host_batch([fh.array] * nb, [a.array for a in pb])
            torch.cuda.synchronize()
            e2e_s = (time.perf_counter() - t0) / nb
        for a in pb:
            a.free()
        if dist is not None:
            t = torch.tensor([e2e_s or 0.0, serial_s], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s, serial_s = (float(t[0].item()) if alloc_ok else None), float(t[1].item())
        one = {"value": unit_scale / serial_s, "ms_per_step": serial_s * 1e3, "steps": ne2e,
               "api": "mg_step_host (upload f, psi; psiOld<-psi; V-cycle; err; download psi), "
                      "nothing overlapped: upload + cycle + download"}
        if e2e_s:
            e2e = {"value": unit_scale / e2e_s, "unit": "V-cycles/s", "h2d_bytes_per_step": 2 * N * elem * world,
                   "d2h_bytes_per_step": (N * elem + 8) * world, "ms_per_step": e2e_s * 1e3, "steps": nb,
                   "api": "mg_step_host_batch: per step upload f, psi; psiOld<-psi; V-cycle; err; download psi -- a batch of "
                          "independent host problems, the copies of neighbouring steps overlapped with the cycle on two copy streams",
                   "pcie_gbs_up_per_gpu": 2 * N * elem / e2e_s / 1e9, "pcie_gbs_down_per_gpu": N * elem / e2e_s / 1e9,
                   "one_call_at_a_time": one}
        else:   # no pinned memory for the batch on some rank: the single-call figure stands
            e2e = {"value": one["value"], "unit": "V-cycles/s", "h2d_bytes_per_step": 2 * N * elem * world,
                   "d2h_bytes_per_step": (N * elem + 8) * world, "ms_per_step": one["ms_per_step"], "steps": ne2e,
                   "api": one["api"], "note": "batch leg skipped: pinned host memory for it could not be allocated"}
        fh.free()
        ph.free()

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port, 1 thread
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and not args.quick:
        import oracle
        oracle.build()
        rate, sample, ncyc = oracle_rate(args.dim, args.size, args.real, 1, budget_s=25.0)
        cpu = {"value": rate, "unit": "V-cycles/s", "cores": 1, "kind": "port",
               "sample": f"{ncyc} V-cycle(s) of the {args.dim}-D {sample}^{args.dim} {args.real} problem on 1 thread "
                         f"(C restatement of cpu-raw.lua; LuaJIT itself is single-threaded), "
                         f"scaled by ({sample}/{args.size})^{args.dim}"}

    # ---- "time to 1e-8 residual" (second half of BASELINE.json's metric): only the cpu.lua variant of the reference
    # (coarse corrections re-zeroed every cycle) converges, only fp64 can represent the tolerance, and the cycle count
    # grows ~3.7x per grid doubling, so it is measured on a bounded grid and named as such.
    ttt = None
    if rank == 0 and world == 1 and not args.quick:
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location(
                "time_to_tolerance", os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "time_to_tolerance.py"))
            tt = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(tt)
            tsize = 64 if args.dim == 3 else 256
            r = tt.solve_gpu(pkg, args.dim, tsize, 1e-8, 10 if args.dim == 3 else 50, 100000)
            ttt = {"seconds": r["seconds"], "cycles": r["cycles"], "tol": 1e-8, "residual_rel": r["residual_rel"],
                   "grid": f"{args.dim}D {tsize}^{args.dim} f64", "quantity": "||f - A psi|| / ||f - A psi_0||",
                   "variant": "cpu.lua: mg_zero_corrections + mg_vcycle per cycle (the cpu-raw.lua variant does not converge)"}
        except Exception as e:  # never let the extra figure break the bench line
            ttt = {"unavailable": repr(e)[:200]}

    if world > 1 and args.dim == 3 and not args.quick:
        # every rank takes part (the residual norm is an all-reduce): the converging (cpu.lua) variant on a 256^3 fp64 grid
        # in z-slabs, towards 1e-8 of the initial residual, bounded by a time budget
        try:
            t = pkg.create_distributed(256, "double", dim=3)
            r0 = t.residual_norm()
            t.zero_corrections(); t.vcycle(); t.residual_norm()
            t.init_cells()
            barrier()
            t0, cyc, r, check, budget = time.perf_counter(), 0, r0, 100, 6.0
            while True:
                for _ in range(check):
                    t.zero_corrections()
                    t.vcycle()
                cyc += check
                r = t.residual_norm()
                stop = torch.tensor([1 if (not (r > 1e-8 * r0) or time.perf_counter() - t0 > budget) else 0], device="cuda")
                dist.all_reduce(stop, op=dist.ReduceOp.MAX)     # every rank leaves the loop in the same iteration
                if int(stop.item()):
                    break
            dt = time.perf_counter() - t0
            ttt = {"seconds": dt, "cycles": cyc, "tol": 1e-8, "residual_rel": r / r0, "reached_tol": bool(r <= 1e-8 * r0),
                   "grid": f"3D 256^3 f64 in {world} z-slabs", "quantity": "||f - A psi|| / ||f - A psi_0||", "budget_s": budget,
                   "ms_per_cycle": 1e3 * dt / cyc,
                   "variant": "cpu.lua: mg_zero_corrections + mg_vcycle per cycle (the cpu-raw.lua variant does not converge); "
                              "omega = 1 needs ~11 400 cycles for 1e-8 on this grid (BASELINE.md 5.4), so the budget usually ends the run"}
            t.close()
        except Exception as e:
            ttt = {"unavailable": repr(e)[:200]}

    # ---- where a rank's cycle goes on N GPUs: the kernels' own timeline of the distributed passes (mg_slab_trace)
    timeline = None
    if world > 1:
        try:
            s.init_cells(); s.zero_corrections()
            s.set_option("slab_trace", 1)
            for _ in range(3):
                lib.mg_vcycle_async(h)         # ghost refresh, graph capture
            barrier()
            s.set_option("slab_trace", 1)      # reset the records (the graph is captured again: the trace pointer is an argument)
            lib.mg_vcycle_async(h)
            barrier()
            s.set_option("slab_trace", 1)
            ncyc = 6
            for _ in range(ncyc):
                lib.mg_vcycle_async(h)
            barrier()
            tr = s.slab_trace().astype(np.int64)
            s.set_option("slab_trace", 0)
            npass = len(tr) // ncyc
            mine = None
            if npass > 0 and len(tr) == npass * ncyc:
                t = tr.reshape(ncyc, npass, 4)[1:]                       # drop the first cycle
                flat = tr.reshape(-1, 4)
                gap = np.zeros(len(flat)); gap[1:] = flat[1:, 0] - flat[:-1, 3]
                gap = gap.reshape(ncyc, npass)[1:]
                mine = {"lo_wait_us": ((t[:, :, 1] - t[:, :, 0]).mean(0) / 1e3).tolist(),
                        "hi_wait_us": (t[:, :, 2].mean(0) / 1e3).tolist(),
                        "pass_us": ((t[:, :, 3] - t[:, :, 0]).mean(0) / 1e3).tolist(),
                        "gap_before_us": (gap.mean(0) / 1e3).tolist(),
                        "cycle_us": float((flat[npass:, 0].reshape(ncyc - 1, npass)[-1, 0] - flat[npass, 0]) / max(ncyc - 2, 1) / 1e3)}
            allr = [None] * world
            dist.all_gather_object(allr, mine)
            if rank == 0 and all(a is not None for a in allr):
                def agg(key, f):
                    return [round(float(f([a[key][j] for a in allr])), 1) for j in range(len(allr[0][key]))]
                timeline = {"passes_per_cycle": len(allr[0]["pass_us"]),
                            "order": "distributed levels top-down (pre-smoothing passes), then bottom-up (post-smoothing passes)",
                            "pass_us_max_over_ranks": agg("pass_us", max), "pass_us_mean": agg("pass_us", np.mean),
                            "lo_wait_us_mean": agg("lo_wait_us", np.mean), "lo_wait_us_max": agg("lo_wait_us", max),
                            "hi_wait_us_mean": agg("hi_wait_us", np.mean), "hi_wait_us_max": agg("hi_wait_us", max),
                            "gap_before_us_mean": agg("gap_before_us", np.mean),
                            "sum_pass_us_mean": round(float(np.mean([sum(a["pass_us"]) for a in allr])), 1),
                            "sum_gap_us_mean": round(float(np.mean([sum(a["gap_before_us"]) for a in allr])), 1),
                            "sum_wait_us_mean": round(float(np.mean([sum(a["lo_wait_us"]) + sum(a["hi_wait_us"]) for a in allr])), 1),
                            "cycle_us_mean": round(float(np.mean([a["cycle_us"] for a in allr])), 1),
                            "note": "globaltimer stamps written by the smoother kernels themselves (CTA 0 on entry / after the wait for "
                                    "the lower neighbour, its deferred wait for the upper neighbour, the last CTA at the end); gap_before "
                                    "= entry minus the previous pass's end: launch gaps, and for the first post-smoothing pass of the "
                                    "last distributed level the epoch fence + the whole replicated coarse sub-cycle"}
        except Exception as e:
            timeline = {"unavailable": repr(e)[:300]}

    mgcg = None
    if rank == 0 and world == 1 and (args.dim, args.size, args.real) == (2, 2048, "double"):
        try:
            mgcg = mg_vs_cg(pkg, args)
        except Exception as e:
            mgcg = {"unavailable": repr(e)[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "V-cycles/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype_name(args.real), "data": "synthetic",
            "config": workload_config(args, world), "roofline": roofline, "vcycle": vcycle,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clk,
            "tuning": {"tb": args.tb, "small_L": args.small_L, "graph": not args.no_graph, "opt": args.opt},
            "time_to_tolerance": ttt, "mg_vs_cg": mgcg, "finite": finite, "err_cycles_1_to_4": first_errs,
            "state_overflowed_during_timed_cycles": diverged,
            "note": "the reference's omega=1 V-cycle diverges after ~5 cycles (BASELINE.md 5.4); kernel time is value independent",
        }
        if world > 1:
            line["parity"] = parity
            line["nvlink"] = nvl
            line["slab_timeline"] = timeline
        print(json.dumps(line))
    s.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dim", type=int, default=None)
    ap.add_argument("--size", type=int, default=None)
    ap.add_argument("--mgpu-size", dest="mgpu_size", type=int, default=1024, help="global grid width for --gpus > 1")
    ap.add_argument("--real", default=None, choices=["float", "double", "float_acc64"])
    ap.add_argument("--tb", type=int, default=-1)
    ap.add_argument("--small-L", dest="small_L", type=int, default=-1)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", dest="no_parity", action="store_true", help="skip the multi-GPU parity check before timing")
    ap.add_argument("--quick", action="store_true", help="tuning runs: skip the end-to-end leg, the CPU baseline and time-to-tolerance")
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS), help="a BASELINE.json configuration (default c3)")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (mg_set_option)")
    args = ap.parse_args()
    preset = CONFIGS[args.config or "c3"]
    for k in ("dim", "size", "real"):
        if getattr(args, k) is None:
            setattr(args, k, preset[k])
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
