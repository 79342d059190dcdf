#!/usr/bin/env python
"""bench.py -- V-cycles/s of the B200 multigrid Poisson V-cycle (BASELINE.json metric).

A "step" is one V-cycle (`twoGrid(1/size, psi, f, size)`, cpu-raw.lua:247) over the synthetic
point-source problem of the reference (cpu-raw.lua:8-20). N=1 workload: 3-D 512^3 fp32
(BASELINE.json configs[2], the configuration the metric is quoted on).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--dim 3 --size 512 --real float] [--tb T --small-L S --no-graph]

Prints ONE JSON line (rank 0). Keys beyond the base contract:
  roofline      dominant kernel: algorithmic (A_op) bytes per launch / CUDA-event duration
  vcycle        whole V-cycle effective bandwidth against A_op
  cpu_baseline  the CPU oracle (C restatement of cpu-raw.lua) timed on this host
  e2e           same metric through mg_step_host with pinned HOST buffers, copies timed
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
    per_step_budget = min(20.0, 150.0 / (steps + warm))
    # choose the sample cube once, then time exactly `steps` V-cycles on it
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
    sample_txt = (f"{steps} V-cycles of the {args.dim}-D {sample}^{args.dim} {args.real} point-source problem, "
                  f"rate scaled by ({sample}/{args.size})^{args.dim} to the {args.size}^{args.dim} workload")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "V-cycles/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": dtype_name(args.real),
        "data": "synthetic", "config": dict(workload_config(args, 1), note=(
            "the CPU path runs the unit workload on this one host whatever --gpus is; the GPU arm at --gpus N > 1 "
            "runs the 1024^3 grid and reports the same unit (V-cycles of this workload per second)")),
        "cpu_baseline": {"value": value, "unit": "V-cycles/s", "cores": nthreads, "kind": "port",
                         "sample": sample_txt},
        "e2e": {"value": value, "unit": "V-cycles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = C restatement of cpu-raw.lua (oracle port); LuaJIT is not available on this image",
    }
    print(json.dumps(line))


def dtype_name(real):
    return {"float": "f32", "double": "f64", "float_acc64": "f32 storage / f64 arithmetic"}[real]


def l2_policy(args):
    mib = args.size ** args.dim * (8 if args.real == "double" else 4) / 2 ** 20
    if 4 * mib > 126:
        return "inputs larger than L2 (each of the 4 top-level fields is %.0f MiB; L2 is 126 MB); no flush needed" % mib
    return ("working set (4 fields x %.2f MiB) FITS in the 126 MB L2: not an HBM measurement; this size is a parity/"
            "latency case, not the roofline workload" % mib)


def workload_config(args, world):
    size = args.size
    return {
        "workload": f"{args.dim}D {size}^{args.dim} {dtype_name(args.real)} Poisson V-cycle, Dirichlet-0, point-source RHS, "
                    f"7+7 Jacobi(omega=1) sweeps per level, {len(level_sizes(args.dim, size))} levels",
        "baseline_config": "configs[2] (3D 512^3 fp32 V-cycle on 1xB200)" if (args.dim, size, args.real) == (3, 512, "float") else "custom",
        "grid": [size] * args.dim, "parallelism": f"slab x{world}" if world > 1 else "single GPU",
        "l2_policy": l2_policy(args),
    }


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

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: W warm-up + exactly K timed V-cycles
    for _ in range(max(args.warmup, 3)):
        lib.mg_vcycle_async(h)
    # The reference's iteration (omega = 1 Jacobi, injection prolongation) DIVERGES after a few
    # cycles (its `err` grows ~16x every 5 cycles from cycle ~5 on, oracle/ and tests/golden), so the
    # state is re-initialised after the warm-up to keep the timed fields finite; kernel time does
    # not depend on the values.
    s.init_cells()
    s.zero_corrections()
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    n0 = s.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
        for _ in range(args.steps):
            rc = lib.mg_vcycle_async(h)
            assert rc == 0, lib.mg_last_error(h)
        e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    launches = s.launch_count() - n0
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
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
    dom_bytes = launch_bytes(dk, args.dim, dL, dsw, elem) / (world if dL >= 64 else 1)  # this rank's slab
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    kname = f"{dk}[L={dL},sweeps={dsw}]"
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json"))).get(
            f"{args.dim}d_{args.size}_{args.real}", {}).get(kname)
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kname, "launches_per_step": dg["n"],
                "avg_launch_ms": dom_ms, "algorithmic_bytes_per_launch": dom_bytes,
                "share_of_step": dg["ms"] / total_prof, "peak_source": peak_src,
                "frac_of_nominal_8000": achieved / NOMINAL_HBM_GBS}
    aop = a_op_bytes(args.dim, gsize, elem)
    v_gbs = aop / (ms_per_step * 1e-3) / 1e9 / world   # per GPU
    vcycle = {"a_op_bytes": aop, "effective_gbs_per_gpu": v_gbs, "frac_of_measured": v_gbs / peak,
              "frac_of_nominal_8000": v_gbs / NOMINAL_HBM_GBS,
              "breakdown_ms": {f"{k}[L={L},sweeps={sw}]x{g['n']}": round(g["ms"], 4)
                               for (k, L, sw), g in sorted(groups.items(), key=lambda kv: -kv[1]["ms"])[:8]},
              "profiled_sum_ms": total_prof}

    # ---- end to end through the C ABI with HOST buffers (pinned), copies inside the timing
    N = gsize ** args.dim // world   # points whose host buffers this rank moves
    dt = np.float64 if args.real == "double" else np.float32
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
    e2e_s = (time.perf_counter() - t0) / ne2e
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": unit_scale / e2e_s, "unit": "V-cycles/s", "h2d_bytes_per_step": 2 * N * elem * world,
           "d2h_bytes_per_step": (N * elem + 8) * world, "ms_per_step": e2e_s * 1e3, "steps": ne2e,
           "api": "mg_step_host (upload f, psi; psiOld<-psi; V-cycle; err; download psi)"}
    fh.free()
    ph.free()

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port, 1 thread
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
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
    if rank == 0 and world == 1:
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

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "V-cycles/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype_name(args.real), "data": "synthetic",
            "config": workload_config(args, world), "roofline": roofline, "vcycle": vcycle,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clk,
            "tuning": {"tb": args.tb, "small_L": args.small_L, "graph": not args.no_graph, "opt": args.opt},
            "time_to_tolerance": ttt, "finite": finite, "err_cycles_1_to_4": first_errs,
            "state_overflowed_during_timed_cycles": diverged,
            "note": "the reference's omega=1 V-cycle diverges after ~5 cycles (BASELINE.md 5.4); kernel time is value independent",
        }
        if world > 1:
            si = s.slab_info()
            line["config"].update({
                "workload": f"{args.dim}D {gsize}^{args.dim} {dtype_name(args.real)} Poisson V-cycle cut into {world} z-slabs "
                            f"({si['own_planes']} planes + {si['ghost']} ghost planes per side per GPU); halo planes are stored "
                            f"into the neighbour's ghost planes by the smoother kernel itself over NVLink peer memory "
                            f"(CUDA IPC), handshake inside the kernel; levels with < 32 planes per GPU replicated (NCCL all-gather)",
                "baseline_config": "configs[3] (3D 1024^3 fp32 slab-decomposed across 2/4/8 B200)",
                "grid": [gsize] * args.dim, "parallelism": f"z-slabs x{world}",
                "unit": f"value counts V-cycles of the N=1 workload ({args.size}^{args.dim}); one {gsize}^{args.dim} V-cycle = {unit_scale:g} units",
                "halo_exchanges_per_step": (si["exchanges"]) // max(1, args.steps + max(args.warmup, 3) + 5 + 1 + ne2e + 1),
            })
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
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--mgpu-size", dest="mgpu_size", type=int, default=1024, help="global grid width for --gpus > 1")
    ap.add_argument("--real", default="float", choices=["float", "double", "float_acc64"])
    ap.add_argument("--tb", type=int, default=-1)
    ap.add_argument("--small-L", dest="small_L", type=int, default=-1)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (mg_set_option)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
