"""Time to a 1e-8 residual (BASELINE.json's second metric) for the converging member of the reference family.

cpu-raw.lua / gpu.lua keep the coarse corrections Vs[L] from cycle to cycle and their iteration does NOT converge
(BASELINE.md 5.4). cpu.lua -- the solver test/converge-multigrid-vs-krylov.lua drives -- starts them from zero in every
cycle (cpu.lua:138) and does converge, slowly (omega = 1 Jacobi leaves the top mode nearly undamped). This script runs
that variant: mg_zero_corrections(); mg_vcycle(); and every `check` cycles the true residual RMS ||f - A psi|| / sqrt(N)
(mg_residual_norm), until it is <= tol x the initial residual. fp64 (fp32 cannot represent a 1e-8 relative residual).

    python tools/time_to_tolerance.py                 # GPU, writes one JSON line per configuration
    python tools/time_to_tolerance.py --oracle         # the same loop on the CPU oracle (small sizes), for the cycle counts
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = [  # (dim, size, check interval)
    (2, 64, 1), (2, 256, 50), (2, 512, 200), (3, 32, 1), (3, 64, 10), (3, 128, 50), (3, 256, 200),
]


def solve_gpu(pkg, dim, size, tol, check, max_cycles, omega=1.0, rezero=True):
    s = pkg.MultigridCUDA(size, "double", dim=dim, out=False)
    if omega != 1.0:
        s.set_omega(omega)      # NOT the reference: weighted Jacobi, a labelled extension (mg_set_omega)
    r0 = s.residual_norm()
    s.zero_corrections(); s.vcycle(); s.residual_norm()      # warm-up: graph capture, first launches
    s.init_cells()
    t0 = time.perf_counter()
    c, r = 0, r0
    while c < max_cycles:
        for _ in range(check):
            if rezero:
                s.zero_corrections()
            s.vcycle()
        c += check
        r = s.residual_norm()                                 # synchronises: 8-byte readback
        if not (r > tol * r0):
            break
    dt = time.perf_counter() - t0
    out = {"cycles": c, "seconds": dt, "residual_rel": r / r0, "linf_psi": s.linf_norm()}
    s.close()
    return out


def solve_oracle(O, dim, size, tol, check, max_cycles, nthreads):
    o = O.Oracle(size, "double", dim, nthreads=nthreads)
    levels, L = [], size // 2
    while L >= 1:
        levels.append(L)
        L //= 2
    r0 = o.residual_rms()
    t0 = time.perf_counter()
    c, r = 0, r0
    while c < max_cycles:
        for _ in range(check):
            for L in levels:
                o.buffer(O.BUF_V, L)[...] = 0
            o.vcycle()
        c += check
        r = o.residual_rms()
        if not (r > tol * r0):
            break
    return {"cycles": c, "seconds": time.perf_counter() - t0, "residual_rel": r / r0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--oracle", action="store_true")
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--max-cycles", type=int, default=400000)
    ap.add_argument("--max-points", type=int, default=1 << 40)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--omega", type=float, default=1.0, help="relaxation weight (1 = the reference; anything else is an extension)")
    ap.add_argument("--keep-corrections", action="store_true", help="cpu-raw.lua form: coarse corrections carried over between cycles")
    a = ap.parse_args()
    if a.oracle:
        import oracle as O
    else:
        from __graft_entry__ import load_package
        pkg = load_package()
    configs = CONFIGS if a.omega == 1.0 else [(2, 64, 1), (2, 256, 1), (2, 2048, 1), (3, 64, 1), (3, 128, 1), (3, 256, 1), (3, 512, 1)]
    for dim, size, check in configs:
        if size ** dim > a.max_points:
            continue
        r = solve_oracle(O, dim, size, a.tol, check, a.max_cycles, a.threads) if a.oracle else \
            solve_gpu(pkg, dim, size, a.tol, check, a.max_cycles if a.omega == 1.0 else min(a.max_cycles, 2000), a.omega, not a.keep_corrections)
        r.update({"dim": dim, "size": size, "real": "double", "tol": a.tol, "check_every": check, "omega": a.omega,
                  "variant": ("cpu.lua (coarse corrections re-zeroed every cycle)" if not a.keep_corrections else
                              "cpu-raw.lua (coarse corrections carried over)") + ("" if a.omega == 1.0 else
                              f"; NON-REFERENCE smoother: weighted Jacobi omega = {a.omega:.6g}, one-sweep-per-launch kernels"),
                  "impl": "oracle" if a.oracle else "cuda"})
        r["ms_per_cycle"] = 1e3 * r["seconds"] / max(r["cycles"], 1)
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
