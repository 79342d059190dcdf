#!/usr/bin/env python
"""cycle_times.py -- time every V-cycle after initCells on its own (CUDA events), to see whether the first cycles from the
reference's point-source state cost more than the later ones (guarded re-runs of passes that meet tiny numerators).

    python tools/cycle_times.py [--size 512] [--real float] [--cycles 16] [--opt name=value ...]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--real", default="float")
    ap.add_argument("--cycles", type=int, default=16)
    ap.add_argument("--opt", action="append", default=[])
    args = ap.parse_args()
    from __graft_entry__ import load_package
    pkg = load_package()
    s = pkg.MultigridCUDA(args.size, args.real, dim=args.dim, out=False)
    for kv in args.opt:
        k, v = kv.split("=")
        s.set_option(k, int(v))
    for _ in range(3):
        s.vcycle()
    out = {"size": args.size, "dim": args.dim, "real": args.real, "opt": args.opt}
    for label, zero in (("cpu_raw_variant", False), ("cpu_lua_variant_rezeroed", True)):
        s.init_cells()
        s.zero_corrections()
        ms = []
        for _ in range(args.cycles):
            if zero:
                s.zero_corrections()
            ms.append(round(s.time_vcycles(1), 4))
        out[label + "_ms_per_cycle"] = ms
    print(json.dumps(out))


if __name__ == "__main__":
    main()
