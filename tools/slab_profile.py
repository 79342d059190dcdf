#!/usr/bin/env python
"""slab_profile.py -- per-launch CUDA-event times of ONE rank's share of a slab V-cycle, measured on a single GPU.

`mg_create_slab_local` puts every slab of an N-rank run on one device and one stream: slab 0's kernels are the
kernels (same grids, same partitions, same planes) rank 0 of a real N-GPU run launches, minus NVLink stores and the
neighbour handshake. Their times show where a rank's cycle goes without spending N GPUs on the question.

    python tools/slab_profile.py [--size 1024] [--slabs 8] [--opt name=value ...]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--slabs", type=int, default=8)
    ap.add_argument("--real", default="float")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--opt", action="append", default=[])
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    from __graft_entry__ import load_package
    pkg = load_package()
    s = pkg.MultigridCUDA(args.size, args.real, dim=3, out=False, local_slabs=args.slabs)
    for kv in args.opt:
        k, v = kv.split("=")
        s.set_option(k, int(v))
    s.vcycle()
    s.init_cells()
    s.zero_corrections()
    recs = [s.profile_vcycle() for _ in range(args.reps)]
    med = []
    for i in range(len(recs[0])):
        r = dict(recs[0][i])
        r["ms"] = sorted(x[i]["ms"] for x in recs)[len(recs) // 2]
        med.append(r)
    groups = {}
    for r in med:
        g = groups.setdefault((r["L"], r["kind"], r["sweeps"]), {"ms": 0.0, "n": 0})
        g["ms"] += r["ms"]
        g["n"] += 1
    levels = {}
    for (L, k, sw), g in groups.items():
        levels[L] = levels.get(L, 0.0) + g["ms"]
    ms_cycle = s.time_vcycles(10) / 10   # all slabs one after the other on this one GPU
    out = {"tool": "slab_profile", "tag": args.tag, "size": args.size, "slabs": args.slabs, "real": args.real, "opt": args.opt,
           "rank0_sum_ms": round(sum(levels.values()), 4),
           "all_slabs_serial_ms_per_cycle": round(ms_cycle, 4), "per_slab_ms": round(ms_cycle / args.slabs, 4),
           "per_level_ms": {str(L): round(v, 4) for L, v in sorted(levels.items(), reverse=True)},
           "launches": {f"{k}[L={L},sweeps={sw}]x{g['n']}": round(g["ms"], 4)
                        for (L, k, sw), g in sorted(groups.items(), key=lambda kv: (-kv[0][0], kv[0][1], kv[0][2]))}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
