"""Turns the ncu artefacts a gpurun call brought back into the committed summaries under profiles/.

    python tools/make_profiles.py launches <launches.csv> <tag> [config]     # per-kernel shares of one V-cycle
    python tools/make_profiles.py full <full.ncu-rep> <tag> <config>         # --set full capture of the fused passes
    python tools/make_profiles.py sass <tag>                                 # SASS mnemonic counts of the built library

  config: 3d_512_float (default) | 2d_4096_float | 2d_2048_double -- the bench.py configuration the capture ran
  * <launches.csv>: `ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file ... python bench.py ...`
  * `full` takes the metrics CSV of the FIRST V-cycle bench.py runs (`ncu --metrics <METRICS below> -k regex:'k_stream3d|k_warp2d|k_small'
    -c N --csv --log-file m.csv ...`); launches are matched, in order, with the pass schedule of a cycle. A `--set full
    --import-source on` capture of the dominant kernel next to it (prof_<tag>_<config>_top.ncu-rep) adds the source-level summary.
`full` writes profiles/<tag>_ncu_full_<config>.md and merges DRAM bytes per launch into profiles/dram_traffic.json,
which bench.py reads for `roofline.traffic` / `roofline.frac`.
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAK = 6544.7
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1,
         "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1}


def short(name):
    name = name.replace("(int)", "").replace("(bool)", "")
    m = re.match(r"void (?:mg::)?(\w+)<(.*?)>\(", name)
    return (m.group(1) + "<" + m.group(2) + ">") if m else name[:60]


def launches(path, tag, config):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    iname, ival, igrid, iblock = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    recs = [(short(r[iname]), r[igrid], r[iblock], float(r[ival].replace(",", ""))) for r in rows[1:]]
    # one V-cycle = the launches between two consecutive launches of the one-CTA small-level kernel (the bottom of the
    # V): the second half of one cycle and the first half of the next -- the same multiset of launches as one cycle
    smalls = [i for i, r in enumerate(recs) if r[0].startswith(("k_small_vcycle", "k_cluster_vcycle"))]
    if len(smalls) >= 3:
        cyc = recs[smalls[1]:smalls[2]]
    else:
        cyc = recs
    tot = sum(t for *_, t in cyc)
    agg = collections.OrderedDict()
    for n, g, bk, t in cyc:
        e = agg.setdefault((n, g, bk), [0, 0.0])
        e[0] += 1
        e[1] += t
    out = [f"# ncu launch list of one V-cycle ({config}, default tuning)",
           "# command: ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file launches.csv "
           "python bench.py ... --steps 3 --warmup 3 --no-cpu   (right after the same command exited 0 without ncu)",
           "# times under ncu are cold-cache and serialised: compare SHARES with bench.py's vcycle.breakdown_ms, not absolutes",
           "kernel,grid,block,launches,total_us,share_pct"]
    for (n, g, bk), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f'"{n}","{g}","{bk}",{c},{t / 1e3:.1f},{100 * t / tot:.2f}')
    out.append(f"TOTAL,,,{len(cyc)},{tot / 1e3:.1f},100")
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_launches_{config}.csv"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[3:]))


def schedule(config):
    """(kind, L, sweeps) of the fused-pass launches of one V-cycle, in launch order (mirrors EngineT::sweeps)."""
    dim, size, real = config.split("_")
    dim, size = int(dim[0]), int(size)
    if dim == 3:
        tb, minL, small = 4, 128, 16
        pre = lambda: [("sweep", 4), ("sweep+residual_restrict", 3)]
        post = lambda: [("sweep+prolong_add", 4), ("sweep", 3)]
    else:
        tb2 = 7 if real == "float" else 4
        minL, small = 128, 64
        if tb2 == 7:
            pre = lambda: [("sweep+residual_restrict", 7)]
            post = lambda: [("sweep+prolong_add", 7)]
        else:
            pre = lambda: [("sweep", 4), ("sweep+residual_restrict", 3)]
            post = lambda: [("sweep+prolong_add", 4), ("sweep", 3)]
    levels = []
    L = size
    while L >= minL:
        levels.append(L)
        L //= 2
    order = [(k, L, s) for L in levels for k, s in pre()]
    if dim == 2:
        order.append(("small_levels", small, 14))
    order += [(k, L, s) for L in reversed(levels) for k, s in post()]
    return dim, size, real, order


def read_metrics_csv(path):
    """`ncu --metrics a,b,c --csv --log-file X`: one row per (launch, metric). Returns [(kernel name, {metric: SI value})]."""
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    out = collections.OrderedDict()
    for r in rows[1:]:
        key = r[ix["ID"]]
        e = out.setdefault(key, {"name": short(r[ix["Kernel Name"]]), "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]], "m": {}})
        try:
            e["m"][r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * SCALE.get(r[ix["Metric Unit"]], 1)
        except ValueError:
            pass
    return list(out.values())


METRICS = ("dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,"
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__cycles_elapsed.max,"
           "launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,"
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum")


def full(rep, tag, config):
    """rep = the metrics CSV of the first V-cycle (`ncu --metrics <METRICS> --csv --log-file ...`)"""
    launches_ = read_metrics_csv(rep)
    dim, size, real, order = schedule(config)
    elem = 8 if real == "double" else 4
    data = []
    for e in launches_:
        kn = e["name"]
        if kn.startswith("k_stream3d"):
            a = [x.strip() for x in kn[kn.index("<") + 1:-1].split(",")]
            if len(a) >= 8 and a[7] == "2":
                continue   # the (empty) guarded re-run kernel behind every branch-free pass
        if kn.startswith(("k_stream3d", "k_warp2d", "k_small", "k_cluster")):
            data.append(e)
    lines = [f"| launch | kernel | grid x block | time (ncu, cold) | DRAM read + write | compulsory | DRAM / compulsory | DRAM GB/s (% of {PEAK:.1f}) | "
             "issue slots busy | regs | smem wavefronts (conflict replays) | LSU-smem busy | L2 hit |", "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    traffic, pipes = {}, {}
    c = 2.0 ** -dim
    for (kind, L, sw), e in zip(order, data):
        name = f"{kind}[L={L},sweeps={sw}]"
        m, kn = e["m"], e["name"]
        rd, wr, t = m.get("dram__bytes_read.sum", 0), m.get("dram__bytes_write.sum", 0), m.get("gpu__time_duration.sum", 1)
        wf = m.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 0)
        bc = m.get("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 0)
        cyc = m.get("sm__cycles_elapsed.max", 1)
        if kind == "small_levels":
            comp = sum((6 + 2 * c) * float(l) ** dim for l in [L >> k for k in range(20)] if l > 1) * elem
        else:
            comp = (3 + (c if "+" in kind else 0)) * elem * float(L) ** dim
        lines.append(f"| `{name}` | `{kn[:kn.index('<') + 34] if '<' in kn else kn}...` | {e['grid']} x {e['block']} | {t * 1e3:.4f} ms | "
                     f"{rd / 1e9:.3f} + {wr / 1e9:.3f} GB | {comp / 1e9:.3f} GB | {(rd + wr) / comp:.2f} | "
                     f"{(rd + wr) / t / 1e9:.0f} ({(rd + wr) / t / 1e9 / PEAK * 100:.0f} %) | "
                     f"{m.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):.1f} % | "
                     f"{int(m.get('launch__registers_per_thread', 0))} | {wf / 1e6:.1f} M ({100 * bc / max(wf, 1):.1f} %) | "
                     f"{100 * wf / 148 / cyc:.0f} % | {m.get('lts__t_sector_hit_rate.pct', 0):.0f} % |")
        traffic[name] = rd + wr
        pipes[name] = {"issue_slots_busy_pct": round(m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0), 1),
                       "lsu_shared_memory_pipe_busy_pct": round(100 * wf / 148 / cyc, 1),
                       "dram_pct_of_measured_peak": round((rd + wr) / t / 1e9 / PEAK * 100, 1)}
    summ = ""
    srcrep = os.path.join(os.path.dirname(rep), f"prof_{tag}_{config}_top.ncu-rep")
    if not os.path.exists(srcrep):
        srcrep = None
    if srcrep:
        src = subprocess.run(["ncu", "-i", srcrep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        tmp = "/tmp/_src_page.csv"
        open(tmp, "w").write(src)
        summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_src_summary.py"), tmp, "12"],
                              capture_output=True, text=True).stdout
    head = [f"# ncu --set full capture of the fused passes of one V-cycle: {config} ({tag})", "",
            "Command (on a B200, right after the same command exited 0 without ncu):", "", "```",
            "ncu --metrics <tools/make_profiles.py METRICS> --clock-control none -k regex:'k_stream3d|k_warp2d|k_small' -c N --csv --log-file m.csv \\",
            "    python bench.py [--config ...] --steps 3 --warmup 3 --no-cpu",
            "ncu --set full --clock-control none --import-source on -k regex:<top kernel> -c 1 -o prof_top ...   (source-level summary below)", "```", "",
            "DRAM = dram__bytes_read.sum + dram__bytes_write.sum of the launch; compulsory = what the fused launch cannot avoid",
            "(read u and f, write u, + the coarse field it reads or writes). The (empty) guarded re-run kernels that follow the",
            "branch-free fp32 passes are left out. Times under ncu are cold-cache and serialised.", ""]
    tail = (["", "## Source-level summary of the dominant launch (`ncu --set full --import-source on`, `--page source`, tools/ncu_src_summary.py)", "", "```"] +
            summ.splitlines() + ["```", ""]) if summ else []
    note_path = os.path.join(ROOT, "profiles", f"{tag}_ncu_reading_{config}.md")
    note = open(note_path).read().splitlines() if os.path.exists(note_path) else []
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_{config}.md"), "w").write("\n".join(head + lines + [""] + note + tail))
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    try:
        tab = json.load(open(tpath))
    except Exception:
        tab = {}
    tab[config] = traffic
    allp = tab.get("_pipes", {})
    allp[config] = pipes     # ncu, same capture: which unit of the SM the launch keeps busiest (bench.py: roofline.binding_resource)
    tab["_pipes"] = allp
    src_note = tab.get("_source", {})
    if not isinstance(src_note, dict):
        src_note = {}
    src_note[config] = f"profiles/{tag}_ncu_full_{config}.md (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum per launch)"
    tab["_source"] = src_note
    json.dump(tab, open(tpath, "w"), indent=1)
    print("\n".join(lines))


def sass(tag):
    lib = os.path.join(ROOT, "lua-multigrid-poisson_b200", "libmgpoisson.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    ops = collections.Counter()
    for line in txt.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(1).split(".")[0]] += 1
    want = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "FADD2", "FFMA2", "FMUL2", "LDS", "STS", "SHFL", "LDTM", "STTM", "LDG", "LDGSTS", "LDGDEPBAR", "STG",
            "DADD", "DFMA", "DMUL", "HMMA", "BAR", "ST", "LD", "ATOMG", "RED", "VIMNMX3", "MUFU"]
    log = os.path.join(ROOT, "lua-multigrid-poisson_b200", "csrc", "build.log")
    spills = regs = ""
    if os.path.exists(log):
        t = open(log).read()
        ents = re.findall(r"Compiling entry function '(\S+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", t)
        hot = [e for e in ents if "k_stream3dIffLi4" in e[0] or "k_stream3dIffLi3" in e[0] or "k_warp2dIffLi7" in e[0] or "k_warp2dIddLi4" in e[0]]
        spills = "\n".join(f"{subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()[:96]:96s} regs {r:>3s} spill stores {ss:>4s} B loads {sl:>4s} B"
                           for n, st, ss, sl, r in hot)
        regs = f"{len(ents)} kernels compiled; {sum(1 for e in ents if int(e[2]) > 0)} with spills (the double-arithmetic 2-D pipelines and S <= 2 variants at 80 registers)"
    out = [f"# SASS summary of libmgpoisson.so ({tag}): cuobjdump -sass | mnemonic counts (static instructions, whole library)", ""]
    out += [f"{k:10s} {ops.get(k, 0)}" for k in want]
    out += ["", "# hot fp32 kernels (ptxas -v): mode 1 = branch-free, 2 = guarded re-run, 0 = guarded", spills, "", regs, ""]
    open(os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt"), "w").write("\n".join(out))
    print("\n".join(out[:30]))


def main():
    cmd = sys.argv[1]
    if cmd == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "3d_512_float")
    elif cmd == "full":
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "3d_512_float")
    elif cmd == "sass":
        sass(sys.argv[2])


if __name__ == "__main__":
    main()
