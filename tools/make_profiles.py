"""Turns the ncu artefacts a gpurun call brought back into the committed summaries under profiles/.

    python tools/make_profiles.py <launches.csv> <full.ncu-rep> [round_tag]

  * <launches.csv>: `ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file ...
                     python bench.py --steps 3 --warmup 3 --no-cpu`
  * <full.ncu-rep>: `ncu --set full --clock-control none --import-source on -k regex:k_stream3d -s 0 -c 12 ...`
                    (the 12 streaming-smoother launches of the first V-cycle at 512^3: L = 512, 256, 128 pre
                    passes, then the post passes at 128, 256, 512)
Writes profiles/<tag>_ncu_launches_3d512_f32.csv, profiles/<tag>_ncu_full_stream3d.md and
profiles/dram_traffic.json (DRAM bytes per launch, read by bench.py for `roofline.traffic`).
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
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}


def short(name):
    m = re.match(r"void (?:mg::)?(\w+)<(.*?)>\(", name)
    return (m.group(1) + "<" + m.group(2) + ">") if m else name[:60]


def launches(path, tag):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5 and r[0] != "ID"]
    recs = [(short(r[4]), r[8], r[7], int(r[14])) for r in rows]
    starts = [i for i, (n, g, b, t) in enumerate(recs) if n.startswith("k_stream3d<float, float, 4, 0, 0") and t > 400000]
    a, b = starts[1], starts[2]
    cyc = recs[a:b]
    tot = sum(t for *_, t in cyc)
    agg = collections.OrderedDict()
    for n, g, bk, t in cyc:
        e = agg.setdefault((n, g, bk), [0, 0])
        e[0] += 1
        e[1] += t
    out = ["# ncu launch list of one V-cycle (3-D 512^3 fp32, default tuning)",
           "# command: ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file launches.csv "
           "python bench.py --steps 3 --warmup 3 --no-cpu   (right after the same command exited 0 without ncu)",
           "# times under ncu are cold-cache and serialised: compare SHARES with bench.py's vcycle.breakdown_ms, not absolutes",
           "kernel,grid,block,launches,total_us,share_pct"]
    for (n, g, bk), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f'"{n}","{g}","{bk}",{c},{t / 1e3:.1f},{100 * t / tot:.2f}')
    out.append(f"TOTAL,,,{len(cyc)},{tot / 1e3:.1f},100")
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_launches_3d512_f32.csv"), "w").write("\n".join(out) + "\n")
    return len(cyc), tot


def full(rep, tag):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]

    def val(r, w):
        i = hdr.index(w)
        return float(r[i]) * SCALE.get(units[i], 1)

    # launch order of the streaming smoother within a V-cycle at 512^3 (tb = 4)
    order = [("sweep", 512, 4), ("sweep+residual_restrict", 512, 3), ("sweep", 256, 4), ("sweep+residual_restrict", 256, 3),
             ("sweep", 128, 4), ("sweep+residual_restrict", 128, 3), ("sweep+prolong_add", 128, 4), ("sweep", 128, 3),
             ("sweep+prolong_add", 256, 4), ("sweep", 256, 3), ("sweep+prolong_add", 512, 4), ("sweep", 512, 3)]
    lines = [f"| launch | grid x block | time (ncu, cold) | DRAM read + write | DRAM / A_op | DRAM GB/s (% of {PEAK:.1f}) | issue slots busy | "
             "regs | smem wavefronts (conflict replays) | LSU-smem busy |", "|---|---|---|---|---|---|---|---|---|---|"]
    traffic = {}
    for (kind, L, sw), r in zip(order, rows[2:]):
        name = f"{kind}[L={L},sweeps={sw}]"
        want_pro, want_res = "prolong" in kind, "restrict" in kind
        kn = r[hdr.index("Kernel Name")]
        m = re.search(r"k_stream3d<float, float, (\d), (\d), (\d)", kn.replace("(int)", "").replace("(bool)", ""))
        if m and (int(m.group(1)) != sw or bool(int(m.group(2))) != want_pro or bool(int(m.group(3))) != want_res):
            name += " (?)"
        rd, wr, t = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum"), val(r, "gpu__time_duration.sum")
        wf = val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
        bc = val(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
        cyc = val(r, "sm__cycles_elapsed.max")
        aop = (3 * sw + (4.125 if (want_pro or want_res) else 0)) * 4 * L ** 3
        lines.append(f"| `{name}` | {r[hdr.index('launch__grid_size')]} x {r[hdr.index('launch__block_size')]} | {t * 1e3:.3f} ms | "
                     f"{rd / 1e9:.3f} + {wr / 1e9:.3f} GB | {(rd + wr) / 1e9:.3f} / {aop / 1e9:.3f} = {(rd + wr) / aop:.2f} | "
                     f"{(rd + wr) / t / 1e9:.0f} ({(rd + wr) / t / 1e9 / PEAK * 100:.0f} %) | "
                     f"{float(r[hdr.index('smsp__issue_active.avg.pct_of_peak_sustained_active')]):.1f} % | "
                     f"{r[hdr.index('launch__registers_per_thread')]} | {wf / 1e6:.1f} M ({100 * bc / max(wf, 1):.1f} %) | "
                     f"{100 * wf / 148 / cyc:.0f} % |")
        traffic[name] = rd + wr
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    tmp = "/tmp/_src_page.csv"
    open(tmp, "w").write(src)
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_src_summary.py"), tmp, "12"],
                          capture_output=True, text=True).stdout
    return lines, traffic, summ


def main():
    lcsv, rep = sys.argv[1], sys.argv[2]
    tag = sys.argv[3] if len(sys.argv) > 3 else "r1"
    n, tot = launches(lcsv, tag)
    lines, traffic, summ = full(rep, tag)
    head = [f"# ncu --set full capture of the streaming smoother `k_stream3d` ({tag}, final kernels)", "",
            "Commands (on a B200, each right after the same command exited 0 without ncu):", "", "```",
            "ncu --set full --clock-control none --import-source on -k regex:k_stream3d -s 0 -c 12 -o prof \\",
            "    python bench.py --steps 3 --warmup 3 --no-cpu", "```", "",
            "3-D 512^3 fp32, default tuning: tb = 4 (pre = [4][3+RES], post = [PRO+4][3]), tile 56 x 40 (+halo 64 x 48),",
            "balanced persistent partition (one CTA per SM), f ring via TMA, packed fp32 arithmetic. The twelve launches are the",
            "streaming-smoother passes of one V-cycle (L = 512, 256, 128). A_op = algorithmic bytes of the reference operators a",
            "launch replaces (SURVEY 8(d)); DRAM = dram__bytes_read.sum + dram__bytes_write.sum of that launch.",
            f"One V-cycle = {n} kernel launches, {tot / 1e6:.3f} ms under ncu (see the launch list next to this file).", ""]
    tail = ["", "## Source-level summary of the first launch (`ncu --page source`, tools/ncu_src_summary.py)", "", "```"] + \
        summ.splitlines() + ["```", ""]
    note_path = os.path.join(ROOT, "profiles", f"{tag}_ncu_reading.md")
    note = open(note_path).read().splitlines() if os.path.exists(note_path) else []
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_stream3d.md"), "w").write("\n".join(head + lines + [""] + note + tail))
    json.dump({"3d_512_float": {k: v for k, v in traffic.items() if "(?)" not in k},
               "_source": f"profiles/{tag}_ncu_full_stream3d.md (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum per launch)"},
              open(os.path.join(ROOT, "profiles", "dram_traffic.json"), "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
