#!/usr/bin/env python
"""ncu_csv.py FILE... : per-launch table of an `ncu --metrics ... --csv` log (one row per kernel launch)."""
import csv, sys, collections
for fn in sys.argv[1:]:
    rows = list(csv.reader(l for l in open(fn) if l.startswith('"')))
    hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
    launches = collections.OrderedDict()
    for r in rows[1:]:
        key = (r[idx["ID"]], r[idx["Kernel Name"]][:60])
        launches.setdefault(key, {})[r[idx["Metric Name"]]] = (r[idx["Metric Value"]], r[idx["Metric Unit"]])
    print("==", fn)
    for (i, name), m in launches.items():
        def g(k, scale=1.0):
            v = m.get(k)
            if not v: return float("nan")
            return float(v[0].replace(",", "")) * scale
        unit = lambda k: m.get(k, ("", ""))[1]
        t = g("gpu__time_duration.sum"); tu = unit("gpu__time_duration.sum")
        t_ms = t / 1e6 if tu in ("ns", "nsecond") else (t / 1e3 if tu in ("us", "usecond") else t)
        def gb(k):
            v = g(k); u = unit(k)
            return v * {"Gbyte": 1, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9, "Tbyte": 1e3}.get(u, 1e-9)
        rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
        print(f"{i:>3} {name[:44]:44s} {t_ms:7.4f} ms  dram rd {rd:6.3f} wr {wr:6.3f} GB -> {(rd+wr)/t_ms:7.1f} GB/s  L2hit {g('lts__t_sector_hit_rate.pct'):5.1f}%  "
              f"issue {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f}%  inst {g('smsp__inst_executed.sum')/1e6:7.1f}M  "
              f"smem wf {g('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum')/1e6:7.1f}M  lsu wf {g('l1tex__data_pipe_lsu_wavefronts.sum')/1e6:7.1f}M  "
              f"L2 bytes {gb('lts__t_bytes.sum'):6.2f} GB  warps {g('sm__warps_active.avg.pct_of_peak_sustained_active'):4.1f}%")
