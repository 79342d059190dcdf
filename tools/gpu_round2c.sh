# (gpurun copies back at most 64 MiB: one .ncu-rep with sources is ~45 MiB, so the fp64 capture is kept as text.)
# Round-2, third session: the full -m gpu tier, the bench lines of every single-GPU BASELINE configuration, and (each only
# after the same command exited 0 without ncu) the ncu tables of the 2-D fp32 cycle with its new ring kernel plus one
# --set full capture of the 2-D fp64 pass; tools/make_profiles.py turns them into profiles/.
TAG=${TAG:-r2d}
M=$(python -c "import sys; sys.path.insert(0,'tools'); import make_profiles as m; print(m.METRICS)")
if [ -z "$SKIP_TESTS" ]; then python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/${TAG}_tests.log; fi
python bench.py > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
for c in c2 c5 c1; do python bench.py --config $c --no-cpu > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err; echo "bench $c rc=$?"; done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_c*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d.get("roofline") or {}
        print(f, round(d["value"],2), "frac", r.get("frac"), r.get("kernel"), "e2e", round(d["e2e"]["value"],1), d["e2e"].get("one_call_at_a_time",{}).get("value"))
    except Exception as e: print(f,"ERR",e)
PY
B="python bench.py --config c2 --steps 3 --warmup 3 --no-cpu --quick"
$B > gpurun_out/${TAG}_plain2.log 2>&1 && \
ncu --metrics $M --clock-control none -k regex:'k_warp2d|k_small' -s 0 -c 13 --csv --log-file gpurun_out/${TAG}_metrics_2d_4096_float.csv $B > gpurun_out/${TAG}_ncu4.log 2>&1; echo "ncu metrics c2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_warp2d' -s 0 -c 1 -f -o gpurun_out/prof_${TAG}_2d_4096_float_top $B > gpurun_out/${TAG}_ncu5.log 2>&1; echo "ncu full c2 top rc=$?"
B="python bench.py --config c5 --steps 3 --warmup 3 --no-cpu --quick"
$B > gpurun_out/${TAG}_plain3.log 2>&1 && \
ncu --section SpeedOfLight --section ComputeWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section Occupancy --section MemoryWorkloadAnalysis --section InstructionStats --clock-control none -k regex:'k_warp2d' -s 0 -c 1 $B > gpurun_out/${TAG}_ncu6_2d_2048_double_warp_details.log 2>&1; echo "ncu full c5 warp rc=$?"
ls -la gpurun_out/*${TAG}*.ncu-rep 2>/dev/null
du -sh gpurun_out
