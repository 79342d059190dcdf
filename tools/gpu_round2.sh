# Round-2 measurement on one B200: bench lines of the BASELINE configs, then (each only after the same command exited 0
# without ncu) the ncu launch list, the per-launch metric tables and the --set full capture of the dominant kernel
# that tools/make_profiles.py summarises into profiles/.
TAG=${TAG:-r2}
M=$(python -c "import sys; sys.path.insert(0,'tools'); import make_profiles as m; print(m.METRICS)")
if [ -z "$SKIP_BENCH" ]; then
python bench.py > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_c3_reference.json 2>/dev/null; echo "reference arm rc=$?"
for c in c2 c5 c1; do python bench.py --config $c --no-cpu > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err; echo "bench $c rc=$?"; done
python tools/time_to_tolerance.py --omega 0.857142857142857 --max-points 140000000 > gpurun_out/${TAG}_ttt_omega67.jsonl 2>&1; echo "ttt omega 6/7 rc=$?"
python tools/time_to_tolerance.py --omega 0.8 --max-points 5000000 > gpurun_out/${TAG}_ttt_omega45_2d.jsonl 2>&1; echo "ttt omega 4/5 rc=$?"
fi
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_c*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d.get("roofline") or {}
        print(f, round(d["value"],2), "frac", r.get("frac"), r.get("kernel"), "e2e", round(d["e2e"]["value"],1))
    except Exception as e: print(f,"ERR",e)
PY
if [ -n "$NCU" ]; then
B="python bench.py --steps 3 --warmup 3 --no-cpu"
$B > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_3d_512_float.csv $B > gpurun_out/${TAG}_ncu1.log 2>&1; echo "ncu launches rc=$?"
ncu --metrics $M --clock-control none -k regex:k_stream3d -s 0 -c 24 --csv --log-file gpurun_out/${TAG}_metrics_3d_512_float.csv $B > gpurun_out/${TAG}_ncu2.log 2>&1; echo "ncu metrics 3d rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_stream3d<float, float, \(int\)4, \(bool\)1" -s 0 -c 1 -f -o gpurun_out/prof_${TAG}_3d_512_float_top $B > gpurun_out/${TAG}_ncu3.log 2>&1; echo "ncu full top rc=$?"
B="python bench.py --config c2 --steps 3 --warmup 3 --no-cpu"
$B > gpurun_out/${TAG}_plain2.log 2>&1 && \
ncu --metrics $M --clock-control none -k regex:'k_warp2d|k_small' -s 0 -c 13 --csv --log-file gpurun_out/${TAG}_metrics_2d_4096_float.csv $B > gpurun_out/${TAG}_ncu4.log 2>&1; echo "ncu metrics c2 rc=$?"
ncu --set full --clock-control none -k regex:'k_warp2d' -s 0 -c 1 -f -o gpurun_out/prof_${TAG}_2d_4096_float_top $B > gpurun_out/${TAG}_ncu5.log 2>&1; echo "ncu full c2 top rc=$?"
B="python bench.py --config c5 --steps 3 --warmup 3 --no-cpu"
$B > gpurun_out/${TAG}_plain3.log 2>&1 && \
ncu --metrics $M --clock-control none -k regex:'k_warp2d|k_small' -s 0 -c 21 --csv --log-file gpurun_out/${TAG}_metrics_2d_2048_double.csv $B > gpurun_out/${TAG}_ncu6.log 2>&1; echo "ncu metrics c5 rc=$?"
ncu --set full --clock-control none -k regex:'k_small' -s 0 -c 1 -f -o gpurun_out/prof_${TAG}_2d_2048_double_top $B > gpurun_out/${TAG}_ncu7.log 2>&1; echo "ncu full small rc=$?"
fi
# gpurun brings back at most 64 MiB: keep the tables, drop the largest reports first
ls -la gpurun_out/*.ncu-rep 2>/dev/null
for f in gpurun_out/prof_${TAG}_2d_2048_double_top.ncu-rep gpurun_out/prof_${TAG}_2d_4096_float_top.ncu-rep gpurun_out/prof_${TAG}_3d_512_float_top.ncu-rep; do
  [ $(du -sm gpurun_out | cut -f1) -gt 55 ] && rm -f $f && echo "dropped $f"
done
du -sh gpurun_out
