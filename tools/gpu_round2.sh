# Round-2 measurement on one B200: bench lines of the BASELINE configs, then (each only after the same command exited 0
# without ncu) the ncu launch list and the --set full captures that tools/make_profiles.py summarises.
TAG=${TAG:-r2}
python bench.py > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 rc=$?"
python bench.py --config c2 --no-cpu > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err; echo "bench c2 rc=$?"
python bench.py --config c5 --no-cpu > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; echo "bench c5 rc=$?"
python bench.py --config c1 --no-cpu > gpurun_out/${TAG}_bench_c1.json 2> gpurun_out/${TAG}_bench_c1.err; echo "bench c1 rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_c*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d["roofline"]
        print(f, round(d["value"],1), "frac", round(r["frac"],3), r["kernel"], "eff", round(r["effective_frac_A_op"],2), "e2e", round(d["e2e"]["value"],1), d.get("mg_vs_cg"))
        print("   ", d["vcycle"]["breakdown_ms"])
    except Exception as e: print(f,"ERR",e)
PY
if [ -n "$NCU" ]; then
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_3d_512_float.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu1.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_stream3d -s 0 -c 20 -f -o gpurun_out/prof_${TAG}_3d_512_float \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu2.log 2>&1; echo "ncu full 3d rc=$?"
python bench.py --config c2 --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_warp2d|k_small' -s 0 -c 13 -f -o gpurun_out/prof_${TAG}_2d_4096_float \
    python bench.py --config c2 --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu3.log 2>&1; echo "ncu full c2 rc=$?"
python bench.py --config c5 --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_warp2d|k_small' -s 0 -c 21 -f -o gpurun_out/prof_${TAG}_2d_2048_double \
    python bench.py --config c5 --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu4.log 2>&1; echo "ncu full c5 rc=$?"
fi
