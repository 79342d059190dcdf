# Round-end measurement on one B200: parity tests, the default bench line, then (each only after the same command
# exited 0 without ncu) the ncu launch list and the --set full capture that tools/make_profiles.py summarises.
TAG=${TAG:-final}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/${TAG}_tests.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu1.log 2>&1; echo "ncu1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_stream3d -s 0 -c 12 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu2.log 2>&1; echo "ncu2 rc=$?"
for cfg in "2 4096 float" "2 2048 double" "2 64 double"; do set -- $cfg
  python bench.py --dim $1 --size $2 --real $3 --no-cpu > gpurun_out/${TAG}_b2_dim$1size$2real$3.json 2>/dev/null; echo "bench $cfg rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_b*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), d["roofline"]["frac"], d["e2e"]["value"])
    except Exception as e: print(f,"ERR",e)
PY
