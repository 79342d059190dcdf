# A/B of the column-parts partition: parity tests of the streaming kernel and the slabs, bench lines, and the
# per-launch profile of one rank's share of the 8-slab 1024^3 cycle on a single GPU (tools/slab_profile.py)
TAG=${TAG:-a1}
python -m pytest tests/test_gpu_stream3d.py tests/test_gpu_slabs.py tests/test_gpu_vcycle.py -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/${TAG}_tests.log
for o in 1 0; do
  python bench.py --steps 30 --warmup 5 --no-cpu --opt colparts=$o > gpurun_out/${TAG}_bench_cp$o.json 2> gpurun_out/${TAG}_bench_cp$o.err; echo "bench colparts=$o rc=$?"
done
for o in 1 0; do
  python tools/slab_profile.py --size 1024 --slabs 8 --opt colparts=$o --tag cp$o > gpurun_out/${TAG}_slab8_cp$o.json 2> gpurun_out/${TAG}_slab8_cp$o.err; echo "slab8 colparts=$o rc=$?"
done
MGPOISSON_SLAB_MIN_PLANES=16 python tools/slab_profile.py --size 1024 --slabs 8 --tag cp1_min16 > gpurun_out/${TAG}_slab8_min16.json 2> gpurun_out/${TAG}_slab8_min16.err; echo "slab8 min16 rc=$?"
python - <<PY
import json
for o in (1,0):
    try:
        d=json.loads(open(f"gpurun_out/${TAG}_bench_cp{o}.json").read().strip().splitlines()[-1]); print("bench cp",o, round(d["value"],1), d["vcycle"]["breakdown_all_ms"])
    except Exception as e: print(o,"ERR",e)
for n in ("cp1","cp0","min16"):
    try:
        d=json.loads(open(f"gpurun_out/${TAG}_slab8_{n}.json").read().strip().splitlines()[-1]); print("slab8",n, d["rank0_sum_ms"], d["per_slab_ms"], d["per_level_ms"]); print(d["launches"])
    except Exception as e: print(n,"ERR",e)
PY
