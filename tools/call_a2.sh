# PDL + shared-memory small-level kernel: parity tests, then A/B bench lines on one box
TAG=${TAG:-a2}
python -m pytest tests/test_gpu_small_smem.py tests/test_gpu_vcycle.py tests/test_gpu_stream3d.py tests/test_gpu_warp2d.py tests/test_gpu_slabs.py -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -n 15 gpurun_out/${TAG}_tests.log | cut -c1-400
run() { # name, args...
  n=$1; shift
  python bench.py --steps 30 --warmup 5 --no-cpu "$@" > gpurun_out/${TAG}_b_$n.json 2> gpurun_out/${TAG}_b_$n.err; echo "bench $n rc=$?"
}
run c3_base --opt pdl=0 --opt small_smem=0
run c3_pdl --opt pdl=1 --opt small_smem=0
run c3_smem --opt pdl=0 --opt small_smem=1
run c3_both
run c3_both_fast128 --opt fast_min_L=128
run c3_both_s64 --opt stream_min_L=64
for c in c1 c2 c5; do
  run ${c}_base --config $c --opt pdl=0 --opt small_smem=0
  run ${c}_both --config $c
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_b_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); b=d["vcycle"]["breakdown_all_ms"]
        print(f.split("_b_")[1][:-5], round(d["value"],1), round(d["ms_per_step"]*1e3,1), {k.replace("sweep","s").replace("prolong_add","P").replace("residual_restrict","R").replace(",sweeps=","/"):v for k,v in b.items() if "L=512" not in k and "L=4096" not in k})
    except Exception as e: print(f,"ERR",e)
PY
