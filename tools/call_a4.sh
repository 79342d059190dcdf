TAG=${TAG:-a4}
python -m pytest tests/test_gpu_small_smem.py tests/test_gpu_vcycle.py -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -n 4 gpurun_out/${TAG}_tests.log | cut -c1-400
python tools/cycle_times.py > gpurun_out/${TAG}_cycles.json 2>&1; cat gpurun_out/${TAG}_cycles.json | cut -c1-900
run() { n=$1; shift
  python bench.py --steps 30 --warmup 5 --no-cpu --quick "$@" > gpurun_out/${TAG}_b_$n.json 2> gpurun_out/${TAG}_b_$n.err; echo "bench $n rc=$?"; }
run b0 --opt block_max_L=0
run b64
run b128 --opt block_max_L=128
python tools/slab_profile.py > gpurun_out/${TAG}_slab8.json 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_b_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); b=d["vcycle"]["breakdown_all_ms"]; lv={}
        for k,v in b.items():
            L=int(k.split("L=")[1].split(",")[0]); lv[L]=lv.get(L,0)+v
        print(f.split("_b_")[1][:-5].ljust(8), round(d["value"],1), round(d["ms_per_step"]*1e3,1), d["gpu_launches"], {L:round(v*1e3,1) for L,v in lv.items()})
    except Exception as e: print(f,"ERR",e)
d=json.loads(open("gpurun_out/${TAG}_slab8.json").read().strip().splitlines()[-1]); print("slab8", d["rank0_sum_ms"], d["per_slab_ms"], d["per_level_ms"])
PY
