TAG=${TAG:-a5}
python -m pytest tests/test_gpu_small_smem.py -m gpu -x -q -k "block" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/${TAG}_tests.log | cut -c1-400
run() { n=$1; shift
  python bench.py --steps 30 --warmup 5 --no-cpu --quick "$@" > gpurun_out/${TAG}_b_$n.json 2> gpurun_out/${TAG}_b_$n.err; echo "bench $n rc=$?"; }
run b0 --opt block_max_L=0
run b32 --opt block_max_L=32
run b64
run c5_tb4 --config c5
run c5_tb5 --config c5 --opt tb2=5
run c5_tb7 --config c5 --opt tb2=7
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_b_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); b=d["vcycle"]["breakdown_all_ms"]; lv={}
        for k,v in b.items():
            L=int(k.split("L=")[1].split(",")[0]); lv[L]=lv.get(L,0)+v
        print(f.split("_b_")[1][:-5].ljust(8), round(d["value"],1), round(d["ms_per_step"]*1e3,1), d["gpu_launches"], {L:round(v*1e3,1) for L,v in lv.items()})
    except Exception as e: print(f,"ERR",e)
PY
