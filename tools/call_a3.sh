# relaxed column parts + fast_min_L=128 + regrouped small-level kernel (PDL off): tests, bench lines
TAG=${TAG:-a3}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -n 5 gpurun_out/${TAG}_tests.log | cut -c1-400
run() { n=$1; shift
  python bench.py --steps 30 --warmup 5 --no-cpu "$@" > gpurun_out/${TAG}_b_$n.json 2> gpurun_out/${TAG}_b_$n.err; echo "bench $n rc=$?"; }
run c3
run c3_fast256 --opt fast_min_L=256
run c3_cp0 --opt colparts=0
run c3_tb2 --tb 2
run c3_tb3 --tb 3
for c in c1 c2 c5; do run $c --config $c; done
python tools/slab_profile.py --size 1024 --slabs 8 --tag new > gpurun_out/${TAG}_slab8.json 2> gpurun_out/${TAG}_slab8.err; echo "slab8 rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_b_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); b=d["vcycle"]["breakdown_all_ms"]; lv={}
        for k,v in b.items():
            L=int(k.split("L=")[1].split(",")[0]); lv[L]=lv.get(L,0)+v
        print(f.split("_b_")[1][:-5].ljust(12), round(d["value"],1), round(d["ms_per_step"]*1e3,1), {L:round(v*1e3,1) for L,v in lv.items()})
    except Exception as e: print(f,"ERR",e)
d=json.loads(open("gpurun_out/${TAG}_slab8.json").read().strip().splitlines()[-1]); print("slab8", d["rank0_sum_ms"], d["per_slab_ms"], d["per_level_ms"])
PY
