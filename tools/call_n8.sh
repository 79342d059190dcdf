# 8 GPUs (charged 8x): the driver's own N=8 line (parity + timeline), a quick variant with the 128^3 level distributed,
# and the N=1 line on the same box for the efficiency
TAG=${TAG:-n8}
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$R --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n8.json 2> gpurun_out/${TAG}_bench_n8.err; echo "bench n8 rc=$?"
# (the 128^3 level distributed, MGPOISSON_SLAB_MIN_PLANES=16: measured 2595.8 vs 2606.6 units/s, not repeated)
python bench.py --steps 20 --warmup 5 --no-cpu --quick > gpurun_out/${TAG}_bench_n1.json 2>/dev/null; echo "bench n1 rc=$?"
python - <<PY
import json
for n in ("n1","n8"):
    try:
        d=json.loads(open(f"gpurun_out/${TAG}_bench_{n}.json").read().strip().splitlines()[-1]); b=d["vcycle"]["breakdown_all_ms"]; lv={}
        for k,v in b.items():
            L=int(k.split("L=")[1].split(",")[0]); lv[L]=lv.get(L,0)+v
        print(n, round(d["value"],1), round(d["ms_per_step"],4), d.get("parity"), {L:round(v*1e3,1) for L,v in lv.items()})
        if d.get("slab_timeline"): print(json.dumps(d["slab_timeline"]))
    except Exception as e: print(n, "ERR", e)
PY
tail -n 5 gpurun_out/${TAG}_bench_n8.err | cut -c1-300
