# N GPUs (N = $NG): the driver's own line at that N, and the N=1 line on the same box
TAG=${TAG:-nx}; NG=${NG:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n$NG.json 2> gpurun_out/${TAG}_bench_n$NG.err; echo "bench n$NG rc=$?"
python bench.py --steps 20 --warmup 5 --no-cpu --quick > gpurun_out/${TAG}_bench_n1.json 2>/dev/null; echo "bench n1 rc=$?"
python - <<PY
import json
for n in ("n1","n$NG"):
    try:
        d=json.loads(open(f"gpurun_out/${TAG}_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"],1), round(d["ms_per_step"],4), d.get("parity"), (d.get("slab_timeline") or {}).get("cycle_us_mean"), (d.get("time_to_tolerance") or {}).get("cycles"))
    except Exception as e: print(n, "ERR", e)
PY
