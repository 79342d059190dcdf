# A/B of library builds on one box: parity tests for the default build, then bench lines for every
# lua-multigrid-poisson_b200/libmgpoisson_<variant>.so named in $VARIANTS (plus the default build as "new").
TAG=${TAG:-ab}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests(new) rc=$?"
for v in ${VARIANTS:-base} new; do
  lib=$PWD/lua-multigrid-poisson_b200/libmgpoisson_$v.so; [ $v = new ] && lib=$PWD/lua-multigrid-poisson_b200/libmgpoisson.so
  MGPOISSON_LIB=$lib python bench.py --steps ${STEPS:-30} --warmup 5 --no-cpu > gpurun_out/${TAG}_b_$v.json 2> gpurun_out/${TAG}_b_$v.err; echo "bench $v rc=$?"
done
python - <<PY
import json
for v in "${VARIANTS:-base} new".split():
    try:
        d=json.loads(open(f"gpurun_out/${TAG}_b_{v}.json").read().strip().splitlines()[-1])
        print(v, round(d["value"],1), [(k.replace("sweep","s").replace("prolong_add","P").replace("residual_restrict","R").replace(",sweeps=","/"),v) for k,v in d["vcycle"]["breakdown_ms"].items()])
    except Exception as e: print(v, "ERR", e)
PY
tail -n 3 gpurun_out/${TAG}_tests*.log
if [ -n "$NCU" ]; then
  ncu --set full --clock-control none --import-source on -k regex:k_stream3d -s 0 -c ${NCU} -o gpurun_out/${TAG}_prof -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
fi
# extra bench lines of the default build with library options: EXTRA_OPTS="name:--opt a=1 --opt b=2;name2:..."
if [ -n "$EXTRA_OPTS" ]; then
  IFS=';' read -ra items <<< "$EXTRA_OPTS"
  for it in "${items[@]}"; do
    nm=${it%%:*}; op=${it#*:}
    MGPOISSON_LIB=${EXTRA_LIB:-$PWD/lua-multigrid-poisson_b200/libmgpoisson.so} python bench.py --steps ${STEPS:-30} --warmup 5 --no-cpu $op > gpurun_out/${TAG}_o_$nm.json 2> gpurun_out/${TAG}_o_$nm.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_o_$nm.json").read().strip().splitlines()[-1])
print("$nm", round(d["value"],1), [(k.replace("sweep","s").replace("prolong_add","P").replace("residual_restrict","R").replace(",sweeps=","/"),v) for k,v in d["vcycle"]["breakdown_ms"].items()][:4])
PY
  done
fi
