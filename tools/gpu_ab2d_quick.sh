# A/B of library builds for the 2-D configurations without the test run: bench lines (--quick) for every variant in
# $VARIANTS plus the default build ("new"); REPS repetitions interleaved so that box drift shows.
TAG=${TAG:-abq}
for rep in $(seq 1 ${REPS:-2}); do
for v in ${VARIANTS:-base} new; do
  lib=$PWD/lua-multigrid-poisson_b200/libmgpoisson_$v.so; [ $v = new ] && lib=$PWD/lua-multigrid-poisson_b200/libmgpoisson.so
  for cfg in ${CFGS:-4096:float 2048:double}; do set -- ${cfg%%:*} ${cfg##*:}
    MGPOISSON_LIB=$lib python bench.py --dim 2 --size $1 --real $2 --steps ${STEPS:-50} --warmup 5 --no-cpu --quick > gpurun_out/${TAG}_b_${v}_$1$2.json 2> gpurun_out/${TAG}_b_${v}_$1$2.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_b_${v}_$1$2.json").read().strip().splitlines()[-1])
    print("$v", "$1", "$2", round(d["value"],1), "V-cycles/s", round(d["ms_per_step"],4), "ms", list(d["vcycle"]["breakdown_ms"].items())[:4])
except Exception as e: print("$v $1 $2 ERR", e)
PY
  done
done
done
