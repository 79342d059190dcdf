# A/B of library builds for the 2-D configurations: parity tests for the default build, then bench lines
# (4096^2 fp32, 2048^2 fp64, 64^2 fp64) for every variant in $VARIANTS plus the default build ("new").
TAG=${TAG:-ab2d}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests(new) rc=$?"; tail -n 3 gpurun_out/${TAG}_tests.log
for v in ${VARIANTS:-base} new; do
  lib=$PWD/lua-multigrid-poisson_b200/libmgpoisson_$v.so; [ $v = new ] && lib=$PWD/lua-multigrid-poisson_b200/libmgpoisson.so
  for cfg in ${CFGS:-4096:float 2048:double 64:double}; do set -- ${cfg%%:*} ${cfg##*:}
    MGPOISSON_LIB=$lib python bench.py --dim 2 --size $1 --real $2 --steps ${STEPS:-50} --warmup 5 --no-cpu > gpurun_out/${TAG}_b_${v}_$1$2.json 2> gpurun_out/${TAG}_b_${v}_$1$2.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_b_${v}_$1$2.json").read().strip().splitlines()[-1])
    print("$v", "$1", "$2", round(d["value"],1), "V-cycles/s", d["ms_per_step"], "ms", list(d["vcycle"]["breakdown_ms"].items())[:4])
except Exception as e: print("$v $1 $2 ERR", e)
PY
  done
done
