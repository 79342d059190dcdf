nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
TAG=c1 VARIANTS="base vx2" RUNS="0:3 0:0 0:2 -1:3 -1:0" NCU="base:0:3 base:0:0 base:-1:3 vx2:0:3 vx2:-1:3" tools/gpu_s3d.sh
