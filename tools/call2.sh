TAG=${TAG:-c2} tools/gpu_s3d.sh
