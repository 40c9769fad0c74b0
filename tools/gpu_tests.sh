#!/bin/bash
# Runs the GPU test files one by one under a timeout (a hung tcgen05 kernel must not eat the whole call).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
make -s -C oracle > gpurun_out/oracle_build.log 2>&1
rc_all=0
for t in "$@"; do
  name=$(basename "$t" .py)
  timeout 600 python -m pytest "$t" -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/$name.log 2>&1
  rc=$?
  echo "== $t rc=$rc"; tail -n 25 gpurun_out/$name.log
  [ $rc -ne 0 ] && rc_all=$rc
done
exit $rc_all
