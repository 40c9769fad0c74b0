#!/bin/bash
# N-GPU sweep of the data-parallel schedule knobs (GGD_DP_OVERLAP, GGD_NCCL_MAX_CTAS)
N=${1:-2}
for cfg in "0 16" "0 32" "1 8" "1 16" "1 32"; do
  set -- $cfg
  GGD_DP_OVERLAP=$1 GGD_NCCL_MAX_CTAS=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 400 --warmup 16 --no-lps --no-e2e > gpurun_out/dp_$1_$2.json 2> gpurun_out/dp_$1_$2.err
  python -c "
import json;d=json.loads(open('gpurun_out/dp_$1_$2.json').read().strip().splitlines()[-1]);print('overlap=$1 maxctas=$2', round(d['value']), round(d['ms_per_step']*1e3,1), {k:round(v['ms_per_step']*1e3,1) for k,v in d['kernels'].items()})"
done
