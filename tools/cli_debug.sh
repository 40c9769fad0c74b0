#!/bin/bash
# debug helper: run both BPtrain binaries on the bundled pfile with timeouts, keep logs
set -x
G=$PWD/tests/golden
T=$(mktemp -d)
python - <<PY
import sys; sys.path.insert(0,'.')
from oracle import oracle as O
ls=[1799,2048,2048,2048,257]
W,b=O.init_weights(ls,seed=4)
O.write_wts("$T/init.wts",ls,W,b)
PY
FLAGS="gpu_used=0 numlayers=5 layersizes=1799,2048,2048,2048,257 bunchsize=128 MLflag=1 shapefactor=1.5 momentum=0.9 weightcost=0.00001 lrate=0.1 fea_dim=257 fea_context=7 traincache=102400 init_randem_seed=27870775 targ_offset=3 initwts_file=$T/init.wts norm_file=$G/train_noisy.norm fea_file=$G/train_noisy.pfile targ_file=$G/train_clean.pfile train_sent_range=0-7 cv_sent_range=8-9 dropoutflag=0 visible_omit=0.1 hid_omit=0.1"
P=speech-enhancement-based-on-a-maximum-likelihood-criterion_b200
( time timeout 120 $P/host/BPtrain_Sigmoid $FLAGS outwts_file=$T/mine.wts log_file=$T/mine.log ) > gpurun_out/cli_mine.out 2>&1
echo "mine rc=$?" >> gpurun_out/cli_mine.out; cat $T/mine.log >> gpurun_out/cli_mine.out
( time timeout 300 oracle/_ref/BPtrain_ref $FLAGS outwts_file=$T/ref.wts log_file=$T/ref.log ) > gpurun_out/cli_ref.out 2>&1
echo "ref rc=$?" >> gpurun_out/cli_ref.out; cat $T/ref.log >> gpurun_out/cli_ref.out
tail -15 gpurun_out/cli_mine.out; tail -15 gpurun_out/cli_ref.out
