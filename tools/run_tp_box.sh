#!/bin/bash
# Everything the multi-GPU box is used for, in ONE gpurun call (N GPUs are charged N x):
#   gpurun --gpus 8 --timeout 900 -- bash tools/run_tp_box.sh 8
# -> gpurun_out/r2_tp/{tp_tests.log, bench_tpN.json, bench_tpN.err, tp_step.md}   (copied to profiles/ by hand)
N=${1:-8}
OUT=gpurun_out/r2_tp
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/env.log 2>&1
nvidia-smi topo -m >> $OUT/env.log 2>&1
KSEL="4 or 8"
[ "$N" -lt 8 ] && KSEL="2 or 4"
timeout 420 python -m pytest tests/test_gpu_tp.py -m gpu -q -s -k "$KSEL" > $OUT/tp_tests.log 2>&1
echo "tp tests rc=$?" >> $OUT/tp_tests.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_tp$N.json 2> $OUT/bench_tp$N.err
echo "bench rc=$?" >> $OUT/bench_tp$N.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  tools/tp_profile.py > $OUT/tp${N}_step.md 2> $OUT/tp${N}_step.err
echo "profile rc=$?" >> $OUT/tp${N}_step.err
tail -3 $OUT/tp_tests.log; cat $OUT/bench_tp$N.json | cut -c1-1500; tail -30 $OUT/tp${N}_step.md
