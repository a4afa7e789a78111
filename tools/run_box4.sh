#!/bin/bash
# One 4-GPU box call (charged 4 x): the tp2/tp4 tests, then four single-GPU jobs in parallel.
#   gpurun --gpus 4 --timeout 420 -- bash tools/run_box4.sh
OUT=gpurun_out/r2_box4
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/env.log 2>&1
# ---- phase 1: all GPUs
timeout 150 python -m pytest tests/test_gpu_tp.py -m gpu -q -s -k "4" > $OUT/tp_tests.log 2>&1
echo "tp tests rc=$?" >> $OUT/tp_tests.log
tail -4 $OUT/tp_tests.log
# ---- phase 2: one GPU each
(
  export CUDA_VISIBLE_DEVICES=0
  timeout 170 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_tp.py > $OUT/pytest_default.log 2>&1
  echo "rc=$?" >> $OUT/pytest_default.log
  OB_BENCH_SKIP_PREFILL=1 timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
     --log-file $OUT/launches_decode_step.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
) &
(
  export CUDA_VISIBLE_DEVICES=1
  OB_GEMM_DEC_CLUSTER=1 timeout 170 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_tp.py > $OUT/pytest_cluster.log 2>&1
  echo "rc=$?" >> $OUT/pytest_cluster.log
  OB_GEMM_DEC_CLUSTER=0 timeout 150 python bench.py --steps 128 --warmup 8 --no-cpu-baseline > $OUT/bench_l2.json 2> $OUT/bench_l2.err
) &
(
  export CUDA_VISIBLE_DEVICES=2
  OB_CL_EXP=1 OB_GEMM_DEC_CLUSTER=0 timeout 100 python tools/gemm_micro.py > $OUT/cluster_ab.log 2>&1
  OB_CL_EXP=1 OB_GEMM_DEC_CLUSTER=1 timeout 100 python tools/gemm_micro.py >> $OUT/cluster_ab.log 2>&1
) &
(
  export CUDA_VISIBLE_DEVICES=3
  OB_GEMM_DEC_CLUSTER=1 timeout 150 python bench.py --steps 128 --warmup 8 --no-cpu-baseline > $OUT/bench_cluster.json 2> $OUT/bench_cluster.err
) &
wait
tail -2 $OUT/pytest_default.log; tail -2 $OUT/pytest_cluster.log; cat $OUT/cluster_ab.log | grep -v "^$" | tail -40
for f in bench_l2 bench_cluster; do python - <<PY
import json
try:
    d=json.loads(open("$OUT/$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["roofline"]["frac"], {k:v["ms_per_layer"] for k,v in d["kernels"].items()})
except Exception as e:
    print("$f", "unreadable", e)
PY
done
