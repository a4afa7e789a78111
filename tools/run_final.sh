#!/bin/bash
# Final single-GPU validation of the shipped defaults:  gpurun --timeout 600 -- bash tools/run_final.sh
OUT=gpurun_out/r2_final
mkdir -p $OUT
timeout 200 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
OB_GEMM_DEC_CLUSTER=0 timeout 120 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity_r2.py tests/test_gpu_model.py tests/test_gpu_moe.py -m gpu -q \
   > $OUT/pytest_l2_splitk.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_l2_splitk.log
tail -2 $OUT/pytest_l2_splitk.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 150 python bench.py --impl reference --steps 16 --warmup 3 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?"
timeout 280 python bench.py > $OUT/bench_ours.json 2> $OUT/bench_ours.err; echo "bench rc=$?"
OB_BENCH_SKIP_PREFILL=1 timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
   --log-file $OUT/launches_decode_step.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
cut -c1-1200 $OUT/bench_ours.json; echo; cut -c1-600 $OUT/bench_reference.json
