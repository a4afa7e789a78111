OUT=gpurun_out/r2_final2
mkdir -p $OUT
timeout 200 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
timeout 200 python bench.py > $OUT/bench_ours.json 2> $OUT/bench_ours.err; echo "bench rc=$?"
cut -c1-400 $OUT/bench_ours.json
python - <<PY
import json
d=json.loads(open("$OUT/bench_ours.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], {k:v["ms_per_layer"] for k,v in d["kernels"].items()})
PY
