mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/t16_gputests.log 2>&1
tail -5 gpurun_out/t16_gputests.log | head -2
for i in 1 2; do
(timeout 600 python bench.py --steps 1000 --warmup 20 --no-baselines --no-batched) > gpurun_out/t16_bench.json 2> gpurun_out/t16_bench.err
python - gpurun_out/t16_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["us_per_launch"], d["roofline"]["frac"], d["gpu_launches"])
PY
done
