mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/t8_gputests.log 2>&1
tail -4 gpurun_out/t8_gputests.log
(timeout 600 python bench.py) > gpurun_out/t8_bench.json 2> gpurun_out/t8_bench.err
tail -c 600 gpurun_out/t8_bench.err
python - gpurun_out/t8_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["us_per_launch"], d["roofline"]["frac"], d["gpu_launches"], d["batched"]["value"], d["cpu_baseline"])
PY
