mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/t15_gputests.log 2>&1
tail -5 gpurun_out/t15_gputests.log | head -2
(timeout 600 python bench.py) > gpurun_out/t15_bench.json 2> gpurun_out/t15_bench.err
python - gpurun_out/t15_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_all_host"]["value"], d["roofline"]["us_per_launch"], d["roofline"]["frac"], d["roofline"]["share_of_step"], d["gpu_launches"], d["batched"]["value"], d["batched"]["roofline"]["frac"], d["cpu_baseline"]["value"], d.get("ref_cuda"))
for k,v in d["other_configs"].items(): print(k, {a:b for a,b in v.items() if isinstance(b,(int,float))})
PY
