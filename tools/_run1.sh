mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/t20_gputests.log 2>&1
tail -5 gpurun_out/t20_gputests.log | head -2
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()") 2>&1 | tail -2
(timeout 600 python bench.py) > gpurun_out/t20_bench.json 2> gpurun_out/t20_bench.err
python - gpurun_out/t20_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_all_host"]["value"], d["roofline"]["us_per_launch"], d["roofline"]["frac"], d["roofline"]["share_of_step"], d["gpu_launches"], d["batched"]["value"], d["batched"]["roofline"]["frac"], d["cpu_baseline"]["value"], d["ref_cuda"]["value"])
PY
(timeout 300 python bench.py --impl reference --steps 3 --warmup 1) 2>/dev/null | tail -1 | cut -c1-400
