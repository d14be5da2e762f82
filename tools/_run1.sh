mkdir -p gpurun_out
for chain in 1 0 1 0; do
(SLAM_GN_PDL_CHAIN=$chain timeout 600 python bench.py --steps 500 --warmup 20 --no-baselines --no-batched) > gpurun_out/t11_bench_$chain.json 2> gpurun_out/t11_bench_$chain.err
python - gpurun_out/t11_bench_$chain.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["us_per_launch"], d["check"])
PY
done
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/t11_gputests.log 2>&1
tail -5 gpurun_out/t11_gputests.log | head -2
