mkdir -p gpurun_out
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 1000 --warmup 20) > gpurun_out/t14_bench_8gpu.json 2> gpurun_out/t14_bench_8gpu.err
tail -c 300 gpurun_out/t14_bench_8gpu.err
python - gpurun_out/t14_bench_8gpu.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["us_per_launch"], d["batched"]["value"], d["batched"]["sequences_per_gpu"], d["batched"]["gpu_launches_per_step"], {k:(v.get("ms_per_frame"),v.get("winner_equals_single_gpu")) for k,v in d["other_configs"].items() if "configs[4]" in k})
PY
