mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpus or sharded or nccl") > gpurun_out/t13_tests2.log 2>&1
tail -3 gpurun_out/t13_tests2.log
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1000 --warmup 20) > gpurun_out/t13_bench_2gpu.json 2> gpurun_out/t13_bench_2gpu.err
tail -c 400 gpurun_out/t13_bench_2gpu.err
python - gpurun_out/t13_bench_2gpu.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["us_per_launch"], d["batched"]["value"], d["batched"]["sequences_per_gpu"], {k:(v.get("ms_per_frame"),v.get("winner_equals_single_gpu")) for k,v in d["other_configs"].items() if "configs[4]" in k})
PY
