mkdir -p gpurun_out/prof
# 1) launch list of the bench command itself (cold, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/prof/launches.csv python bench.py --steps 20 --warmup 3 --no-baselines --no-batched > gpurun_out/prof/launches_bench.log 2>&1
# 2) full captures: one frame's three kernels (prepare, cluster, fine levels) after warm-up
ncu --set full --import-source on --clock-control none -k regex:"k_prepare_frame|k_gn_persistent" -s 30 -c 3 -o gpurun_out/prof/r02_split python bench.py --steps 10 --warmup 3 --no-baselines --no-batched > gpurun_out/prof/full_bench.log 2>&1
ls -la gpurun_out/prof
