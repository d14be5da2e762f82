mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q -k "sequential_split or split_launch or batch_equals") > gpurun_out/t12_tests.log 2>&1
tail -15 gpurun_out/t12_tests.log
for B in 2 3 4 8 12 16; do
  timeout 120 python tools/batch_time.py $B 40 2>&1 | tail -1
  SLAM_GN_SEQ_MAX=16 timeout 120 python tools/batch_time.py $B 40 2>&1 | tail -1 | sed 's/^/   seq_max 16: /'
  SLAM_GN_SEQ_MAX=1 timeout 120 python tools/batch_time.py $B 40 2>&1 | tail -1 | sed 's/^/   seq_max  1: /'
done
