(SLAM_ODOM_DEBUG_TIMING=1 timeout 300 python bench.py --steps 1100 --warmup 20 --no-baselines --no-batched) 2>&1 >/dev/null | grep -i "frame host\|timing\|prep\b\|us" | head -12
