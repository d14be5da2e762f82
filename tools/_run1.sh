(SLAM_GN_PHASES=3 QT_DEVICE_ONLY=1 timeout 200 python tools/quick_time.py) 2>&1 | head -2 | cut -c1-700
(timeout 600 python -m pytest tests -m gpu -x -q -k "split or device_loop or closed_loop_sequence or sequential or frame_call or sensor or prepared") 2>&1 | tail -2
(timeout 600 python bench.py --steps 1000 --warmup 20 --no-baselines --no-batched) 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['roofline']['us_per_launch'],2), round(d['e2e']['value'],1))"
