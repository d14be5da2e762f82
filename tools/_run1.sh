for pd in 500 700 900 1100 1300; do
SLAM_GN_POLL_DELAY=$pd timeout 300 python bench.py --steps 400 --warmup 20 --no-baselines --no-batched 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('poll delay $pd:', round(d['value'],1), round(d['roofline']['us_per_launch'],2), round(d['e2e']['value'],1))"
done
