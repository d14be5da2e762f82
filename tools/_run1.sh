(SLAM_GN_PAIR_GATE=0 timeout 300 python -m pytest tests -m gpu -x -q -k "two_handles") 2>&1 | tail -4
