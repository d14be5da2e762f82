#!/bin/bash
# Development aid: the standard GPU round trip (parity tests, kernel timeline, quick timing) -> gpurun_out/<tag>_*.log
tag=${1:-chk}; shift
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q "$@") > gpurun_out/${tag}_gputests.log 2>&1
timeout 120 python tools/timeline.py full > gpurun_out/${tag}_timeline.log 2>&1
QT_DEVICE_ONLY=1 timeout 120 python tools/quick_time.py > gpurun_out/${tag}_quick.log 2>&1
tail -4 gpurun_out/${tag}_gputests.log; tail -3 gpurun_out/${tag}_quick.log
