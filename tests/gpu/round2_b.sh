#!/bin/bash
# GPU call B: radix-sort broad phase with static/dynamic split -- parity + perf of all configs
set -x
O=gpurun_out/r02b; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
timeout 900 python tests/perf_scenes.py > $O/perf_scenes.jsonl 2> $O/perf_scenes.err
cat $O/perf_scenes.jsonl | cut -c1-400
tail -3 $O/perf_scenes.err
