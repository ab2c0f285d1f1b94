#!/bin/bash
# final build of the round on hardware: full GPU suite (sensors, compound bodies, substep groups, coupled joints included), then the
# large-island config whose r02u numbers were lost with the container
set -x
O=gpurun_out/r02w; mkdir -p $O
timeout 115 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 45 python tests/perf_scenes.py pyramid3_50 > $O/perf_scenes.jsonl 2> $O/perf_scenes.err
cut -c1-600 $O/perf_scenes.jsonl
tail -2 $O/perf_scenes.err
