#!/bin/bash
# coupled joint axes + substep solve-groups on hardware: the full GPU suite, then a short headline bench of the same build
set -x
O=gpurun_out/r02v; mkdir -p $O
timeout 330 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log
timeout 120 python bench.py --steps 150 --warmup 20 > $O/bench.json 2> $O/bench.err
cut -c1-700 $O/bench.json
tail -2 $O/bench.err
