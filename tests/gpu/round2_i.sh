#!/bin/bash
# validation of HEAD after the z-strip broad phase: GPU suite, secondary configs, bench
set -x
O=gpurun_out/r02i; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()"
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
timeout 900 python tests/perf_scenes.py > $O/perf_scenes.jsonl 2> $O/perf_scenes.err
cat $O/perf_scenes.jsonl | cut -c1-520
tail -3 $O/perf_scenes.err
timeout 300 python bench.py > $O/bench.json 2> $O/bench.err; cut -c1-600 $O/bench.json; tail -2 $O/bench.err
