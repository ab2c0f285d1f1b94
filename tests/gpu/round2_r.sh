#!/bin/bash
# convex narrow phase after the Gauss-map pruning: convex tests + the convex_polyhedron3 line
set -x
O=gpurun_out/r02r; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()"
timeout 600 python -m pytest tests -m gpu -q -k "convex or variant" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 600 python tests/perf_scenes.py convex_polyhedron3 > $O/perf_scenes.jsonl 2> $O/perf_scenes.err
cat $O/perf_scenes.jsonl | cut -c1-560
tail -3 $O/perf_scenes.err
