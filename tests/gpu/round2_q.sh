#!/bin/bash
# dominance, warmstart_joints, convex polyhedra on hardware: GPU suite, secondary configs (+ convex_polyhedron3), bench
set -x
O=gpurun_out/r02q; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()"
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
timeout 300 python bench.py > $O/bench.json 2> $O/bench.err; cut -c1-400 $O/bench.json; tail -2 $O/bench.err
timeout 900 python tests/perf_scenes.py > $O/perf_scenes.jsonl 2> $O/perf_scenes.err
cat $O/perf_scenes.jsonl | cut -c1-560
tail -3 $O/perf_scenes.err
