#!/bin/bash
# warp-cooperative polyhedron manifolds + parallel colour stage order: tests, perf line, section timeline
set -x
O=gpurun_out/r02t; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 600 python tests/perf_scenes.py convex_polyhedron3 keva3_5 pyramid3_50 b3d_joint_grid_100 > $O/perf_scenes.jsonl 2> $O/perf_scenes.err
cat $O/perf_scenes.jsonl | cut -c1-560
tail -3 $O/perf_scenes.err
for s in convex keva5; do timeout 200 python tests/prof_collide_phases.py $s 200 2>&1 | tail -2; done | tee $O/collide_phases.txt
timeout 300 python bench.py --cpu-seconds 1 > $O/bench.json 2> $O/bench.err; cut -c1-260 $O/bench.json
