#!/bin/bash
# colouring over a pending list + grid-wide colour sort of item 0: full GPU suite, the dynamic configs, section timelines
set -x
O=gpurun_out/r02u; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 600 python tests/perf_scenes.py convex_polyhedron3 keva3_5 pyramid3_50 b3d_joint_grid_100 > $O/perf_scenes.jsonl 2> $O/perf_scenes.err
cat $O/perf_scenes.jsonl | cut -c1-560
tail -3 $O/perf_scenes.err
for s in "p3_50 40" "convex 200"; do timeout 250 python tests/prof_collide_phases.py $s 2>&1 | tail -2; done | tee $O/collide_phases.txt
