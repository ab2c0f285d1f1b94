#!/bin/bash
# GPU call C: profiles of the large-island path after the split
set -x
O=gpurun_out/r02c; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()"
timeout 900 python tests/perf_scenes.py > $O/perf_scenes.jsonl 2> $O/perf_scenes.err
cat $O/perf_scenes.jsonl | cut -c1-420
for sc in keva5 p3_50 jg100; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 30 --csv --log-file $O/launches_$sc.csv python tests/prof_scene.py $sc 40 > /dev/null 2>&1
python tests/ncu_summary.py $O/launches_$sc.csv
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_solve_large -s 30 -c 1 -o $O/k_solve_large_keva5 python tests/prof_scene.py keva5 35 > $O/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_solve_large -s 30 -c 1 -o $O/k_solve_large_p3_50 python tests/prof_scene.py p3_50 35 > $O/ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_collide -s 30 -c 1 -o $O/k_collide_p3_50 python tests/prof_scene.py p3_50 35 > $O/ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_solve_large -s 30 -c 1 -o $O/k_solve_large_jg100 python tests/prof_scene.py jg100 35 > $O/ncu4.log 2>&1
ls -la $O
