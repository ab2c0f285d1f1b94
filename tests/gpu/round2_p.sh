#!/bin/bash
# HEAD (CCD + bullets, events, kinematic, limits/motors, capsules, Coulomb): GPU suite, bench, launch list, ncu --set full
set -x
O=gpurun_out/r02p; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()"
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -8 $O/pytest_gpu.log
timeout 300 python bench.py > $O/bench.json 2> $O/bench.err; cut -c1-700 $O/bench.json; tail -2 $O/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 20 --warmup 3 --cpu-seconds 0.2 > $O/bench_under_ncu.log 2>&1
python tests/ncu_summary.py $O/launches_bench.csv > $O/launches_bench.txt; cat $O/launches_bench.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_solve_coop_big -s 10 -c 1 -o $O/k_solve_coop_big python tests/prof_scene.py 80x20 14 > $O/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_collide -s 12 -c 1 -o $O/k_collide_80x20 python tests/prof_scene.py 80x20 14 > $O/ncu2.log 2>&1
ls -la $O
