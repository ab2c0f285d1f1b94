#!/bin/bash
# GPU call A of round 2: everything that had never run on hardware + fresh profiles of the shipped binary.
set -x
mkdir -p gpurun_out/r02a
O=gpurun_out/r02a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
which cargo rustc > $O/rust_toolchain.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 600 python tests/perf_scenes.py > $O/perf_scenes.jsonl 2> $O/perf_scenes.err
timeout 300 python tests/protocol_1000.py b3d_many_pyramids_80x20 16 > $O/protocol_80x20.json 2>&1
timeout 300 python tests/protocol_1000.py falling_pile_2000 8 > $O/protocol_falling_pile.json 2>&1
timeout 300 python bench.py > $O/bench.json 2> $O/bench.err
timeout 300 python bench.py --impl reference --steps 60 --warmup 3 > $O/bench_ref.json 2>&1
# profiles of the shipped binary
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file $O/launches_80x20.csv python tests/prof_scene.py 80x20 40 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_solve_coop_big -s 10 -c 2 -o $O/k_solve_coop_big python tests/prof_scene.py 80x20 10 > $O/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_collide -s 12 -c 1 -o $O/k_collide_keva5 python tests/prof_scene.py keva5 6 > $O/ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_collide -s 12 -c 1 -o $O/k_collide_jg100 python tests/prof_scene.py jg100 6 > $O/ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_collide -s 12 -c 1 -o $O/k_collide_80x20 python tests/prof_scene.py 80x20 6 > $O/ncu4.log 2>&1
ls -la $O
