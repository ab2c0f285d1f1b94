#!/bin/bash
# where does k_collide<1> spend its time on the convex heap?
set -x
O=gpurun_out/r02s; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_collide -s 90 -c 1 -o $O/k_collide_convex python tests/prof_scene.py convex 30 > $O/ncu1.log 2>&1
tail -3 $O/ncu1.log
