#!/bin/bash
# Coulomb friction model on hardware + first full ncu captures of k_solve_large (the grid-wide island kernel)
set -x
O=gpurun_out/r02j; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "coulomb" > $O/pytest_coulomb.log 2>&1; echo "pytest rc=$?" >> $O/pytest_coulomb.log
tail -8 $O/pytest_coulomb.log
for sc in keva5 p3_50 jg100; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_solve_large -s 4 -c 1 -o $O/k_solve_large_$sc python tests/prof_scene.py $sc 4 > $O/ncu_$sc.log 2>&1
  tail -2 $O/ncu_$sc.log
done
ls -la $O
