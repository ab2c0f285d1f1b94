#!/bin/bash
# events + kinematic bodies + CCD + Coulomb on hardware, full GPU suite, headline
set -x
O=gpurun_out/r02n; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()"
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -12 $O/pytest_gpu.log
timeout 300 python bench.py --cpu-seconds 1 > $O/bench.json 2> $O/bench.err; tail -2 $O/bench.err
python -c "
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print('HEADLINE', round(d['value']), d['stage_ms'], round(d['e2e']['value']), d['roofline']['frac'])"
