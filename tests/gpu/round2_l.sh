#!/bin/bash
# CCD motion clamping on hardware + its effect on the headline (in-kernel call of the clamp)
set -x
O=gpurun_out/r02l; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()"
timeout 300 python bench.py --cpu-seconds 1 > $O/bench.json 2> $O/bench.err; cut -c1-400 $O/bench.json; tail -2 $O/bench.err
python -c "
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print('HEADLINE', round(d['value']), d['stage_ms'], round(d['e2e']['value']), d['roofline']['frac'])"
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "ccd or coulomb or variants" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -8 $O/pytest.log
