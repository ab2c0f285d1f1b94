#!/bin/bash
# headline experiments: PDL on/off, 1 vs 2 lanes per constraint in the big launch shape
set -x
O=gpurun_out/r02h; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()"
for cfg in "0 1" "1 1" "1 2" "0 2"; do
  set -- $cfg
  RB_PDL=$1 RB_BIG_LANES=$2 timeout 300 python bench.py --steps 300 --warmup 30 --cpu-seconds 1 > $O/bench_pdl$1_lanes$2.json 2> $O/err.txt
  python -c "
import json,sys
d=json.loads(open('$O/bench_pdl$1_lanes$2.json').read().strip().splitlines()[-1])
print('PDL=$1 LANES=$2', round(d['value']), round(d['ms_per_step'],4), d['stage_ms'], round(d['e2e']['value']))"
done
RB_PDL=1 RB_BIG_LANES=2 timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "headline or streamed or many_pyramids_full_size_matches" 2>&1 | tail -3
