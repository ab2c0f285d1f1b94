#!/bin/bash
# A/B of k_collide / solve times: baseline (r02i sources), + Coulomb, + CCD (current)
O=gpurun_out/r02m; mkdir -p $O
for rep in ${REPS:-1 2}; do
for lib in ${LIBS:-librapier_b200_base.so librapier_b200_coul.so librapier_b200.so}; do
  RAPIER_B200_DEBUG_LIB=$lib timeout 300 python bench.py --steps 300 --warmup 30 --cpu-seconds 1 > $O/bench_$lib.$rep.json 2> $O/err.txt
  python -c "
import json
d=json.loads(open('$O/bench_$lib.$rep.json').read().strip().splitlines()[-1])
print('$lib', $rep, round(d['value']), {k: round(v,4) for k,v in d['stage_ms'].items()}, round(d['e2e']['value']))"
done; done
