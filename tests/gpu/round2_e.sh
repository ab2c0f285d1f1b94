#!/bin/bash
# multi-GPU: halo-based sharding, N GPUs (run with gpurun --gpus N)
set -x
N=${1:-2}
O=gpurun_out/r02e; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()"
for n in $(seq 1 $N); do :; done
if [ "$N" -ge 1 ]; then timeout 300 python bench.py --gpus 1 --steps 300 --warmup 30 > $O/bench_n1.json 2> $O/bench_n1.err; cut -c1-300 $O/bench_n1.json; fi
for n in 2 4 8; do
  if [ "$n" -le "$N" ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 300 --warmup 30 > $O/bench_n$n.json 2> $O/bench_n$n.err
    grep '^{' $O/bench_n$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['shard_check'], d['config']['parallelism'][:200])"
    tail -3 $O/bench_n$n.err
  fi
done
