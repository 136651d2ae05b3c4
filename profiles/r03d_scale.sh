#!/bin/bash
# c4 (domain decomposition, sweep replayed as a CUDA graph) and optionally c2 on N GPUs with the final kernels
N=${1:-2}; TAG=${2:-r03d}; WHAT=${3:-c4}
mkdir -p gpurun_out
run() {
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N $2 > gpurun_out/${TAG}_n${N}_$1.json 2> gpurun_out/${TAG}_n${N}_$1.err
  echo "bench N=$N $1 exit $?"
  python - gpurun_out/${TAG}_n${N}_$1.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms/step %.3f' % d['ms_per_step'], 'value %.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], d.get('energy_check'))
except Exception as e:
    print(sys.argv[1], 'parse failed', e)
PY
}
run c4 "--workload c4 --steps 40 --warmup 3 --dd-skin 0.2 --no-cpu-baseline"
if [ "$WHAT" = all ]; then run c2_graph "--workload c2 --steps 40 --warmup 3 --train-graph --no-cpu-baseline"; fi
