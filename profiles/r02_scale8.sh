#!/bin/bash
# one 8 x B200 box: c4 domain decomposition (graph replay and eager) and the c2 data-parallel training step
N=${1:-8}
TAG=${2:-r02n}
mkdir -p gpurun_out
run() {  # name, extra flags
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --no-cpu-baseline $2 > gpurun_out/${TAG}_n${N}_$1.json 2> gpurun_out/${TAG}_n${N}_$1.err
  echo "bench N=$N $1 exit $?"; grep "dd step phases" gpurun_out/${TAG}_n${N}_$1.err | tail -1
  python - gpurun_out/${TAG}_n${N}_$1.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms/step %.3f' % d['ms_per_step'], 'value %.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], d.get('energy_check', d.get('loss_check')), d['config']['parallelism'])
except Exception as e:
    print(sys.argv[1], 'parse failed', e)
PY
}
run c4 "--workload c4 --steps 40 --warmup 3"
run c2_graph "--workload c2 --steps 40 --warmup 3 --train-graph"
