#!/bin/bash
# domain decomposition check on N GPUs of one box: eager sweep vs the sweep replayed as one CUDA graph (plan kept within a skin)
N=${1:-2}
TAG=${2:-r02m}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dd_session" > gpurun_out/${TAG}_pytest_dd.log 2>&1; echo "pytest dd_session exit $?"; tail -2 gpurun_out/${TAG}_pytest_dd.log
for MODE in graph eager; do
FLAG=""; [ $MODE = graph ] && FLAG="--dd-graph"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --dd-skin 0.2 $FLAG > gpurun_out/${TAG}_bench_n${N}_${MODE}.json 2> gpurun_out/${TAG}_bench_n${N}_${MODE}.err
echo "bench N=$N $MODE exit $?"; tail -3 gpurun_out/${TAG}_bench_n${N}_${MODE}.err
python - gpurun_out/${TAG}_bench_n${N}_${MODE}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms/step %.2f' % d['ms_per_step'], 'value %.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], 'E', d['energy_check'], d['config']['parallelism'])
except Exception as e:
    print(sys.argv[1], 'parse failed', e)
PY
done
