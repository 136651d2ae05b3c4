#!/bin/bash
# domain decomposition check on N GPUs of one box: skin re-use vs rebuild every step
N=${1:-2}
TAG=${2:-r02j}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dd_session" > gpurun_out/${TAG}_pytest_dd.log 2>&1; echo "pytest dd_session exit $?"; tail -2 gpurun_out/${TAG}_pytest_dd.log
for SKIN in 0.2 0.0; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --dd-timing --dd-skin $SKIN > gpurun_out/${TAG}_bench_n${N}_skin${SKIN}.json 2> gpurun_out/${TAG}_bench_n${N}_skin${SKIN}.err
echo "bench N=$N skin=$SKIN exit $?"
grep "dd step phases" gpurun_out/${TAG}_bench_n${N}_skin${SKIN}.err | tail -1
python - gpurun_out/${TAG}_bench_n${N}_skin${SKIN}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms/step %.2f' % d['ms_per_step'], 'value %.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], 'E', d['energy_check'], d['config']['parallelism'])
except Exception as e:
    print(sys.argv[1], 'parse failed', e)
PY
done
