#!/bin/bash
# training sweep with its large dense layers on the tensor cores: parity tests, c2 bench (graph), launch list
TAG=${1:-r02w}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_training_gpu.py -m gpu -x -q -s 2>&1 | grep -E "passed|failed|rror|flags|stress|branches" | tail -12
for MODE in tensor fp32; do
FLAG=""; [ $MODE = fp32 ] && FLAG="--fp32-simt"
timeout 300 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline --train-graph $FLAG > gpurun_out/${TAG}_bench_c2_${MODE}.json 2> gpurun_out/${TAG}_bench_c2_${MODE}.err; echo "bench c2 $MODE exit $?"; tail -2 gpurun_out/${TAG}_bench_c2_${MODE}.err
python - gpurun_out/${TAG}_bench_c2_${MODE}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('c2 ms/step %.3f' % d['ms_per_step'], 'value %.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], d['loss_check'], d['phases_ms'])
except Exception as e:
    print(sys.argv[1], 'parse failed', e)
PY
done
