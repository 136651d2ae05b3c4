#!/bin/bash
TAG=${1:-r02p}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_training_gpu.py tests/test_gpu_parity.py -m gpu -x -q -k "training or graph or dd_session" > gpurun_out/${TAG}_pytest_train.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/${TAG}_pytest_train.log
for MODE in graph eager; do
FLAG=""; [ $MODE = graph ] && FLAG="--train-graph"
timeout 300 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline $FLAG > gpurun_out/${TAG}_bench_c2_${MODE}.json 2> gpurun_out/${TAG}_bench_c2_${MODE}.err; echo "bench c2 $MODE exit $?"; tail -2 gpurun_out/${TAG}_bench_c2_${MODE}.err
python - gpurun_out/${TAG}_bench_c2_${MODE}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('c2 ms/step %.3f' % d['ms_per_step'], 'value %.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], d['loss_check'], 'launches', d['gpu_launches'])
except Exception as e:
    print(sys.argv[1], 'parse failed', e)
PY
done
