#!/bin/bash
# training seam on one B200: GPU parity tests, c2 bench line, kernel launch list of one training step
TAG=${1:-r02k}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_training_gpu.py -m gpu -x -q -s > gpurun_out/${TAG}_pytest_train.log 2>&1; echo "pytest train exit $?"; tail -6 gpurun_out/${TAG}_pytest_train.log
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err; echo "bench c2 exit $?"
tail -3 gpurun_out/${TAG}_bench_c2.err
python - gpurun_out/${TAG}_bench_c2.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('ms/step %.3f' % d['ms_per_step'], 'value %.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], d['phases_ms'], d['loss_check'], 'launches', d['gpu_launches'], 'roof', d['roofline']['achieved'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'])
except Exception as e:
    print(sys.argv[1], 'parse failed', e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches_c2.csv python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_c2.log 2>&1; echo "ncu exit $?"
