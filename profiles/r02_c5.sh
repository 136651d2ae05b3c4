#!/bin/bash
TAG=${1:-r02s}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_optional_branches_gpu.py -m gpu -x -q -s -k "c5" > gpurun_out/${TAG}_pytest_c5.log 2>&1; echo "pytest c5 exit $?"; grep "c5 (" gpurun_out/${TAG}_pytest_c5.log; tail -3 gpurun_out/${TAG}_pytest_c5.log
for MODE in tensor fp32; do
FLAG=""; [ $MODE = fp32 ] && FLAG="--fp32-simt"
timeout 300 python bench.py --workload c5 --steps 20 --warmup 3 --no-cpu-baseline $FLAG > gpurun_out/${TAG}_bench_c5_${MODE}.json 2> gpurun_out/${TAG}_bench_c5_${MODE}.err; echo "bench c5 $MODE exit $?"; tail -2 gpurun_out/${TAG}_bench_c5_${MODE}.err
python - gpurun_out/${TAG}_bench_c5_${MODE}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('c5 ms/step %.3f' % d['ms_per_step'], 'value %.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], d['config']['filter_gemm'], 'E', d['energy_check'])
    print('   ', {n: round(v, 3) for n, v in d['kernel_ms_per_step'].items()})
except Exception as e:
    print(sys.argv[1], 'parse failed', e)
PY
done
timeout 900 python -m pytest tests/ -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
python - gpurun_out/${TAG}_bench_c4.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('c4 ms/step %.2f' % d['ms_per_step'], 'E', d['energy_check'])
PY
