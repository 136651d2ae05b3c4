#!/bin/bash
TAG=${1:-r02u}
mkdir -p gpurun_out
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
python - gpurun_out/${TAG}_bench_c4.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms/step %.2f' % d['ms_per_step'], 'E', d['energy_check'])
    print('   ', {n: round(v, 2) for n, v in d['kernel_ms_per_step'].items()})
except Exception as e:
    print(sys.argv[1], 'parse failed', e)
PY
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_optional_branches_gpu.py tests/test_zzz_reference_model_gpu.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest.log
