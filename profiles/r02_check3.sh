#!/bin/bash
# one B200: smoke, c4 bench, tensor-kernel phase cycles, c2 bench, GPU suite
TAG=${1:-r02l}
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/${TAG}_smoke.log
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
timeout 300 python profiles/tc_phase_timing.py > gpurun_out/${TAG}_phases.log 2>&1; tail -3 gpurun_out/${TAG}_phases.log
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err; echo "bench c2 exit $?"
python - gpurun_out/${TAG}_bench_c2.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('c2 ms/step %.3f' % d['ms_per_step'], 'value %.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], d['phases_ms'], d['loss_check'], 'launches', d['gpu_launches'], 'roof', d['roofline']['achieved'])
except Exception as e:
    print(sys.argv[1], 'parse failed', e)
PY
timeout 1500 python -m pytest tests/ -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --md-steps 30 --md-skin 0.5 > gpurun_out/${TAG}_bench_c4_md.json 2> gpurun_out/${TAG}_bench_c4_md.err
python - gpurun_out/${TAG}_bench_c4_md.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('md', d['md'])
except Exception as e:
    print(sys.argv[1], 'parse failed', e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches_c2.csv python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_c2.log 2>&1; echo "ncu c2 exit $?"
