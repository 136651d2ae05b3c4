#!/bin/bash
# Round-2 device check, one gpurun call: layout self-tests + tensor-mode parity, per-phase cycles of the tensor kernels
# (timing build), a short c4 bench.  Every step under its own timeout so that a hung kernel cannot eat the box.
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_smi.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log
tail -4 gpurun_out/${TAG}_smoke.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tcgen05 or (tensor_mode and not 132)" > gpurun_out/${TAG}_pytest_tc.log 2>&1
echo "pytest tc exit $?" >> gpurun_out/${TAG}_pytest_tc.log
tail -5 gpurun_out/${TAG}_pytest_tc.log
timeout 300 python profiles/tc_phase_timing.py > gpurun_out/${TAG}_phases.log 2>&1
echo "phases exit $?" >> gpurun_out/${TAG}_phases.log
tail -4 gpurun_out/${TAG}_phases.log
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err
echo "bench exit $?"
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/${TAG}_bench_c4.json').read().strip().splitlines()[-1])
    print('ms/step', d['ms_per_step'], 'value', d['value'], 'energy', d['energy_check'])
    print(d['kernel_ms_per_step'])
    for r in [d['roofline']] + d['roofline_others']:
        print(r['kernel'], round(r['frac'], 3), round(r['ms_per_layer_scope'], 3))
except Exception as e:
    print('bench parse failed', e)
PY
