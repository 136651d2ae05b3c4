#!/bin/bash
TAG=${1:-r02d}
mkdir -p gpurun_out
summ() {
python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms/step %.2f' % d['ms_per_step'], 'E', d['energy_check'])
    k = d['kernel_ms_per_step']
    print('   ', {n: round(k[n], 2) for n in ('edge_fwd', 'aggregate', 'alpha_bwd', 'qk_stream', 'edge_bwd', 'pair_records', 'graph')})
except Exception as e:
    print(sys.argv[1], 'parse failed', e)
PY
}
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/${TAG}_smoke.log
B="python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline"
timeout 300 $B > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; summ gpurun_out/${TAG}_bench_default.json
SO3K_LIB_PATH=$PWD/profiles/_scratch/libso3k_oldagg.so timeout 300 $B > gpurun_out/${TAG}_bench_oldagg.json 2> gpurun_out/${TAG}_bench_oldagg.err; summ gpurun_out/${TAG}_bench_oldagg.json
timeout 300 python profiles/tc_phase_timing.py > gpurun_out/${TAG}_phases.log 2>&1; tail -3 gpurun_out/${TAG}_phases.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_mode or minimum_image or domain_decomposition" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest.log
