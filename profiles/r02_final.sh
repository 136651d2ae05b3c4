#!/bin/bash
# final state of the round on one B200: smoke, default bench lines (with CPU baseline), reference arm, GPU suite
TAG=${1:-r02t}
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err; echo "bench c4 exit $?"
timeout 600 python bench.py --workload c2 --steps 20 --warmup 3 --train-graph > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err; echo "bench c2 exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_c4_reference.json 2> gpurun_out/${TAG}_ref.err; echo "reference arm exit $?"
python - gpurun_out/${TAG}_bench_c4.json gpurun_out/${TAG}_bench_c2.json gpurun_out/${TAG}_bench_c4_reference.json <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get('roofline') or {}
        print(f, 'ms/step %.3f' % d['ms_per_step'], 'value %.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], 'cpu', (d.get('cpu_baseline') or {}).get('value'),
              'roof', r.get('kernel', '')[:30], r.get('frac'), 'traffic', r.get('traffic'), 'launches', d.get('gpu_launches'), d.get('clocks'))
    except Exception as e:
        print(f, 'parse failed', e)
PY
timeout 1200 python -m pytest tests/ -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest.log
