#!/bin/bash
# final state of the round on one B200: GPU suite, smoke, bench lines (c4 default with CPU baseline, c3 MD, c1, c5, c2), and the
# ncu --set full re-capture of the five roofline kernels (their DRAM traffic -> profiles/ncu_traffic.json)
TAG=${1:-r03b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -m gpu -x -q > gpurun_out/${TAG}_gpu_suite.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/${TAG}_gpu_suite.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err; echo "bench c4 exit $?"
timeout 300 python bench.py --workload c3 --steps 20 --warmup 3 --no-cpu-baseline --md-steps 50 --md-graph > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err; echo "bench c3 exit $?"
timeout 300 python bench.py --workload c1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c1.json 2> gpurun_out/${TAG}_bench_c1.err; echo "bench c1 exit $?"
timeout 300 python bench.py --workload c5 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; echo "bench c5 exit $?"
timeout 300 python bench.py --workload c2 --steps 20 --warmup 3 --train-graph > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err; echo "bench c2 exit $?"
python - gpurun_out/${TAG}_bench_c4.json gpurun_out/${TAG}_bench_c3.json gpurun_out/${TAG}_bench_c1.json gpurun_out/${TAG}_bench_c5.json gpurun_out/${TAG}_bench_c2.json <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get('roofline') or {}
        md = d.get('md') or {}
        print(f.split('/')[-1], 'ms/step %.3f' % d['ms_per_step'], 'value %.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], 'cpu', (d.get('cpu_baseline') or {}).get('value'),
              'roof', r.get('kernel', '')[:24], r.get('frac'), 'md ms', md.get('ms_per_step'))
    except Exception as e:
        print(f, 'parse failed', e)
PY
C4="python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline"
for K in k_filter_tc k_edge_bwd2_tc k_alpha_aggregate_scan k_qk_bwd_stream k_alpha_bwd; do
  ncu --set full --clock-control none -k regex:$K -s 3 -c 1 -f -o /tmp/${TAG}_$K $C4 > /tmp/ncu_$K.log 2>&1
  echo "ncu $K exit $?"
  ncu -i /tmp/${TAG}_$K.ncu-rep --page raw --csv > gpurun_out/${TAG}_${K}_raw.csv 2>/dev/null
done
