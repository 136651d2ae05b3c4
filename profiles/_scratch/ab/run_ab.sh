#!/bin/bash
# A/B of libso3k builds that differ only in so3k_agg.cu compile-time switches (SO3K_W_POLICY, SO3K_AGG_MINB)
mkdir -p gpurun_out
for v in "$@"; do
  SO3K_LIB_PATH=$PWD/profiles/_scratch/ab/libso3k_$v.so timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  echo "variant $v exit $?"
  python - gpurun_out/ab_$v.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d['kernel_ms_per_step']
    print(sys.argv[1], 'ms/step', round(d['ms_per_step'],3), 'agg', round(k['aggregate'],3), 'qk', round(k['qk_stream'],3), 'abwd', round(k['alpha_bwd'],3), 'bwd2', round(k['edge_bwd'],3), 'E', d['energy_check'])
except Exception as e:
    print(sys.argv[1], 'parse failed', e)
PY
done
