#!/bin/bash
# ncu evidence of round 2 on one B200 (recipe of /opt/skills/guides/B200_PROFILING.md: every ncu run follows a plain run of the
# same command that exited 0; --clock-control none).  The .ncu-rep files stay on the box (/tmp); their `--page raw --csv`
# exports and the launch lists come back in gpurun_out/.
TAG=${1:-r02r}
mkdir -p gpurun_out
C4="python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline"
C2="python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline"
$C4 > gpurun_out/${TAG}_plain_c4.log 2>&1 || { echo "plain c4 failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches_c4.csv $C4 > /tmp/ncu_l.log 2>&1; echo "launch list c4 exit $?"
for K in k_filter_tc k_edge_bwd2_tc k_alpha_aggregate_scan k_qk_bwd_stream k_alpha_bwd k_featurise k_pairs; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o /tmp/${TAG}_$K $C4 > /tmp/ncu_$K.log 2>&1
  echo "ncu $K exit $?"
  ncu -i /tmp/${TAG}_$K.ncu-rep --page raw --csv > gpurun_out/${TAG}_${K}_raw.csv 2>/dev/null
done
$C2 > gpurun_out/${TAG}_plain_c2.log 2>&1 || { echo "plain c2 failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches_c2.csv $C2 > /tmp/ncu_l2.log 2>&1; echo "launch list c2 exit $?"
for K in k_tr_gemm_tn "k_tr_gemm\b"; do
  N=$(echo $K | tr -cd 'a-z_')
  ncu --set full --clock-control none --import-source on -k "regex:$K" -s 40 -c 1 -f -o /tmp/${TAG}_$N $C2 > /tmp/ncu_$N.log 2>&1
  echo "ncu $N exit $?"
  ncu -i /tmp/${TAG}_$N.ncu-rep --page raw --csv > gpurun_out/${TAG}_${N}_raw.csv 2>/dev/null
done
ls -la gpurun_out/ | grep ${TAG}
