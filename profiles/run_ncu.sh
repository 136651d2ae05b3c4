#!/bin/bash
# ncu recipe of /opt/skills/guides/B200_PROFILING.md, run on the GPU box through gpurun:
#   gpurun --timeout 1500 -- 'bash profiles/run_ncu.sh r01'
# Writes gpurun_out/<tag>_launches.csv (every launch with its device time) and gpurun_out/<tag>_<kernel>.ncu-rep
# (--set full captures of the hot kernels).  Each ncu run follows a plain run of the same command that exited 0.
TAG=${1:-r01}
EXTRA=${2:-}
set -x
CMD="python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline $EXTRA"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
for K in ${KERNELS:-k_edge_bwd k_edge_fwd k_aggregate k_alpha_bwd}; do
  $CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/${TAG}_$K $CMD > gpurun_out/${TAG}_ncu_$K.log 2>&1
done
ls -la gpurun_out/
