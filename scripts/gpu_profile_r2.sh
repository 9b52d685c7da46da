#!/bin/bash
# Round-2 evidence pass (one GPU): launch list + DRAM bytes of EVERY kernel of one eager training step (NVTX range
# "dp_step"), then `--set full` captures of the conv / BN / loss kernel families of the same step.
# usage: scripts/gpu_profile_r2.sh [tag]     (files land in gpurun_out/<tag>_*)
TAG=${1:-r2}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --profile-steps 0 --no-cpu-baseline --caller eager --nvtx-step"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --nvtx --nvtx-include "dp_step/" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/${TAG}_step_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?" > gpurun_out/${TAG}_rc.txt
ncu --nvtx --nvtx-include "dp_step/" --set full --clock-control none --import-source on \
    -k regex:'tc_gather_gemm|wgrad_tc_kernel|bn_act|col_reduce|loss_kernel' -o gpurun_out/${TAG}_prof_step $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/${TAG}_rc.txt
ncu -i gpurun_out/${TAG}_prof_step.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_step_raw.csv 2>/dev/null
cat gpurun_out/${TAG}_rc.txt
