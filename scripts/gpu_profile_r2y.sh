#!/bin/bash
# refresh of the evidence pass after the BatchNorm kernel changes: launch list + DRAM bytes of every kernel of one step,
# `--set full` of the BatchNorm families only (the conv captures of r2s stand: those kernels did not change)
TAG=r2y
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --profile-steps 0 --no-cpu-baseline --caller eager --nvtx-step"
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/${TAG}_step_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?" > gpurun_out/${TAG}_rc.txt
ncu --profile-from-start off --set full --clock-control none \
    -k regex:'bn_act|col_reduce|bn_finalize|wgrad_reduce' -o /tmp/${TAG}_prof_step $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/${TAG}_rc.txt
ncu -i /tmp/${TAG}_prof_step.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_step_raw.csv 2>/dev/null
gzip -f gpurun_out/${TAG}_prof_step_raw.csv
du -sh gpurun_out; cat gpurun_out/${TAG}_rc.txt
