#!/bin/bash
# Round-2 evidence pass (one GPU): launch list + DRAM bytes of EVERY kernel of one eager training step (between
# cudaProfilerStart/Stop in bench.py --nvtx-step: forward and backward threads alike), a `--set full` capture of the conv / BN / loss kernel families of the same step (exported to CSV, the
# report itself is too large to bring back), and one `--set full --import-source on` capture of the dominant launch
# (the conv2 spatial data gradient) whose .ncu-rep is kept.
# usage: scripts/gpu_profile_r2.sh [tag]     (files land in gpurun_out/<tag>_*)
TAG=${1:-r2}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --profile-steps 0 --no-cpu-baseline --caller eager --nvtx-step"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/${TAG}_step_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?" > gpurun_out/${TAG}_rc.txt
ncu --profile-from-start off --set full --clock-control none \
    -k regex:'tc_gather_gemm|wgrad_tc_kernel|bn_act|col_reduce|loss_kernel' -o /tmp/${TAG}_prof_step $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/${TAG}_rc.txt
ncu -i /tmp/${TAG}_prof_step.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_step_raw.csv 2>/dev/null
ls -la /tmp/${TAG}_prof_step.ncu-rep gpurun_out/${TAG}_prof_step_raw.csv
# one dominant launch with source counters: the third gather launch of the step = conv2.block1.conv1.spatio_conv forward
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:tc_gather_gemm --launch-skip 2 -c 1 -o gpurun_out/${TAG}_prof_top $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
echo "ncu top rc=$?" >> gpurun_out/${TAG}_rc.txt
ncu -i gpurun_out/${TAG}_prof_top.ncu-rep --page source --csv > gpurun_out/${TAG}_prof_top_source.csv 2>/dev/null
gzip -f gpurun_out/${TAG}_prof_step_raw.csv
du -sh gpurun_out
cat gpurun_out/${TAG}_rc.txt
