#!/bin/bash
# bounded `ncu --set full` refresh of the LARGE BatchNorm launches of one step on the final build (the r2y attempt to
# capture all 187 BatchNorm launches ran into the call's time limit: ~8 s per captured launch with 20 GB resident):
# the first 6 forward applies (stem + conv2 stage), the last 8 backward applies and the last 8 backward reductions
TAG=r2x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --profile-steps 0 --no-cpu-baseline --caller eager --nvtx-step"
cap() { name=$1; shift
  timeout 420 ncu --profile-from-start off --set full --clock-control none "$@" -o /tmp/${TAG}_$name $CMD > gpurun_out/${TAG}_ncu_$name.log 2>&1
  echo "$name rc=$?" >> gpurun_out/${TAG}_rc.txt
  ncu -i /tmp/${TAG}_$name.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_${name}_raw.csv 2>/dev/null
  gzip -f gpurun_out/${TAG}_prof_${name}_raw.csv
}
: > gpurun_out/${TAG}_rc.txt
cap apply -k regex:bn_act_apply_kernel --launch-count 6
cap bwd_apply -k regex:bn_act_bwd_apply_kernel --launch-skip 24 --launch-count 8
cap bwd_reduce -k regex:col_reduce2_kernel --launch-skip 19 --launch-count 8
du -sh gpurun_out; cat gpurun_out/${TAG}_rc.txt
