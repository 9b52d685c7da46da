#!/bin/bash
# full GPU test pass, bench, ncu launch list and full captures of the conv / BN kernel families
# usage: scripts/gpu_profile.sh [tag]   (files land in gpurun_out/<tag>_*)
TAG=${1:-cur}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_t_gpu.log 2>&1; echo "pytest rc=$?" > gpurun_out/${TAG}_rc.txt
tail -3 gpurun_out/${TAG}_t_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.log 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?" >> gpurun_out/${TAG}_rc.txt
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --profile-steps 0 --no-cpu-baseline --no-graph"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1600 -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/${TAG}_rc.txt
ncu --set full --clock-control none --import-source on -k regex:tc_gather_gemm -s 66 -c 8 -o gpurun_out/${TAG}_prof_gather $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu gather rc=$?" >> gpurun_out/${TAG}_rc.txt
ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 32 -c 8 -o gpurun_out/${TAG}_prof_wgrad $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
echo "ncu wgrad rc=$?" >> gpurun_out/${TAG}_rc.txt
ncu --set full --clock-control none --import-source on -k regex:bn_act -s 100 -c 6 -o gpurun_out/${TAG}_prof_bn $CMD > gpurun_out/${TAG}_ncu4.log 2>&1
echo "ncu bn rc=$?" >> gpurun_out/${TAG}_rc.txt
cat gpurun_out/${TAG}_rc.txt
