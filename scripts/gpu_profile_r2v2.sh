#!/bin/bash
# source-level ncu capture (stall samples per line) of the drain-bound gather kernels on the final build:
# launches 4..10 of scripts/role_profile.py = conv2 spatial fwd x2, its dgrad x2, conv2 temporal fwd x2, its dgrad (first run)
TAG=r2v2
mkdir -p gpurun_out
B=64 timeout 500 ncu --set full --import-source on --clock-control none -k regex:tc_gather_gemm --launch-skip 4 --launch-count 7 \
    -o /tmp/${TAG}_src python scripts/role_profile.py > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?" > gpurun_out/${TAG}_rc.txt
ncu -i /tmp/${TAG}_src.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_src.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_source.csv 2>gpurun_out/${TAG}_source.err || \
ncu -i /tmp/${TAG}_src.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2>>gpurun_out/${TAG}_source.err
gzip -f gpurun_out/${TAG}_raw.csv gpurun_out/${TAG}_source.csv
ls -la gpurun_out/${TAG}_*; cat gpurun_out/${TAG}_rc.txt; tail -3 gpurun_out/${TAG}_ncu.log
