mkdir -p gpurun_out
timeout 300 python scripts/split_bench.py 2>&1 | tee gpurun_out/r2z4_split_bench.txt | tail -14
