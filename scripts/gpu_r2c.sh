#!/bin/bash
mkdir -p gpurun_out
T=r2c
python -m pytest tests/test_gpu_parity_extra.py tests/test_gpu_paths.py tests/test_gpu_slowfast.py -q -s > gpurun_out/${T}_t_extra.log 2>&1; echo "extra rc=$?" > gpurun_out/${T}_rc.txt
tail -8 gpurun_out/${T}_t_extra.log
python -m pytest tests/test_gpu_kernels.py -q -x > gpurun_out/${T}_t_kernels.log 2>&1; echo "kernels rc=$?" >> gpurun_out/${T}_rc.txt
tail -3 gpurun_out/${T}_t_kernels.log
timeout 900 python scripts/error_attribution.py --tag ${T} > gpurun_out/${T}_attr.log 2>&1; echo "attr rc=$?" >> gpurun_out/${T}_rc.txt
tail -8 gpurun_out/${T}_attr.log
timeout 900 python tests/convergence_ab.py --out gpurun_out/${T}_convergence.json > gpurun_out/${T}_conv.log 2>&1; echo "conv rc=$?" >> gpurun_out/${T}_rc.txt
tail -60 gpurun_out/${T}_conv.log
for w in infer slowfast multimodal loss; do
timeout 400 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/${T}_bench_$w.log 2> gpurun_out/${T}_bench_$w.err; echo "$w rc=$?" >> gpurun_out/${T}_rc.txt
echo "== $w"; tail -c 1200 gpurun_out/${T}_bench_$w.log; tail -3 gpurun_out/${T}_bench_$w.err
done
cat gpurun_out/${T}_rc.txt
