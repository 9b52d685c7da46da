#!/bin/bash
mkdir -p gpurun_out
T=r2b
python -m pytest tests/test_gpu_parity_extra.py -q -x -s > gpurun_out/${T}_t_extra.log 2>&1; echo "extra rc=$?" > gpurun_out/${T}_rc.txt
tail -30 gpurun_out/${T}_t_extra.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err; echo "bench rc=$?" >> gpurun_out/${T}_rc.txt
timeout 300 python bench.py --caller stock --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_stock.log 2> gpurun_out/${T}_bench_stock.err; echo "stock rc=$?" >> gpurun_out/${T}_rc.txt
timeout 300 python bench.py --caller eager --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_eager.log 2> gpurun_out/${T}_bench_eager.err; echo "eager rc=$?" >> gpurun_out/${T}_rc.txt
timeout 400 python bench.py --mode fp32 --steps 3 --warmup 3 --no-cpu-baseline --caller eager > gpurun_out/${T}_bench_fp32.log 2> gpurun_out/${T}_bench_fp32.err; echo "fp32 rc=$?" >> gpurun_out/${T}_rc.txt
timeout 400 python bench.py --workload infer --steps 5 --warmup 2 > gpurun_out/${T}_bench_infer.log 2> gpurun_out/${T}_bench_infer.err; echo "infer rc=$?" >> gpurun_out/${T}_rc.txt
timeout 300 python bench.py --workload slowfast --steps 10 --warmup 3 > gpurun_out/${T}_bench_slowfast.log 2> gpurun_out/${T}_bench_slowfast.err; echo "slowfast rc=$?" >> gpurun_out/${T}_rc.txt
timeout 300 python bench.py --workload multimodal --steps 10 --warmup 3 > gpurun_out/${T}_bench_multimodal.log 2> gpurun_out/${T}_bench_multimodal.err; echo "multimodal rc=$?" >> gpurun_out/${T}_rc.txt
timeout 300 python bench.py --workload loss --steps 10 --warmup 3 > gpurun_out/${T}_bench_loss.log 2> gpurun_out/${T}_bench_loss.err; echo "loss rc=$?" >> gpurun_out/${T}_rc.txt
for f in bench bench_stock bench_eager bench_fp32 bench_infer bench_slowfast bench_multimodal bench_loss; do echo "== $f"; tail -c 1500 gpurun_out/${T}_$f.log; tail -3 gpurun_out/${T}_$f.err; done
cat gpurun_out/${T}_rc.txt
