#!/bin/bash
# 2-GPU pass: multi-rank parity on real NCCL, 2-GPU bench lines (train + multimodal), plus the single-GPU tests added since
mkdir -p gpurun_out
T=r2d
python -m pytest tests/test_gpu_parity_extra.py tests/test_gpu_multimodal.py -q -s > gpurun_out/${T}_t_extra.log 2>&1; echo "extra rc=$?" > gpurun_out/${T}_rc.txt
grep -E "passed|failed|teacher|features rel" gpurun_out/${T}_t_extra.log | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --check --gpus 2 > gpurun_out/${T}_check2.log 2> gpurun_out/${T}_check2.err; echo "check rc=$?" >> gpurun_out/${T}_rc.txt
tail -c 3000 gpurun_out/${T}_check2.log; tail -5 gpurun_out/${T}_check2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${T}_bench2.log 2> gpurun_out/${T}_bench2.err; echo "bench2 rc=$?" >> gpurun_out/${T}_rc.txt
tail -c 1500 gpurun_out/${T}_bench2.log; tail -3 gpurun_out/${T}_bench2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --workload multimodal --steps 10 --warmup 3 > gpurun_out/${T}_mm2.log 2> gpurun_out/${T}_mm2.err; echo "mm2 rc=$?" >> gpurun_out/${T}_rc.txt
tail -c 1200 gpurun_out/${T}_mm2.log; tail -3 gpurun_out/${T}_mm2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --workload infer --steps 5 --warmup 2 > gpurun_out/${T}_inf2.log 2> gpurun_out/${T}_inf2.err; echo "inf2 rc=$?" >> gpurun_out/${T}_rc.txt
tail -c 600 gpurun_out/${T}_inf2.log; tail -3 gpurun_out/${T}_inf2.err
DP_BENCH_DUMP=gpurun_out/${T}_step_dump.txt python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench1.log 2>&1
python bench.py --workload multimodal --steps 10 --warmup 3 > gpurun_out/${T}_mm1.log 2> gpurun_out/${T}_mm1.err; echo "mm1 rc=$?" >> gpurun_out/${T}_rc.txt
tail -c 800 gpurun_out/${T}_mm1.log
cat gpurun_out/${T}_rc.txt
