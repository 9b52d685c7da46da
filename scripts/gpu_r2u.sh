#!/bin/bash
# round-2 final evidence, two GPUs: multi-rank parity on real NCCL, train / stock caller / multimodal / inference lines
mkdir -p gpurun_out
T=r2u
run2() { n=$1; port=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 "$@" > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err; echo "$n rc=$?" >> gpurun_out/${T}_rc.txt
  tail -c 900 gpurun_out/${T}_$n.json | head -c 500; echo; tail -2 gpurun_out/${T}_$n.err
}
: > gpurun_out/${T}_rc.txt
timeout 600 python -m pytest tests/test_gpu_multirank.py -q -s --timeout 500 2>&1 | tail -5
run2 multirank_check_2gpu 29531 --check
run2 bench_train_2gpu 29532 --steps 10 --warmup 3
run2 bench_train_stock_caller_2gpu 29533 --caller stock --steps 10 --warmup 3
run2 bench_multimodal_2gpu 29534 --workload multimodal --steps 10 --warmup 3
run2 bench_infer_2gpu 29535 --workload infer --steps 5 --warmup 2
cat gpurun_out/${T}_rc.txt
