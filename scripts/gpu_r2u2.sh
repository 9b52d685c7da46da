#!/bin/bash
# two GPUs on the final build: multi-rank parity check on real NCCL and the training line
mkdir -p gpurun_out
T=r2u2
: > gpurun_out/${T}_rc.txt
run2() { n=$1; port=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 "$@" > gpurun_out/${T}_$n.json 2> gpurun_out/${T}_$n.err; echo "$n rc=$?" >> gpurun_out/${T}_rc.txt
  tail -c 1200 gpurun_out/${T}_$n.json | head -c 600; echo; tail -2 gpurun_out/${T}_$n.err
}
run2 multirank_check_2gpu 29541 --check
run2 bench_train_2gpu 29542 --steps 10 --warmup 3 --no-cpu-baseline
cat gpurun_out/${T}_rc.txt
