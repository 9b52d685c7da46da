#!/bin/bash
mkdir -p gpurun_out
T=r2e
python -m pytest tests/test_gpu_multirank.py -q -s > gpurun_out/${T}_t_multirank.log 2>&1; echo "multirank rc=$?" > gpurun_out/${T}_rc.txt
tail -40 gpurun_out/${T}_t_multirank.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${T}_bench2.log 2> gpurun_out/${T}_bench2.err; echo "bench2 rc=$?" >> gpurun_out/${T}_rc.txt
tail -c 1800 gpurun_out/${T}_bench2.log; tail -3 gpurun_out/${T}_bench2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --workload multimodal --steps 10 --warmup 3 > gpurun_out/${T}_mm2.log 2> gpurun_out/${T}_mm2.err; echo "mm2 rc=$?" >> gpurun_out/${T}_rc.txt
tail -c 1200 gpurun_out/${T}_mm2.log; tail -3 gpurun_out/${T}_mm2.err
python bench.py --workload multimodal --steps 10 --warmup 3 > gpurun_out/${T}_mm1.log 2> gpurun_out/${T}_mm1.err; echo "mm1 rc=$?" >> gpurun_out/${T}_rc.txt
tail -c 800 gpurun_out/${T}_mm1.log; tail -3 gpurun_out/${T}_mm1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --caller stock --steps 10 --warmup 3 > gpurun_out/${T}_stock2.log 2> gpurun_out/${T}_stock2.err; echo "stock2 rc=$?" >> gpurun_out/${T}_rc.txt
tail -c 600 gpurun_out/${T}_stock2.log | head -c 300; tail -3 gpurun_out/${T}_stock2.err
cat gpurun_out/${T}_rc.txt
