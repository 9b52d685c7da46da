#!/bin/bash
# round-2 final evidence on the final build, one GPU: smoke(), every bench line (train graph / eager / stock caller / fp32 mode, configs 3, 4, 5, loss sweep),
# the per-launch dump of one step and the per-layer roofline table
mkdir -p gpurun_out
T=r2w
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
DP_BENCH_DUMP=gpurun_out/${T}_step_dump.txt timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_train.json 2> gpurun_out/${T}_bench_train.err; echo "train rc=$?" > gpurun_out/${T}_rc.txt
timeout 300 python bench.py --caller stock --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_train_stock_caller.json 2> gpurun_out/${T}_stock.err; echo "stock rc=$?" >> gpurun_out/${T}_rc.txt
timeout 300 python bench.py --caller eager --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_train_eager.json 2> gpurun_out/${T}_eager.err; echo "eager rc=$?" >> gpurun_out/${T}_rc.txt
timeout 400 python bench.py --mode fp32 --steps 3 --warmup 3 --no-cpu-baseline --caller eager > gpurun_out/${T}_bench_train_fp32_mode.json 2> gpurun_out/${T}_fp32.err; echo "fp32 rc=$?" >> gpurun_out/${T}_rc.txt
for w in infer slowfast multimodal loss; do
timeout 500 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_$w.err; echo "$w rc=$?" >> gpurun_out/${T}_rc.txt
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_ref.err; echo "ref rc=$?" >> gpurun_out/${T}_rc.txt
python scripts/layer_roofline.py gpurun_out/${T}_step_dump.txt gpurun_out/${T}_layer_roofline.md | tail -3
for f in train train_stock_caller train_eager train_fp32_mode infer slowfast multimodal loss reference_arm; do echo "== $f"; tail -c 700 gpurun_out/${T}_bench_$f.json | head -c 400; echo; done
cat gpurun_out/${T}_rc.txt
