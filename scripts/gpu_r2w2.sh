#!/bin/bash
# final build after the wide-tile statistics change: every GPU test, smoke(), the default bench line (with the per-launch dump
# for the per-layer table) and the SlowFast / multimodal / inference lines
mkdir -p gpurun_out
T=r2w2
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
DP_BENCH_DUMP=gpurun_out/${T}_step_dump.txt timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_train.json 2> gpurun_out/${T}_bench_train.err; echo "train rc=$?" > gpurun_out/${T}_rc.txt
for w in slowfast multimodal infer; do
timeout 500 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_$w.err; echo "$w rc=$?" >> gpurun_out/${T}_rc.txt
done
python scripts/layer_roofline.py gpurun_out/${T}_step_dump.txt gpurun_out/${T}_layer_roofline.md | tail -3
for f in train slowfast multimodal infer; do echo "== $f"; head -c 330 gpurun_out/${T}_bench_$f.json; echo; done
cat gpurun_out/${T}_rc.txt
