# staged-tile statistics kept in registers across tiles (N = 144 forward convs): kernel tests, per-layer times, whole-step A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q --timeout 300 -k "channel_split or tcgen05 or fused_finalize" 2>&1 | tail -3
OPT=tc_stats_keep timeout 300 python scripts/split_bench.py 2>&1 | tee gpurun_out/r2z5_stats_keep_bench.txt | tail -13
timeout 600 python scripts/option_ab.py --reps 4 --out gpurun_out/r2z5_option_ab.txt \
  base "tc_stats_keep=0" keep "tc_stats_keep=1" base "tc_stats_keep=0" keep "tc_stats_keep=1" 2>&1 | tail -5
