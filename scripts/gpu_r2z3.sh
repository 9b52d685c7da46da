# forward output-channel split (resident weights for the 64 -> 144 spatial convs): kernel tests, then in-process A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q --timeout 300 -k "channel_split or tcgen05 or fused_finalize or dgrad_bnstats" 2>&1 | tail -4
timeout 600 python scripts/option_ab.py --reps 4 --out gpurun_out/r2z3_option_ab.txt \
  base "tc_nsplit=0" nsplit "tc_nsplit=1" base "tc_nsplit=0" nsplit "tc_nsplit=1" 2>&1 | tail -6
