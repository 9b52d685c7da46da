# programmatic dependent launch of the small finalize / split-reduce kernels: interleaved in-process A/B on the whole step
mkdir -p gpurun_out
timeout 600 python scripts/option_ab.py --reps 5 --out gpurun_out/r2z2_option_ab.txt \
  base "pdl_small=0" pdl_small "pdl_small=1" base "pdl_small=0" pdl_small "pdl_small=1" base "pdl_small=0" pdl_small "pdl_small=1" 2>&1 | tail -8
