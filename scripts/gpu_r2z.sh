# traversal direction of the BatchNorm passes (L2 carry-over between consecutive kernels), streaming loads, evict-first
# activation loads in the gather kernel, programmatic dependent launch of the small kernels: kernel tests, then an
# in-process A/B on the whole step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q --timeout 180 -k "bn_act or fused_finalize or dgrad_bnstats or bn_" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_model.py -x -q --timeout 600 2>&1 | tail -3
timeout 600 python scripts/option_ab.py --out gpurun_out/r2z_option_ab.txt \
  base "bn_sweep=0" apply "bn_sweep=1" apply_red "bn_sweep=3" full "bn_sweep=7" \
  full_cs "bn_sweep=7,bn_cs=3" full_cs1 "bn_sweep=7,bn_cs=1" full_hint "bn_sweep=7,tc_l2hint=1" \
  full_all "bn_sweep=7,bn_cs=3,tc_l2hint=1" full_pdl "bn_sweep=7,pdl_small=1" base2 "bn_sweep=0" full2 "bn_sweep=7" 2>&1 | tail -14
