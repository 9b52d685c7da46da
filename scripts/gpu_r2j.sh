# decomposition of the stage-1/2 gather kernels: no epilogue movement (1), no TMA loads (2), one MMA per load (4)
VARIANTS="{};{'tc_dbg_skip':(1,0)};{'tc_dbg_skip':(2,0)};{'tc_dbg_skip':(4,0)};{'tc_dbg_skip':(3,0)};{'tc_dbg_skip':(5,0)};{'tc_dbg_skip':(6,0)};{'tc_dbg_skip':(7,0)}" python scripts/role_variants.py 2>&1 | tee gpurun_out/r2j_variants.txt
python -m pytest tests/test_gpu_kernels.py -x -q -k "fused_finalize" 2>&1 | tail -2
