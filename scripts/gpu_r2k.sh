VARIANTS="{};{'tc_mt':(1,0)};{'tc_mt':(2,0)};{'tc_mt':(1,0),'tc_st_bufs':(1,2)}" python scripts/role_variants.py 2>&1 | tee gpurun_out/r2k_variants.txt
