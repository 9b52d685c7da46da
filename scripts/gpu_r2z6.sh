mkdir -p gpurun_out
B=64 timeout 300 python scripts/role_profile.py 2>&1 | grep -v "^\[tc_" | tee gpurun_out/r2z6_role_cycles.txt | grep -A3 "conv3.spatial\|conv3.temporal"
