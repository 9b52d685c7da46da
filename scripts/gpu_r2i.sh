# whole-step planner option sweep (graph replay, no e2e leg) + the failing fused-finalize test in detail
python -m pytest tests/test_gpu_kernels.py -x -q -k "fused_finalize" 2>&1 | grep -E "Error|assert|passed|failed|^E " | head -20
run() { # name, env...
  n=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2i_bench_$n.log 2>gpurun_out/r2i_bench_$n.err
  python - "$n" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2i_bench_{n}.log").read().strip().splitlines()[-1])
    print(f"{n:14s}", d["ms_per_step"], d["gpu_launches"], {k[:12]:round(v["ms_per_step"],3) for k,v in d["kernels"].items()}, flush=True)
except Exception as e:
    print(n, "FAILED", e); print(open(f"gpurun_out/r2i_bench_{n}.err").read()[-800:])
PY
}
run base A=1
run mt1 DP_OPTIONS=tc_mt=1
run mt2 DP_OPTIONS=tc_mt=2
run acc4_0 DP_OPTIONS=tc_acc4=0
run stbuf1 DP_OPTIONS=tc_st_bufs=1
run lps4 DP_OPTIONS=tc_lps_max=4
run lps2 DP_OPTIONS=tc_lps_max=2
run dual0 DP_OPTIONS=tc_dual_mma=0
run resid0 DP_OPTIONS=tc_resident=0
run tail0 DP_OPTIONS=tc_tail=0
run bws0 DP_OPTIONS=tc_bwd_stats_max=0
run bws32 DP_OPTIONS=tc_bwd_stats_max=32
run bws96 DP_OPTIONS=tc_bwd_stats_max=96
run cls0 DP_OPTIONS=tc_classes=0
run stack0 DP_OPTIONS=wg_stack=0
run wghalo0 DP_OPTIONS=wg_halo=0
run regst0 DP_OPTIONS=tc_reg_stats=0
