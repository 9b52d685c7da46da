# sub-tile staging ring with per-tile / per-sub-tile hand-off: kernel tests, per-layer timing, whole-step threshold sweep
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q --timeout 180 2>&1 | tail -3
VARIANTS="{}" timeout 300 python scripts/role_variants.py 2>&1 | tee gpurun_out/r2o_variants.txt
run() { n=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2o_bench_$n.log 2>gpurun_out/r2o_bench_$n.err
  python - "$n" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2o_bench_{n}.log").read().strip().splitlines()[-1])
    print(f"{n:10s}", d["ms_per_step"], d["gpu_launches"], d["config"]["final_loss"], {k[:12]:round(v["ms_per_step"],3) for k,v in d["kernels"].items()}, flush=True)
except Exception as e:
    print(n, "FAILED", e); print(open(f"gpurun_out/r2o_bench_{n}.err").read()[-800:])
PY
}
run default A=1
run sub0 DP_OPTIONS=tc_pub_sub_min=0
run sub48 DP_OPTIONS=tc_pub_sub_min=48
run sub96 DP_OPTIONS=tc_pub_sub_min=96
run never DP_OPTIONS=tc_pub_sub_min=999
run default2 A=1
