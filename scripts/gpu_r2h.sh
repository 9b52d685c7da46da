# fused BatchNorm finalisation (last CTA done): kernel tests, model tests, step-time A/B; wgrad items per SM A/B
python -m pytest tests/test_gpu_kernels.py -x -q -k "fused_finalize or dgrad_bnstats or bn_act" 2>&1 | tail -4
python -m pytest tests/test_gpu_model.py tests/test_gpu_paths.py tests/test_gpu_parity_extra.py -x -q 2>&1 | tail -3
run() { # name, env...
  n=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench_$n.log 2>gpurun_out/r2h_bench_$n.err
  python - "$n" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2h_bench_{n}.log").read().strip().splitlines()[-1])
    print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["config"]["final_loss"], {k:v["ms_per_step"] for k,v in d["kernels"].items()})
except Exception as e:
    print(n, "FAILED", e); print(open(f"gpurun_out/r2h_bench_{n}.err").read()[-1500:])
PY
}
run fin DP_FUSE_FIN=1
run nofin DP_FUSE_FIN=0
run wg1 DP_OPTIONS=wg_items=1
run wg3 DP_OPTIONS=wg_items=3
