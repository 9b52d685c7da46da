# PDL A/B (graph replay), correctness of the model tests under PDL, per-role cycle accounting of the stage-1/2 layers
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_base.log 2>gpurun_out/r2g_bench_base.err
DP_OPTIONS=pdl=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_pdl.log 2>gpurun_out/r2g_bench_pdl.err
python - <<'PY'
import json
for n in ("base", "pdl"):
    try:
        d=json.loads(open(f"gpurun_out/r2g_bench_{n}.log").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["config"]["final_loss"])
    except Exception as e:
        print(n, "FAILED", e); print(open(f"gpurun_out/r2g_bench_{n}.err").read()[-1500:])
PY
DP_OPTIONS=pdl=1 python -m pytest tests/test_gpu_model.py tests/test_gpu_paths.py -x -q 2>&1 | tail -3
python scripts/role_profile.py 2>&1 | grep -v "^\[tc_gather\]\|^\[wgrad" > gpurun_out/r2g_roles.txt; cat gpurun_out/r2g_roles.txt
