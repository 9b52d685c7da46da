# sub-tile staging ring: kernel tests (with a per-test timeout), per-layer timing, whole step
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q --timeout 180 2>&1 | tail -5
VARIANTS="{}" timeout 300 python scripts/role_variants.py 2>&1 | tee gpurun_out/r2n_variants.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench.log 2>gpurun_out/r2n_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2n_bench.log").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["config"]["final_loss"], {k:v["ms_per_step"] for k,v in d["kernels"].items()})
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/r2n_bench.err").read()[-1500:])
PY
