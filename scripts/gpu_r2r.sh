# row-pair packed stem: kernel test, role profile, model / inference / slowfast tests, step time
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q --timeout 180 -k "stem" 2>&1 | tail -3
timeout 200 python scripts/role_stem.py 2>&1 | grep -v "^\[" | tee gpurun_out/r2r_role_stem.txt
timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_paths.py tests/test_gpu_slowfast.py tests/test_gpu_parity_extra.py -x -q --timeout 600 2>&1 | tail -3
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2r_bench$i.log 2>gpurun_out/r2r_bench$i.err
python - $i <<'PY'
import json, sys
i=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2r_bench{i}.log").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["config"]["final_loss"], {k:v["ms_per_step"] for k,v in d["kernels"].items()})
except Exception as e:
    print("FAILED", e); print(open(f"gpurun_out/r2r_bench{i}.err").read()[-1500:])
PY
done
