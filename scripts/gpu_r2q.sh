# full GPU suite on the new planner (staging ring, fine stages), per-layer timing, step time
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -4
VARIANTS="{}" timeout 300 python scripts/role_variants.py 2>&1 | tee gpurun_out/r2q_variants.txt
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_bench$i.log 2>gpurun_out/r2q_bench$i.err
python - $i <<'PY'
import json, sys
i=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2q_bench{i}.log").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["config"]["final_loss"], {k:v["ms_per_step"] for k,v in d["kernels"].items()})
except Exception as e:
    print("FAILED", e); print(open(f"gpurun_out/r2q_bench{i}.err").read()[-1500:])
PY
done
