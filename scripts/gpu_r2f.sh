# round 2, session 3: state check — all GPU tests, bench with the per-launch dump, per-layer table
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
DP_BENCH_DUMP=gpurun_out/r2f_step_dump.txt python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.log 2>gpurun_out/r2f_bench.err
tail -3 gpurun_out/r2f_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2f_bench.log").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], {k:v["ms_per_step"] for k,v in d["kernels"].items()})
PY
python scripts/layer_roofline.py gpurun_out/r2f_step_dump.txt gpurun_out/r2f_layer_roofline.md 2>&1 | tail -3
