python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -x 2>&1 | tail -4
DP_DEBUG_PLAN=1 python scripts/strided_dgrad_bench.py 2>&1 | grep -E "x128 src C=128 taps=4|x256 src|x288 src|x480 src|per-class|sum:" | awk '!seen[$0]++'
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench.log 2>gpurun_out/r2l_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2l_bench.log").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], {k:v["ms_per_step"] for k,v in d["kernels"].items()})
PY
tail -3 gpurun_out/r2l_bench.err
