# BatchNorm passes with raw loads hoisted (more rows / vectors in flight): kernel tests + step time (twice)
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q --timeout 180 -k "bn_act or fused_finalize or dgrad_bnstats" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_extra.py -x -q --timeout 600 2>&1 | tail -3
for i in 1 2 3; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2v_bench$i.log 2>gpurun_out/r2v_bench$i.err
python - $i <<'PY'
import json, sys
i=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r2v_bench{i}.log").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["gpu_launches"], d["config"]["final_loss"], {k:(v["ms_per_step"], v["gbs"]) for k,v in d["kernels"].items()})
except Exception as e:
    print("FAILED", e); print(open(f"gpurun_out/r2v_bench{i}.err").read()[-1500:])
PY
done
